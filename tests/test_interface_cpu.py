"""CPU: Interface orchestration (chunking, edge anchors, padding, codebook stacking, time stretch, feedback passes)
against the oracle's restatement of reference vampnet/interface.py:328-562.  Both sides are driven by the same
deterministic stand-in for generate() — a pure function of (start_tokens, mask) — so no CUDA is needed and any
difference is in the host logic.  The GPU twin of this file (tests/test_gpu_interface.py) uses the real generate."""
import pytest
import torch

from oracle import vampnet_oracle as vo
from vampnet_b200.interface import Interface

MASK_TOKEN = 1024


def fake_generate(start_tokens, mask, salt):
    """Deterministic stand-in: a masked slot becomes a hash of (its position, the chunk's kept tokens, salt); kept
    slots pass through.  Depends on the whole chunk so wrong chunk boundaries or anchors change the result."""
    B, C, T = start_tokens.shape
    if mask is None:
        mask = torch.ones_like(start_tokens)
    kept = (start_tokens * (1 - mask)).sum(dim=(1, 2), keepdim=True)
    pos = torch.arange(C)[None, :, None] * 131 + torch.arange(T)[None, None, :] * 17
    fill = (kept * 7 + pos + salt + mask.sum(dim=(1, 2), keepdim=True) * 3) % 1024
    return torch.where(mask.bool(), fill, start_tokens)


class StubModel:
    """Duck-types the attributes Interface reads from a VampNet (interface.py:328-452)."""

    def __init__(self, n_codebooks, n_conditioning_codebooks, salt):
        self.n_codebooks, self.n_conditioning_codebooks, self.salt = n_codebooks, n_conditioning_codebooks, salt
        self.mask_token = MASK_TOKEN
        self.chunk_size_s = None
        self.calls = []

    def to(self, device):
        return self

    def generate(self, codec=None, time_steps=None, start_tokens=None, mask=None, return_signal=True, **kwargs):
        assert return_signal is False
        self.calls.append(dict(time_steps=time_steps, shape=tuple(start_tokens.shape), kwargs=kwargs))
        return fake_generate(start_tokens, mask, self.salt)


class StubCodec:
    sample_rate, hop_length = 44100, 768

    def to(self, device):
        return self


def make_iface(coarse_s=0.6, c2f_s=0.25):
    coarse, c2f = StubModel(4, 0, salt=5), StubModel(14, 4, salt=9)
    return Interface.from_models(StubCodec(), coarse, c2f, device="cpu", coarse_chunk_size_s=coarse_s,
                                 coarse2fine_chunk_size_s=c2f_s)


def rand_case(B, T, seed, keep_every=7):
    g = torch.Generator().manual_seed(seed)
    z = torch.randint(0, 1024, (B, 14, T), generator=g)
    mask = torch.ones_like(z)
    mask[:, :, ::keep_every] = 0
    return z, mask


@pytest.mark.parametrize("T", [1, 34, 35, 36, 83, 140])
def test_coarse_vamp_matches_oracle(T):
    iface = make_iface()
    z, mask = rand_case(2, T, seed=T)
    if T > 70:
        mask[:, :, 70:] = 1  # a fully-masked chunk gets no edge anchors (interface.py:410)
    chunk = iface.s2t(0.6)
    assert chunk == 35
    want, want_start = vo.coarse_vamp(z, mask, 4, chunk, MASK_TOKEN, lambda s, m: fake_generate(s, m, 5))
    got, got_start = iface.coarse_vamp(z, mask, return_mask=True, temperature=0.7)
    assert torch.equal(got, want) and torch.equal(got_start, want_start)
    assert torch.equal(iface.coarse_vamp(z, mask), want)
    assert torch.equal(got[:, 4:], z[:, 4:])
    assert len(iface.coarse.calls) == 2 * -(-T // chunk)
    assert all(c["time_steps"] == chunk and c["kwargs"] == {"temperature": 0.7} for c in iface.coarse.calls[:-(-T // chunk)])


def test_coarse_vamp_gen_fn_override():
    iface = make_iface()
    z, mask = rand_case(1, 50, seed=1)
    seen = []

    def gen_fn(codec, time_steps, start_tokens, mask, return_signal, **kw):
        seen.append(start_tokens.shape[-1])
        return start_tokens.clamp(max=1023)
    out = iface.coarse_vamp(z, mask, gen_fn=gen_fn)
    assert seen == [35, 15] and out.shape == z.shape and not iface.coarse.calls


@pytest.mark.parametrize("T,n_in", [(15, 14), (29, 14), (30, 4), (47, 14), (1, 4)])
@pytest.mark.parametrize("with_mask", [True, False])
def test_coarse_to_fine_matches_oracle(T, n_in, with_mask):
    iface = make_iface()
    z, mask = rand_case(2, T, seed=100 + T)
    z = z[:, :n_in]
    mask = mask if with_mask else None
    chunk = iface.s2t(0.25)
    assert chunk == 15
    gen = lambda s, m: fake_generate(s, m, 9)  # noqa: E731
    if with_mask:
        want, want_start = vo.coarse_to_fine(z, mask, 14, 4, chunk, MASK_TOKEN, gen)
        got, got_start = iface.coarse_to_fine(z, mask=mask, return_mask=True)
        assert torch.equal(got_start, want_start)
    else:
        # without a mask the reference cannot build the returned start tokens (apply_mask(None) asserts)
        pad = (-T) % chunk
        zp = torch.nn.functional.pad(z, (0, pad))
        if n_in < 14:
            zp = torch.cat([zp, torch.zeros(2, 14 - n_in, zp.shape[-1], dtype=torch.long)], 1)
        want = torch.cat([gen(zp[..., i:i + chunk], None) for i in range(0, zp.shape[-1], chunk)], -1)[..., :T]
        got = iface.coarse_to_fine(z, mask=None)
    assert torch.equal(got, want) and got.shape == (2, 14, T)
    assert all(c["time_steps"] == chunk and c["shape"][-1] == chunk and c["kwargs"] == {"cfg_guidance": None}
               for c in iface.c2f.calls)


def test_coarse_to_fine_requires_model():
    iface = make_iface()
    iface.c2f = None
    with pytest.raises(AssertionError, match="No coarse2fine model"):
        iface.coarse_to_fine(torch.zeros(1, 14, 4, dtype=torch.long))


@pytest.mark.parametrize("batch,feedback,stretch,T", [(1, 1, 1, 83), (3, 1, 1, 40), (2, 2, 1, 61), (2, 3, 2, 37),
                                                       (1, 1, 3, 20)])
def test_vamp_matches_oracle(batch, feedback, stretch, T):
    iface = make_iface()
    z, mask = rand_case(1, T, seed=7 * T + batch)
    want, want_mask = vo.vamp(z, mask, batch, feedback, stretch, 4, 14, 4, 35, 15, MASK_TOKEN,
                              lambda s, m: fake_generate(s, m, 5), lambda s, m: fake_generate(s, m, 9))
    got, got_mask = iface.vamp(z, mask, batch_size=batch, feedback_steps=feedback, time_stretch_factor=stretch,
                               return_mask=True, temperature=1.3)
    assert got.shape == (batch, 14, T * stretch)
    assert torch.equal(got, want) and torch.equal(got_mask, want_mask)
    assert torch.equal(iface.vamp(z, mask, batch_size=batch, feedback_steps=feedback, time_stretch_factor=stretch), want)
    # kwargs reach the coarse generate only; the fine stage is pinned (interface.py:545-551)
    assert all(c["kwargs"] == {"temperature": 1.3} for c in iface.coarse.calls[:1])
    assert all(c["kwargs"] == {"cfg_guidance": None, "typical_filtering": True, "_sampling_steps": 2}
               for c in iface.c2f.calls)


def test_units_and_unsupported_entry_points():
    iface = make_iface()
    assert iface.s2t(10) == 575 and iface.s2t(3) == 173
    assert abs(iface.s2t2s(1.0) - 58 * 768 / 44100) < 1e-12
    iface.set_chunk_size(4)
    assert iface.coarse.chunk_size_s == 4
    for fn in (Interface.default, iface.make_beat_mask):  # no cached checkpoints here; no beat tracker on this path
        with pytest.raises(RuntimeError):
            fn()
