"""GPU: Interface orchestration (A1-A4: chunked coarse_vamp, padded/chunked coarse_to_fine, vamp, decode) against
the oracle's restatement of reference vampnet/interface.py:328-562, both driven by the SAME generate function
(the CUDA path, greedy so that it is deterministic) — so any difference is in chunking / masking / stitching.
Plus the full encode -> build_mask -> vamp -> decode path of hello.py / app.py on synthetic weights."""
import math

import pytest
import torch

from oracle import dac_oracle as do
from oracle import vampnet_oracle as vo

pytestmark = pytest.mark.gpu

COARSE = dict(n_heads=4, n_layers=2, n_codebooks=4, n_conditioning_codebooks=0, embedding_dim=256)
C2F = dict(n_heads=4, n_layers=1, n_codebooks=14, n_conditioning_codebooks=4, embedding_dim=256)
CODEC = do.CodecConfig(encoder_dim=32, decoder_dim=512)
GREEDY = dict(sample_cutoff=-1.0, mask_temperature=0.0)


@pytest.fixture(scope="module")
def iface():
    from vampnet_b200.codec import DAC
    from vampnet_b200.interface import Interface
    from vampnet_b200.modules.transformer import VampNet
    codec = DAC(encoder_dim=CODEC.encoder_dim, encoder_rates=CODEC.encoder_rates, decoder_dim=CODEC.decoder_dim)
    codec.load_flat(do.make_codec_weights(CODEC, seed=0))
    coarse, c2f = VampNet(**COARSE), VampNet(**C2F)
    coarse.load_state_dict(vo.make_state_dict(vo.OracleConfig(**COARSE), seed=0), strict=False)
    c2f.load_state_dict(vo.make_state_dict(vo.OracleConfig(**C2F), seed=1), strict=False)
    return Interface.from_models(codec, coarse, c2f, device="cuda", coarse_chunk_size_s=0.6, coarse2fine_chunk_size_s=0.25)


def test_units(iface):
    assert iface.s2t(10) == 575 and iface.s2t(3) == 173  # hop 768 @ 44.1 kHz
    assert abs(iface.t2s(575) - 575 * 768 / 44100) < 1e-9
    assert iface.coarse.chunk_size_s == 0.6 and iface.c2f.chunk_size_s == 0.25


def test_coarse_vamp_chunking_matches_oracle(iface):
    g = torch.Generator().manual_seed(0)
    T = 83  # chunk_len = s2t(0.6) = 35 -> chunks 35 + 35 + 13
    z = torch.randint(0, 1024, (2, 14, T), generator=g).cuda()
    mask = torch.ones_like(z)
    mask[:, :, ::7] = 0
    mask[:, :, 70:] = 1  # last chunk fully masked: no edge anchors there (interface.py:410)
    chunk_len = iface.s2t(iface.coarse.chunk_size_s)
    assert chunk_len == 35

    def gen(cm, mc):
        return iface.coarse.generate(iface.codec, start_tokens=cm, mask=mc, _sampling_steps=3, return_signal=False,
                                     seed=3, **GREEDY)
    want, want_masked = vo.coarse_vamp(z.cpu(), mask.cpu(), 4, chunk_len, 1024,
                                       lambda cm, mc: gen(cm.cuda(), mc.cuda()).cpu())
    got, got_masked = iface.coarse_vamp(z, mask, return_mask=True, _sampling_steps=3, seed=3, **GREEDY)
    assert torch.equal(got.cpu(), want) and torch.equal(got_masked.cpu(), want_masked)
    assert torch.equal(got[:, 4:], z[:, 4:])  # fine codebooks ride along


def test_coarse_to_fine_padding_and_chunks_match_oracle(iface):
    g = torch.Generator().manual_seed(1)
    T = 40  # chunk_len = s2t(0.25) = 15 -> padded to 45, 3 chunks
    z = torch.randint(0, 1024, (2, 4, T), generator=g).cuda()  # only coarse codebooks: c2f appends 10 zero books
    mask = torch.ones(2, 4, T, dtype=torch.long).cuda()
    mask[:, :, ::5] = 0
    chunk_len = iface.s2t(iface.c2f.chunk_size_s)
    assert chunk_len == 15

    def gen(chunk, mc):
        return iface.c2f.generate(iface.codec, start_tokens=chunk, mask=mc, _sampling_steps=2, return_signal=False,
                                  seed=4, **GREEDY)
    # the reference pads a 4-codebook mask with value 1 on time only; codebook padding of z does not touch the mask,
    # so generate sees a (B,4,T) mask against (B,14,T) tokens -> the reference would fail; Interface.vamp always
    # passes a 14-codebook mask, which is what we test.
    mask14 = torch.ones(2, 14, T, dtype=torch.long).cuda()
    mask14[:, :4] = mask
    want, want_masked = vo.coarse_to_fine(z.cpu(), mask14.cpu(), 14, 4, chunk_len, 1024,
                                          lambda c, m: gen(c.cuda(), m.cuda()).cpu())
    got, got_masked = iface.coarse_to_fine(z, mask=mask14, return_mask=True, _sampling_steps=2, seed=4, **GREEDY)
    assert got.shape == (2, 14, T)
    assert torch.equal(got.cpu(), want) and torch.equal(got_masked.cpu(), want_masked)
    assert torch.equal(got[:, :4], z)  # conditioning codebooks untouched


def test_full_path_encode_vamp_decode(iface):
    from vampnet_b200.audio import AudioSignal
    torch.manual_seed(0)
    sr = 44100
    t = torch.arange(int(sr * 0.9)) / sr
    sig = AudioSignal((0.3 * torch.sin(2 * math.pi * 220 * t) + 0.05 * torch.randn_like(t))[None, None], sr)
    codes = iface.encode(sig)
    assert codes.dtype == torch.int64 and codes.shape == (1, 14, math.ceil(sig.signal_length / 768))
    mask = iface.build_mask(codes, sig, periodic_prompt=7, upper_codebook_mask=3)
    assert mask.shape == codes.shape and set(mask.unique().tolist()) <= {0, 1}
    z, mask_z = iface.vamp(codes, mask, batch_size=2, return_mask=True, _sampling_steps=4, seed=11)
    assert z.shape == (2, 14, codes.shape[-1]) and not (z == 1024).any()
    keep = (mask == 0).expand(2, -1, -1) & (torch.arange(14, device=z.device)[None, :, None] < 4)
    # chunk-edge frames may additionally be kept; frames the mask kept must be unchanged in the coarse books
    assert torch.equal(z[keep], codes.expand(2, -1, -1)[keep])
    assert mask_z.device.type == "cpu" and mask_z.shape == z.shape
    out = iface.decode(z)
    assert out.sample_rate == sr and out.samples.shape == (2, 1, codes.shape[-1] * 768)
    assert torch.isfinite(out.samples).all() and out.samples.abs().max() <= 1.0
    # decode == oracle decode of the same tokens (from_latents(from_codes) -> decoder)
    w = do.make_codec_weights(CODEC, seed=0)
    lat = torch.cat([w[f"quantizer.quantizers.{i}.codebook.weight"][z[:, i].cpu()].transpose(1, 2) for i in range(14)], 1)
    ref = do.decode(do.rvq_from_latents(lat, w, CODEC)[0], w, CODEC)["audio"]
    assert (out.samples.cpu() - ref).abs().max() < 5e-4


def test_chunk_loops_do_not_synchronise(iface):
    """between the first chunk's generate() and the end of coarse_vamp / coarse_to_fine nothing
    may wait for the device (the reference syncs twice per chunk in apply_mask plus once for the edge-anchor test).
    torch's sync debug mode turns any synchronising torch call into an error from the first chunk on."""
    g = torch.Generator().manual_seed(3)
    T = 83  # three coarse chunks (35 + 35 + 13), six c2f chunks of 15
    z = torch.randint(0, 1024, (2, 14, T), generator=g).cuda()
    mask = torch.ones_like(z)
    mask[:, :, ::7] = 0
    kw = dict(_sampling_steps=2, seed=1)
    iface.coarse_vamp(z, mask, **kw)                      # warm: workspaces, graphs, packed weights
    iface.coarse_to_fine(z, mask=mask, **kw)
    torch.cuda.synchronize()
    calls = []

    def armed(fn):
        def run(**k):
            if not calls:
                torch.cuda.set_sync_debug_mode("error")
            calls.append(1)
            return fn(**k)
        return run

    try:
        out = iface.coarse_vamp(z, mask, gen_fn=armed(iface.coarse.generate), **kw)
    finally:
        torch.cuda.set_sync_debug_mode("default")
    assert len(calls) == 3 and out.shape == z.shape
    calls.clear()
    orig = iface.c2f.generate
    iface.c2f.generate = armed(orig)
    try:
        fine = iface.coarse_to_fine(z, mask=mask, **kw)
    finally:
        torch.cuda.set_sync_debug_mode("default")
        del iface.c2f.generate
    assert len(calls) == 6 and fine.shape == z.shape
