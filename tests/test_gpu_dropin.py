"""GPU: the reference's hello.py:1-47 / app.py:160-250 call sequence, executed through the ``vampnet`` import names
against a synthetic model cache in the reference's on-disk layout (lac-layout codec checkpoint included): default()
-> load_finetuned -> to(device) -> encode -> build_mask -> vamp -> decode, and the result is checked against the
oracles (codec oracle for the tokens / waveform, Interface restatement for the token bookkeeping)."""
import math
import sys

import pytest
import torch

from oracle import dac_oracle as do

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cache(tmp_path_factory):
    import os
    from tests.dropin_cache import write_cache
    root = tmp_path_factory.mktemp("cache") / "models" / "vampnet"
    codec, sds = write_cache(root)
    old = os.environ.get("VAMPNET_MODELS_DIR")
    os.environ["VAMPNET_MODELS_DIR"] = str(root)
    for k in [k for k in sys.modules if k == "vampnet" or k.startswith("vampnet.")]:
        del sys.modules[k]
    yield root, codec, sds
    if old is None:
        os.environ.pop("VAMPNET_MODELS_DIR", None)
    else:
        os.environ["VAMPNET_MODELS_DIR"] = old


def test_hello_py_sequence(cache):
    root, codec_ref, sds = cache
    from tests.dropin_cache import CODEC
    import vampnet                                                     # hello.py:2
    from vampnet.interface import AudioSignal
    interface = vampnet.interface.Interface.default()                  # hello.py:6
    choices = interface.available_models()                             # hello.py:9
    assert "default" in choices and "opera" in choices
    interface.load_finetuned("default")                                # hello.py:23
    interface.to("cuda")                                               # app.py:163
    torch.manual_seed(0)
    sr = 44100
    t = torch.arange(int(sr * 1.3)) / sr
    signal = AudioSignal((0.3 * torch.sin(2 * math.pi * 330 * t) + 0.05 * torch.randn_like(t))[None, None], sr)
    codes = interface.encode(signal)                                   # hello.py:30
    T = math.ceil(signal.signal_length / 768)
    assert codes.shape == (1, 14, T) and codes.dtype == torch.int64 and codes.is_cuda
    # tokens == the codec oracle's on the same preprocessed audio, weights read back from the loaded (lac-layout) file
    from vampnet_b200.codec import _layout
    w = {k: interface.codec.params.get(k).detach().cpu() for k in _layout(interface.codec._cfg)}
    pre = interface._preprocess(signal.clone().to("cuda"))
    ref_codes = do.encode(pre.samples.cpu(), w, CODEC)["codes"]
    agree = (codes.cpu() == ref_codes).float().mean().item()
    assert agree > 0.97, agree        # split-bf16 tensor-core convolutions: an occasional near-tie in the RVQ search
    mask = interface.build_mask(codes, signal, periodic_prompt=13, upper_codebook_mask=3)   # hello.py:33-36
    assert mask.shape == codes.shape
    out = interface.vamp(codes, mask, return_mask=False, temperature=1.0, typical_filtering=False, debug=True,
                         _sampling_steps=4, seed=3)                    # hello.py:39-44 (fewer steps: tiny models)
    assert out.shape == codes.shape and not (out == 1024).any()
    keep = (mask == 0)
    keep[:, 3:] = False
    assert torch.equal(out[keep], codes[keep])                         # prompt frames of the coarse books survive
    sig = interface.decode(out)                                        # hello.py:47
    assert sig.sample_rate == sr and sig.samples.shape == (1, 1, T * 768)
    lat = torch.cat([w[f"quantizer.quantizers.{i}.codebook.weight"][out[:, i].cpu()].transpose(1, 2) for i in range(14)], 1)
    ref = do.decode(do.rvq_from_latents(lat, w, CODEC)[0], w, CODEC)["audio"]
    assert (sig.samples.cpu() - ref).abs().max() < 1e-3                # north-star: waveforms within 1e-3
    # fine-tuned checkpoints hot-swap into the live models (app.py:181) and change the result
    interface.load_finetuned("opera")
    out2 = interface.vamp(codes, mask, _sampling_steps=4, seed=3)
    assert out2.shape == out.shape and not torch.equal(out2, out)
    interface.load_finetuned("default")
    assert torch.equal(interface.vamp(codes, mask, _sampling_steps=4, seed=3, temperature=1.0), out)


def test_app_py_sequence_with_batch_and_mask_return(cache):
    """app.py:160-250: _preprocess, encode, build_mask with onset/dropout arguments at their UI defaults, vamp with
    batch_size=2 / feedback_steps / return_mask=True, decode, loudness restore."""
    import vampnet
    from vampnet.interface import Interface, signal_concat, AudioSignal
    from vampnet import mask as pmask
    interface = Interface.default(device="cuda")
    sr = 44100
    g = torch.Generator().manual_seed(1)
    sig = AudioSignal(0.2 * torch.randn(1, 2, int(sr * 0.8), generator=g), sr).to_mono()
    loudness = sig.loudness()
    sig = interface._preprocess(sig)
    codes = interface.encode(sig)
    mask = interface.build_mask(codes, sig=sig, periodic_prompt=7, onset_mask_width=0, _dropout=0.0, upper_codebook_mask=3)
    mask = pmask.codebook_mask(pmask.mask_and(mask, pmask.full_mask(codes)), 3)
    interface.set_chunk_size(10.0)
    z, mask_z = interface.vamp(codes, mask, batch_size=2, feedback_steps=2, _sampling_steps=3, time_stretch_factor=1,
                               return_mask=True, temperature=1.0, typical_filtering=True, typical_mass=0.15,
                               typical_min_tokens=64, top_p=None, seed=5, sample_cutoff=1.0)
    assert z.shape == (2, 14, codes.shape[-1]) and mask_z.device.type == "cpu"
    out = interface.decode(z).normalize(loudness)
    both = signal_concat([out, out])
    assert both.samples.shape[-1] == 2 * out.samples.shape[-1] and torch.isfinite(both.samples).all()
