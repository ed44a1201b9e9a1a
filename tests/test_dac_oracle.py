"""CPU: the codec oracle (oracle/dac_oracle.py) against the in-image HF DacModel stand-in of the same
family (the reference's own codec, lac, is an absent third-party dependency: parity unpinned against it)."""
import pytest
import torch

from oracle import dac_oracle as do

transformers = pytest.importorskip("transformers")


@pytest.fixture(scope="module")
def pair():
    from transformers.models.dac import DacConfig, DacModel
    cfg = do.CodecConfig(encoder_dim=16, decoder_dim=128)
    w = do.make_codec_weights(cfg, seed=0)
    hc = DacConfig(encoder_hidden_size=cfg.encoder_dim, downsampling_ratios=list(cfg.encoder_rates),
                   decoder_hidden_size=cfg.decoder_dim, n_codebooks=cfg.n_codebooks, codebook_size=cfg.codebook_size,
                   codebook_dim=cfg.codebook_dim, sampling_rate=cfg.sample_rate)
    m = DacModel(hc).eval()
    m.load_state_dict(do.to_hf_state_dict(w, cfg), strict=True)
    return cfg, w, m


def test_hop_and_shapes(pair):
    cfg, w, m = pair
    assert cfg.hop_length == 768 and cfg.latent_dim == 256
    x = torch.zeros(1, 1, 1000)
    xp, n = do.preprocess(x, cfg)
    assert xp.shape[-1] == 1536 and n == 1000


def test_encode_decode_vs_standin(pair):
    cfg, w, m = pair
    x = torch.randn(2, 1, 768 * 9, generator=torch.Generator().manual_seed(1)) * 0.3
    with torch.no_grad():
        enc = m.encode(x)
        dec = m.decode(enc.quantized_representation).audio_values
    out = do.encode(x, w, cfg)
    assert torch.equal(out["codes"], enc.audio_codes)
    assert (out["z"] - enc.quantized_representation).abs().max() < 1e-5
    assert (out["latents"] - enc.projected_latents).abs().max() < 1e-5
    audio = do.decode(out["z"], w, cfg)["audio"]
    assert audio.shape == (2, 1, 768 * 9)
    assert (audio.squeeze(1) - dec).abs().max() < 1e-4
    assert 0.05 < audio.abs().mean() < 0.8  # not saturated: the comparison is sensitive


def test_from_latents_and_from_codes_vs_standin(pair):
    cfg, w, m = pair
    x = torch.randn(1, 1, 768 * 5, generator=torch.Generator().manual_seed(2)) * 0.3
    with torch.no_grad():
        enc = m.encode(x)
        zl, _ = m.quantizer.from_latents(enc.projected_latents)
        zc = m.quantizer.from_codes(enc.audio_codes)[0]
    assert (do.rvq_from_latents(enc.projected_latents, w, cfg)[0] - zl).abs().max() < 1e-5
    assert (do.rvq_from_codes(enc.audio_codes, w, cfg) - zc).abs().max() < 1e-5
    # the VampNet.decode path (reference transformer.py:672): from_latents(from_codes-latents) reproduces from_codes
    lat = torch.cat([w[f"quantizer.quantizers.{i}.codebook.weight"][enc.audio_codes[:, i]].transpose(1, 2)
                     for i in range(cfg.n_codebooks)], 1)
    assert (do.rvq_from_latents(lat, w, cfg)[0] - zc).abs().max() < 1e-5
