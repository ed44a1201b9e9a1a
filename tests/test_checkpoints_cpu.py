"""CPU: checkpoint ingestion — the audiotools BaseModel file layout the reference loads
(interface.py:34, 70): {"state_dict": ..., "metadata": {"kwargs": ...}}, LoRA overlays (interface.py:37-45),
weight-norm pairs, and that the packed (folded) weights equal an unfused evaluation."""
import math

import pytest
import torch

from oracle import dac_oracle as do
from oracle import vampnet_oracle as vo

CFG = dict(n_heads=2, n_layers=2, n_codebooks=4, n_conditioning_codebooks=0, embedding_dim=128)


def test_vampnet_load_roundtrip_and_lora_overlay(tmp_path):
    from vampnet_b200.modules.transformer import VampNet
    cfg = vo.OracleConfig(**CFG)
    base = vo.make_state_dict(cfg, seed=0, lora=False)
    path = tmp_path / "coarse.pth"
    torch.save({"state_dict": base, "metadata": {"kwargs": dict(CFG, flash_attn=False, dropout=0.1)}}, path)
    m = VampNet.load(path, map_location="cpu", strict=False)
    assert (m.n_heads, m.n_layers, m.n_codebooks, m.embedding_dim) == (2, 2, 4, 128)
    sd = m.state_dict()
    for k, v in base.items():
        assert torch.equal(sd[k], v), k
    # LoRA-only overlay file, as train.py:399-406 writes and _load_model applies with strict=False
    full = vo.make_state_dict(cfg, seed=0, lora=True)
    lora_only = {k: v for k, v in full.items() if "lora_" in k}
    assert lora_only
    res = m.load_state_dict(lora_only, strict=False)
    assert not res.unexpected_keys
    w = m.transformer.layers[1].feed_forward.w_1
    n = "transformer.layers.1.feed_forward.w_1"
    want = vo.fold_lora({n + ".weight": base[n + ".weight"], n + ".lora_A": full[n + ".lora_A"],
                         n + ".lora_B": full[n + ".lora_B"]}, n)
    assert torch.allclose(w.folded(), want, atol=1e-6)
    assert not torch.allclose(w.folded(), w.weight)  # the overlay really changes the effective weight


def test_pack_weights_layout_cpu():
    """pack_weights is pure tensor algebra: check the documented layouts without a GPU."""
    from vampnet_b200.modules.transformer import VampNet, relative_position_bucket, REL_SAT
    cfgd = dict(n_heads=4, n_layers=1, n_codebooks=14, n_conditioning_codebooks=4, embedding_dim=256)
    cfg = vo.OracleConfig(**cfgd)
    sd = vo.make_state_dict(cfg, seed=3, lora=True)
    m = VampNet(**cfgd)
    m.load_state_dict(sd, strict=False)
    cb = vo.make_codebooks(14, seed=1)

    class Codec:
        pass
    import types
    codec = types.SimpleNamespace(quantizer=types.SimpleNamespace(
        quantizers=[types.SimpleNamespace(codebook=types.SimpleNamespace(weight=cb[i])) for i in range(14)]))
    p = m.pack_weights(codec)
    d, V, Cp = 256, 1024, 10
    assert p["emb_table"].shape == (14, 1025, 8) and torch.equal(p["emb_table"][3, 1024], sd["embedding.special.MASK"][3])
    assert p["emb_w3"].shape == (d, 3 * 128) and p["emb_w3"].dtype == torch.bfloat16
    w = sd["embedding.out_proj.weight"].squeeze(-1)
    hi, lo = p["emb_w3"][:, :112].float(), p["emb_w3"][:, 128:240].float()
    assert torch.equal(p["emb_w3"][:, 256:368].float(), hi) and (hi + lo - w).abs().max() < 2e-5 * w.abs().max()
    assert not p["emb_w3"][:, 112:128].any() and not p["emb_w3"][:, 240:256].any()
    n1 = sd["transformer.layers.0.norm_1.weight"]
    wq = vo.fold_lora(sd, "transformer.layers.0.self_attn.w_qs") * n1[None, :]
    assert torch.equal(p["wqkv"][0, :d], wq.to(torch.bfloat16))
    wk = sd["transformer.layers.0.self_attn.w_ks.weight"] * n1[None, :]
    assert torch.equal(p["wqkv"][0, d:2 * d], wk.to(torch.bfloat16))
    # FFN-up rows interleaved per 256-row tile: [128 value | 128 gate]
    w1 = vo.fold_lora(sd, "transformer.layers.0.feed_forward.w_1") * sd["transformer.layers.0.norm_3.weight"][None, :]
    assert torch.equal(p["w1"][0, 256:384], w1[128:256].to(torch.bfloat16))            # tile 1 value rows
    assert torch.equal(p["w1"][0, 384:512], w1[2 * d + 128: 2 * d + 256].to(torch.bfloat16))  # tile 1 gate rows
    # classifier: weight-norm fold, final-norm fold, channel p*Cp + c -> row c*V + p
    w = vo.weight_norm_fold(sd["classifier.layers.0.weight_g"], sd["classifier.layers.0.weight_v"]).squeeze(-1)
    w = w * sd["transformer.norm.weight"][None, :]
    pp, c = 777, 6
    assert torch.equal(p["wcls"][c * V + pp], w[pp * Cp + c].to(torch.bfloat16))
    assert p["bcls"][c * V + pp] == sd["classifier.layers.0.bias"][pp * Cp + c]
    # relative bias table: Toeplitz over key-query in [-128, 128]
    lut = vo.relative_position_bucket_lut(3072)
    rel = torch.arange(-REL_SAT, REL_SAT + 1)
    assert torch.equal(relative_position_bucket(rel), lut[rel + 3071])
    # ... cut at the saturation distance found from the bucket function (91: SURVEY.md §A.3)
    assert m._rel_sat == 91 and p["rel_bias"].shape == (2 * 91 + 1, 4)
    E = sd["transformer.layers.0.self_attn.relative_attention_bias.weight"]
    assert torch.equal(p["rel_bias"][0], E[15]) and torch.equal(p["rel_bias"][-1], E[31]) and torch.equal(p["rel_bias"][91], E[0])


def test_codec_load_roundtrip_with_weight_norm(tmp_path):
    from vampnet_b200.codec import DAC
    cfg = do.CodecConfig(encoder_dim=32, decoder_dim=512)
    w = do.make_codec_weights(cfg, seed=0)
    # store one conv as a weight-norm pair, as DAC checkpoints do
    name = "decoder.block.0.res_unit1.conv1"
    v = w.pop(name + ".weight")
    g = v.flatten(1).norm(dim=1).view(-1, 1, 1) * 1.3
    w[name + ".weight_v"], w[name + ".weight_g"] = v, g
    path = tmp_path / "codec.pth"
    torch.save({"state_dict": w, "metadata": {"kwargs": dict(encoder_dim=32, encoder_rates=[2, 4, 8, 12], decoder_dim=512,
                                                               n_codebooks=14, codebook_size=1024, codebook_dim=8,
                                                               sample_rate=44100)}}, path)
    m = DAC.load(path)
    assert m.hop_length == 768 and m.sample_rate == 44100 and m.latent_dim == 512
    got = m.params.get(name + ".weight")
    assert torch.allclose(got, v * 1.3, atol=1e-6)
    assert torch.equal(m.quantizer.quantizers[5].codebook.weight, w["quantizer.quantizers.5.codebook.weight"])
    x, n = m.preprocess(torch.zeros(1, 1, 1000), 44100)
    assert x.shape[-1] == 1536 and n == 1000
    with pytest.raises(KeyError):
        DAC(encoder_dim=32, decoder_dim=512).load_flat({"encoder.conv1.weight": torch.zeros(32, 1, 7)})


def _save(path, cfg_kwargs, seed, lora):
    sd = vo.make_state_dict(vo.OracleConfig(**cfg_kwargs), seed=seed, lora=lora)
    torch.save({"state_dict": sd, "metadata": {"kwargs": dict(cfg_kwargs)}}, path)
    return sd


def test_swap_checkpoint_in_place_resets_adapters_and_rejects_other_architectures(tmp_path):
    """VampNet.swap_checkpoint (f-4): same architecture -> parameters overwritten in place (module identity and
    parameter storage kept); going from a LoRA checkpoint to a plain one must leave NO adapter behind (the reference
    rebuilds the model in reload(), interface.py:146-174, so its lora_B is zero again); different architecture ->
    False and nothing changes."""
    from vampnet_b200.modules.transformer import VampNet
    tuned = _save(tmp_path / "tuned.pth", CFG, seed=0, lora=True)
    plain = _save(tmp_path / "plain.pth", CFG, seed=5, lora=False)
    other = dict(CFG, n_layers=3)
    _save(tmp_path / "other.pth", other, seed=1, lora=False)
    m = VampNet.load(tmp_path / "tuned.pth")
    w = m.transformer.layers[0].self_attn.w_qs
    storage = w.weight.data_ptr()
    assert w.lora_B.abs().sum() > 0
    assert m.swap_checkpoint(tmp_path / "plain.pth") is True
    assert w.weight.data_ptr() == storage  # in place
    assert torch.equal(w.weight, plain["transformer.layers.0.self_attn.w_qs.weight"])
    assert w.lora_B.abs().sum() == 0 and torch.equal(w.folded(), w.weight.float())
    before = {k: v.clone() for k, v in m.state_dict().items()}
    assert m.swap_checkpoint(tmp_path / "other.pth") is False
    assert all(torch.equal(v, before[k]) for k, v in m.state_dict().items())
    assert m.swap_checkpoint(tmp_path / "tuned.pth") is True
    assert torch.equal(w.lora_B, tuned["transformer.layers.0.self_attn.w_qs.lora_B"])
    # a checkpoint whose metadata omits a key means the constructor default (n_layers 16 != 2 here)
    sd = vo.make_state_dict(vo.OracleConfig(**CFG), seed=2)
    torch.save({"state_dict": sd, "metadata": {"kwargs": {k: v for k, v in CFG.items() if k != "n_layers"}}},
               tmp_path / "implicit.pth")
    assert m.swap_checkpoint(tmp_path / "implicit.pth") is False


def test_interface_model_cache_discovery_and_reload_policy(tmp_path, monkeypatch):
    """default / available_models / load_finetuned resolve against the local cache layout the reference uses
    (vampnet/__init__.py:13-76: {coarse,c2f,codec}.pth, loras/<name>/{coarse,c2f}.pth); reload() hot-swaps when the
    live model accepts the checkpoint and rebuilds otherwise, skipping paths already loaded."""
    from vampnet_b200 import interface as mod
    from vampnet_b200.interface import Interface
    monkeypatch.setenv("VAMPNET_MODELS_DIR", str(tmp_path))
    assert Interface.available_models() == ["default"]
    with pytest.raises(RuntimeError, match="local model cache"):
        Interface.default()
    for name in ("jazz", "broken", "ambient"):
        (tmp_path / "loras" / name).mkdir(parents=True)
        (tmp_path / "loras" / name / "coarse.pth").touch()
    (tmp_path / "loras" / "jazz" / "c2f.pth").touch()
    (tmp_path / "loras" / "ambient" / "c2f.pth").touch()
    assert Interface.available_models() == ["ambient", "jazz", "default"]  # "broken" lacks c2f.pth

    class Live:
        chunk_size_s = 7

        def __init__(self, accept):
            self.accept, self.swapped = accept, []

        def swap_checkpoint(self, ckpt):
            self.swapped.append(Path(ckpt))
            return self.accept

        def to(self, device):
            return self

    from pathlib import Path
    built = []

    def fake_load(ckpt, lora_ckpt=None, device="cpu", chunk_size_s=10):
        built.append((Path(ckpt), chunk_size_s))
        return Live(True)
    monkeypatch.setattr(mod, "_load_model", fake_load)
    codec = type("Codec", (), {"sample_rate": 44100, "hop_length": 768, "to": lambda self, d: self})()
    coarse, c2f = Live(accept=True), Live(accept=False)
    iface = Interface.from_models(codec, coarse, c2f, device="cpu", coarse_chunk_size_s=7, coarse2fine_chunk_size_s=2)
    with pytest.raises(AssertionError, match="not a valid model name"):
        iface.load_finetuned("broken")
    iface.load_finetuned("jazz")
    jazz = tmp_path / "loras" / "jazz"
    assert coarse.swapped == [jazz / "coarse.pth"] and iface.coarse is coarse       # hot-swapped, same object
    assert c2f.swapped == [jazz / "c2f.pth"] and built == [(jazz / "c2f.pth", 2)]   # refused -> rebuilt, chunk kept
    assert iface.c2f is not c2f and (iface.coarse_path, iface.c2f_path) == (jazz / "coarse.pth", jazz / "c2f.pth")
    iface.load_finetuned("jazz")  # already loaded: nothing happens
    assert coarse.swapped == [jazz / "coarse.pth"] and len(built) == 1
    with pytest.raises(RuntimeError, match="local model cache"):
        iface.load_finetuned("default")  # coarse.pth / c2f.pth not in the cache root
