"""Shared by the drop-in tests: write a synthetic model cache in the REFERENCE's on-disk layout
($VAMPNET_MODELS_DIR/{codec,coarse,c2f}.pth and loras/<name>/{coarse,c2f}.pth, vampnet/__init__.py:13-59):
  * codec.pth   — a descript-audio-codec / lac state_dict (nn.Sequential key names, weight_g/weight_v, (1,C,1) alphas)
                  produced by a torch module tree (tests/test_codec_lac_layout_cpu.py), metadata kwargs included;
  * coarse.pth / c2f.pth — reference VampNet state_dicts (oracle.make_state_dict) + metadata kwargs;
  * loras/<name>/ — fine-tuned variants (other seeds, LoRA tensors present).
Returns the oracle-side handles needed to check results."""
import torch

from oracle import dac_oracle as do
from oracle import vampnet_oracle as vo

COARSE = dict(n_heads=4, n_layers=2, n_codebooks=4, n_conditioning_codebooks=0, embedding_dim=256)
C2F = dict(n_heads=4, n_layers=1, n_codebooks=14, n_conditioning_codebooks=4, embedding_dim=256)
CODEC = do.CodecConfig(encoder_dim=32, decoder_dim=512)   # channel widths stay multiples of 32 (tensor-core codec)


def write_cache(root, lora_name="opera"):
    from tests.test_codec_lac_layout_cpu import DescriptLayoutCodec
    root.mkdir(parents=True, exist_ok=True)
    torch.manual_seed(0)
    codec = DescriptLayoutCodec(CODEC).eval()
    with torch.no_grad():
        for n, p in codec.named_parameters():
            if n.endswith("alpha"):
                p.copy_(0.5 + torch.rand_like(p))
            elif n.endswith("weight_g"):
                p.mul_(0.6 + 0.3 * torch.rand_like(p))
            elif n.startswith("decoder.model") and n.endswith("bias"):
                p.mul_(0.1)
    kwargs = dict(encoder_dim=CODEC.encoder_dim, encoder_rates=list(CODEC.encoder_rates), decoder_dim=CODEC.decoder_dim,
                  decoder_rates=list(CODEC.decoder_rates), n_codebooks=CODEC.n_codebooks, codebook_size=CODEC.codebook_size,
                  codebook_dim=CODEC.codebook_dim, sample_rate=CODEC.sample_rate, quantizer_dropout=0.5)
    torch.save({"state_dict": codec.state_dict(), "metadata": {"kwargs": kwargs}}, root / "codec.pth")

    def vampnet_file(path, cfgd, seed, lora):
        sd = vo.make_state_dict(vo.OracleConfig(**cfgd), seed=seed, lora=lora)
        path.parent.mkdir(parents=True, exist_ok=True)
        torch.save({"state_dict": sd, "metadata": {"kwargs": dict(cfgd, flash_attn=False, dropout=0.1, vocab_size=1024,
                                                                   latent_dim=8, noise_mode="mask")}}, path)
        return sd

    sds = dict(coarse=vampnet_file(root / "coarse.pth", COARSE, 0, False), c2f=vampnet_file(root / "c2f.pth", C2F, 1, False),
               lora_coarse=vampnet_file(root / "loras" / lora_name / "coarse.pth", COARSE, 5, True),
               lora_c2f=vampnet_file(root / "loras" / lora_name / "c2f.pth", C2F, 6, True))
    (root / "loras" / "incomplete").mkdir(parents=True, exist_ok=True)   # no c2f.pth: must not be listed
    torch.save({}, root / "loras" / "incomplete" / "coarse.pth")
    return codec, sds
