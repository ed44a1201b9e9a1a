"""Secondary measurement (not the driver's bench.py): BASELINE.json configs[3] on one GPU's share —
DAC encode -> vamp -> DAC decode end to end on 10 s 44.1 kHz clips (32 clips per GPU = 256 over 8 GPUs),
through the public Interface API, host audio in / host audio out inside the timed region.

    python tools/bench_e2e.py [--batch 32] [--reps 3]

Two variants: "app" = Interface.vamp exactly as app.py drives it (coarse 12 steps on 10 s chunks, c2f forced to
2 steps on 3 s chunks, interface.py:545-551); "cfg3" = coarse 12 + c2f 24 steps unchunked (SURVEY.md §8d cfg 4).
Weights are random-init (no checkpoints, no network); codec = DAC-family stand-in, tensor-core path.
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vampnet_b200.audio import AudioSignal  # noqa: E402
from vampnet_b200.codec import DAC  # noqa: E402
from vampnet_b200.interface import Interface  # noqa: E402
from vampnet_b200.modules.transformer import VampNet  # noqa: E402

COARSE = dict(n_heads=20, n_layers=20, n_codebooks=4, n_conditioning_codebooks=0, embedding_dim=1280)
C2F = dict(n_heads=20, n_layers=16, n_codebooks=14, n_conditioning_codebooks=4, embedding_dim=1280)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    with torch.device(dev):
        coarse, c2f = VampNet(**COARSE), VampNet(**C2F)
    codec = DAC()
    iface = Interface.from_models(codec, coarse, c2f, device="cuda")
    B, sr, n = args.batch, 44100, 441000
    t = torch.arange(n) / sr
    clips = (0.3 * torch.sin(2 * torch.pi * (110.0 + 20.0 * torch.arange(B)[:, None]) * t[None, :])
             + 0.05 * torch.randn(B, n))[:, None, :].pin_memory()

    def run(variant):
        sig = AudioSignal(clips.to(dev, non_blocking=True), sr)        # H2D
        sig.samples, _ = iface.codec.preprocess(sig.samples, sr)       # pad to hop
        codes = iface.codec.encode(sig.samples, sr)["codes"]           # (B, 14, 575)
        mask = iface.build_mask(codes, sig, periodic_prompt=7, upper_codebook_mask=3)
        if variant == "app":
            zc = iface.coarse_vamp(codes, mask, _sampling_steps=12)
            z = iface.coarse_to_fine(zc, mask=mask, typical_filtering=True, _sampling_steps=2)
        else:
            zc = iface.coarse_vamp(codes, mask, _sampling_steps=12)
            iface.c2f.chunk_size_s = 10  # s2t(10) = 575 frames = the whole clip: unchunked
            z = iface.coarse_to_fine(zc, mask=mask, _sampling_steps=24)
            iface.c2f.chunk_size_s = 3
        out = iface.decode(z)
        return out.samples.cpu()                                        # D2H

    res = {}
    for variant in ("app", "cfg3"):
        run(variant)
        run(variant)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            audio = run(variant)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.reps
        assert audio.shape == (B, 1, 441600) and torch.isfinite(audio).all()
        res[variant] = {"s_per_batch": dt, "rtf": B * 10.0 / dt, "tokens_per_s": B * 575 * 14 / dt}
    print(json.dumps({"workload": "configs[3] share of one GPU: encode->vamp->decode, 10 s clips", "batch": B,
                      "h2d_bytes": clips.numel() * 4, "d2h_bytes": B * 441600 * 4, **res}))


if __name__ == "__main__":
    main()
