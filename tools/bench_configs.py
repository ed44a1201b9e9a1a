"""Secondary measurements for the other BASELINE.json configs on ONE GPU (the driver's bench.py measures configs[2]):
  cfg2: coarse generate, 12 steps, T=768, B=8
  cfg5 share: long-context coarse, T=3072, 24 steps, B=8 per GPU (64 over 8 GPUs)
Prints one JSON line per config with tokens/s, real-time factor and achieved TFLOP/s (algorithmic).
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vampnet_b200.modules.transformer import VampNet  # noqa: E402

COARSE = dict(n_heads=20, n_layers=20, n_codebooks=4, n_conditioning_codebooks=0, embedding_dim=1280)


class _Codec:
    def __init__(self, cb):
        import types
        self.quantizer = types.SimpleNamespace(quantizers=[types.SimpleNamespace(
            codebook=types.SimpleNamespace(weight=cb[i])) for i in range(cb.shape[0])])


def fwd_flops(T, d=1280, L=20, Cn=4, Cp=4):
    return T * (L * (20 * d * d + 4 * T * d) + 2 * 8 * Cn * d + 2 * d * 1024 * Cp)


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    with torch.device(dev):
        model = VampNet(**COARSE)
        cb = torch.randn(4, 1024, 8)
    codec = _Codec(cb)
    for name, B, T, steps, reps in (("cfg2 coarse 12 steps T=768 B=8", 8, 768, 12, 5),
                                    ("cfg5 share: coarse 24 steps T=3072 B=8", 8, 3072, 24, 2)):
        z = torch.randint(0, 1024, (B, 4, T), device=dev)
        mask = torch.ones_like(z)
        mask[:, :, ::7] = 0
        for _ in range(2):
            model.generate(codec, start_tokens=z, mask=mask, _sampling_steps=steps, return_signal=False, seed=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            model.generate(codec, start_tokens=z, mask=mask, _sampling_steps=steps, return_signal=False, seed=2 + i)
        e1.record()
        torch.cuda.synchronize()
        s = e0.elapsed_time(e1) / reps * 1e-3
        print(json.dumps({"config": name, "s_per_call": s, "tokens_per_s": B * T * 4 / s, "rtf": B * T * 768 / 44100 / s,
                          "tflops": fwd_flops(T) * B * steps / s / 1e12}), flush=True)


if __name__ == "__main__":
    main()
