"""Short workload for `ncu --set full`: one sampling iteration of the bench's coarse model at the bench's shape
(B=32, T=768) run three times, plus one encode/decode of a 10 s clip.  Same kernels and shapes as bench.py, a few
hundred launches instead of 27 000."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vampnet_b200.codec import DAC  # noqa: E402
from vampnet_b200.modules.transformer import VampNet  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
with torch.device(dev):
    model = VampNet(n_heads=20, n_layers=20, n_codebooks=4, n_conditioning_codebooks=0, embedding_dim=1280)
    cb = torch.randn(4, 1024, 8)
import types
codec = types.SimpleNamespace(quantizer=types.SimpleNamespace(
    quantizers=[types.SimpleNamespace(codebook=types.SimpleNamespace(weight=cb[i])) for i in range(4)]))
model.use_cuda_graph = False
z = torch.randint(0, 1024, (32, 4, 768), device=dev)
mask = torch.ones_like(z)
mask[:, :, ::7] = 0
for i in range(3):
    model.generate(codec, start_tokens=z, mask=mask, _sampling_steps=1, return_signal=False, seed=i)
torch.cuda.synchronize()
if "--codec" in sys.argv:
    dac = DAC().to(dev)
    x = torch.randn(1, 1, 441600, device=dev) * 0.3
    for _ in range(2):
        enc = dac.encode(x)
        dac.decode(enc["z"])
    torch.cuda.synchronize()
print("done")
