"""TEST INFRASTRUCTURE — numpy Philox4x32-10, the counter-based generator the CUDA
sampler uses (vampnet_b200/csrc/sampler.cu: philox4x32_10), restated so the oracle
can draw exactly the same noise as the kernel.

The reference draws its noise from torch's global generator (torch.multinomial,
Tensor.uniform_: transformer.py:28-30, 1025); that stream cannot be reproduced by any
custom kernel, so parity under sampling is defined against this shared stream
(SURVEY.md §7 "RNG parity").

Stream layout (must match sampler.cu):
  token noise  : counter = (s, b, step, 0), output lanes 0 and 1   (TWO uniforms per row: a two-level inverse-CDF
                 draw, lane 0 picks the 128-entry vocabulary tile, lane 1 the entry inside it)
  remask noise : counter = (s, b, step, 1), output lane 0
  key          = (seed_lo, seed_hi)
  uniform      = ((x >> 9) + 0.5) * 2**-23     -> strictly inside (0, 1), exact in fp32
                 (24 bits + 0.5 needs 25 mantissa bits and would round up to 1.0)
"""
from __future__ import annotations

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint32)
    c1 = np.asarray(c1, dtype=np.uint32)
    c2 = np.asarray(c2, dtype=np.uint32)
    c3 = np.asarray(c3, dtype=np.uint32)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * _M0
            p1 = c2.astype(np.uint64) * _M1
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def to_uniform(x: np.ndarray) -> np.ndarray:
    return ((x >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)


def uniform_bsv(key, step: int, B: int, S: int, V: int) -> np.ndarray:
    """(B, S, V) fp32 uniforms for the categorical draw."""
    assert V % 4 == 0
    rows = np.arange(B * S, dtype=np.uint32)[:, None]
    v4 = np.arange(V // 4, dtype=np.uint32)[None, :]
    o = philox4x32_10(v4, rows, np.uint32(step), np.uint32(0), key[0], key[1])
    u = np.stack([to_uniform(x) for x in o], axis=-1)  # (B*S, V/4, 4)
    return u.reshape(B, S, V)


def uniform_bs(key, step: int, B: int, S: int, stream: int = 1, word: int = 0) -> np.ndarray:
    """(B, S) fp32 uniforms: stream 1 = re-mask Gumbel noise, stream 0 = categorical draw (words 0 and 1)."""
    s = np.arange(S, dtype=np.uint32)[None, :]
    b = np.arange(B, dtype=np.uint32)[:, None]
    o = philox4x32_10(s, b, np.uint32(step), np.uint32(stream), key[0], key[1])
    return to_uniform(o[word])
