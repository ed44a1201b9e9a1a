"""GPU: kernel variants that compile and are wired behind vnb_set_option but have NOT yet been measured / validated
on a B200 (the round's GPU budget ran out).  They are off by default, so they are not on the product path; this
file is how they get validated: run with VNB_TEST_EXPERIMENTAL=1.  Each variant must reproduce the default kernel
bit for bit (same operands, same accumulation order), and a timing line is printed for the bench shape.

  "attn_p_tmem"     attention probabilities through tensor memory (tcgen05.st + A-from-TMEM tcgen05.mma)
  "pair_arrive_cta" CTA-pair GEMM: accumulator-drained arrival without the GPU-scope fence
  "attn_v2"         second attention design (two query tiles in ping-pong, 128-key blocks, one thread per row, P in
                    TMEM, one-pass softmax); different blocking, so it is compared within tolerance, not bit for bit
"""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("VNB_TEST_EXPERIMENTAL") != "1",
                                 reason="experimental kernel variants: set VNB_TEST_EXPERIMENTAL=1")]


@pytest.fixture(scope="module")
def L():
    from vampnet_b200 import _lib
    _lib.lib()
    return _lib


def set_opt(L, name, value):
    prev = L.C.c_int32()
    L.check(L.lib().vnb_get_option(name, L.C.byref(prev)))
    L.check(L.lib().vnb_set_option(name, value))
    return prev.value


def attention_inputs(B, T, H, seed):
    d, sat = H * 64, 128
    g = torch.Generator().manual_seed(seed)
    q, k, v = (torch.randn(B, T, d, generator=g).bfloat16().cuda() for _ in range(3))
    rel = (torch.randn(2 * sat + 1, H, generator=g) * 0.5).cuda()
    rel[:36] = rel[36]
    rel[-36:] = rel[-37]
    Tpad = (T + 7) // 8 * 8
    qk = torch.cat([q, k], dim=-1).contiguous()
    vT = torch.zeros(B, d, Tpad, device="cuda", dtype=torch.bfloat16)
    vT[:, :, :T] = v.permute(0, 2, 1)
    return q, k, v, rel, sat, qk, vT, Tpad


def attention_ref(q, k, v, rel, sat, H):
    B, T, d = q.shape
    qf, kf, vf = (x.float().view(B, T, H, 64).permute(0, 2, 1, 3) for x in (q, k, v))
    s = qf @ kf.transpose(-1, -2) * 0.125
    ar = torch.arange(T, device=q.device)
    s = s + rel[(ar[None, :] - ar[:, None]).clamp(-sat, sat) + sat].permute(2, 0, 1)[None]
    e = torch.exp(s - s.amax(-1, keepdim=True))
    o = (e.to(torch.bfloat16).float() @ vf) / e.sum(-1, keepdim=True)
    return o.permute(0, 2, 1, 3).reshape(B, T, d)


def run_attention(L, qk, vT, rel, sat, B, T, Tpad, H):
    out = torch.full((B, T, H * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(L.lib().vnb_op_attention(L.ptr(qk), L.ptr(vT), L.ptr(out), L.ptr(rel), sat, B, T, Tpad, H, L.stream_ptr()))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("B,T,H", [(1, 64, 1), (2, 100, 4), (1, 3, 2), (2, 768, 4), (1, 1000, 2)])
def test_attention_p_through_tmem(L, B, T, H):
    q, k, v, rel, sat, qk, vT, Tpad = attention_inputs(B, T, H, seed=T)
    prev = set_opt(L, b"attn_p_tmem", 0)
    try:
        base = run_attention(L, qk, vT, rel, sat, B, T, Tpad, H)
        L.check(L.lib().vnb_set_option(b"attn_p_tmem", 1))
        got = run_attention(L, qk, vT, rel, sat, B, T, Tpad, H)
    finally:
        set_opt(L, b"attn_p_tmem", prev)
    ref = attention_ref(q, k, v, rel, sat, H)
    assert (got.float() - ref).abs().max() < 3e-2
    assert torch.equal(got, base)  # same P values, same MMA accumulation order


@pytest.mark.parametrize("B,T,H", [(1, 64, 1), (1, 128, 1), (1, 129, 2), (2, 100, 4), (1, 3, 2), (1, 256, 2), (1, 257, 1),
                                   (2, 768, 4), (1, 1000, 2), (1, 3072, 1)])
def test_attention_v2(L, B, T, H):
    """Ragged T (partial last key block, a CTA with a single query tile, T < one block), the lookup / constant bias
    regimes (T > 2*sat) and many blocks (O rescale path: the first block's max is rarely the row max)."""
    q, k, v, rel, sat, qk, vT, Tpad = attention_inputs(B, T, H, seed=T + 1)
    prev = set_opt(L, b"attn_v2", 0)
    try:
        base = run_attention(L, qk, vT, rel, sat, B, T, Tpad, H)
        L.check(L.lib().vnb_set_option(b"attn_v2", 1))
        got = run_attention(L, qk, vT, rel, sat, B, T, Tpad, H)
    finally:
        set_opt(L, b"attn_v2", prev)
    ref = attention_ref(q, k, v, rel, sat, H)
    e_ref, e_base = (got.float() - ref).abs().max().item(), (got.float() - base.float()).abs().max().item()
    print(f"attn_v2 B={B} T={T} H={H}: vs torch {e_ref:.3e}, vs first design {e_base:.3e}")
    assert not torch.isnan(got.float()).any()
    assert e_ref < 3e-2 and e_base < 3e-2


def test_attention_v2_large_logits(L):
    """Scores with a wide dynamic range force the lazy-rescale path (row max grows by more than 2^8 after block 0)."""
    B, T, H = 1, 640, 2
    q, k, v, rel, sat, qk, vT, Tpad = attention_inputs(B, T, H, seed=5)
    qk = qk.clone()
    qk[:, 400:, H * 64:] *= 6.0   # keys of the later blocks produce much larger scores
    prev = set_opt(L, b"attn_v2", 1)
    try:
        got = run_attention(L, qk, vT, rel, sat, B, T, Tpad, H)
    finally:
        set_opt(L, b"attn_v2", prev)
    d = H * 64
    ref = attention_ref(qk[..., :d], qk[..., d:], v, rel, sat, H)
    assert (got.float() - ref).abs().max() < 5e-2


def test_attention_v2_timing(L):
    B, T, H = 32, 768, 20
    q, k, v, rel, sat, qk, vT, Tpad = attention_inputs(B, T, H, seed=1)
    out = torch.empty(B, T, H * 64, device="cuda", dtype=torch.bfloat16)
    prev = set_opt(L, b"attn_v2", 0)
    try:
        for mode in (0, 1):
            L.check(L.lib().vnb_set_option(b"attn_v2", mode))
            ms = time_op(lambda: L.lib().vnb_op_attention(L.ptr(qk), L.ptr(vT), L.ptr(out), L.ptr(rel), sat, B, T, Tpad,
                                                          H, L.stream_ptr()))
            print(f"attention attn_v2={mode}: {ms * 1e3:.1f} us  {4.0 * B * H * T * T * 64 / ms / 1e9:.0f} TFLOP/s")
    finally:
        set_opt(L, b"attn_v2", prev)


def time_op(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def test_attention_p_through_tmem_timing(L):
    B, T, H = 32, 768, 20
    q, k, v, rel, sat, qk, vT, Tpad = attention_inputs(B, T, H, seed=1)
    out = torch.empty(B, T, H * 64, device="cuda", dtype=torch.bfloat16)
    prev = set_opt(L, b"attn_p_tmem", 0)
    try:
        for mode in (0, 1):
            L.check(L.lib().vnb_set_option(b"attn_p_tmem", mode))
            ms = time_op(lambda: L.lib().vnb_op_attention(L.ptr(qk), L.ptr(vT), L.ptr(out), L.ptr(rel), sat, B, T, Tpad,
                                                          H, L.stream_ptr()))
            print(f"attention attn_p_tmem={mode}: {ms * 1e3:.1f} us  {4.0 * B * H * T * T * 64 / ms / 1e9:.0f} TFLOP/s")
    finally:
        set_opt(L, b"attn_p_tmem", prev)


@pytest.mark.parametrize("M,N,K", [(300, 512, 256), (24576, 1280, 1280), (40000, 512, 128)])
def test_pair_arrive_cta(L, M, N, K):
    import math
    g = torch.Generator().manual_seed(M)
    A = torch.randn(M, K, generator=g).bfloat16().cuda()
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16().cuda()
    x0 = torch.randn(M, N, generator=g).cuda()
    prev_pair, prev = set_opt(L, b"gemm_pair", 1), set_opt(L, b"pair_arrive_cta", 0)
    try:
        outs, times = [], []
        for mode in (0, 1):
            L.check(L.lib().vnb_set_option(b"pair_arrive_cta", mode))
            out = x0.clone()
            call = lambda: L.check(L.lib().vnb_op_gemm(L.EPI_RESID, L.ptr(A), L.ptr(W), M, N, K, L.ptr(out), None, None,  # noqa: E731
                                                       1, 8, L.stream_ptr()))
            call()
            torch.cuda.synchronize()
            outs.append(out.clone())
            times.append(time_op(call))
        print(f"resid gemm M={M} N={N} K={K}: release.cluster {times[0] * 1e3:.1f} us, cta-scope {times[1] * 1e3:.1f} us")
    finally:
        set_opt(L, b"gemm_pair", prev_pair)
        set_opt(L, b"pair_arrive_cta", prev)
    assert torch.equal(outs[0], outs[1])
    assert (outs[1] - (x0 + A.float() @ W.float().t())).abs().max() < 2e-4


def test_experimental_variants_in_the_full_stack(L):
    """Both options on: logits and generated tokens must equal the default configuration's."""
    from tests.test_gpu_parity import TINY_COARSE, build
    cfg, sd, model, cb, codec = build(TINY_COARSE)
    z = torch.randint(0, 1025, (3, 4, 200), generator=torch.Generator().manual_seed(0)).cuda()
    kw = dict(start_tokens=z.clamp(max=1023), _sampling_steps=3, seed=1, return_signal=False)
    prev_a, prev_p = set_opt(L, b"attn_p_tmem", 0), set_opt(L, b"pair_arrive_cta", 0)
    try:
        base_logits, base_tokens = model.forward_codes(z, codec).clone(), model.generate(codec, **kw)
        L.check(L.lib().vnb_set_option(b"attn_p_tmem", 1))
        L.check(L.lib().vnb_set_option(b"pair_arrive_cta", 1))
        assert torch.equal(model.forward_codes(z, codec), base_logits)
        assert torch.equal(model.generate(codec, **kw), base_tokens)
    finally:
        set_opt(L, b"attn_p_tmem", prev_a)
        set_opt(L, b"pair_arrive_cta", prev_p)
