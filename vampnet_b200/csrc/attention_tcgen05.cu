// vampnet_b200 — fused bidirectional self-attention with T5-style relative-position bias on the sm_100a tensor cores.
//
// Replaces MultiHeadRelativeAttention.forward between the projections (reference vampnet/modules/transformer.py:234-254):
// scores = q.k^T / sqrt(64) + bias[h, k - q]; softmax over keys; out = P.v; heads merged as "b l (head v)".  The
// reference materialises (H,B,T,T) scores in HBM three times per layer; here they live only in TMEM / registers.  The
// position bias (compute_bias, :183-209) is Toeplitz in (k - q) and saturates beyond |k - q| >= sat, so it is a
// (2*sat+1)-entry table per head held in shared memory.
//
// One CTA = (batch, head, 128 queries), 64-key blocks, two CTAs per SM (256 TMEM columns each: S[2] | O | P[2]):
//   warp 0      TMA producer (Q once; K_j and V^T_j through a 3-stage ring); owns the TMEM allocation
//   warp 1      MMA issuer: S[j&1] = Q.K_j^T two blocks ahead of the softmax; O += P_j.V_j with the A operand (P) read
//               from TENSOR memory (tcgen05.mma A-from-TMEM: no shared-memory store of P, no generic->async proxy fence)
//   warps 2..9  softmax, two threads per query row (32 keys of every block each), ONE pass per block:
//               tcgen05.ld S -> FFMA2 (scale, bias and the reference max folded into one packed multiply-add) -> exp2
//               -> row sum -> bf16 pack -> tcgen05.st P.  The exponentials (MUFU, the binding unit at d_head 64:
//               512 MUFU cycles against 256 tensor cycles per 128 x 64 tile) and the FMA/ALU work sit in the same
//               straight-line loop, so they overlap inside every warp; nothing but 32 raw scores and the packed P
//               is live, so the addresses and loop state stay in registers (an earlier two-phase version kept 64
//               logits per thread and executed ~3 rematerialised integer instructions per useful one).
// The softmax is OPTIMISTIC: P is computed against the running reference max m_ref (set by a max-only pre-pass over
// block 0) while the block maximum is tracked on the side; the two threads of a row exchange it through shared memory
// once per block, and only when a row grew by more than 2^8 is the reference moved, O and l rescaled and the block's P
// recomputed from the score tile (still in TMEM: it is released together with P).
//
// A persistent variant (CTAs walking work items, next item's Q prefetched, first Q.K^T of the next item overlapping the
// O write-back) was measured too: correct, but 207 us at T=768 and 686 us at T=3072 — SLOWER than one item per CTA (the
// hardware's dynamic CTA dispatch balances the two co-resident CTAs of an SM better than a static item walk, and 1920
// items over 296 resident CTAs quantise to 7 vs 6.49 rounds), so the grid stays (query tile, head, batch).
//
// Likewise 128-key blocks with a single score buffer (all that fits next to O and P in 256 TMEM columns): 202 / 615 us.
//
// Measured on B200 at B=32, T=768, H=20 (profiles/attention_r2_variants.txt): this kernel 190.5 us (507 TFLOP/s; T=3072,
// B=8: 573 us, 674 TFLOP/s); the round-1
// kernel (two-phase, P through shared memory) 222 us; P through TMEM alone 205 us; two 128-query tiles per CTA with
// 128-key blocks 240-258 us; evaluating 25-50 % of the exponentials as a cubic polynomial on the FMA pipe did not help
// in any of them (the kernels are issue/latency-bound at 20 warps per SM, not MUFU-bound), so that path was removed.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vnb {
namespace att {

constexpr int AQ = 128, AK = 64, DH = 64, KV_STAGES = 3;
constexpr int Q_BYTES = AQ * DH * 2;        // 16 KiB
constexpr int K_BYTES = AK * DH * 2;        // 8 KiB
constexpr int V_BYTES = DH * AK * 2;        // 8 KiB
constexpr int MAX_SAT = 128;
constexpr int PAD = 64;                     // a lookup chunk is 32 rows x 32 keys: |rel| < sat + 62
constexpr int TAB = 2 * (MAX_SAT + PAD) + 2;
constexpr int THREADS = 320;
constexpr int XCH = 2 /*slot*/ * 2 /*half*/ * AQ;   // row-max / row-sum exchange, floats
constexpr int SMEM = Q_BYTES + KV_STAGES * (K_BYTES + V_BYTES) + TAB * 4 + XCH * 4 + 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 domain

struct Args {
  __nv_bfloat16* out;
  const float* rel;
  int sat, B, T, H, d;
};

// One pass over this thread's 32 keys of a block: t' = score * c + bias - m_ref (exp2 domain; keys beyond T -> -inf),
// block maximum of t', and (DO_EXP) P = exp2(t'), row sum, bf16 pack.  Constant-bias chunks take bias - m_ref as one
// scalar, so scale + bias + reference are a single packed FFMA2 per pair of keys.
template <bool TAIL, bool LOOKUP, bool DO_EXP>
__device__ __forceinline__ float chunk_pass(const uint32_t (&s)[32], uint32_t (&pk)[16], uint64_t (&psum2)[2], float c,
                                            float add, float m_ref, uint32_t bias_addr, int valid) {
  float mx[2] = {-INFINITY, -INFINITY};
  const uint64_t c2 = pack2(c, c), a2 = pack2(add, add), nm2 = pack2(-m_ref, -m_ref);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    uint64_t t2;
    const uint64_t s2 = pack2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]));
    if constexpr (LOOKUP) {
      const uint64_t b2 = pack2(lds_f32(bias_addr + 8 * i), lds_f32(bias_addr + 8 * i + 4));
      t2 = fadd2(ffma2(s2, c2, b2), nm2);
    } else {
      t2 = ffma2(s2, c2, a2);
    }
    float t0, t1;
    unpack2(t2, t0, t1);
    if constexpr (TAIL) {
      if (2 * i >= valid) t0 = -INFINITY;
      if (2 * i + 1 >= valid) t1 = -INFINITY;
    }
    mx[i & 1] = fmaxf(mx[i & 1], fmaxf(t0, t1));
    if constexpr (DO_EXP) {
      const float p0 = fast_exp2(t0), p1 = fast_exp2(t1);
      psum2[i & 1] = fadd2(psum2[i & 1], pack2(p0, p1));
      pk[i] = pack_bf16x2(p0, p1);
    }
  }
  return fmaxf(mx[0], mx[1]);
}

__device__ __forceinline__ void named_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// shared-memory map, byte offsets from the 1024-aligned base (every address below is base + constant)
constexpr uint32_t OFF_Q = 0;
constexpr uint32_t OFF_K = OFF_Q + Q_BYTES;
constexpr uint32_t OFF_V = OFF_K + KV_STAGES * K_BYTES;
constexpr uint32_t OFF_BIAS = OFF_V + KV_STAGES * V_BYTES;
constexpr uint32_t OFF_X = OFF_BIAS + TAB * 4;           // [slot][half][128 rows] floats
constexpr uint32_t OFF_BAR = (OFF_X + XCH * 4 + 7) & ~7u;
constexpr uint32_t BAR_Q_FULL = OFF_BAR;
constexpr uint32_t BAR_KV_FULL = OFF_BAR + 8;                      // [KV_STAGES]
constexpr uint32_t BAR_KV_EMPTY = BAR_KV_FULL + 8 * KV_STAGES;     // [KV_STAGES]
constexpr uint32_t BAR_S_FULL = BAR_KV_EMPTY + 8 * KV_STAGES;      // [2] S[j&1] = Q.K_j^T complete
constexpr uint32_t BAR_P_FULL = BAR_S_FULL + 16;                   // [2] P[j&1] written, S[j&1] consumed (256 arrivals)
constexpr uint32_t BAR_PV_LAST = BAR_P_FULL + 16;                  // P.V of the last block retired: O is final
constexpr uint32_t BAR_PV_PREV = BAR_PV_LAST + 8;                  // P.V of the last-but-one block retired
constexpr uint32_t OFF_TMEM_SLOT = BAR_PV_PREV + 8;
static_assert(OFF_TMEM_SLOT + 16 <= SMEM, "shared-memory map exceeds the allocation");


// Per-thread state of a softmax thread (all members live in registers: every method is force-inlined).
struct Softmax {
  static constexpr uint32_t X_SLOT = 2 * AQ * 4;    // bytes between the two exchange slots
  uint32_t sb, tS, tO, tP, x_own, x_par;
  float bias_lo, bias_hi, m_ref, l;
  int pair_bar, kq_row, kq_warp, koff, sat, T, nblk;
  uint64_t psum2[2];

  // dispatch on the bias regime / tail of this warp's 32 x 32 patch of block j
  template <bool DO_EXP>
  __device__ __forceinline__ float pass(int j, const uint32_t (&sn)[32], uint32_t (&pk)[16]) {
    const float c = 0.125f * LOG2E;                             // 1/sqrt(64) folded with log2(e)
    const int rel0 = j * AK + kq_warp;                          // key - query at the patch's (first key, first row)
    const bool hi = rel0 - 31 >= sat;
    const bool is_const = hi || (rel0 + 31 <= -sat);
    const int valid = T - (j * AK + koff);
    const bool tail = valid < 32;
    const float add = (hi ? bias_hi : bias_lo) - m_ref;
    const uint32_t bias_addr = sb + OFF_BIAS + 4u * static_cast<uint32_t>(j * AK + kq_row + sat + PAD);
    if (is_const)
      return tail ? chunk_pass<true, false, DO_EXP>(sn, pk, psum2, c, add, m_ref, bias_addr, valid)
                  : chunk_pass<false, false, DO_EXP>(sn, pk, psum2, c, add, m_ref, bias_addr, valid);
    return tail ? chunk_pass<true, true, DO_EXP>(sn, pk, psum2, c, add, m_ref, bias_addr, valid)
                : chunk_pass<false, true, DO_EXP>(sn, pk, psum2, c, add, m_ref, bias_addr, valid);
  }
  // maximum over both threads of the row (one 64-thread named barrier)
  __device__ __forceinline__ float row_max(int use, float mx) {
    const uint32_t slot = static_cast<uint32_t>(use & 1) * X_SLOT;
    sts_f32(x_own + slot, mx);
    named_barrier(pair_bar, 64);
    return fmaxf(mx, lds_f32(x_par + slot));
  }
  __device__ __forceinline__ void load_scores(int j, uint32_t (&sn)[32]) {
    tmem_ld_x32(tS + static_cast<uint32_t>(j & 1) * 64, sn);
    tmem_wait_ld();
  }
  __device__ __forceinline__ void block(int j) {
    const uint32_t jb = static_cast<uint32_t>(j & 1);
    uint32_t sn[32], pk[16];
    mbar_wait_a(sb + BAR_S_FULL + 8 * jb, (j >> 1) & 1, 740);
    tc_fence_after();
    load_scores(j, sn);
    if (j == 0) {  // no reference yet: a max-only pass over block 0 sets it (use 0 of the exchange slots)
      m_ref = row_max(0, pass<false>(0, sn, pk));
    }
    // P[j&1] is free: it was last read by the P.V of block j-2, and Q.K^T of block j (whose completion was just
    // waited for) was issued after that P.V on the in-order tensor pipe — no barrier of its own (waiting for one in
    // place cost 5 % of the kernel: ~100 cycles of mbarrier round trip per block and thread).
    psum2[0] = pack2(0.f, 0.f);
    psum2[1] = pack2(0.f, 0.f);
    const float mx = row_max(j + 1, pass<true>(j, sn, pk));
    const float delta = mx > RESCALE_THRESHOLD ? mx : 0.f;
    if (__any_sync(0xffffffffu, delta != 0.f)) {
      // rare: some row of this warp outgrew its reference by more than 2^8.  Move the reference, rescale O and l
      // (every P.V issued so far must have retired) and recompute this block's P from the score tile in TMEM.
      m_ref += delta;
      if (j > 0) {
        // every P.V up to block j-1 must have retired: Q.K^T of block j+1 was issued right after P.V of block j-1
        // (in-order pipe), so its completion barrier says so; the last block has no successor and uses its own
        if (j + 1 < nblk) mbar_wait_a(sb + BAR_S_FULL + 8 * (jb ^ 1u), ((j + 1) >> 1) & 1, 750);
        else mbar_wait_a(sb + BAR_PV_PREV, 0, 751);
        tc_fence_after();
        const float alpha = fast_exp2(-delta);             // exactly 1 for rows that keep their reference
        uint32_t o[32];
        tmem_ld_x32(tO, o);                                // each half rescales its own 32 output columns
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st_x32(tO, o);
        tmem_wait_st();
        l *= alpha;
      }
      load_scores(j, sn);
      psum2[0] = pack2(0.f, 0.f);
      psum2[1] = pack2(0.f, 0.f);
      pass<true>(j, sn, pk);
    }
    tmem_st_x16(tP + jb * 32, pk);  // keys [32 half, 32 half + 32) of block j -> bf16 pairs in 16 columns
    float ps0, ps1, ps2, ps3;
    unpack2(psum2[0], ps0, ps1);
    unpack2(psum2[1], ps2, ps3);
    l += (ps0 + ps1) + (ps2 + ps3);
    tmem_wait_st();
    tc_fence_before();
    mbar_arrive_a(sb + BAR_P_FULL + 8 * jb);
  }
};

__global__ void __launch_bounds__(THREADS, 2)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmVT, const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];   // 128B-swizzled tiles need the 1024-byte alignment
  const uint32_t sb = smem_u32(smem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int nblk = (a.T + AK - 1) / AK;

  if (warp == 1 && lane == 0) {
    mbar_init_a(sb + BAR_Q_FULL, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init_a(sb + BAR_KV_FULL + 8 * s, 1);
      mbar_init_a(sb + BAR_KV_EMPTY + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init_a(sb + BAR_S_FULL + 8 * i, 1);
      mbar_init_a(sb + BAR_P_FULL + 8 * i, 2 * AQ);
    }
    mbar_init_a(sb + BAR_PV_LAST, 1);
    mbar_init_a(sb + BAR_PV_PREV, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmVT);
    }
    __syncwarp();
    tmem_alloc<256>(reinterpret_cast<uint32_t*>(smem + OFF_TMEM_SLOT));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM_SLOT);
  const uint32_t tmem_S = tmem_base;          // two score buffers: columns [0,64) and [64,128)
  const uint32_t tmem_O = tmem_base + 128;    // columns [128, 192)
  const uint32_t tmem_P = tmem_base + 192;    // two bf16 P tiles, 32 columns each

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx_a(sb + BAR_Q_FULL, Q_BYTES);
      tma_load_3d_a(sb + OFF_Q, &tmQ, sb + BAR_Q_FULL, h * DH, q0, b);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait_a(sb + BAR_KV_EMPTY + 8 * st, ph ^ 1, 700 + st);
        mbar_expect_tx_a(sb + BAR_KV_FULL + 8 * st, K_BYTES + V_BYTES);
        tma_load_3d_a(sb + OFF_K + st * K_BYTES, &tmK, sb + BAR_KV_FULL + 8 * st, a.d + h * DH, j * AK, b);
        tma_load_3d_a(sb + OFF_V + st * V_BYTES, &tmVT, sb + BAR_KV_FULL + 8 * st, j * AK, h * DH, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(AQ, AK);  // M=128, N=64 for both products
      auto issue_qk = [&](int j) {  // S[j&1] = Q . K_j^T
        const int st = j % KV_STAGES;
        mbar_wait_a(sb + BAR_KV_FULL + 8 * st, (j / KV_STAGES) & 1, 710 + st);
        tc_fence_after();
        const uint32_t aK = sb + OFF_K + st * K_BYTES;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_bf16(tmem_S + (j & 1) * 64, umma_desc_sw128(sb + OFF_Q + k * 32), umma_desc_sw128(aK + k * 32), idesc, k != 0);
        umma_commit_a(sb + BAR_S_FULL + 8 * (j & 1));
      };
      mbar_wait_a(sb + BAR_Q_FULL, 0, 709);
      issue_qk(0);
      if (nblk > 1) issue_qk(1);
      for (int j = 0; j < nblk; ++j) {
        mbar_wait_a(sb + BAR_P_FULL + 8 * (j & 1), (j >> 1) & 1, 720);   // P_j is in TMEM, S[j&1] has been consumed
        tc_fence_after();
        const uint32_t aV = sb + OFF_V + (j % KV_STAGES) * V_BYTES;
#pragma unroll
        for (int k = 0; k < AK / 16; ++k)  // 16 keys = 8 TMEM columns of bf16 pairs
          umma_bf16_ts(tmem_O, tmem_P + (j & 1) * 32 + k * 8, umma_desc_sw128(aV + k * 32), idesc, (j | k) != 0);
        umma_commit_a(sb + BAR_KV_EMPTY + 8 * (j % KV_STAGES));
        if (j == nblk - 1) umma_commit_a(sb + BAR_PV_LAST);
        else if (j == nblk - 2) umma_commit_a(sb + BAR_PV_PREV);
        if (j + 2 < nblk) issue_qk(j + 2);
      }
    }
  } else {
    // ===================== softmax: two threads per query row, 32 keys of every block each =====================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int sat = a.sat;
    const int T = a.T;
    // bias table of this head, times log2(e): entry [rel + sat + PAD], saturated outside [-sat, sat].  Filled by the
    // softmax warps only (the producer / MMA warps are already loading and multiplying).
    {
      float* sBias = reinterpret_cast<float*>(smem + OFF_BIAS);
      for (int i = threadIdx.x - 64; i < 2 * (sat + PAD) + 1; i += THREADS - 64) {
        int r = i - PAD;
        r = r < 0 ? 0 : (r > 2 * sat ? 2 * sat : r);
        sBias[i] = a.rel[r * a.H + h] * LOG2E;
      }
    }
    named_barrier(9, THREADS - 64);
    Softmax sm;
    sm.sb = sb;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    sm.tS = tmem_S + half * 32 + lane_off;
    sm.tO = tmem_O + half * 32 + lane_off;
    sm.tP = tmem_P + half * 16 + lane_off;
    sm.bias_lo = lds_f32(sb + OFF_BIAS);
    sm.bias_hi = lds_f32(sb + OFF_BIAS + 8 * (sat + PAD));
    sm.x_own = sb + OFF_X + 4u * static_cast<uint32_t>(half * AQ + row);
    sm.x_par = sb + OFF_X + 4u * static_cast<uint32_t>((half ^ 1) * AQ + row);
    sm.pair_bar = 1 + quad;                    // the two warps that share these 32 rows
    sm.kq_row = half * 32 - (q0 + row);
    sm.kq_warp = half * 32 - (q0 + quad * 32);
    sm.koff = half * 32;
    sm.sat = sat; sm.T = T; sm.nblk = nblk;
    sm.m_ref = 0.f; sm.l = 0.f;
    for (int j = 0; j < nblk; ++j) sm.block(j);
    float l = sm.l;
    const uint32_t x_own = sm.x_own, x_par = sm.x_par, tO = sm.tO;
    const int pair_bar = sm.pair_bar;
    constexpr uint32_t X_SLOT = 2 * AQ * 4;
    // ---- finalize: O / l -> bf16 -> (B, T, d) at [b, q, h*64 + half*32 ..]
    {  // exchange use number nblk + 1 (uses 0 .. nblk were the block maxima)
      const uint32_t slot = static_cast<uint32_t>((nblk + 1) & 1) * X_SLOT;
      sts_f32(x_own + slot, l);
      named_barrier(pair_bar, 64);
      l += lds_f32(x_par + slot);
    }
    mbar_wait_a(sb + BAR_PV_LAST, 0, 760);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int q = q0 + row;
    __nv_bfloat16* orow = a.out + (static_cast<size_t>(b) * T + q) * a.d + h * DH + half * 32;
    uint32_t o[32];
    tmem_ld_x32(tO, o);
    tmem_wait_ld();
    if (q < T) {
      uint4* o4 = reinterpret_cast<uint4*>(orow);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l);
        w.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l);
        w.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l);
        w.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l);
        o4[i] = w;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_base);
}

}  // namespace att

cudaError_t launch_attention(const AttnPlan& p, cudaStream_t st) {
  static PerDeviceOnce once;
  int dev;
  if (once.need(&dev)) {
    cudaError_t e = cudaFuncSetAttribute(att::attention_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, att::SMEM);
    if (e != cudaSuccess) return e;
    once.mark(dev);
  }
  if (p.sat > att::MAX_SAT || p.sat < 1) return cudaErrorInvalidValue;
  att::Args a;
  a.out = reinterpret_cast<__nv_bfloat16*>(p.out);
  a.rel = p.rel;
  a.sat = p.sat; a.B = p.B; a.T = p.T; a.H = p.H; a.d = p.H * att::DH;
  dim3 grid((p.T + att::AQ - 1) / att::AQ, p.H, p.B);
  att::attention_tcgen05_kernel<<<grid, att::THREADS, att::SMEM, st>>>(p.tmQ, p.tmK, p.tmVT, a);
  return cudaGetLastError();
}

}  // namespace vnb
