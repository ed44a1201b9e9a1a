// vampnet_b200 — fused bidirectional self-attention with T5-style relative-position bias, second design
// (option "attn_v2").
//
// Same contract as attention_tcgen05.cu (reference vampnet/modules/transformer.py:234-254: softmax(q.k^T/8 +
// bias[h, k-q]) . v, heads merged).  At d_head = 64 the exponentials, not the tensor core, bound the kernel
// (per 128 x 128 score tile: 512 tensor cycles, 1024 MUFU.EX2 cycles), so the design is organised around keeping the
// MUFU pipe of every SM sub-partition fed:
//
//   * one CTA = (batch, head, 256 queries) = TWO 128-query tiles that ping-pong on the tensor core, 128-key blocks;
//   * 16 softmax warps (4 per sub-partition): two threads per query row, 64 keys each, so that one warp's TMEM
//     loads / barrier waits hide behind another warp's exponentials;
//   * two phases per block with the 64 logits of a thread held in registers: phase 1 reads S from tensor memory,
//     applies scale + bias - reference max in ONE packed FFMA2 and finds the block maximum; S is released to the
//     tensor core right after the read ("s_free"), so Q.K^T of the NEXT block runs while phase 2 (exp2, row sum,
//     bf16 pack) of this one is still going: the softmax warps never wait for the tensor core in steady state;
//   * P never touches shared memory: tcgen05.st into tensor memory, P.V issued with the A operand read from TMEM;
//   * the row reference max is only moved when a block maximum exceeds it by more than 2^8 (O and l are then
//     rescaled in place); the two threads of a row agree through a 2-slot shared-memory exchange + a 64-thread
//     named barrier per block.
//
// TMEM (512 columns, one CTA per SM): S0 S1 (128 fp32 columns each) | O0 O1 (64 each) | P0 P1 (64 columns = 128 bf16).
// Warps: 0..7 softmax of tile 0 (0..3 keys [0,64) of a block, 4..7 keys [64,128)), 8..15 softmax of tile 1,
// 16 TMA producer + TMEM owner, 17 MMA issuer.  (setmaxnreg was tried to move the producer / MMA warps' registers to
// the softmax warpgroups: ptxas then spills MORE in the softmax branch, so every warp runs with the launch value.)
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vnb {
namespace a2 {

constexpr int AQ = 128, AK = 128, DH = 64, KV_STAGES = 3;
constexpr int Q_BYTES = AQ * DH * 2;        // 16 KiB per query tile
constexpr int K_BYTES = AK * DH * 2;        // 16 KiB
constexpr int V_BYTES = DH * AK * 2;        // 16 KiB: two (64 dh x 64 keys) 128B-swizzled halves
constexpr int MAX_SAT = 128;
constexpr int PAD = 64;                     // a lookup chunk is 32 rows x 32 keys: |rel| < sat + 62
constexpr int TAB = 2 * (MAX_SAT + PAD) + 2;
constexpr int SOFTMAX_THREADS = 512;
constexpr int THREADS = SOFTMAX_THREADS + 64;
constexpr int PRODUCER_WARP = 16, MMA_WARP = 17;
constexpr int XCH = 2 /*tile*/ * 2 /*slot*/ * 2 /*half*/ * AQ;   // row-max / row-sum exchange, floats
constexpr int SMEM = 1024 + 2 * Q_BYTES + KV_STAGES * (K_BYTES + V_BYTES) + TAB * 4 + XCH * 4 + 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 domain

struct Args {
  __nv_bfloat16* out;
  const float* rel;
  int sat, B, T, H, d;
};

// 32 raw scores of one row -> t' = score * c + bias - m_ref (exp2 domain; keys beyond T -> -inf); returns max t'.
// Constant-bias chunks take bias - m_ref as one scalar: a single packed FFMA2 per pair of keys.
template <bool TAIL, bool LOOKUP>
__device__ __forceinline__ float logits32(uint32_t* sr, float c, float add, float m_ref, uint32_t bias_addr,
                                          int valid) {
  float mx = -INFINITY;
  const uint64_t c2 = pack2(c, c);
  const uint64_t a2 = pack2(add, add);
  const uint64_t nm2 = pack2(-m_ref, -m_ref);
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    uint64_t t2;
    if constexpr (LOOKUP) {
      const uint64_t b2 = pack2(lds_f32(bias_addr + 4 * i), lds_f32(bias_addr + 4 * i + 4));
      t2 = fadd2(ffma2(pack2(__uint_as_float(sr[i]), __uint_as_float(sr[i + 1])), c2, b2), nm2);
    } else {
      t2 = ffma2(pack2(__uint_as_float(sr[i]), __uint_as_float(sr[i + 1])), c2, a2);
    }
    float t0, t1;
    unpack2(t2, t0, t1);
    if constexpr (TAIL) {
      if (i >= valid) t0 = -INFINITY;
      if (i + 1 >= valid) t1 = -INFINITY;
    }
    sr[i] = __float_as_uint(t0);
    sr[i + 1] = __float_as_uint(t1);
    mx = fmaxf(mx, fmaxf(t0, t1));
  }
  return mx;
}

// Two exp2 evaluated on the FMA / ALU pipes instead of MUFU (which is the binding unit of this kernel): round to the
// nearest integer with the 1.5 * 2^23 trick, cubic minimax polynomial for 2^f on [-0.5, 0.5] (max relative error
// 7.5e-5, a thirtieth of the bf16 rounding P gets anyway), integer part added into the exponent field.  Inputs are
// clamped at -126 so the exponent cannot wrap (exp2 of anything below is 0 to bf16 precision next to a row max of 1).
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& p0, float& p1) {
  x0 = fmaxf(x0, -126.f);
  x1 = fmaxf(x1, -126.f);
  const uint64_t x2 = pack2(x0, x1);
  const uint64_t xr2 = fadd2(x2, pack2(12582912.f, 12582912.f));          // mantissa low bits = round(x)
  const uint64_t xi2 = fadd2(xr2, pack2(-12582912.f, -12582912.f));       // round(x) as a float
  const uint64_t f2 = ffma2(xi2, pack2(-1.f, -1.f), x2);                  // x - round(x) in [-0.5, 0.5]
  uint64_t q2 = ffma2(f2, pack2(0.05517132207751274f, 0.05517132207751274f), pack2(0.24261054396629333f, 0.24261054396629333f));
  q2 = ffma2(q2, f2, pack2(0.6932609677314758f, 0.6932609677314758f));
  q2 = ffma2(q2, f2, pack2(0.9999281167984009f, 0.9999281167984009f));
  float q0, q1, r0, r1;
  unpack2(q2, q0, q1);
  unpack2(xr2, r0, r1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(r0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(r1) << 23));
}
// which of the 16 key pairs of a 32-key chunk take the polynomial: NPOLY of them (0, 4, 8 or 12), evenly spread
template <int NPOLY>
__device__ __forceinline__ constexpr bool pair_is_poly(int i) {
  return NPOLY == 8 ? (i & 1) != 0 : NPOLY == 4 ? (i & 3) == 3 : NPOLY == 12 ? (i & 3) != 0 : false;
}

__device__ __forceinline__ void named_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// 96 registers per thread: the register file is handed out per 4 warps, so 18 warps count as 20 (65536 / 640 = 102);
// __maxnreg__(112) compiles without spills but the launch is refused ("too many resources").
template <int NPOLY>
__global__ void __launch_bounds__(THREADS, 1)
attention2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmVT, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                              // [2] query tiles
  uint8_t* sK = sQ + 2 * Q_BYTES;                  // [KV_STAGES]
  uint8_t* sV = sK + KV_STAGES * K_BYTES;          // [KV_STAGES][2 halves]
  float* sBias = reinterpret_cast<float*>(sV + KV_STAGES * V_BYTES);
  float* sX = sBias + TAB;                         // [tile][slot][half][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + XCH);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                    // [KV_STAGES]
  uint64_t* kv_empty = kv_full + KV_STAGES;        // [KV_STAGES]
  uint64_t* s_full = kv_empty + KV_STAGES;         // [2] S_g(j) complete
  uint64_t* s_free = s_full + 2;                   // [2] S_g(j) is in registers (256 arrivals)
  uint64_t* p_full = s_free + 2;                   // [2] P_g(j) written (256 arrivals)
  uint64_t* pv_done = p_full + 2;                  // [2] P.V_g(j) retired: O_g stable up to block j, P_g reusable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * AQ);
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int nblk = (a.T + AK - 1) / AK;
  const int ntile = (q0 + AQ < a.T) ? 2 : 1;       // a CTA at the end of the sequence may own a single query tile

  if (warp == MMA_WARP && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&s_free[g], 2 * AQ);
      mbar_init(&p_full[g], 2 * AQ);
      mbar_init(&pv_done[g], 1);
    }
    mbar_fence_init();
  }
  if (warp == PRODUCER_WARP) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmVT);
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;          // + g * 128
  const uint32_t tmem_O = tmem_base + 256;    // + g * 64
  const uint32_t tmem_P = tmem_base + 384;    // + g * 64   (64 columns = 128 bf16 per row)

  if (warp == PRODUCER_WARP) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * Q_BYTES);     // a tile beyond T arrives as zeros
      tma_load_3d(sQ, &tmQ, q_full, h * DH, q0, b);
      tma_load_3d(sQ + Q_BYTES, &tmQ, q_full, h * DH, q0 + AQ, b);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1, 700 + st);
        mbar_expect_tx(&kv_full[st], K_BYTES + V_BYTES);
        tma_load_3d(sK + st * K_BYTES, &tmK, &kv_full[st], a.d + h * DH, j * AK, b);
        tma_load_3d(sV + st * V_BYTES, &tmVT, &kv_full[st], j * AK, h * DH, b);
        tma_load_3d(sV + st * V_BYTES + V_BYTES / 2, &tmVT, &kv_full[st], j * AK + 64, h * DH, b);
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(AQ, AK);   // S = Q . K^T : M128 N128, K = 64 in 4 steps
      constexpr uint32_t idesc_pv = umma_idesc_bf16(AQ, DH);   // O += P . V : M128 N64,  K = 128 in 8 steps
      auto issue_qk = [&](int g, int j) {
        const uint32_t aQ = smem_u32(sQ + g * Q_BYTES);
        const uint32_t aK = smem_u32(sK + (j % KV_STAGES) * K_BYTES);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_bf16(tmem_S + g * 128, umma_desc_sw128(aQ + k * 32), umma_desc_sw128(aK + k * 32), idesc_qk, k != 0);
        umma_commit(&s_full[g]);
      };
      auto issue_pv = [&](int g, int j) {
        const uint32_t aV = smem_u32(sV + (j % KV_STAGES) * V_BYTES);
#pragma unroll
        for (int k = 0; k < AK / 16; ++k)  // 16 keys = 8 TMEM columns of bf16 pairs; V halves hold 64 keys each
          umma_bf16_ts(tmem_O + g * 64, tmem_P + g * 64 + k * 8,
                       umma_desc_sw128(aV + (k >> 2) * (V_BYTES / 2) + (k & 3) * 32), idesc_pv, (j | k) != 0);
        umma_commit(&pv_done[g]);
      };
      mbar_wait(q_full, 0, 709);
      mbar_wait(&kv_full[0], 0, 710);
      tc_fence_after();
      // The two tiles are started half a block apart ON PURPOSE.  Started together they stay in lockstep (same work,
      // fair arbitration): all 16 softmax warps read S / scale / exchange maxima at the same time (MUFU idle) and then
      // all queue exponentials at the same time (MUFU saturated, everything else idle).  Tile 1's first Q.K^T is
      // issued only once tile 0 has pulled its first score block into registers, so one tile's phase 1 overlaps the
      // other's exponentials from then on.
      issue_qk(0, 0);
      if (ntile == 2) {
        mbar_wait(&s_free[0], 0, 713);
        tc_fence_after();
        issue_qk(1, 0);
      }
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) {
          // the score tiles were copied to registers: Q.K^T of the next block overlaps this block's exponentials
          mbar_wait(&kv_full[(j + 1) % KV_STAGES], ((j + 1) / KV_STAGES) & 1, 711);
          for (int g = 0; g < ntile; ++g) {
            mbar_wait(&s_free[g], j & 1, 715 + g);
            tc_fence_after();
            issue_qk(g, j + 1);
          }
        }
        for (int g = 0; g < ntile; ++g) {
          mbar_wait(&p_full[g], j & 1, 720 + g);
          tc_fence_after();
          issue_pv(g, j);
        }
        umma_commit(&kv_empty[j % KV_STAGES]);  // K_j, V_j consumed by every tile
      }
    }
  } else if (warp < PRODUCER_WARP) {
    // ===================== softmax: two threads per query row, 64 keys of every block each =====================
    const int g = warp >> 3;
    const int half = (warp >> 2) & 1;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const int q = q0 + g * AQ + row;
    const int qw = q0 + g * AQ + quad * 32;    // first row of this warp
    const int sat = a.sat;
    // bias table of this head, times log2(e): entry [rel + sat + PAD], saturated outside [-sat, sat].  Filled by the
    // softmax warps only (the producer / MMA warps are already loading and multiplying).
    for (int i = threadIdx.x; i < 2 * (sat + PAD) + 1; i += SOFTMAX_THREADS) {
      int r = i - PAD;
      r = r < 0 ? 0 : (r > 2 * sat ? 2 * sat : r);
      sBias[i] = a.rel[r * a.H + h] * LOG2E;
    }
    named_barrier(9, SOFTMAX_THREADS);
    if (g < ntile) {
      const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
      const uint32_t tS = tmem_S + g * 128 + half * 64 + lane_off;
      const uint32_t tO = tmem_O + g * 64 + half * 32 + lane_off;
      const uint32_t tP = tmem_P + g * 64 + half * 32 + lane_off;
      const float c = 0.125f * LOG2E;            // 1/sqrt(64) folded with log2(e)
      const uint32_t sBias_addr = smem_u32(sBias);
      const float bias_lo = sBias[0], bias_hi = sBias[2 * (sat + PAD)];
      // exchange slots of this row: own [slot][half][row], partner [slot][half ^ 1][row]
      const uint32_t x_own = smem_u32(sX) + 4u * static_cast<uint32_t>(((g * 2) * 2 + half) * AQ + row);
      const uint32_t x_par = smem_u32(sX) + 4u * static_cast<uint32_t>(((g * 2) * 2 + (half ^ 1)) * AQ + row);
      constexpr uint32_t X_SLOT = 2 * AQ * 4;    // bytes between the two slots
      const int pair_bar = 1 + g * 4 + quad;     // the two warps that share these 32 rows
      float m_ref = 0.f, l = 0.f;

      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&s_full[g], j & 1, 740 + g);
        tc_fence_after();
        uint32_t t[64];
        tmem_ld_x32(tS, *reinterpret_cast<uint32_t(*)[32]>(&t[0]));
        tmem_ld_x32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&t[32]));
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(&s_free[g]);
        // ---- phase 1: t' = score*c + bias - m_ref for this thread's 64 keys, block maximum
        float mx = -INFINITY;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const int k0 = j * AK + half * 64 + ch * 32;
          const int rel_lo = k0 - (qw + 31), rel_hi = k0 + 31 - qw;   // key - query over this warp's 32 x 32 patch
          const bool is_const = (rel_lo >= sat) || (rel_hi <= -sat);
          const bool tail = k0 + 32 > a.T;
          const float add = (rel_lo >= sat ? bias_hi : bias_lo) - m_ref;
          const uint32_t bias_addr = sBias_addr + 4u * static_cast<uint32_t>(k0 - q + sat + PAD);
          const int valid = a.T - k0;
          uint32_t* tc = &t[ch * 32];
          float m;
          if (is_const)
            m = tail ? logits32<true, false>(tc, c, add, m_ref, bias_addr, valid)
                     : logits32<false, false>(tc, c, add, m_ref, bias_addr, valid);
          else
            m = tail ? logits32<true, true>(tc, c, add, m_ref, bias_addr, valid)
                     : logits32<false, true>(tc, c, add, m_ref, bias_addr, valid);
          mx = fmaxf(mx, m);
        }
        // ---- the two threads of a row agree on the block maximum
        const uint32_t slot = static_cast<uint32_t>(j & 1) * X_SLOT;
        sts_f32(x_own + slot, mx);
        named_barrier(pair_bar, 64);
        mx = fmaxf(mx, lds_f32(x_par + slot));
        // block 0 sets the reference; later blocks move it only when the maximum grew by more than 2^8
        const float delta = (j == 0) ? mx : (mx > RESCALE_THRESHOLD ? mx : 0.f);
        if (__any_sync(0xffffffffu, delta != 0.f)) {
          const uint64_t nd2 = pack2(-delta, -delta);
#pragma unroll
          for (int i = 0; i < 64; i += 2) {
            float t0, t1;
            unpack2(fadd2(pack2(__uint_as_float(t[i]), __uint_as_float(t[i + 1])), nd2), t0, t1);
            t[i] = __float_as_uint(t0);
            t[i + 1] = __float_as_uint(t1);
          }
          m_ref += delta;
          if (j > 0) {
            mbar_wait(&pv_done[g], (j - 1) & 1, 750 + g);   // O_g is stable: every P.V up to block j-1 has retired
            tc_fence_after();
            const float alpha = fast_exp2(-delta);           // exactly 1 for rows that keep their reference
            uint32_t o[32];
            tmem_ld_x32(tO, o);                              // each half rescales its own 32 output columns
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_x32(tO, o);
            tmem_wait_st();
            l *= alpha;
          }
        }
        // ---- phase 2: P = exp2(t'), row sum, bf16 pack into tensor memory
        if (j > 0) {
          mbar_wait(&pv_done[g], (j - 1) & 1, 755 + g);     // P_g is free again (normally long retired)
          tc_fence_after();
        }
        uint64_t psum2 = pack2(0.f, 0.f);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p0, p1;
            if (pair_is_poly<NPOLY>(i)) {
              exp2_poly2(__uint_as_float(t[ch * 32 + 2 * i]), __uint_as_float(t[ch * 32 + 2 * i + 1]), p0, p1);
            } else {
              p0 = fast_exp2(__uint_as_float(t[ch * 32 + 2 * i]));
              p1 = fast_exp2(__uint_as_float(t[ch * 32 + 2 * i + 1]));
            }
            psum2 = fadd2(psum2, pack2(p0, p1));
            pk[i] = pack_bf16x2(p0, p1);
          }
          tmem_st_x16(tP + ch * 16, pk);  // keys [32 ch, 32 ch + 32) of this half -> bf16 pairs in 16 columns
        }
        float ps0, ps1;
        unpack2(psum2, ps0, ps1);
        l += ps0 + ps1;
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&p_full[g]);
      }
      // ---- finalize: O / l -> bf16 -> (B, T, d) at [b, q, h*64 + half*32 ..]
      {
        const uint32_t slot = static_cast<uint32_t>(nblk & 1) * X_SLOT;
        sts_f32(x_own + slot, l);
        named_barrier(pair_bar, 64);
        l += lds_f32(x_par + slot);
      }
      mbar_wait(&pv_done[g], (nblk - 1) & 1, 760 + g);
      tc_fence_after();
      const float inv_l = 1.0f / l;
      __nv_bfloat16* orow = a.out + (static_cast<size_t>(b) * a.T + q) * a.d + h * DH + half * 32;
      uint32_t o[32];
      tmem_ld_x32(tO, o);
      tmem_wait_ld();
      if (q < a.T) {
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == PRODUCER_WARP) tmem_dealloc<512>(tmem_base);
}

}  // namespace a2

// "attn_v2": 0 = first design (attention_tcgen05.cu), 1 = this kernel.  vnb_set_option, else environment VNB_ATTN_V2.
static int g_attn_v2 = -1;
void set_attn_v2(int v) { g_attn_v2 = v ? 1 : 0; }
int get_attn_v2() {
  if (g_attn_v2 < 0) {
    const char* e = getenv("VNB_ATTN_V2");
    g_attn_v2 = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return g_attn_v2;
}

// "attn_poly": how many of every 16 exponentials run as a polynomial on the FMA pipe (0, 4, 8, 12).  Environment
// VNB_ATTN_POLY, default 8.
static int g_attn_poly = -1;
static int get_attn_poly() {
  if (g_attn_poly < 0) {
    const char* e = getenv("VNB_ATTN_POLY");
    const int v = e ? atoi(e) : 8;
    g_attn_poly = (v == 0 || v == 4 || v == 8 || v == 12) ? v : 8;
  }
  return g_attn_poly;
}

template <int NPOLY>
static cudaError_t launch_a2(const AttnPlan& p, cudaStream_t st) {
  static PerDeviceOnce once;
  int dev;
  if (once.need(&dev)) {
    cudaError_t e = cudaFuncSetAttribute(a2::attention2_kernel<NPOLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, a2::SMEM);
    if (e != cudaSuccess) return e;
    once.mark(dev);
  }
  a2::Args a;
  a.out = reinterpret_cast<__nv_bfloat16*>(p.out);
  a.rel = p.rel;
  a.sat = p.sat; a.B = p.B; a.T = p.T; a.H = p.H; a.d = p.H * a2::DH;
  dim3 grid((p.T + 2 * a2::AQ - 1) / (2 * a2::AQ), p.H, p.B);
  a2::attention2_kernel<NPOLY><<<grid, a2::THREADS, a2::SMEM, st>>>(p.tmQ, p.tmK128, p.tmVT, a);
  return cudaGetLastError();
}

cudaError_t launch_attention2(const AttnPlan& p, cudaStream_t st) {
  if (p.sat > a2::MAX_SAT || p.sat < 1) return cudaErrorInvalidValue;
  switch (get_attn_poly()) {
    case 0: return launch_a2<0>(p, st);
    case 4: return launch_a2<4>(p, st);
    case 12: return launch_a2<12>(p, st);
    default: return launch_a2<8>(p, st);
  }
}

}  // namespace vnb
