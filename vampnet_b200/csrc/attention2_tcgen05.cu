// vampnet_b200 — attention, second design (option "attn_v2"; EXPERIMENTAL: compiles and is wired, not yet run on a
// GPU — the first design in attention_tcgen05.cu is the measured default).
//
// Same contract as attention_tcgen05.cu (reference vampnet/modules/transformer.py:234-254: softmax(q.k^T/8 +
// bias[h, k-q]) . v, heads merged).  What changes is the shape of the pipeline, following what the ncu captures of
// the first design showed (issue slots half idle, 14 % of stalls on the generic->async proxy fence
// after writing P to shared memory, a bar.sync per block to exchange row maxima between the two threads of a row):
//
//   * one CTA = (batch, head, 256 queries) = TWO 128-query tiles that ping-pong on the tensor core: while the softmax
//     warps of tile 0 work on block j, the tensor core runs P.V / Q.K^T of tile 1, and vice versa;
//   * 128-key blocks, ONE thread per query row (no cross-thread row-max exchange, half as many synchronisation points
//     per key);
//   * P never touches shared memory: softmax threads write bf16 P into tensor memory (tcgen05.st) and P.V is issued
//     with the A operand read from TMEM (no proxy fence, no STS);
//   * optimistic softmax: P is computed against the running reference max in ONE pass over S; only when a row's max
//     grew by more than 2^8 is the block redone for that warp after rescaling O (block 0 finds its max first).
//
// TMEM (512 columns, one CTA per SM): S0 S1 (128 fp32 columns each) | O0 O1 (64 each) | P0 P1 (64 columns = 128 bf16).
// Warps: 0 TMA producer + TMEM owner, 1 MMA issuer, 2..5 softmax of tile 0, 6..9 softmax of tile 1.
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
constexpr int THREADS = 320;
constexpr int SMEM = 1024 + 2 * Q_BYTES + KV_STAGES * (K_BYTES + V_BYTES) + TAB * 4 + 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 8.0f;   // log2 domain

struct Args {
  __nv_bfloat16* out;
  const float* rel;
  int sat, B, T, H, d;
};

// 32 raw scores of one row -> exp2-domain logits (scale, Toeplitz bias, keys beyond T masked); returns their max.
template <bool TAIL, bool LOOKUP>
__device__ __forceinline__ float logits32(uint32_t (&sr)[32], float c, float bconst, uint32_t bias_addr, int valid) {
  float mx = -INFINITY;
  const uint64_t c2 = pack2(c, c);
  const uint64_t b2c = pack2(bconst, bconst);
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    uint64_t b2 = b2c;
    if constexpr (LOOKUP) b2 = pack2(lds_f32(bias_addr + 4 * i), lds_f32(bias_addr + 4 * i + 4));
    const uint64_t t2 = ffma2(pack2(__uint_as_float(sr[i]), __uint_as_float(sr[i + 1])), c2, b2);
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

__global__ void __launch_bounds__(THREADS, 1)
attention2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmVT, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                              // [2] query tiles
  uint8_t* sK = sQ + 2 * Q_BYTES;                  // [KV_STAGES]
  uint8_t* sV = sK + KV_STAGES * K_BYTES;          // [KV_STAGES][2 halves]
  float* sBias = reinterpret_cast<float*>(sV + KV_STAGES * V_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + TAB);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                    // [KV_STAGES]
  uint64_t* kv_empty = kv_full + KV_STAGES;        // [KV_STAGES]
  uint64_t* s_full = kv_empty + KV_STAGES;         // [2] S_g(j) complete (implies P.V_g(j-1) retired: same issue queue)
  uint64_t* p_full = s_full + 2;                   // [2] P_g(j) written, S_g(j) consumed (128 arrivals)
  uint64_t* o_final = p_full + 2;                  // [2] last P.V of tile g retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_final + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * AQ);
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int nblk = (a.T + AK - 1) / AK;
  const int ntile = (q0 + AQ < a.T) ? 2 : 1;       // a CTA at the end of the sequence may own a single query tile

  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], AQ);
      mbar_init(&o_final[g], 1);
    }
    mbar_fence_init();
  }
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmVT);
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  // bias table of this head, times log2(e): entry [rel + sat + PAD], saturated outside [-sat, sat]
  for (int i = threadIdx.x; i < 2 * (a.sat + PAD) + 1; i += THREADS) {
    int r = i - PAD;
    r = r < 0 ? 0 : (r > 2 * a.sat ? 2 * a.sat : r);
    sBias[i] = a.rel[r * a.H + h] * LOG2E;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;          // + g * 128
  const uint32_t tmem_O = tmem_base + 256;    // + g * 64
  const uint32_t tmem_P = tmem_base + 384;    // + g * 64   (64 columns = 128 bf16 per row)

  if (warp == 0) {
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
  } else if (warp == 1) {
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
      };
      mbar_wait(q_full, 0, 709);
      mbar_wait(&kv_full[0], 0, 710);
      tc_fence_after();
      for (int g = 0; g < ntile; ++g) issue_qk(g, 0);
      for (int j = 0; j < nblk; ++j) {
        for (int g = 0; g < ntile; ++g) {
          mbar_wait(&p_full[g], j & 1, 720 + g);
          tc_fence_after();
          issue_pv(g, j);
          if (g == ntile - 1) umma_commit(&kv_empty[j % KV_STAGES]);  // K_j, V_j consumed by every tile
          if (j + 1 < nblk) {
            if (g == 0) {
              mbar_wait(&kv_full[(j + 1) % KV_STAGES], ((j + 1) / KV_STAGES) & 1, 711);
              tc_fence_after();
            }
            issue_qk(g, j + 1);
          } else {
            umma_commit(&o_final[g]);
          }
        }
      }
    }
  } else if (((warp - 2) >> 2) < ntile) {
    // ===================== softmax: one thread per query row =====================
    const int g = (warp - 2) >> 2;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const int q = q0 + g * AQ + row;
    const int qw = q0 + g * AQ + quad * 32;    // first row of this warp
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_S + g * 128 + lane_off, tO = tmem_O + g * 64 + lane_off, tP = tmem_P + g * 64 + lane_off;
    const float c = 0.125f * LOG2E;            // 1/sqrt(64) folded with log2(e)
    const int sat = a.sat;
    const uint32_t sBias_addr = smem_u32(sBias);
    const float bias_lo = sBias[0], bias_hi = sBias[2 * (sat + PAD)];
    float m_ref = -INFINITY, l = 0.f;

    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&s_full[g], j & 1, 740 + g);
      tc_fence_after();
      // 32 keys of this row: raw scores out of TMEM -> logits in sr, returns their max
      auto load_chunk = [&](int ch, uint32_t (&sr)[32]) -> float {
        tmem_ld_x32(tS + ch * 32, sr);
        tmem_wait_ld();
        const int k0 = j * AK + ch * 32;
        const int rel_lo = k0 - (qw + 31), rel_hi = k0 + 31 - qw;   // key - query over this warp's 32 x 32 patch
        const bool is_const = (rel_lo >= sat) || (rel_hi <= -sat);
        const bool tail = k0 + 32 > a.T;
        const float bconst = rel_lo >= sat ? bias_hi : bias_lo;
        const uint32_t bias_addr = sBias_addr + 4u * static_cast<uint32_t>(k0 - q + sat + PAD);
        const int valid = a.T - k0;
        if (is_const)
          return tail ? logits32<true, false>(sr, c, bconst, bias_addr, valid)
                      : logits32<false, false>(sr, c, bconst, bias_addr, valid);
        return tail ? logits32<true, true>(sr, c, bconst, bias_addr, valid)
                    : logits32<false, true>(sr, c, bconst, bias_addr, valid);
      };
      if (j == 0) {  // no reference yet: one extra pass over S finds the row max
        float mx = -INFINITY;
#pragma unroll 1
        for (int ch = 0; ch < AK / 32; ++ch) {
          uint32_t sr[32];
          mx = fmaxf(mx, load_chunk(ch, sr));
        }
        m_ref = mx;
      }
      float psum;
      bool redo;
      do {
        float mx_new = -INFINITY;
        uint64_t psum2 = pack2(0.f, 0.f);
        const uint64_t negm2 = pack2(-m_ref, -m_ref);
#pragma unroll 1
        for (int ch = 0; ch < AK / 32; ++ch) {
          uint32_t sr[32];
          mx_new = fmaxf(mx_new, load_chunk(ch, sr));
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint64_t d2 = fadd2(pack2(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1])), negm2);
            float d0, d1;
            unpack2(d2, d0, d1);
            const float p0 = fast_exp2(d0), p1 = fast_exp2(d1);
            psum2 = fadd2(psum2, pack2(p0, p1));
            pk[i] = pack_bf16x2(p0, p1);
          }
          tmem_st_x16(tP + ch * 16, pk);  // keys [32 ch, 32 ch + 32) -> bf16 pairs in columns [16 ch, 16 ch + 16)
        }
        float ps0, ps1;
        unpack2(psum2, ps0, ps1);
        psum = ps0 + ps1;
        redo = false;
        if (j > 0) {
          const bool grow = mx_new > m_ref + RESCALE_THRESHOLD;
          if (__any_sync(0xffffffffu, grow)) {
            // rare: O_g is stable here (P.V_g(j-1) retired before S_g(j) was signalled; P.V_g(j) waits for p_full)
            const float m_new = grow ? mx_new : m_ref;
            const float alpha = fast_exp2(m_ref - m_new);  // exactly 1 for rows that keep their reference
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t o[32];
              tmem_ld_x32(tO + hh * 32, o);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_x32(tO + hh * 32, o);
            }
            tmem_wait_st();
            l *= alpha;
            m_ref = m_new;
            redo = true;  // recompute this block's P against the new reference (no row can grow again)
          }
        }
      } while (redo);
      l += psum;
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[g]);
    }
    // ---- finalize: O / l -> bf16 -> (B, T, d) at [b, q, h*64 ..]
    mbar_wait(&o_final[g], 0, 760 + g);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    __nv_bfloat16* orow = a.out + (static_cast<size_t>(b) * a.T + q) * a.d + h * DH;
#pragma unroll 1
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t o[32];
      tmem_ld_x32(tO + hh * 32, o);
      tmem_wait_ld();
      if (q < a.T) {
        uint4* o4 = reinterpret_cast<uint4*>(orow + hh * 32);
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
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

}  // namespace a2

// "attn_v2": 0 = first design (measured default), 1 = this kernel.  vnb_set_option, else environment VNB_ATTN_V2.
static int g_attn_v2 = -1;
void set_attn_v2(int v) { g_attn_v2 = v ? 1 : 0; }
int get_attn_v2() {
  if (g_attn_v2 < 0) {
    const char* e = getenv("VNB_ATTN_V2");
    g_attn_v2 = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return g_attn_v2;
}

cudaError_t launch_attention2(const AttnPlan& p, cudaStream_t st) {
  static PerDeviceOnce once;
  int dev;
  if (once.need(&dev)) {
    cudaError_t e = cudaFuncSetAttribute(a2::attention2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, a2::SMEM);
    if (e != cudaSuccess) return e;
    once.mark(dev);
  }
  if (p.sat > a2::MAX_SAT || p.sat < 1) return cudaErrorInvalidValue;
  a2::Args a;
  a.out = reinterpret_cast<__nv_bfloat16*>(p.out);
  a.rel = p.rel;
  a.sat = p.sat; a.B = p.B; a.T = p.T; a.H = p.H; a.d = p.H * a2::DH;
  dim3 grid((p.T + 2 * a2::AQ - 1) / (2 * a2::AQ), p.H, p.B);
  a2::attention2_kernel<<<grid, a2::THREADS, a2::SMEM, st>>>(p.tmQ, p.tmK128, p.tmVT, a);
  return cudaGetLastError();
}

}  // namespace vnb
