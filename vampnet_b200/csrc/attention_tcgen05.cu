// vampnet_b200 — fused bidirectional self-attention with T5-style relative-position bias on the
// sm_100a tensor cores.
//
// Replaces MultiHeadRelativeAttention.forward between the projections (reference
// vampnet/modules/transformer.py:234-254): scores = q.k^T / sqrt(64) + bias[h, k - q]; softmax over
// keys; out = P.v; heads merged as "b l (head v)".  The reference materialises (H,B,T,T) scores in
// HBM three times per layer; here they live only in TMEM/registers.  The position bias
// (compute_bias, :183-209) is Toeplitz in (k - q) and saturates beyond |k - q| >= sat, so it is a
// (2*sat+1)-entry table per head held in shared memory.
//
// One CTA = one (batch, head, 128-query tile); key/value blocks of 64:
//   warp 0      TMA producer (Q once; K_j and V^T_j through a 3-stage ring); owns the TMEM allocation
//   warp 1      MMA issuer   S[j&1] = Q.K_j^T (tcgen05.mma M128 N64 K16 x4), issued one block AHEAD of the
//                            softmax into a double-buffered TMEM score tile ; O += P_j.V_j (same shape) from a
//                            double-buffered P tile, so softmax(j+1) never waits for P.V(j)
//   warps 2..9  softmax      two threads per query row (32 keys each): tcgen05.ld S -> scale+bias -> row max ->
//                            exp2 -> bf16 P into 128B-swizzled smem (A operand of P.V) ; O stays in TMEM
//                            and is rescaled in place only when a row's reference max grows by > 2^8
//                            (lazy rescale: P may exceed 1 by that factor, harmless in bf16/fp32).
// Roofline: tensor-bound in FLOPs (4*T^2*64 per (b,h)), but at d_head = 64 the per-block TMEM read
// (128x64 fp32) and MUFU.EX2 cost as much as the two MMAs; see DESIGN.md.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vnb {

constexpr int AQ = 128, AK = 64, DH = 64;
constexpr int KV_STAGES = 3;
constexpr int Q_BYTES = AQ * DH * 2;   // 16 KiB
constexpr int K_BYTES = AK * DH * 2;   // 8 KiB
constexpr int V_BYTES = DH * AK * 2;   // 8 KiB
constexpr int P_BYTES = AQ * AK * 2;   // 16 KiB
constexpr int ATT_MAX_SAT = 128;
constexpr int ATT_PAD = 160;  // a lookup block spans 128 query rows + 32 keys: pad the table so indices never clamp
constexpr int ATT_TAB = 2 * (ATT_MAX_SAT + ATT_PAD) + 2;
constexpr int ATT_SMEM_TILES = Q_BYTES + KV_STAGES * (K_BYTES + V_BYTES) + 2 * P_BYTES;  // 96 KiB (P double buffered)
constexpr int ATT_SMEM = ATT_SMEM_TILES + 1024 /*align*/ + ATT_TAB * 4 + 2 * 2 * AQ * 4 /*row-max exchange*/ +
                         256 /*barriers*/;
constexpr int ATT_THREADS = 320;  // producer + MMA + 8 softmax warps (two threads per query row)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 domain: O is rescaled only when a row max grows by > 2^8

struct AttnArgs {
  __nv_bfloat16* out;
  const float* rel;
  int sat, B, T, H, d;
};

// Half a 64-key block (32 keys) of one query row: turn raw scores (TMEM) into exp2-domain logits, return the
// max over these 32.  Scale+bias runs as packed FFMA2 (two keys per issue slot).
template <bool TAIL, bool LOOKUP>
__device__ __forceinline__ float scores_to_logits(uint32_t (&sr)[32], float c, float bconst, uint32_t bias_addr,
                                                  int valid) {
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

__device__ __forceinline__ void pair_barrier(int id) {  // the two warps that share a TMEM lane quadrant
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

// PTMEM = true (option "attn_p_tmem", experimental): the probabilities never touch shared memory.  Each softmax thread
// writes its 32 bf16 P values into tensor memory (tcgen05.st, columns [192,256): two 128 x 64 bf16 tiles) and P.V is
// issued with the A operand read from TMEM.  This removes the 8 x STS.128 per thread and block, the shared-memory
// reads of P by the tensor core, and the generic->async proxy fence (MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC: 14 % of the
// kernel's stall samples under ncu, profiles/ncu_attn_r1_final.txt / DESIGN.md §8).
template <bool PTMEM>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmVT, const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;                  // KV_STAGES stages
  uint8_t* sV = sK + KV_STAGES * K_BYTES;      // KV_STAGES stages
  uint8_t* sP = sV + KV_STAGES * V_BYTES;
  float* sBias = reinterpret_cast<float*>(sP + 2 * P_BYTES);
  float* sMx = sBias + ATT_TAB;  // [2 buffers][2 halves][128 rows] row-max / row-sum exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMx + 2 * 2 * AQ);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;                    // [KV_STAGES]
  uint64_t* kv_empty = kv_full + KV_STAGES;        // [KV_STAGES]
  uint64_t* s_full = kv_empty + KV_STAGES;         // [2]
  uint64_t* p_full = s_full + 2;                   // [2] P buffer written (per buffer: a lagging MMA thread can
                                                   //     never fall two phases behind on the same barrier)
  uint64_t* p_free = p_full + 2;                   // [2] P buffer consumed by its P.V (also: O updated)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int nblk = (a.T + AK - 1) / AK;

  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(&p_full[0], 256);
    mbar_init(&p_full[1], 256);
    mbar_init(&p_free[0], 1);
    mbar_init(&p_free[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmVT);
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }
  // bias table for this head, pre-multiplied by log2(e)
  // entry [rel + sat + ATT_PAD] for rel in [-(sat+PAD), sat+PAD], saturated outside [-sat, sat]
  for (int i = threadIdx.x; i < 2 * (a.sat + ATT_PAD) + 1; i += ATT_THREADS) {
    int r = i - ATT_PAD;
    r = r < 0 ? 0 : (r > 2 * a.sat ? 2 * a.sat : r);
    sBias[i] = a.rel[r * a.H + h] * LOG2E;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;         // two score buffers: columns [0,64) and [64,128)
  const uint32_t tmem_O = tmem_base + 128;   // columns [128, 192)
  const uint32_t tmem_P = tmem_base + 192;   // PTMEM: two bf16 P tiles, 32 columns each

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, Q_BYTES);
      tma_load_3d(sQ, &tmQ, q_full, h * DH, q0, b);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1, 500 + st);
        mbar_expect_tx(&kv_full[st], K_BYTES + V_BYTES);
        tma_load_3d(sK + st * K_BYTES, &tmK, &kv_full[st], a.d + h * DH, j * AK, b);
        tma_load_3d(sV + st * V_BYTES, &tmVT, &kv_full[st], j * AK, h * DH, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(AQ, AK);  // M=128, N=64 for both products
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      auto issue_qk = [&](int j) {  // S[j&1] = Q . K_j^T, one block ahead of the softmax
        const int st = j % KV_STAGES;
        mbar_wait(&kv_full[st], (j / KV_STAGES) & 1, 510 + st);
        tc_fence_after();
        const uint32_t aK = smem_u32(sK + st * K_BYTES);
        const uint32_t dS = tmem_S + (j & 1) * 64;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_bf16(dS, umma_desc_sw128(aQ + k * 32), umma_desc_sw128(aK + k * 32), idesc, k != 0);
        umma_commit(&s_full[j & 1]);
      };
      mbar_wait(q_full, 0, 509);
      issue_qk(0);
      if (nblk > 1) issue_qk(1);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % KV_STAGES;
        mbar_wait(&p_full[j & 1], (j >> 1) & 1, 520);  // P_j is in smem, S[j&1] has been consumed
        tc_fence_after();
        const uint32_t aV = smem_u32(sV + st * V_BYTES);
        const uint32_t aPj = aP + (j & 1) * P_BYTES;
#pragma unroll
        for (int k = 0; k < AK / 16; ++k) {
          if constexpr (PTMEM)  // 16 keys = 8 columns of bf16 pairs
            umma_bf16_ts(tmem_O, tmem_P + (j & 1) * 32 + k * 8, umma_desc_sw128(aV + k * 32), idesc, (j | k) != 0);
          else
            umma_bf16(tmem_O, umma_desc_sw128(aPj + k * 32), umma_desc_sw128(aV + k * 32), idesc, (j | k) != 0);
        }
        umma_commit(&kv_empty[st]);
        umma_commit(&p_free[j & 1]);
        if (j + 2 < nblk) issue_qk(j + 2);
      }
    }
  } else {
    // ===================== softmax warps: two threads per query row (32 keys of each block each) ==========
    const int quad = warp & 3;                 // TMEM lane quadrant of this warp
    const int half = (warp - 2) >> 2;          // 0: keys [0,32) of the block, 1: keys [32,64)
    const int row = quad * 32 + lane;
    const int q = q0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const float c = 0.125f * LOG2E;  // 1/sqrt(64) folded with log2(e)
    const int sat = a.sat;
    float m_ref = -INFINITY, l = 0.f;
    const int sw = row & 7;
    const uint32_t sBias_addr = smem_u32(sBias), sMx_addr = smem_u32(sMx), sP_addr = smem_u32(sP);
    const float bias_lo = sBias[0], bias_hi = sBias[2 * (sat + ATT_PAD)];

    for (int j = 0; j < nblk; ++j) {
      const int k0 = j * AK + half * 32;
      const uint32_t prow = sP_addr + (j & 1) * P_BYTES + row * 128;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1, 540 + (j & 1));
      tc_fence_after();
      uint32_t sr[32];
      tmem_ld_x32(tmem_S + (j & 1) * 64 + half * 32 + lane_off, sr);
      tmem_wait_ld();
      // scale + bias (+ mask of keys beyond T), all in the log2 domain
      // key - query range of THIS WARP's 32 rows x 32 keys: beyond +-sat the bias is one constant
      const int qw = q0 + quad * 32;
      const int rel_lo = k0 - (qw + 31), rel_hi = k0 + 31 - qw;
      const bool is_const = (rel_lo >= sat) || (rel_hi <= -sat);
      const bool tail = k0 + 32 > a.T;
      const float bconst = rel_lo >= sat ? bias_hi : bias_lo;
      // lookup blocks satisfy |k0 - q0| < sat + 160, so rel + sat + ATT_PAD stays inside the padded table
      const uint32_t bias_addr = sBias_addr + 4u * static_cast<uint32_t>(k0 - q + sat + ATT_PAD);
      const int valid = a.T - k0;
      float mx;
      if (is_const) {
        mx = tail ? scores_to_logits<true, false>(sr, c, bconst, bias_addr, valid)
                  : scores_to_logits<false, false>(sr, c, bconst, bias_addr, valid);
      } else {
        mx = tail ? scores_to_logits<true, true>(sr, c, bconst, bias_addr, valid)
                  : scores_to_logits<false, true>(sr, c, bconst, bias_addr, valid);
      }
      // row max over both halves (partner = same lane of warp +-4)
      const uint32_t ex = sMx_addr + static_cast<uint32_t>((j & 1) * 2 * AQ) * 4u;
      sts_f32(ex + (half * AQ + row) * 4u, mx);
      pair_barrier(1 + quad);
      mx = fmaxf(mx, lds_f32(ex + ((half ^ 1) * AQ + row) * 4u));
      if (j == 0) {
        m_ref = mx;
      } else {
        const bool grow = mx > m_ref + RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, grow)) {
          // rare: O must be stable, i.e. the P.V of the previous block has retired
          mbar_wait(&p_free[(j - 1) & 1], ((j - 1) >> 1) & 1, 550);
          tc_fence_after();
          const float m_new = grow ? mx : m_ref;
          const float alpha = fast_exp2(m_ref - m_new);  // exactly 1 for rows that keep their reference
          uint32_t o[32];
          tmem_ld_x32(tmem_O + lane_off + half * 32, o);  // each half rescales its own 32 output columns
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_x32(tmem_O + lane_off + half * 32, o);
          tmem_wait_st();
          l *= alpha;
          m_ref = m_new;
        }
      }
      // this block's P buffer was last read by the P.V of block j-2
      if (j >= 2) mbar_wait(&p_free[j & 1], ((j - 2) >> 1) & 1, 555);
      uint64_t psum2 = pack2(0.f, 0.f);
      const uint64_t negm2 = pack2(-m_ref, -m_ref);
      [[maybe_unused]] uint32_t pall[16];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint64_t d2 = fadd2(pack2(__uint_as_float(sr[ch * 8 + 2 * i]), __uint_as_float(sr[ch * 8 + 2 * i + 1])), negm2);
          float d0, d1;
          unpack2(d2, d0, d1);
          const float p0 = fast_exp2(d0), p1 = fast_exp2(d1);
          psum2 = fadd2(psum2, pack2(p0, p1));
          pk[i] = pack_bf16x2(p0, p1);
        }
        if constexpr (PTMEM) {
#pragma unroll
          for (int i = 0; i < 4; ++i) pall[ch * 4 + i] = pk[i];
        } else {
          sts_v4(prow + (((half * 4 + ch) ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
      }
      float ps0, ps1;
      unpack2(psum2, ps0, ps1);
      l += ps0 + ps1;
      if constexpr (PTMEM) {
        // keys [half*32, half*32+32) of this row -> columns [half*16, half*16+16) of the P tile
        tmem_st_x16(tmem_P + (j & 1) * 32 + half * 16 + lane_off, pall);
        tmem_wait_st();
      } else {
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      }
      tc_fence_before();
      mbar_arrive(&p_full[j & 1]);
    }
    // ---- finalize: O / l -> bf16 -> (B, T, d) at [b, q, h*64 + half*32 ..]
    {
      const uint32_t ex = sMx_addr + static_cast<uint32_t>((nblk & 1) * 2 * AQ) * 4u;
      sts_f32(ex + (half * AQ + row) * 4u, l);
      pair_barrier(1 + quad);
      l += lds_f32(ex + ((half ^ 1) * AQ + row) * 4u);
    }
    mbar_wait(&p_free[(nblk - 1) & 1], ((nblk - 1) >> 1) & 1, 560);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    __nv_bfloat16* orow = a.out + (static_cast<size_t>(b) * a.T + q) * a.d + h * DH + half * 32;
    {
      uint32_t o[32];
      tmem_ld_x32(tmem_O + lane_off + half * 32, o);
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
  if (warp == 0) tmem_dealloc<256>(tmem_base);
}

// "attn_p_tmem": 0 = P through shared memory (validated default), 1 = P through tensor memory (experimental until
// measured).  vnb_set_option, else environment VNB_ATTN_P_TMEM.
static int g_attn_p_tmem = -1;
void set_attn_p_tmem(int v) { g_attn_p_tmem = v ? 1 : 0; }
int get_attn_p_tmem() {
  if (g_attn_p_tmem < 0) {
    const char* e = getenv("VNB_ATTN_P_TMEM");
    g_attn_p_tmem = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return g_attn_p_tmem;
}

cudaError_t launch_attention(const AttnPlan& p, cudaStream_t st) {
  if (get_attn_v2()) return launch_attention2(p, st);
  static PerDeviceOnce once;
  int dev;
  if (once.need(&dev)) {
    cudaError_t e = cudaFuncSetAttribute(attention_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         ATT_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(attention_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e != cudaSuccess) return e;
    once.mark(dev);
  }
  if (p.sat > ATT_MAX_SAT || p.sat < 1) return cudaErrorInvalidValue;
  AttnArgs a;
  a.out = reinterpret_cast<__nv_bfloat16*>(p.out);
  a.rel = p.rel;
  a.sat = p.sat; a.B = p.B; a.T = p.T; a.H = p.H; a.d = p.H * DH;
  dim3 grid((p.T + AQ - 1) / AQ, p.H, p.B);
  if (get_attn_p_tmem())
    attention_tcgen05_kernel<true><<<grid, ATT_THREADS, ATT_SMEM, st>>>(p.tmQ, p.tmK, p.tmVT, a);
  else
    attention_tcgen05_kernel<false><<<grid, ATT_THREADS, ATT_SMEM, st>>>(p.tmQ, p.tmK, p.tmVT, a);
  return cudaGetLastError();
}

}  // namespace vnb
