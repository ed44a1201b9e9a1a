// vampnet_b200 — bf16 GEMM on the sm_100a tensor cores (tcgen05.mma, TMEM accumulators, TMA operands).
//
//   out = epilogue( A (M,K) row-major bf16  x  W (N,K)^T row-major bf16 ),  fp32 accumulation
//
// One kernel family covers every dense contraction of VampNet.forward (reference
// vampnet/modules/transformer.py): QKV projection (:229-231), attention output projection (:255),
// FFN up-projection with the GatedGELU fused (:81-83, activations.py:16-35), FFN down-projection
// with the residual add fused (:84, :367), the classifier (:632) and, in the generate loop, the classifier with
// sample_from_logits (:952-1034) fused into its epilogue (EPI_SAMPLE), and the embedding out_proj (layers.py:162).
//
// Structure (one CTA per SM, persistent over output tiles; default: CTA pairs on 256 x 256 tiles, see the note above the
// kernel; single-CTA variant: 128 x 256):
//   warp 0       TMA producer   : ring of {A 128x64, W 256x64 (pair: W half 128x64)} bf16 tiles, 128B-swizzled, 4 / 6 stages
//   warp 1       MMA issuer     : one elected thread, tcgen05.mma (.cta_group::2) M=128 (256) N=256 K=16, 4 per k-block
//   warp 2       TMEM allocator : 512 columns = two 128x256 fp32 accumulators (double buffered)
//   warps 4..    epilogue       : tcgen05.ld 32 lanes x 32 columns -> registers -> fused op -> global (4 warps; 8 for the
//                                 residual and the sampling epilogues)
// The epilogue of tile i overlaps the mainloop of tile i+1 through the two accumulators.
//
// Roofline: tensor-bound.  Algorithmic work = 2*M*N*K flop per launch; bytes (A+W+out) are a few
// MB against > 10 GFLOP, far right of the ridge.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

#ifndef VNB_GEMM_PAIR_DEFAULT
#define VNB_GEMM_PAIR_DEFAULT true
#endif

namespace vnb {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;  // 16 KiB
constexpr int B_BYTES = BN * BK * 2;  // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int GEMM_STG_BYTES = 8 * 32 * 36 * 4 - 3072;  // epilogue staging: 4 warps x 32x36 floats, or 8 warps x 32x33 (RESID)
constexpr int RING_BYTES = STAGES * STAGE_BYTES;  // 192 KiB: 4 x {A 16K, W 32K}, or (CTA pair) 6 x {A 16K, W-half 16K}
constexpr int GEMM_SMEM = RING_BYTES + GEMM_STG_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
// The residual epilogue is bound by memory-level parallelism (residual rows must be fetched before they can be
// updated): it gets 8 epilogue warps, two per TMEM lane quadrant, splitting the 32-column chunks even / odd.
template <int EPI> constexpr int gemm_epi_warps() { return (EPI == VNB_EPI_RESID || EPI == VNB_EPI_SAMPLE) ? 8 : 4; }
template <int EPI> constexpr int gemm_threads() { return 128 + 32 * gemm_epi_warps<EPI>(); }
constexpr int STG_PITCH_RESID = 33;  // scalar, conflict-free; 8 x 32 x 33 floats fit beside the 4-stage ring

struct GemmArgs {
  int M, N, K;
  int epi;
  void* out;          // see VNB_EPI_*
  void* out2;         // vT for EPI_QKV
  const float* bias;  // EPI_BIAS_F32
  int T, Tpad;        // EPI_QKV: rows m = b*T + t
  int d2;             // EPI_QKV: 2*d_model (column where V starts)
  // ---- fused RMSNorm (reference transformer.py:43-58), see DESIGN.md §4 ----
  __nv_bfloat16* out_bf16;  // EPI_RESID: bf16 copy of the updated residual stream (A operand of the next GEMM)
  float* ss_out;            // EPI_RESID: (N/256, M) per-n-tile partial row sums of squares of the updated rows
  const float* ss_in;       // consumers: partial row sums of squares of THEIR A operand; null = no row scaling
  int ss_parts;             // number of partials to add (fixed order: deterministic)
  float inv_d, eps;         // row scale = rsqrt(sum * inv_d + eps)
  // ---- EPI_SAMPLE (see the epilogue) ----
  const int32_t* zcur;
  const SampleDyn* dyn;
  float4* partials;
  int C, ncc, V, mask_token;
};

__device__ __forceinline__ float gelu_tanh(float x) {
  // 0.5*x*(1+tanh(sqrt(2/pi)*(x+0.044715*x^3)))   (activations.py:16-26); tanh(y) = 1 - 2/(1+exp(2y))
  const float y = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  const float e = __expf(2.0f * y);
  const float t = 1.0f - __fdividef(2.0f, 1.0f + e);
  return 0.5f * x * (1.0f + t);
}

// ---- epilogue ---------------------------------------------------------------------------------
// tcgen05.ld hands every lane one accumulator ROW (32 consecutive columns).  Writing that straight to
// global memory makes each warp instruction touch 32 different 128-byte lines.  Instead each warp
// transposes its 32x32 fp32 chunk through a padded shared-memory tile (row pitch 36 floats: the
// 16-byte stores and loads below are bank-conflict free) so that a warp instruction covers whole
// rows: 128 contiguous bytes per 8 lanes for fp32 outputs, 64 for bf16.
constexpr int STG_PITCH = 36;
constexpr int STG_FLOATS_PER_WARP = 32 * STG_PITCH;

// `stg` is a shared-space byte address (smem_u32): the tile base passes through an integer alignment, which would make
// plain pointers generic (LD.E / ST.E instead of LDS / STS).
__device__ __forceinline__ void stage_rows(uint32_t stg, int lane, const float (&v)[32]) {
  const uint32_t dst = stg + 4u * (lane * STG_PITCH);
#pragma unroll
  for (int i = 0; i < 8; ++i) sts_f4(dst + 16u * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
  __syncwarp();
}
__device__ __forceinline__ void stage_rows33(uint32_t stg, int lane, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) sts_f32(stg + 4u * (lane * STG_PITCH_RESID + j), v[j]);
  __syncwarp();
}

// Sum of squares of four consecutive outputs, in a fixed operation order (the fused RMSNorm statistics are
// deterministic).
__device__ __forceinline__ float sumsq4(float x, float y, float z, float w) {
  return __fmaf_rn(w, w, __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x))));
}

// fp32 destinations: lane -> (row = it*4 + lane/8, 4 columns at (lane%8)*4).
// The addend (residual rows or bias) is fetched by prefetch_addend() BEFORE the accumulator chunk is pulled
// out of TMEM, so the global-load latency overlaps the tcgen05.ld and the shared-memory transpose, and all
// eight loads are in flight before the first store.
template <int EPI>
__device__ __forceinline__ void prefetch_addend(const GemmArgs& g, int lane, int row_base, int col0, float4 (&add)[8]) {
  const int c4 = (lane & 7) * 4;
  const int r0 = row_base + (lane >> 3);
  if constexpr (EPI == VNB_EPI_RESID) {
    const float* base = reinterpret_cast<const float*>(g.out) + static_cast<size_t>(r0) * g.N + col0 + c4;
    const size_t step = static_cast<size_t>(4) * g.N;
#pragma unroll
    for (int it = 0; it < 8; ++it)
      add[it] = (r0 + it * 4 < g.M) ? __ldcg(reinterpret_cast<const float4*>(base + it * step))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + col0 + c4));
#pragma unroll
    for (int it = 0; it < 8; ++it) add[it] = b4;
  }
}
template <bool FUSED, int PITCH>
__device__ __forceinline__ void drain_f32(const GemmArgs& g, uint32_t stg, int lane, int row_base, int col0,
                                          const float4 (&add)[8], float (&ssacc)[8]) {
  const int c4 = (lane & 7) * 4;
  const int r0 = row_base + (lane >> 3);
  float* const base = reinterpret_cast<float*>(g.out) + static_cast<size_t>(r0) * g.N + col0 + c4;
  const size_t step = static_cast<size_t>(4) * g.N;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const uint32_t sp = stg + 4u * ((it * 4 + (lane >> 3)) * PITCH + c4);
    float4 a = (PITCH % 4 == 0) ? lds_f4(sp) : make_float4(lds_f32(sp), lds_f32(sp + 4), lds_f32(sp + 8), lds_f32(sp + 12));
    a.x = __fadd_rn(a.x, add[it].x); a.y = __fadd_rn(a.y, add[it].y);
    a.z = __fadd_rn(a.z, add[it].z); a.w = __fadd_rn(a.w, add[it].w);
    if (r0 + it * 4 < g.M) {
      *reinterpret_cast<float4*>(base + it * step) = a;
      if constexpr (FUSED) {
        uint2 w;
        w.x = pack_bf16x2(a.x, a.y);
        w.y = pack_bf16x2(a.z, a.w);
        *reinterpret_cast<uint2*>(g.out_bf16 + static_cast<size_t>(r0 + it * 4) * g.N + col0 + c4) = w;
        ssacc[it] = __fadd_rn(ssacc[it], sumsq4(a.x, a.y, a.z, a.w));
      }
    }
  }
  __syncwarp();
}

// bf16 destinations: lane -> (row = it*8 + lane/4, 8 columns at (lane%4)*8); `pitch` = output row pitch
__device__ __forceinline__ void drain_bf16(__nv_bfloat16* out, int pitch, int M, uint32_t stg, int lane,
                                           int row_base, int col0) {
  const int c8 = (lane & 3) * 8;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2);
    const int row = row_base + r;
    if (row < M) {
      const float4 a = lds_f4(stg + 4u * (r * STG_PITCH + c8));
      const float4 b = lds_f4(stg + 4u * (r * STG_PITCH + c8 + 4));
      uint4 w;
      w.x = pack_bf16x2(a.x, a.y);
      w.y = pack_bf16x2(a.z, a.w);
      w.z = pack_bf16x2(b.x, b.y);
      w.w = pack_bf16x2(b.z, b.w);
      *reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * pitch + col0 + c8) = w;
    }
  }
  __syncwarp();
}

// PAIR = true: the CTA-pair variant.  A cluster of two CTAs (one TPC) owns a 256 x 256 output tile; each CTA stages its
// own 128 rows of A and HALF of the W tile (128 of the 256 rows), the rank-0 CTA issues tcgen05.mma.cta_group::2 with
// M = 256, and each CTA drains its own 128 accumulator rows.  Per SM this halves the W bytes pulled from L2 into
// shared memory and read by the tensor core per flop (the kernels run power-capped: bytes moved per flop is what
// sets the clock), and the smaller stage buys a 6-deep ring in the same 192 KiB.
//
// The accumulator-drained arrival of the epilogue warps on the MMA-issuing CTA's barrier uses the default (.cta-scope)
// release: the waiter only depends on TMEM reads that have already completed (tcgen05.wait::ld), so the
// MEMBAR.ALL.GPU that .release.cluster compiles to (11 % of the residual epilogue's stall samples in round 1) is not
// needed.  Measured bit-identical and 3 % faster on the residual GEMMs (profiles/bench_r2_variants.txt).
//
// A residual epilogue that moved its x tiles by TMA (tile in, add in shared memory, tile out) was measured in round 1:
// bit-identical but 8-30 % slower than the register-staged prefetch below (profiles/bench_r1_resid_tma*.json); removed.
template <int EPI, bool PAIR>
__global__ void __launch_bounds__(gemm_threads<EPI>(), 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmArgs g) {
  constexpr int NSTAGE = PAIR ? 6 : STAGES;
  constexpr int W_BYTES = PAIR ? B_BYTES / 2 : B_BYTES;
  constexpr int STAGE_SZ = A_BYTES + W_BYTES;
  constexpr int RING = NSTAGE * STAGE_SZ;
  static_assert(RING == RING_BYTES, "ring size");
  constexpr int EPI_REGION = GEMM_STG_BYTES;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms are 1024 B: align the tile ring to 1024.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* stg_all = reinterpret_cast<float*>(smem + RING);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + RING + EPI_REGION);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* tfull_bar = empty_bar + NSTAGE;  // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;      // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;            // 0 = the CTA that issues the MMAs
  const int worker = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int num_workers = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  constexpr int TM = PAIR ? 2 * BM : BM;                          // output rows per tile (per CTA: always BM)
  const int num_m = (g.M + TM - 1) / TM;
  const int num_n = g.N / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = g.K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      // one elected lane of each epilogue warp (of both CTAs in a pair: rank 0's barrier gates the next MMA)
      mbar_init(&tempty_bar[a], gemm_epi_warps<EPI>() * (PAIR ? 2 : 1));
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) tmem_alloc_pair<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // the peer's barriers must be initialised before anything signals them
  __syncthreads();  // (pair: redundant after the cluster barrier, but it is the barrier compute-sanitizer's racecheck
                    //  understands between tcgen05.alloc's write of tmem_slot and the reads below)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers) {
        // n-fastest rasterisation: the N/256 tiles that share an A row-block run concurrently, so A is fetched from
        // HBM once (the weights, <= 13 MB, stay L2-resident).  m-fastest order re-read A 3-4x (ncu dram__bytes).
        const int m0 = (tile / num_n) * TM + static_cast<int>(rank) * BM;
        const int n0 = (tile % num_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
          uint8_t* sa = smem + stage * STAGE_SZ;
          uint8_t* sb = sa + A_BYTES;
          if constexpr (PAIR) {
            // both CTAs' bytes are credited to rank 0's barrier, which the MMA thread waits on
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_SZ);
            const uint32_t bar0 = mapa_u32(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_pair(sa, &tmA, bar0, kb * BK, m0);
            tma_load_2d_pair(sb, &tmB, bar0, kb * BK, n0 + static_cast<int>(rank) * (BN / 2));
          } else {
            mbar_expect_tx(&full_bar[stage], STAGE_SZ);
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);
          }
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
      if constexpr (PAIR) {
        // tail: the multicast commits that free the last stages are remote arrivals into THIS CTA's barriers; they
        // must have landed before the CTA may exit
        for (int s = 0; s < NSTAGE; ++s) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 150 + stage);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        if constexpr (PAIR) mbar_wait_cluster(&tempty_bar[acc], acc_phase ^ 1, 200 + acc);
        else mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 200 + acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase, 300 + stage);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_SZ);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128B swizzle span
            if constexpr (PAIR)
              umma_bf16_pair(d_tmem, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), idesc,
                             (kb | k) != 0 ? 1u : 0u);
            else
              umma_bf16(d_tmem, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), idesc,
                        (kb | k) != 0 ? 1u : 0u);
          }
          // frees the smem stage (in both CTAs of a pair) when these MMAs retire
          if constexpr (PAIR) umma_commit_pair(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        // accumulator complete
        if constexpr (PAIR) umma_commit_pair(&tfull_bar[acc]);
        else umma_commit(&tfull_bar[acc]);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    int it = 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = (tile / num_n) * TM + static_cast<int>(rank) * BM;
      const int n0 = (tile % num_n) * BN;
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < g.M;
      int b_idx = 0, t_idx = 0;
      if constexpr (EPI == VNB_EPI_QKV) {
        b_idx = row / g.T;
        t_idx = row - b_idx * g.T;
      }
      constexpr bool kWide = gemm_epi_warps<EPI>() == 8;
      const int half = kWide ? ((warp - 4) >> 2) : 0;  // 8-warp epilogue: this warp takes chunks with (c & 1) == half
      const uint32_t stg = smem_u32(kWide ? stg_all + (warp - 4) * (32 * STG_PITCH_RESID) : stg_all + quad * STG_FLOATS_PER_WARP);
      const int row_base = m0 + quad * 32;
      float rs = 1.0f;  // fused RMSNorm of the A operand: rsqrt(mean(x^2) + eps) of this thread's row
      if (g.ss_in != nullptr && row_ok) {
        float t = 0.f;
        for (int p = 0; p < g.ss_parts; ++p) t += __ldg(g.ss_in + static_cast<size_t>(p) * g.M + row);
        rs = rsqrtf(t * g.inv_d + g.eps);
      }
      float4 pre0[8];  // residual / bias of the first chunk: fetched while the mainloop of this tile still runs
      if constexpr (EPI == VNB_EPI_RESID || EPI == VNB_EPI_BIAS_F32)
        prefetch_addend<EPI>(g, lane, row_base, n0 + (gemm_epi_warps<EPI>() == 8 ? ((warp - 4) >> 2) : 0) * 32, pre0);
      mbar_wait(&tfull_bar[acc], acc_phase, 400 + acc);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
      if constexpr (EPI == VNB_EPI_GEGLU) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32], v2[32];
          tmem_ld_x32(t_addr + c * 32, v);
          tmem_ld_x32(t_addr + 128 + c * 32, v2);
          tmem_wait_ld();
          float r[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = (__uint_as_float(v[j]) * rs) * gelu_tanh(__uint_as_float(v2[j]) * rs);
          stage_rows(stg, lane, r);
          drain_bf16(reinterpret_cast<__nv_bfloat16*>(g.out), g.N / 2, g.M, stg, lane, row_base, (n0 >> 1) + c * 32);
        }
      } else if constexpr (EPI == VNB_EPI_SAMPLE) {
        // The classifier of the generate loop (transformer.py:632-634 followed by sample_from_logits, :952-1034): the
        // logits of a still-masked position are consumed where they are produced.  A thread owns one row (tcgen05.ld
        // hands every lane an accumulator row) and one 128-column strip = one 128-entry tile of one codebook's
        // vocabulary; three sweeps over the strip in TMEM: max / arg-max, sum of exp((x - max) / temperature), and the
        // inverse-CDF draw inside the strip with this row's second uniform.  What leaves the SM is 16 bytes per (row,
        // strip); sample_combine_kernel picks the strip with the first uniform.  Same arithmetic for the logit as the
        // materialising epilogue (acc * row scale, + bias), so vnb_forward_* shows exactly what was sampled from.
        constexpr float LOG2E_F = 1.4426950408889634f;
        const int et = static_cast<int>(threadIdx.x) - 128;              // 0..255 over the eight epilogue warps
        const uint32_t sbias = smem_u32(stg_all + (it & 1) * BN);        // this tile's bias, double-buffered by tile
        sts_f32(sbias + 4u * et, __ldg(g.bias + n0 + et));
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const int strip = (warp - 4) >> 2;                               // columns [128 strip, 128 strip + 128) of the tile
        const int col0 = n0 + strip * 128;
        const int cp = col0 / g.V, v0 = col0 - cp * g.V;
        const int Cp = g.C - g.ncc;
        const bool active = row_ok && __ldg(g.zcur + static_cast<size_t>(row) * g.C + g.ncc + cp) == g.mask_token;
        if (__any_sync(0xffffffffu, active)) {
          const uint32_t t_strip = t_addr + strip * 128;
          const uint32_t sb4 = sbias + 4u * (strip * 128);
          const float inv_temp = g.dyn->inv_temp;
          const int do_sample = g.dyn->do_sample;
          // sweep 1: maximum and arg-max (lowest index on ties) of the logits
          float mx = -INFINITY;
          int am = 0;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld_x32(t_strip + c * 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 b4 = lds_f4(sb4 + 16u * (c * 8 + j4));
              const float bj[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float x = __fadd_rn(__fmul_rn(__uint_as_float(v[j4 * 4 + j]), rs), bj[j]);
                if (x > mx) { mx = x; am = c * 32 + j4 * 4 + j; }
              }
            }
          }
          // sweep 2: s = sum over the strip of e = 2^((x - mx) * c1), c1 = log2(e) / temperature, in column order
          const float c1 = __fmul_rn(inv_temp, LOG2E_F);
          const float c0 = -__fmul_rn(mx, c1);
          float ssum = 0.f;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld_x32(t_strip + c * 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 b4 = lds_f4(sb4 + 16u * (c * 8 + j4));
              const float bj[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float x = __fadd_rn(__fmul_rn(__uint_as_float(v[j4 * 4 + j]), rs), bj[j]);
                ssum += fast_exp2(__fmaf_rn(x, c1, c0));
              }
            }
          }
          // sweep 3 (sampling steps only): first column whose running sum exceeds u2 * s; the arg-max if rounding
          // leaves none.  Greedy steps take the arg-max.
          int cand = am;
          float xc = mx;
          if (do_sample) {
            const int b_idx = row / g.T, t_idx = row - b_idx * g.T;
            uint32_t r4[4];
            philox4x32_10(static_cast<uint32_t>(t_idx * Cp + cp), static_cast<uint32_t>(b_idx),
                          static_cast<uint32_t>(g.dyn->step), 0u, g.dyn->seed_lo, g.dyn->seed_hi, r4);
            const float target = u01(r4[1]) * ssum;
            float run = 0.f;
            int found = -1;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t v[32];
              tmem_ld_x32(t_strip + c * 32, v);
              tmem_wait_ld();
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 b4 = lds_f4(sb4 + 16u * (c * 8 + j4));
                const float bj[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float x = __fadd_rn(__fmul_rn(__uint_as_float(v[j4 * 4 + j]), rs), bj[j]);
                  run += fast_exp2(__fmaf_rn(x, c1, c0));
                  if (run > target && found < 0) { found = c * 32 + j4 * 4 + j; xc = x; }
                }
              }
            }
            if (found >= 0) cand = found;
            else xc = mx;
          }
          if (active)
            g.partials[(static_cast<size_t>(row) * Cp + cp) * (g.V >> 7) + (v0 >> 7)] =
                make_float4(mx, ssum, xc, __uint_as_float(static_cast<uint32_t>(v0 + cand) |
                                                          (static_cast<uint32_t>(v0 + am) << 16)));
        }
      } else if constexpr (EPI == VNB_EPI_RESID || EPI == VNB_EPI_BIAS_F32) {
        // software-pipelined: the residual rows of chunk c+1 are in flight while chunk c is pulled out of TMEM,
        // transposed and stored (the global-load latency would otherwise be paid 8 times per tile, serially)
        const bool fused_out = g.out_bf16 != nullptr;  // residual GEMMs, and the embedding projection (BIAS_F32)
        constexpr int CSTEP = kWide ? 2 : 1;  // chunks owned by this warp: half, half + CSTEP, ...
        float ssacc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ssacc[i] = 0.f;
        auto process = [&](int c, const float4 (&add)[8]) {
          uint32_t v[32];
          tmem_ld_x32(t_addr + c * 32, v);
          tmem_wait_ld();
          float r[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __fmul_rn(__uint_as_float(v[j]), rs);  // rs == 1 unless a norm is fused in
          if constexpr (kWide) {
            stage_rows33(stg, lane, r);
            if (fused_out) drain_f32<true, STG_PITCH_RESID>(g, stg, lane, row_base, n0 + c * 32, add, ssacc);
            else drain_f32<false, STG_PITCH_RESID>(g, stg, lane, row_base, n0 + c * 32, add, ssacc);
          } else {
            stage_rows(stg, lane, r);
            if (fused_out) drain_f32<true, STG_PITCH>(g, stg, lane, row_base, n0 + c * 32, add, ssacc);
            else drain_f32<false, STG_PITCH>(g, stg, lane, row_base, n0 + c * 32, add, ssacc);
          }
        };
        float4 addA[8], addB[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) addA[i] = pre0[i];
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2 * CSTEP) {
          prefetch_addend<EPI>(g, lane, row_base, n0 + (c + CSTEP) * 32, addB);
          process(c, addA);
          if (c + 2 * CSTEP < BN / 32) prefetch_addend<EPI>(g, lane, row_base, n0 + (c + 2 * CSTEP) * 32, addA);
          process(c + CSTEP, addB);
        }
        if (fused_out) {
          // per-row sum of squares over this warp's columns of the tile: 8 lanes share a row; fixed reduction order
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            float v = ssacc[it];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            const int rr = row_base + it * 4 + (lane >> 3);
            const int part = (n0 / BN) * (kWide ? 2 : 1) + half;
            if ((lane & 7) == 0 && rr < g.M) g.ss_out[static_cast<size_t>(part) * g.M + rr] = v;
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int col0 = n0 + c * 32;
          uint32_t v[32];
          tmem_ld_x32(t_addr + c * 32, v);
          tmem_wait_ld();
          if constexpr (EPI == VNB_EPI_QKV) {
            if (col0 >= g.d2) {
              // v : transposed (B, d, Tpad) so that attention's P.V B-operand is K-major over keys.
              // lane == row == consecutive t: each of the 32 stores is one contiguous 64-byte segment.
              if (row_ok) {
                const int d = g.N - g.d2;
                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(g.out2) +
                                   (static_cast<size_t>(b_idx) * d + (col0 - g.d2)) * g.Tpad + t_idx;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  o[static_cast<size_t>(j) * g.Tpad] = __float2bfloat16_rn(__uint_as_float(v[j]) * rs);
              }
              continue;
            }
          }
          float r[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __uint_as_float(v[j]) * rs;
          stage_rows(stg, lane, r);
          if constexpr (EPI == VNB_EPI_BF16) {
            drain_bf16(reinterpret_cast<__nv_bfloat16*>(g.out), g.N, g.M, stg, lane, row_base, col0);
          } else {
            drain_bf16(reinterpret_cast<__nv_bfloat16*>(g.out), g.d2, g.M, stg, lane, row_base, col0);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) {
          const uint32_t bar0 = mapa_u32(smem_u32(&tempty_bar[acc]), 0);
          mbar_arrive_remote_cta(bar0);  // .cta-scope release, see the note above the kernel
        }
        else mbar_arrive(&tempty_bar[acc]);
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // neither CTA may free TMEM / exit while the other still uses the pair
  else __syncthreads();
  if (warp == 2) {
    if constexpr (PAIR) tmem_dealloc_pair<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
static int g_num_sms = 0;

// Single-CTA (128 x 256) or CTA-pair (256 x 256, cta_group::2) kernel: vnb_set_option("gemm_pair", 0|1), else the
// environment variable VNB_GEMM_PAIR, else the compiled default.
static int g_gemm_pair = -1;
void set_gemm_pair(int on) { g_gemm_pair = on ? 1 : 0; }
static bool gemm_pair_enabled() {
  if (g_gemm_pair < 0) {
    const char* e = getenv("VNB_GEMM_PAIR");
    g_gemm_pair = e != nullptr ? (e[0] == '1') : (VNB_GEMM_PAIR_DEFAULT ? 1 : 0);
  }
  return g_gemm_pair == 1;
}
int get_gemm_pair() { return gemm_pair_enabled() ? 1 : 0; }

// Once per device and epilogue: opt in to the large dynamic shared memory for both tile variants and ask how many CTA
// pairs can be co-resident (one CTA per SM, both SMs of a TPC).  Called eagerly by prepare_gemm() (model creation), so
// that none of this runs inside a stream capture.
static int g_max_clusters[6][64];
template <int EPI>
static cudaError_t init_epi() {
  static PerDeviceOnce once;
  int dev;
  if (!once.need(&dev)) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<EPI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tcgen05_kernel<EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t q = {};
  q.gridDim = dim3(2 * device_sm_count());
  q.blockDim = dim3(gemm_threads<EPI>());
  q.dynamicSmemBytes = GEMM_SMEM;
  cudaLaunchAttribute qa[1];
  qa[0].id = cudaLaunchAttributeClusterDimension;
  qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
  q.attrs = qa; q.numAttrs = 1;
  int n = 0;
  e = cudaOccupancyMaxActiveClusters(&n, gemm_tcgen05_kernel<EPI, true>, &q);
  if (e != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = device_sm_count() / 2; }
  if (dev >= 0 && dev < 64) g_max_clusters[EPI][dev] = n;
  once.mark(dev);
  return cudaSuccess;
}
cudaError_t prepare_gemm() {
  cudaError_t e;
  if ((e = init_epi<VNB_EPI_BF16>()) != cudaSuccess) return e;
  if ((e = init_epi<VNB_EPI_QKV>()) != cudaSuccess) return e;
  if ((e = init_epi<VNB_EPI_RESID>()) != cudaSuccess) return e;
  if ((e = init_epi<VNB_EPI_GEGLU>()) != cudaSuccess) return e;
  if ((e = init_epi<VNB_EPI_BIAS_F32>()) != cudaSuccess) return e;
  return init_epi<VNB_EPI_SAMPLE>();
}

int get_gemm_max_clusters() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (prepare_gemm() != cudaSuccess || dev < 0 || dev >= 64) return 0;
  return g_max_clusters[VNB_EPI_RESID][dev];
}

template <int EPI>
static cudaError_t launch_epi(const GemmPlan& p, const GemmArgs& g, cudaStream_t st) {
  cudaError_t e = init_epi<EPI>();
  if (e != cudaSuccess) return e;
  int dev = 0;
  g_num_sms = device_sm_count();
  if (gemm_pair_enabled()) {
    cudaGetDevice(&dev);
    const int cap = (dev >= 0 && dev < 64 && g_max_clusters[EPI][dev] > 0) ? g_max_clusters[EPI][dev] : g_num_sms / 2;
    const int tiles = ((g.M + 2 * BM - 1) / (2 * BM)) * (g.N / BN);
    const int clusters = tiles < cap ? tiles : cap;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(gemm_threads<EPI>());
    cfg.dynamicSmemBytes = GEMM_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<EPI, true>, p.tmA, p.tmBh, g);
  }
  const int tiles = ((g.M + BM - 1) / BM) * (g.N / BN);
  const int grid = tiles < g_num_sms ? tiles : g_num_sms;
  gemm_tcgen05_kernel<EPI, false><<<grid, gemm_threads<EPI>(), GEMM_SMEM, st>>>(p.tmA, p.tmB, g);
  return cudaGetLastError();
}

cudaError_t launch_gemm(const GemmPlan& p, cudaStream_t st) {
  GemmArgs g;
  g.M = p.M; g.N = p.N; g.K = p.K; g.epi = p.epi; g.out = p.out; g.out2 = p.out2; g.bias = p.bias;
  g.T = p.T; g.Tpad = p.Tpad; g.d2 = p.d2;
  g.out_bf16 = reinterpret_cast<__nv_bfloat16*>(p.out_bf16); g.ss_out = p.ss_out; g.ss_in = p.ss_in;
  g.ss_parts = p.ss_parts; g.inv_d = p.inv_d; g.eps = p.eps;
  g.zcur = p.zcur; g.dyn = p.dyn; g.partials = reinterpret_cast<float4*>(p.partials);
  g.C = p.C; g.ncc = p.ncc; g.V = p.V; g.mask_token = p.mask_token;
  switch (p.epi) {
    case VNB_EPI_BF16: return launch_epi<VNB_EPI_BF16>(p, g, st);
    case VNB_EPI_QKV: return launch_epi<VNB_EPI_QKV>(p, g, st);
    case VNB_EPI_RESID: return launch_epi<VNB_EPI_RESID>(p, g, st);
    case VNB_EPI_GEGLU: return launch_epi<VNB_EPI_GEGLU>(p, g, st);
    case VNB_EPI_BIAS_F32: return launch_epi<VNB_EPI_BIAS_F32>(p, g, st);
    case VNB_EPI_SAMPLE:
      if (!p.zcur || !p.dyn || !p.partials || !p.bias || p.V % 128 != 0 || p.V > 1024) return cudaErrorInvalidValue;
      return launch_epi<VNB_EPI_SAMPLE>(p, g, st);
    default: return cudaErrorInvalidValue;
  }
}

// ------------------------------------------------------------------------------------------------
// Naive SIMT GEMM (test-only bisecting aid; never on the product path).
__global__ void gemm_ref_kernel(const __nv_bfloat16* A, const __nv_bfloat16* W, int M, int N, int K, float* out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k)
    acc += __bfloat162float(A[static_cast<size_t>(m) * K + k]) * __bfloat162float(W[static_cast<size_t>(n) * K + k]);
  out[static_cast<size_t>(m) * N + n] = acc;
}
cudaError_t launch_gemm_ref(const void* A, const void* W, int M, int N, int K, float* out, cudaStream_t st) {
  dim3 grid((N + 127) / 128, M);
  gemm_ref_kernel<<<grid, 128, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(A),
                                        reinterpret_cast<const __nv_bfloat16*>(W), M, N, K, out);
  return cudaGetLastError();
}

}  // namespace vnb
