// vampnet_b200 — internal launcher declarations shared by the .cu translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vampnet_b200.h"

namespace vnb {

// Function attributes (max dynamic smem) are per device: remember them per (kernel, device), not per process.
struct PerDeviceOnce {
  unsigned long long done = 0;  // bit d = attribute already set on device d
  bool need(int* dev_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    *dev_out = dev;
    return dev >= 64 || !((done >> dev) & 1ull);
  }
  void mark(int dev) { if (dev < 64) done |= 1ull << dev; }
};
inline int device_sm_count() {
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 64) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }
  if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
  return sms[dev];
}

// kernels launched by this library so far (vnb_launch_count; graph replays add their node count)
void count_launch(unsigned long long n = 1);

// ---- TMA tensor maps (driver entry point fetched at run time; no link-time libcuda dependency) ----
// 2-D bf16 row-major (rows, cols) with a (box_rows x 64) box, 128B swizzle.
bool make_tmap_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                  uint32_t box_cols);
// 3-D bf16 (batch, rows, cols) row-major with row pitch `pitch_elems`; box (1, box_rows, 64), 128B swizzle.
bool make_tmap_3d(CUtensorMap* tm, const void* base, uint64_t batch, uint64_t rows, uint64_t cols,
                  uint64_t pitch_elems, uint32_t box_rows, uint32_t box_cols);
const char* tmap_error();
}  // namespace vnb
// sets the thread-local error string from a CUDA error code; returns 1
extern "C" int32_t vnb_set_error_cuda(const char* what, int32_t cuda_error);
namespace vnb {

struct SampleDyn;

// ---- GEMM ----
struct GemmPlan {
  CUtensorMap tmA, tmB;
  CUtensorMap tmBh;  // W with a 128-row box: the half tile each CTA of a pair stages (gemm_tcgen05.cu, PAIR)
  int M = 0, N = 0, K = 0, epi = 0;
  void* out = nullptr;
  void* out2 = nullptr;
  const float* bias = nullptr;
  int T = 1, Tpad = 1, d2 = 0;
  // fused RMSNorm plumbing (see GemmArgs in gemm_tcgen05.cu)
  void* out_bf16 = nullptr;
  float* ss_out = nullptr;
  const float* ss_in = nullptr;
  int ss_parts = 0;
  float inv_d = 0.f, eps = 0.f;
  // VNB_EPI_SAMPLE (the classifier of the generate loop): the logits are sampled in the epilogue instead of being stored
  const int32_t* zcur = nullptr;     // (B, T, C) current tokens: only still-masked positions are sampled
  const SampleDyn* dyn = nullptr;    // this step's scalars (device memory: graph replay safe)
  void* partials = nullptr;          // (M * Cp * V/128) float4 records, see sample_combine_kernel
  int C = 0, ncc = 0, V = 0, mask_token = 0;
};
// Fills the tensor maps; A (M,K) bf16, W (N,K) bf16.
bool make_gemm_plan(GemmPlan* p, int epi, const void* A, const void* W, int M, int N, int K, void* out, void* out2,
                    const float* bias, int T, int Tpad, int d2);
bool gemm_plan_set_fused_out(GemmPlan* p, void* out_bf16, float* ss_out);
cudaError_t launch_gemm(const GemmPlan& p, cudaStream_t st);
cudaError_t prepare_gemm();  // per-device kernel attributes; call outside stream capture
void set_gemm_pair(int on);  // 1: CTA-pair (cta_group::2) GEMM tiles, 0: single-CTA tiles
int get_gemm_pair();
int get_gemm_max_clusters();  // co-resident CTA pairs of the pair kernel on the current device
cudaError_t launch_gemm_ref(const void* A, const void* W, int M, int N, int K, float* out, cudaStream_t st);

// ---- attention ----
struct AttnPlan {
  CUtensorMap tmQ;   // (B, T, 2d)  box (1, 128 rows, 64 cols)
  CUtensorMap tmK;   // (B, T, 2d)  box (1, 64 rows, 64 cols)
  CUtensorMap tmVT;  // (B, d, Tpad) box (1, 64 rows, 64 cols)
  void* out = nullptr;         // (B, T, d) bf16
  const float* rel = nullptr;  // (2*sat+1, H)
  int sat = 0, B = 0, T = 0, Tpad = 0, H = 0;
};
bool make_attn_plan(AttnPlan* p, const void* qk, const void* vT, void* out, const float* rel, int sat, int B, int T,
                    int Tpad, int H);
cudaError_t launch_attention(const AttnPlan& p, cudaStream_t st);

// ---- elementwise / gather ----
// codes_btc (B*T, C) int32 (or latents (B, K, T) fp32 when codes_btc is null) -> A (M, 3*Kp) bf16 = [hi | hi | lo] of the
// gathered latents, the A operand of the out_proj contraction; zeroes ss partials [zero_from, ss_parts) of every row
cudaError_t launch_embed_gather(const int32_t* codes_btc, const float* latents, const float* table, void* A, int M, int T,
                                int C, int V1, int K, int Kp, float* ss, int zero_from, int ss_parts, cudaStream_t st);

// ---- generate-loop state kernels ----
// z (B,C,T) int64, mask (B,C,T) int32|null -> zcur (B,T,C) int32 (masked), zorig (B,T,C) int32; n0 += count(MASK)
cudaError_t launch_gen_init(const int64_t* z, const int32_t* mask, int32_t* zcur, int32_t* zorig, int32_t* n0, int B,
                            int C, int T, int ncc, int mask_token, cudaStream_t st);
// tokens (B, T, Cp) int32 + zorig cond -> out (B, C, T) int64
cudaError_t launch_gen_finish(const int32_t* tokens, const int32_t* zorig, int64_t* out, int B, int C, int T, int ncc,
                              cudaStream_t st);

// per-step dynamic scalars, read from DEVICE memory so that a captured CUDA graph can be replayed with
// new temperatures / seeds (the schedule values are computed on the host with the reference's fp32
// expressions: mask.py:8-9, transformer.py:831-834, 917-919)
struct SampleDyn {
  float inv_temp, gamma, temp_eff;
  int do_sample, is_last, step;
  uint32_t seed_lo, seed_hi;
  float top_p;  // <= 0 or >= 1: disabled (reference transformer.py:1001-1016)
};
struct SampleArgs {
  const float* logits;  // (B*S, V)
  int32_t* zcur;        // (B, T, C) int32, predicted codebooks at c >= ncc ; updated in place by the remask kernel
  const int32_t* zorig; // (B, T, C) int32 or null (then conditioning codebooks are left untouched)
  int32_t* tokens;      // (B, T, Cp) sampled_z
  float* conf;          // (B, S)
  const int32_t* n0;    // device scalar
  int B, T, C, ncc, V, mask_token;
};
// use_top_p selects the kernel variant at launch time (it is baked into a captured graph: part of the graph key)
cudaError_t launch_sample_step_dev(const SampleArgs& a, const SampleDyn* dyn_dev, cudaStream_t st, bool use_top_p);
// fused path: the classifier's sampling epilogue wrote `partials`; picks the tile, writes tokens + confidences, re-masks
cudaError_t launch_sample_combine_dev(const SampleArgs& a, const void* partials, const SampleDyn* dyn_dev, cudaStream_t st);

}  // namespace vnb
