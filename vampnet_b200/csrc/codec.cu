// vampnet_b200 — DAC-family codec kernels: strided / dilated 1-D
// convolutions with the Snake activation fused into the operand load, transposed convolutions as
// per-phase 2-tap convolutions, and the residual vector quantiser (encode, from_codes, from_latents).
//
// The reference's codec (lac.model.lac.LAC) is an un-vendored third-party DAC fork; the architecture
// follows the published DAC as in transformers/models/dac/modeling_dac.py (cited per kernel), see
// oracle/dac_oracle.py.  Call sites replaced: codec.encode (reference vampnet/interface.py:223),
// codec.quantizer.from_latents + codec.decode (vampnet/modules/transformer.py:671-675).
//
// fp32 CUDA-core kernels on (B, C, T) channels-first tensors (coalesced along T), per the north-star;
// the wide decoder layers are FLOP-bound on the fp32 pipe, the narrow late layers HBM-bound.
#include "common.cuh"
#include "kernels.h"

namespace vnb {

// ------------------------------------------------------------------------------------------------
// Generic direct convolution:
//   y[b, co, q*out_stride + out_off] = bias[co] + sum_ci sum_j W[co, ci, j] * act(x[b, ci, q*stride + j*dil - pad])
//   (+ residual[b, co, same index]) (tanh optional), act = snake_alpha[ci] or identity.
// Conv1d: out_stride=1, out_off=0.  ConvTranspose1d phase r: stride=1, dil=-1, pad=0, K=2, out_stride=s,
// out_off = r - pad_t, weights repacked per phase on the host (codec.py).
// Block = 64 output channels x 128 output positions; 256 threads, each 8 channels x 4 positions.
constexpr int CV_CO = 64, CV_TT = 128, CV_CI = 8;

struct ConvArgs {
  const float* x; const float* w; const float* bias; const float* alpha; const float* resid; float* y;
  int B, Cin, Tin, Cout, Tout, K, stride, dil, pad, out_stride, out_off, nq, do_tanh;
};

__device__ __forceinline__ float snake_f(float v, float a) {
  const float s = sinf(a * v);
  return v + s * s / (a + 1e-9f);
}

__global__ void __launch_bounds__(256) conv1d_kernel(const ConvArgs a) {
  extern __shared__ float sm[];
  const int span = (CV_TT - 1) * a.stride + (a.K - 1) * abs(a.dil) + 1;  // x positions needed per channel
  float* xs = sm;                                  // [CV_CI][span]
  float* ws = sm + CV_CI * span;                   // [CV_CI][K][CV_CO]
  const int q0 = blockIdx.x * CV_TT;
  const int co0 = blockIdx.y * CV_CO;
  const int b = blockIdx.z;
  const int tx = threadIdx.x & 31;                 // lane: positions tx, tx+32, tx+64, tx+96 (bank-conflict free)
  const int ty = threadIdx.x >> 5;                 // 8 channel groups of 8
  // first x index of the tile: min over taps of q0*stride + j*dil - pad
  const int x_base = q0 * a.stride - a.pad + (a.dil < 0 ? (a.K - 1) * a.dil : 0);
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int ci0 = 0; ci0 < a.Cin; ci0 += CV_CI) {
    const int nci = min(CV_CI, a.Cin - ci0);
    __syncthreads();
    for (int i = threadIdx.x; i < nci * span; i += 256) {
      const int c = i / span, p = i - c * span;
      const int xi = x_base + p;
      float v = 0.f;
      if (xi >= 0 && xi < a.Tin) {
        v = a.x[(static_cast<size_t>(b) * a.Cin + ci0 + c) * a.Tin + xi];
        if (a.alpha) v = snake_f(v, a.alpha[ci0 + c]);  // zero padding is applied AFTER the activation
      }
      xs[c * span + p] = v;
    }
    for (int i = threadIdx.x; i < nci * a.K * CV_CO; i += 256) {
      const int co = i % CV_CO, r = i / CV_CO;  // r = c*K + j
      const int c = r / a.K, j = r - c * a.K;
      float v = 0.f;
      if (co0 + co < a.Cout) v = a.w[(static_cast<size_t>(co0 + co) * a.Cin + ci0 + c) * a.K + j];
      ws[r * CV_CO + co] = v;
    }
    __syncthreads();
    for (int c = 0; c < nci; ++c) {
      for (int j = 0; j < a.K; ++j) {
        const int off = j * a.dil - (a.dil < 0 ? (a.K - 1) * a.dil : 0);  // relative to x_base + q*stride
        float xv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) xv[t] = xs[c * span + (t * 32 + tx) * a.stride + off];
        const float4 w0 = *reinterpret_cast<const float4*>(&ws[(c * a.K + j) * CV_CO + ty * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&ws[(c * a.K + j) * CV_CO + ty * 8 + 4]);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int t = 0; t < 4; ++t) acc[i][t] = fmaf(wv[i], xv[t], acc[i][t]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int co = co0 + ty * 8 + i;
    if (co >= a.Cout) continue;
    const float bv = a.bias ? a.bias[co] : 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int q = q0 + t * 32 + tx;
      const int yo = q * a.out_stride + a.out_off;
      if (q < a.nq && yo >= 0 && yo < a.Tout) {
        const size_t o = (static_cast<size_t>(b) * a.Cout + co) * a.Tout + yo;
        float v = acc[i][t] + bv;
        if (a.resid) v += a.resid[o];
        if (a.do_tanh) v = tanhf(v);
        a.y[o] = v;
      }
    }
  }
}

cudaError_t launch_conv1d(const float* x, const float* w, const float* bias, const float* alpha, const float* resid,
                          float* y, int B, int Cin, int Tin, int Cout, int Tout, int K, int stride, int dil, int pad,
                          int out_stride, int out_off, int nq, int do_tanh, cudaStream_t st) {
  ConvArgs a{x, w, bias, alpha, resid, y, B, Cin, Tin, Cout, Tout, K, stride, dil, pad, out_stride, out_off, nq, do_tanh};
  const int span = (CV_TT - 1) * stride + (K - 1) * abs(dil) + 1;
  const size_t smem = (static_cast<size_t>(CV_CI) * span + static_cast<size_t>(CV_CI) * K * CV_CO) * sizeof(float);
  if (smem > 48 * 1024) {  // idempotent; per call so that it holds on whichever device is current
    cudaError_t e = cudaFuncSetAttribute(conv1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  dim3 grid((nq + CV_TT - 1) / CV_TT, (Cout + CV_CO - 1) / CV_CO, B);
  conv1d_kernel<<<grid, 256, smem, st>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Residual vector quantiser (modeling_dac.py:102-169, 271-402).  One CTA handles RQ_T frames of one clip
// and walks the levels sequentially with the residual held in shared memory.
//   mode 0 (encode)      : in = z (B, D, T)        -> codes (B, L, T) int64, zq (B, D, T), latents (B, 8L, T)
//   mode 1 (from_latents): in = latents (B, 8L, T) -> zq (B, D, T)            [VampNet.decode path]
//   mode 2 (from_codes)  : in = codes (B, L, T)    -> zq (B, D, T)
constexpr int RQ_T = 8;
constexpr int RQ_MAXD = 1024;

struct RvqArgs {
  const float* in_f; const int64_t* in_codes;
  const float* win;   // (L, 8, D)
  const float* bin;   // (L, 8)
  const float* wout;  // (L, D, 8)
  const float* bout;  // (L, D)
  const float* cb;    // (L, V, 8) raw codebooks
  const float* cbn;   // (L, V, 8) L2-normalised codebooks (F.normalize, eps 1e-12)
  int64_t* codes; float* zq; float* latents;
  int B, D, T, L, V, mode;
  int cl;                       // 1: in_f (mode 0) and zq are channels-last (B, T, D) -- tensor-core codec path
  __nv_bfloat16* zq_hi; __nv_bfloat16* zq_lo;  // optional split-bf16 copy of zq (channels-last only)
};

__global__ void __launch_bounds__(256) rvq_kernel(const RvqArgs a) {
  __shared__ float res[RQ_T][RQ_MAXD];
  __shared__ float e[RQ_T][8], en[RQ_T][8], qv[RQ_T][8];
  __shared__ int sel[RQ_T];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * RQ_T;
  const int nt = min(RQ_T, a.T - t0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = a.D;
  // residual <- z (encode) ; zq accumulates in registers-free fashion directly in `acc` smem alias: reuse res for
  // the residual and accumulate zq in global at the end via (z - residual) for encode, explicit sum otherwise.
  if (a.mode == 0) {
    for (int i = threadIdx.x; i < D * RQ_T; i += 256) {
      const int c = i / RQ_T, t = i - c * RQ_T;
      res[t][c] = t < nt ? (a.cl ? a.in_f[(static_cast<size_t>(b) * a.T + t0 + t) * D + c]
                               : a.in_f[(static_cast<size_t>(b) * D + c) * a.T + t0 + t]) : 0.f;
    }
  } else {
    for (int i = threadIdx.x; i < D * RQ_T; i += 256) res[i / D][i % D] = 0.f;  // here `res` accumulates zq
  }
  __syncthreads();
  for (int l = 0; l < a.L; ++l) {
    // ---- (1) 8-d latent of every frame
    if (a.mode == 0) {
      // e[t][j] = bin[j] + <win[l][j][:], res[t][:]> : warp w owns frame w (RQ_T == 8 warps), lanes split D
      const int t = warp;
      float part[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) part[j] = 0.f;
      for (int c = lane; c < D; c += 32) {
        const float r = res[t][c];
#pragma unroll
        for (int j = 0; j < 8; ++j) part[j] = fmaf(a.win[(static_cast<size_t>(l) * 8 + j) * D + c], r, part[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float s = warp_sum(part[j]);
        if (lane == 0) e[t][j] = s + a.bin[l * 8 + j];
      }
    } else if (a.mode == 1) {
      if (threadIdx.x < RQ_T * 8) {
        const int t = threadIdx.x >> 3, j = threadIdx.x & 7;
        e[t][j] = t < nt ? a.in_f[(static_cast<size_t>(b) * a.L * 8 + l * 8 + j) * a.T + t0 + t] : 0.f;
      }
    }
    __syncthreads();
    if (a.mode != 2) {
      if (a.mode == 0 && threadIdx.x < RQ_T * 8) {
        const int t = threadIdx.x >> 3, j = threadIdx.x & 7;
        if (t < nt) a.latents[(static_cast<size_t>(b) * a.L * 8 + l * 8 + j) * a.T + t0 + t] = e[t][j];
      }
      // ---- (2) L2-normalise (F.normalize: x / max(|x|, 1e-12))
      if (threadIdx.x < RQ_T) {
        const int t = threadIdx.x;
        float n2 = 0.f;
        for (int j = 0; j < 8; ++j) n2 += e[t][j] * e[t][j];
        const float n = fmaxf(sqrtf(n2), 1e-12f);
        for (int j = 0; j < 8; ++j) en[t][j] = e[t][j] / n;
      }
      __syncthreads();
      // ---- (3) nearest code: argmax_v  -(|en|^2 - 2 en.cn_v) + |cn_v|^2   (modeling_dac.py:163-166)
      // warp w owns frame w; lanes split the V codes; first index wins ties
      {
        const int t = warp;
        float en2 = 0.f;
        for (int j = 0; j < 8; ++j) en2 += en[t][j] * en[t][j];
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int v = lane; v < a.V; v += 32) {
          const float4 c0 = *reinterpret_cast<const float4*>(a.cbn + (static_cast<size_t>(l) * a.V + v) * 8);
          const float4 c1 = *reinterpret_cast<const float4*>(a.cbn + (static_cast<size_t>(l) * a.V + v) * 8 + 4);
          const float cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
          float dot = 0.f, c2 = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) { dot = fmaf(en[t][j], cv[j], dot); c2 = fmaf(cv[j], cv[j], c2); }
          const float dist = -(en2 - 2.f * dot) + c2;
          if (dist > bv) { bv = dist; bi = v; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) sel[t] = bi;
      }
      __syncthreads();
    } else {
      if (threadIdx.x < RQ_T) sel[threadIdx.x] = threadIdx.x < nt ? static_cast<int>(a.in_codes[(static_cast<size_t>(b) * a.L + l) * a.T + t0 + threadIdx.x]) : 0;
      __syncthreads();
    }
    // ---- (4) chosen code vector; from_latents feeds out_proj with chunk + (q - chunk) (straight-through form)
    if (threadIdx.x < RQ_T * 8) {
      const int t = threadIdx.x >> 3, j = threadIdx.x & 7;
      const float q = a.cb[(static_cast<size_t>(l) * a.V + sel[t]) * 8 + j];
      qv[t][j] = a.mode == 1 ? e[t][j] + (q - e[t][j]) : q;
      if (a.mode == 0 && j == 0 && t < nt) a.codes[(static_cast<size_t>(b) * a.L + l) * a.T + t0 + t] = sel[t];
    }
    __syncthreads();
    // ---- (5) out_proj and residual / accumulator update
    for (int c = threadIdx.x; c < D; c += 256) {
      const float4 w0 = *reinterpret_cast<const float4*>(a.wout + (static_cast<size_t>(l) * D + c) * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(a.wout + (static_cast<size_t>(l) * D + c) * 8 + 4);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float bo = a.bout[static_cast<size_t>(l) * D + c];
#pragma unroll
      for (int t = 0; t < RQ_T; ++t) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s = fmaf(wv[j], qv[t][j], s);
        s += bo;
        if (a.mode == 0) res[t][c] -= s; else res[t][c] += s;
      }
    }
    __syncthreads();
  }
  // ---- write zq: encode -> z - final residual ; others -> accumulated sum
  for (int i = threadIdx.x; i < D * RQ_T; i += 256) {
    const int c = i / RQ_T, t = i - c * RQ_T;
    if (t >= nt) continue;
    const size_t o = a.cl ? (static_cast<size_t>(b) * a.T + t0 + t) * D + c : (static_cast<size_t>(b) * D + c) * a.T + t0 + t;
    const float v = a.mode == 0 ? a.in_f[o] - res[t][c] : res[t][c];
    a.zq[o] = v;
    if (a.zq_hi) {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      a.zq_hi[o] = h;
      a.zq_lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

cudaError_t launch_rvq(int mode, const float* in_f, const int64_t* in_codes, const float* win, const float* bin,
                       const float* wout, const float* bout, const float* cb, const float* cbn, int64_t* codes,
                       float* zq, float* latents, int B, int D, int T, int L, int V, cudaStream_t st, int cl = 0,
                       void* zq_hi = nullptr, void* zq_lo = nullptr) {
  if (D > RQ_MAXD) return cudaErrorInvalidValue;
  RvqArgs a{in_f, in_codes, win, bin, wout, bout, cb, cbn, codes, zq, latents, B, D, T, L, V, mode, cl,
            reinterpret_cast<__nv_bfloat16*>(zq_hi), reinterpret_cast<__nv_bfloat16*>(zq_lo)};
  dim3 grid((T + RQ_T - 1) / RQ_T, B);
  rvq_kernel<<<grid, 256, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace vnb

using namespace vnb;
extern "C" {
int32_t vnb_codec_conv1d(const float* x, const float* w, const float* bias, const float* snake_alpha,
                         const float* residual, float* y, int32_t B, int32_t Cin, int32_t Tin, int32_t Cout,
                         int32_t Tout, int32_t K, int32_t stride, int32_t dil, int32_t pad, int32_t out_stride,
                         int32_t out_off, int32_t nq, int32_t do_tanh, void* stream) {
  cudaError_t e = launch_conv1d(x, w, bias, snake_alpha, residual, y, B, Cin, Tin, Cout, Tout, K, stride, dil, pad,
                                out_stride, out_off, nq, do_tanh, reinterpret_cast<cudaStream_t>(stream));
  if (e == cudaSuccess) count_launch();
  return e == cudaSuccess ? 0 : vnb_set_error_cuda("vnb_codec_conv1d", static_cast<int>(e));
}
int32_t vnb_codec_rvq(int32_t mode, const float* in_f, const int64_t* in_codes, const float* win, const float* bin,
                      const float* wout, const float* bout, const float* cb, const float* cbn, int64_t* codes, float* zq,
                      float* latents, int32_t B, int32_t D, int32_t T, int32_t L, int32_t V, int32_t channels_last,
                      void* zq_hi, void* zq_lo, void* stream) {
  cudaError_t e = launch_rvq(mode, in_f, in_codes, win, bin, wout, bout, cb, cbn, codes, zq, latents, B, D, T, L, V,
                             reinterpret_cast<cudaStream_t>(stream), channels_last, zq_hi, zq_lo);
  if (e == cudaSuccess) count_launch();
  return e == cudaSuccess ? 0 : vnb_set_error_cuda("vnb_codec_rvq", static_cast<int>(e));
}
}
