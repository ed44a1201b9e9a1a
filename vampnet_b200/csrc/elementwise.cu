// vampnet_b200 — HBM-bound helper kernels of the forward pass: codebook-embedding gather + 1x1
// projection, RMSNorm, and the (B,C,T) int64 <-> (B,T,C) int32 state conversions of generate().
#include "common.cuh"
#include "kernels.h"

namespace vnb {

// ------------------------------------------------------------------------------------------------
// RMSNorm (reference vampnet/modules/transformer.py:43-58): y = w * (x * rsqrt(mean(x^2) + eps)).
// fp32 in (the residual stream), bf16 out (the A operand of the next GEMM).  One warp per row,
// float4 loads; algorithmic bytes = 6 B per element (4 read + 2 write).
__global__ void __launch_bounds__(256) rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      __nv_bfloat16* __restrict__ y, int M, int d, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * d);
  const float4* wr = reinterpret_cast<const float4*>(w);
  const int n4 = d >> 2;
  float4 v[16];  // d <= 2048
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const int i = lane + 32 * c;
    if (i < n4) {
      v[c] = xr[i];
      ss += v[c].x * v[c].x + v[c].y * v[c].y + v[c].z * v[c].z + v[c].w * v[c].w;
    }
  }
  ss = warp_sum(ss);
  const float r = rsqrtf(ss / static_cast<float>(d) + eps);
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * d);
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const int i = lane + 32 * c;
    if (i < n4) {
      const float4 ww = __ldg(wr + i);
      uint2 o;
      o.x = pack_bf16x2(ww.x * (v[c].x * r), ww.y * (v[c].y * r));
      o.y = pack_bf16x2(ww.z * (v[c].z * r), ww.w * (v[c].w * r));
      yr[i] = o;
    }
  }
}

cudaError_t launch_rmsnorm(const float* x, const float* w, void* y, int M, int d, float eps, cudaStream_t st) {
  if (d % 4 != 0 || d > 2048) return cudaErrorInvalidValue;
  const int rows_per_block = 8;
  rmsnorm_kernel<<<(M + rows_per_block - 1) / rows_per_block, rows_per_block * 32, 0, st>>>(
      x, w, reinterpret_cast<__nv_bfloat16*>(y), M, d, eps);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// CodebookEmbedding.from_codes + out_proj (reference vampnet/modules/layers.py:134-162):
//   latent[m, c*8 + j] = table[c][code[m, c]][j]        (code == V selects the learned MASK row)
//   x[m, n] = bias[n] + sum_k Wt[k, n] * latent[m, k]    (Conv1d kernel 1)
// K = 8*C is 32 (coarse) or 112 (c2f): a CUDA-core fp32 contraction, negligible next to the GEMMs.
constexpr int EMB_ROWS = 16;
constexpr int EMB_MAXK = 128;

template <bool FROM_CODES>
__global__ void __launch_bounds__(256) embed_kernel(const int32_t* __restrict__ codes, const float* __restrict__ lat_in,
                                                    const float* __restrict__ table, const float* __restrict__ wt,
                                                    const float* __restrict__ bias, float* __restrict__ x, int M, int T,
                                                    int C, int V1, int K, int d, __nv_bfloat16* __restrict__ xb,
                                                    float* __restrict__ ss, int ss_parts) {
  __shared__ __align__(16) float lat[EMB_ROWS][EMB_MAXK];
  __shared__ float red[8][EMB_ROWS];
  const int m0 = blockIdx.x * EMB_ROWS;
  for (int i = threadIdx.x; i < EMB_ROWS * K; i += blockDim.x) {
    const int r = i / K, k = i - r * K;
    const int m = m0 + r;
    float v = 0.f;
    if (m < M) {
      if constexpr (FROM_CODES) {
        const int c = k >> 3, j = k & 7;
        const int code = codes[static_cast<size_t>(m) * C + c];
        v = table[(static_cast<size_t>(c) * V1 + code) * 8 + j];
      } else {
        const int b = m / T, t = m - b * T;  // latents (B, K, T)
        v = lat_in[(static_cast<size_t>(b) * K + k) * T + t];
      }
    }
    lat[r][k] = v;
  }
  __syncthreads();
  float ssr[EMB_ROWS];
#pragma unroll
  for (int r = 0; r < EMB_ROWS; ++r) ssr[r] = 0.f;
  for (int n = threadIdx.x; n < d; n += blockDim.x) {
    float acc[EMB_ROWS];
    const float bn = bias[n];
#pragma unroll
    for (int r = 0; r < EMB_ROWS; ++r) acc[r] = 0.f;
    for (int k = 0; k < K; k += 4) {  // K = 8*C is a multiple of 4: one LDS.128 feeds four FMAs per row
      const float w0 = __ldg(wt + static_cast<size_t>(k) * d + n), w1 = __ldg(wt + static_cast<size_t>(k + 1) * d + n);
      const float w2 = __ldg(wt + static_cast<size_t>(k + 2) * d + n), w3 = __ldg(wt + static_cast<size_t>(k + 3) * d + n);
#pragma unroll
      for (int r = 0; r < EMB_ROWS; ++r) {
        const float4 l = *reinterpret_cast<const float4*>(&lat[r][k]);
        acc[r] = fmaf(w0, l.x, acc[r]);
        acc[r] = fmaf(w1, l.y, acc[r]);
        acc[r] = fmaf(w2, l.z, acc[r]);
        acc[r] = fmaf(w3, l.w, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < EMB_ROWS; ++r) {
      if (m0 + r < M) {
        const float v = acc[r] + bn;
        x[static_cast<size_t>(m0 + r) * d + n] = v;
        if (xb) xb[static_cast<size_t>(m0 + r) * d + n] = __float2bfloat16_rn(v);
        ssr[r] += v * v;
      }
    }
  }
  if (ss) {  // row sums of squares for the RMSNorm fused into the next GEMM (fixed reduction order)
#pragma unroll
    for (int r = 0; r < EMB_ROWS; ++r) {
      const float v = warp_sum(ssr[r]);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][r] = v;
    }
    __syncthreads();
    if (threadIdx.x < EMB_ROWS && m0 + threadIdx.x < M) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
      ss[m0 + threadIdx.x] = t;
      for (int p = 1; p < ss_parts; ++p) ss[static_cast<size_t>(p) * M + m0 + threadIdx.x] = 0.f;
    }
  }
}

cudaError_t launch_embed_codes(const int32_t* codes_btc, const float* table, const float* wt, const float* b, float* x,
                               int M, int C, int V1, int d, cudaStream_t st, void* xb, float* ss, int ss_parts) {
  if (C * 8 > EMB_MAXK) return cudaErrorInvalidValue;
  embed_kernel<true><<<(M + EMB_ROWS - 1) / EMB_ROWS, 256, 0, st>>>(
      codes_btc, nullptr, table, wt, b, x, M, 1, C, V1, C * 8, d, reinterpret_cast<__nv_bfloat16*>(xb), ss, ss_parts);
  return cudaGetLastError();
}
cudaError_t launch_embed_latents(const float* lat, const float* wt, const float* b, float* x, int B, int T, int K,
                                 int d, cudaStream_t st, void* xb, float* ss, int ss_parts) {
  if (K > EMB_MAXK) return cudaErrorInvalidValue;
  const int M = B * T;
  embed_kernel<false><<<(M + EMB_ROWS - 1) / EMB_ROWS, 256, 0, st>>>(
      nullptr, lat, nullptr, wt, b, x, M, T, K / 8, 0, K, d, reinterpret_cast<__nv_bfloat16*>(xb), ss, ss_parts);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// generate() set-up (reference transformer.py:749-766): z_masked = z.masked_fill(mask, MASK);
// N0 = count(z_masked == MASK) over the WHOLE batch.  State is kept as (B, T, C) int32 so that the
// codes of one frame are contiguous for the embedding gather and "b (t c)" flattening (util.py:39)
// of the predicted codebooks is a plain stride.
__global__ void gen_init_kernel(const int64_t* __restrict__ z, const int32_t* __restrict__ mask,
                                int32_t* __restrict__ zcur, int32_t* __restrict__ zorig, int32_t* __restrict__ n0,
                                int B, int C, int T, int ncc, int mask_token) {
  const int total = B * C * T;
  int local = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int t = i % T;
    const int c = (i / T) % C;
    const int b = i / (T * C);
    const int v = static_cast<int>(z[i]);
    const int mk = mask ? mask[i] : (c >= ncc ? 1 : 0);  // default mask, transformer.py:749-751
    const int vm = mk ? mask_token : v;
    const size_t o = (static_cast<size_t>(b) * T + t) * C + c;
    zorig[o] = v;
    zcur[o] = vm;
    local += (vm == mask_token);
  }
  // block reduce then one atomic
  __shared__ int sh[32];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0 && v) atomicAdd(n0, v);
  }
}

cudaError_t launch_gen_init(const int64_t* z, const int32_t* mask, int32_t* zcur, int32_t* zorig, int32_t* n0, int B,
                            int C, int T, int ncc, int mask_token, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(n0, 0, sizeof(int32_t), st);
  if (e != cudaSuccess) return e;
  const int total = B * C * T;
  int grid = (total + 255) / 256;
  if (grid > 1184) grid = 1184;
  gen_init_kernel<<<grid, 256, 0, st>>>(z, mask, zcur, zorig, n0, B, C, T, ncc, mask_token);
  return cudaGetLastError();
}

// sampled_z (B, T, Cp) + conditioning codebooks of the ORIGINAL z -> (B, C, T) int64 (transformer.py:935-938)
__global__ void gen_finish_kernel(const int32_t* __restrict__ tokens, const int32_t* __restrict__ zorig,
                                  int64_t* __restrict__ out, int B, int C, int T, int ncc) {
  const int total = B * C * T;
  const int Cp = C - ncc;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int t = i % T;
    const int c = (i / T) % C;
    const int b = i / (T * C);
    const size_t bt = static_cast<size_t>(b) * T + t;
    out[i] = c < ncc ? zorig[bt * C + c] : tokens[bt * Cp + (c - ncc)];
  }
}
cudaError_t launch_gen_finish(const int32_t* tokens, const int32_t* zorig, int64_t* out, int B, int C, int T, int ncc,
                              cudaStream_t st) {
  const int total = B * C * T;
  int grid = (total + 255) / 256;
  if (grid > 1184) grid = 1184;
  gen_finish_kernel<<<grid, 256, 0, st>>>(tokens, zorig, out, B, C, T, ncc);
  return cudaGetLastError();
}

}  // namespace vnb
