/* vampnet_b200 — C ABI of the B200-native VampNet masked-token generation hot path.
 *
 * The reference (hugofloresgarcia/vampnet) is pure Python/PyTorch and has no FFI of its own; its
 * boundary is the Python class surface.  These entry points are what a binding
 * for that surface binds (INTEGRATION.md shows the ctypes stub); each cites the reference
 * function it replaces.  Plain pointers and sizes only: device pointers are raw CUDA addresses,
 * `stream` is a cudaStream_t passed as void*.  All functions return 0 on success, non-zero on
 * error; vnb_last_error() returns a thread-local message.  Nothing here aborts the process.
 *
 * Handles are not thread-safe (the reference is called from one worker thread at a time:
 * app.py:730 demo.queue()).
 */
#ifndef VAMPNET_B200_H
#define VAMPNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VNB_ABI_VERSION 2

typedef struct vnb_model vnb_model;

/* Mirrors VampNet.__init__ (reference vampnet/modules/transformer.py:536-552). */
typedef struct vnb_config {
  int32_t n_heads;
  int32_t n_layers;
  int32_t n_codebooks;
  int32_t n_conditioning_codebooks;
  int32_t latent_dim; /* 8 */
  int32_t d_model;    /* embedding_dim */
  int32_t vocab_size; /* 1024; mask token id == vocab_size */
} vnb_config;

/* Packed weights, all DEVICE pointers owned by the caller and kept alive while the model exists.
 * Packing (LoRA fold, weight-norm fold, bf16 rounding, row permutations) is done by the host
 * side (vampnet_b200/modules/transformer.py: pack_weights). */
typedef struct vnb_weights {
  const float* emb_table;  /* (C, V+1, 8)  codec codebooks with the learned MASK row appended (layers.py:145-150) */
  const void* emb_w3;      /* (d, 3*Kp)    bf16, embedding.out_proj.weight (d, 8C) zero-padded to Kp = 8C rounded up to 64
                                           and split w = hi + lo: rows are [hi | lo | hi], the B operand of the
                                           split-bf16 contraction [a_hi | a_hi | a_lo] . [w_hi | w_lo | w_hi]^T
                                           (fp32-grade: the dropped a_lo.w_lo term is 2^-18 relative) (layers.py:132,162) */
  const float* emb_b;      /* (d) */
  const float* norm1;      /* (L, d)       norm_1.weight (informational; the forward uses the folded form) */
  const void* wqkv;        /* (L, 3d, d)   bf16, rows = [w_qs | w_ks | w_vs] (transformer.py:109-114), columns scaled
                                           by norm_1.weight: RMSNorm (transformer.py:43-58) is fused, the kernels
                                           apply rsqrt(mean(x^2)+eps) as a row scale of the GEMM result */
  const void* wo;          /* (L, d, d)    bf16, fc */
  const float* norm3;      /* (L, d)       norm_3.weight */
  const void* w1;          /* (L, 4d, d)   bf16, feed_forward.w_1 (columns scaled by norm_3.weight) with rows interleaved per 256-row tile:
                                           [128 value rows | 128 gate rows] (activations.py:33-35) */
  const void* w2;          /* (L, d, 2d)   bf16, feed_forward.w_2 */
  const float* norm_f;     /* (d)          transformer.norm.weight */
  const void* wcls;        /* (Cp*V, d)    bf16, classifier g*v/|v| (columns scaled by transformer.norm.weight) with rows
                                           permuted to c*V + p (transformer.py:634) */
  const float* bcls;       /* (Cp*V)       permuted the same way */
  const float* rel_bias;   /* (2*rel_sat+1, H) fp32: bias for clamp(key-query, -rel_sat, rel_sat) (transformer.py:123-209) */
  int32_t rel_sat;
} vnb_weights;

/* Mirrors the keyword arguments of VampNet.generate that are live on this path
 * (transformer.py:687-710; SURVEY.md §A.6 lists the dead ones).  The per-step schedule arrays
 * are computed by the host with the reference's own fp32 expressions (mask.py:8-9,
 * transformer.py:831-834, 903, 917-919) so that floor(gamma*N0) matches bit for bit. */
typedef struct vnb_gen_params {
  int32_t sampling_steps;
  float temperature;        /* <=0 means softmax(logits) without scaling (transformer.py:1019-1023) */
  const float* gamma;       /* host, [steps]: _gamma((i+1)/steps) as fp32 */
  const float* temp_eff;    /* host, [steps]: mask_temperature * (1 - r_i) as fp32 */
  const int32_t* do_sample; /* host, [steps]: (i/steps) <= sample_cutoff */
  uint32_t seed_lo, seed_hi; /* Philox key */
  int32_t use_graph;        /* 1: capture the whole loop once per shape and replay it as a CUDA graph */
  float top_p;              /* nucleus filtering on the raw logits (transformer.py:1001-1016); <=0 or >=1: off */
} vnb_gen_params;

int32_t vnb_abi_version(void);
const char* vnb_last_error(void);

/* VampNet.__init__ + load (interface.py:27-50): builds tensor maps / workspace lazily per (B, T). */
int32_t vnb_model_create(const vnb_config* cfg, const vnb_weights* w, vnb_model** out);
void vnb_model_destroy(vnb_model* m);

/* embedding.from_codes + VampNet.forward (layers.py:134-162, transformer.py:617-639).
 * codes: (B, C, T) int64 device.  logits: (B, T*Cp, V) fp32 device, i.e. the reference's
 * (B, V, T*Cp) output transposed (the layout generate() permutes to at transformer.py:849). */
int32_t vnb_forward_codes(vnb_model* m, const int64_t* codes, int32_t B, int32_t T, float* logits, void* stream);
/* VampNet.forward on caller-supplied latents (B, 8C, T) fp32 (transformer.py:617). */
int32_t vnb_forward_latents(vnb_model* m, const float* latents, int32_t B, int32_t T, float* logits, void* stream);
/* VampNet.forward(return_activations=True) (transformer.py:617-639, 443-461): as vnb_forward_latents, and the fp32
 * residual stream after EVERY layer is copied to acts (n_layers, B, T, d_model) — the reference's
 * torch.stack(activations). */
int32_t vnb_forward_latents_acts(vnb_model* m, const float* latents, int32_t B, int32_t T, float* logits, float* acts,
                                 void* stream);
/* Debug tap: copy the fp32 residual stream (B*T, d) after the last layer of the last forward. */
int32_t vnb_get_hidden(vnb_model* m, float* out, void* stream);

/* VampNet.generate(return_signal=False) (transformer.py:686-946).
 * z: (B, C, T) int64; mask: (B, C, T) int32 or NULL (default mask, transformer.py:749-751);
 * out: (B, C, T) int64. */
int32_t vnb_generate(vnb_model* m, const int64_t* z, const int32_t* mask, int32_t B, int32_t T,
                     const vnb_gen_params* p, int64_t* out, void* stream);
/* One sampling iteration on caller-supplied logits (B, S, V) fp32 — sample_from_logits +
 * mask_by_random_topk + the where()s around them (transformer.py:849-932).  State is explicit:
 * zflat (B, S) int32 in "t c" order (util.py:39) is updated in place; tokens_out (B, S) int32
 * receives sampled_z; conf_out (B, S) fp32 receives the confidences (debug).
 * n0: device pointer to the whole-batch initial mask count (transformer.py:766). */
int32_t vnb_sample_step(const float* logits, int32_t* zflat, int32_t* tokens_out, float* conf_out,
                        const int32_t* n0, int32_t B, int32_t S, int32_t V, int32_t mask_token, int32_t step,
                        int32_t is_last, int32_t do_sample, float temperature, float gamma, float temp_eff,
                        uint32_t seed_lo, uint32_t seed_hi, void* stream);

/* ---- measurement hooks (bench.py) --------------------------------------------------------------
 * vnb_launch_count: kernels launched by this library so far in this process (a graph replay adds the
 * number of kernel nodes it contains).
 * vnb_graph_capture_count: generate graphs captured so far (a weight hot swap or a repeated call must not
 * add to it: graphs are cached per (workspace, steps, mask, top_p)).
 * vnb_profile_begin/end: between the two calls every launch of forward/generate is bracketed by CUDA
 * events on the launching stream (graph replay is bypassed so that the events can be recorded);
 * end() returns the summed device time and launch count per kernel family:
 *   0 embed, 1 rmsnorm, 2 gemm_qkv, 3 attention, 4 gemm_attn_out, 5 gemm_ffn_up, 6 gemm_ffn_down,
 *   7 gemm_classifier, 8 sample+remask, 9 state init/finish. */
#define VNB_NUM_FAMILIES 10
uint64_t vnb_launch_count(void);
uint64_t vnb_graph_capture_count(void);

/* ---- tuning options ----------------------------------------------------------------------------
 * "gemm_pair": 1 = dense contractions run as CTA pairs (tcgen05.mma.cta_group::2, 256 x 256 tiles, each SM stages
 *              half of the weight tile), 0 = one CTA per 128 x 256 tile.  Results are bit-identical (same
 *              accumulation order per output element).  Initial value: environment VNB_GEMM_PAIR, else the
 *              compiled default.  Generate graphs are cached per value.
 * "fused_sampler": 1 (default) = vnb_generate samples inside the classifier GEMM's epilogue (VNB_EPI_SAMPLE: the logits of
 *   the generate loop never reach HBM), 0 = from a materialised fp32 logits tensor (sample_rows_kernel).  Both draw
 *   with the same two-level inverse CDF and the same Philox stream; nucleus (top-p) sampling always materialises.
 * "gemm_pair_max_clusters" (get only): CTA pairs that can be co-resident on the current device. */
int32_t vnb_set_option(const char* name, int32_t value);
int32_t vnb_get_option(const char* name, int32_t* value);
int32_t vnb_profile_begin(vnb_model* m);
int32_t vnb_profile_end(vnb_model* m, float* ms_per_family, int32_t* launches_per_family, int32_t n_families);

/* ---- unit-level entry points (parity tests bisect with these) ------------------------------- */
enum {
  VNB_EPI_BF16 = 0,     /* out bf16 (M, N) */
  VNB_EPI_QKV = 1,      /* cols < 2d -> qk bf16 (M, 2d); cols >= 2d -> vT bf16 (B, d, Tpad) */
  VNB_EPI_RESID = 2,    /* out fp32 (M, N) += acc */
  VNB_EPI_GEGLU = 3,    /* out bf16 (M, N/2) = value * gelu_tanh(gate) */
  VNB_EPI_BIAS_F32 = 4, /* out fp32 (M, N) = acc + bias[n] */
  VNB_EPI_SAMPLE = 5    /* internal to vnb_generate: acc + bias[n] sampled per 128-column strip, nothing stored but
                           16 bytes per (row, strip); not accepted by vnb_op_gemm */
};
/* out = A (M,K) bf16 row-major  x  W (N,K)^T bf16 row-major, fp32 accumulate in TMEM.
 * N % 256 == 0, K % 64 == 0.  For VNB_EPI_QKV: out = qk, out2 = vT, T/Tpad describe the batch split. */
int32_t vnb_op_gemm(int32_t epi, const void* A, const void* W, int32_t M, int32_t N, int32_t K, void* out,
                    void* out2, const float* bias, int32_t T, int32_t Tpad, void* stream);
/* Fused self-attention with relative-position bias (transformer.py:234-254).
 * qk (B, T, 2d) bf16 [q | k], vT (B, d, Tpad) bf16, out (B, T, d) bf16, d = H*64. */
int32_t vnb_op_attention(const void* qk, const void* vT, void* out, const float* rel_bias, int32_t rel_sat,
                         int32_t B, int32_t T, int32_t Tpad, int32_t H, void* stream);
/* Naive SIMT GEMM used only to bisect the tcgen05 path in tests: out fp32 (M, N) = A x W^T. */
int32_t vnb_dbg_gemm_ref(const void* A, const void* W, int32_t M, int32_t N, int32_t K, float* out, void* stream);

/* ---- codec (DAC family; reference call sites: interface.py:223 codec.encode, transformer.py:671-675
 *      codec.quantizer.from_latents + codec.decode).  fp32, (B, C, T) channels-first. ------------------------
 * Generic 1-D convolution with the Snake activation fused on the input:
 *   y[b, co, q*out_stride + out_off] = bias[co] + sum_ci sum_j w[co, ci, j] * act(x[b, ci, q*stride + j*dil - pad])
 *                                      (+ residual) (tanh)
 * q in [0, nq).  snake_alpha (Cin) or NULL; residual (same shape as y) or NULL.  A ConvTranspose1d with stride s
 * is s launches of the K=2, dil=-1 form with out_stride = s (one per output phase). */
int32_t vnb_codec_conv1d(const float* x, const float* w, const float* bias, const float* snake_alpha,
                         const float* residual, float* y, int32_t B, int32_t Cin, int32_t Tin, int32_t Cout,
                         int32_t Tout, int32_t K, int32_t stride, int32_t dil, int32_t pad, int32_t out_stride,
                         int32_t out_off, int32_t nq, int32_t do_tanh, void* stream);
/* Residual vector quantiser.  mode 0 encode: in_f = z (B,D,T) -> codes (B,L,T) int64, zq (B,D,T), latents (B,8L,T).
 * mode 1 from_latents: in_f = latents (B,8L,T) -> zq.  mode 2 from_codes: in_codes (B,L,T) -> zq.
 * win (L,8,D), bin (L,8), wout (L,D,8), bout (L,D), cb (L,V,8) raw and cbn (L,V,8) L2-normalised codebooks. */
int32_t vnb_codec_rvq(int32_t mode, const float* in_f, const int64_t* in_codes, const float* win, const float* bin,
                      const float* wout, const float* bout, const float* cb, const float* cbn, int64_t* codes, float* zq,
                      float* latents, int32_t B, int32_t D, int32_t T, int32_t L, int32_t V, int32_t channels_last,
                      void* zq_hi, void* zq_lo, void* stream);
/* Tensor-core codec path (tcgen05, split-bf16 operands = fp32-grade products; see csrc/conv_tcgen05.cu).
 * Activations are channels-last (B, T, C) and travel as hi/lo bf16 pairs (x = hi + lo).
 *   y[b, q, n] = bias[n % bias_mod] + sum_tap sum_ci W[n, tap, ci] * a[b, q*s + tap*dil - pad, ci]   (+ resid)
 * w_hi/w_lo: (N, taps * ceil(Cin/64) * 64) bf16, tap-major, channel blocks zero-padded to 64.
 * Outputs (each optional): out_f32 = y, out_hi/out_lo = split(snake_alpha(y)) (alpha NULL: identity).
 * Output element (q, n) of batch b lands at b*out_batch_stride + q*N + n + out_offset, and is dropped unless that
 * flat index (without the batch term) is in [0, out_limit): this is how a transposed convolution with N = s*Cout
 * phase-major columns writes its (T*s, Cout) result. */
int32_t vnb_codec_conv_tc(const void* a_hi, const void* a_lo, int32_t B, int32_t Tin, int32_t Cin, int32_t s,
                          const void* w_hi, const void* w_lo, int32_t N, int32_t taps, int32_t dil, int32_t pad,
                          int32_t Tq, const float* bias, int32_t bias_mod, const float* alpha, int32_t alpha_mod,
                          const float* resid, float* out_f32, void* out_hi, void* out_lo, int64_t out_batch_stride,
                          int64_t out_offset, int64_t out_limit, int32_t do_tanh, void* stream);
/* encoder.conv1 (Cin = 1): x (B,1,T) -> fp32 (B,T,C) + split snake_alpha(y);  decoder.conv2 (Cout = 1) + tanh. */
int32_t vnb_codec_conv_in(const float* x, const float* w, const float* bias, const float* alpha, float* out_f32,
                          void* out_hi, void* out_lo, int32_t B, int32_t T, int32_t C, int32_t K, int32_t pad,
                          void* stream);
int32_t vnb_codec_conv_out(const void* a_hi, const void* a_lo, const float* w, const float* bias, float* audio, int32_t B,
                           int32_t T, int32_t C, int32_t K, int32_t pad, void* stream);
/* internal helper exported for the other translation units */
int32_t vnb_set_error_cuda(const char* what, int32_t cuda_error);

#ifdef __cplusplus
}
#endif
#endif /* VAMPNET_B200_H */
