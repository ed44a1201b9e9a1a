// vampnet_b200 — C ABI (include/vampnet_b200.h): model handle, per-(B,T) workspaces with their TMA
// tensor maps, the forward pass, and the graph-captured generate() loop.
#include <cuda_bf16.h>
#include <cudaTypedefs.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "kernels.h"

namespace vnb {

static thread_local std::string g_err;
static thread_local std::string g_tmap_err;

static int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// ---------------------------------------------------------------------------------- tensor maps
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    g_tmap_err = std::string("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: ") + cudaGetErrorString(e);
    return nullptr;
  }
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  return fn;
}
const char* tmap_error() { return g_tmap_err.c_str(); }

bool make_tmap_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                  uint32_t box_cols) {
  auto enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[256];
    snprintf(b, sizeof(b), "cuTensorMapEncodeTiled(2d rows=%llu cols=%llu box=%ux%u) -> %d", (unsigned long long)rows,
             (unsigned long long)cols, box_rows, box_cols, (int)r);
    g_tmap_err = b;
    return false;
  }
  return true;
}
bool make_tmap_3d(CUtensorMap* tm, const void* base, uint64_t batch, uint64_t rows, uint64_t cols,
                  uint64_t pitch_elems, uint32_t box_rows, uint32_t box_cols) {
  auto enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {pitch_elems * 2, rows * pitch_elems * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[256];
    snprintf(b, sizeof(b), "cuTensorMapEncodeTiled(3d batch=%llu rows=%llu cols=%llu pitch=%llu box=%ux%u) -> %d",
             (unsigned long long)batch, (unsigned long long)rows, (unsigned long long)cols,
             (unsigned long long)pitch_elems, box_rows, box_cols, (int)r);
    g_tmap_err = b;
    return false;
  }
  return true;
}

bool make_gemm_plan(GemmPlan* p, int epi, const void* A, const void* W, int M, int N, int K, void* out, void* out2,
                    const float* bias, int T, int Tpad, int d2) {
  if (N % 256 != 0 || K % 64 != 0 || M < 1) {
    g_tmap_err = "gemm: need N % 256 == 0 and K % 64 == 0";
    return false;
  }
  p->M = M; p->N = N; p->K = K; p->epi = epi; p->out = out; p->out2 = out2; p->bias = bias;
  p->T = T; p->Tpad = Tpad; p->d2 = d2;
  return make_tmap_2d(&p->tmA, A, M, K, 128, 64) && make_tmap_2d(&p->tmB, W, N, K, 256, 64) &&
         make_tmap_2d(&p->tmBh, W, N, K, 128, 64);
}

// RESID plans that also produce the next GEMM's operand: bf16 copy of the updated rows + row sums of squares.
bool gemm_plan_set_fused_out(GemmPlan* p, void* out_bf16, float* ss_out) {
  p->out_bf16 = out_bf16;
  p->ss_out = ss_out;
  return true;
}

bool make_attn_plan(AttnPlan* p, const void* qk, const void* vT, void* out, const float* rel, int sat, int B, int T,
                    int Tpad, int H) {
  const int d = H * 64;
  p->out = out; p->rel = rel; p->sat = sat; p->B = B; p->T = T; p->Tpad = Tpad; p->H = H;
  return make_tmap_3d(&p->tmQ, qk, B, T, 2 * d, 2 * d, 128, 64) && make_tmap_3d(&p->tmK, qk, B, T, 2 * d, 2 * d, 64, 64) &&
         make_tmap_3d(&p->tmVT, vT, B, d, Tpad, Tpad, 64, 64);
}

// ---------------------------------------------------------------------------------- model
struct DevBuf {
  void* p = nullptr;
  size_t n = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes, bool zero = false) {
    n = bytes;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e == cudaSuccess && zero) e = cudaMemset(p, 0, bytes);
    return e;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct GraphKey {  // graphs bake pointers, so generate() stages z/mask/out in workspace-owned buffers
  int steps;
  bool has_mask;
  bool top_p;  // selects the sampler kernel variant
  int variant;  // GEMM kernel variant baked into the graph (vnb_set_option "gemm_pair")
  bool fused;   // sampler fused into the classifier epilogue (vnb_set_option "fused_sampler")
  bool operator<(const GraphKey& o) const {
    return std::tie(steps, has_mask, top_p, variant, fused) < std::tie(o.steps, o.has_mask, o.top_p, o.variant, o.fused);
  }
};

struct Workspace {
  int B = 0, T = 0, Tpad = 0, M = 0;
  unsigned long long last_use = 0;
  DevBuf x, y, qk, vT, att, h, logits, zcur, zorig, tokens, conf, n0, dyn, z_in, mask_in, z_out, ssA, ssB, embA, partials;
  int embKp = 0;
  int ss_parts = 0;
  std::vector<GemmPlan> qkv, wo, up, down;
  GemmPlan cls, emb;
  GemmPlan cls_sample;  // the classifier with the sampling epilogue (generate loop)
  bool can_fuse = false;
  AttnPlan attn;
  std::map<GraphKey, cudaGraphExec_t> graphs;
  std::map<GraphKey, unsigned long long> graph_kernels;
  ~Workspace() { for (auto& kv : graphs) cudaGraphExecDestroy(kv.second); }
};

}  // namespace vnb

namespace vnb {
enum { FAM_EMBED = 0, FAM_RMSNORM, FAM_GEMM_QKV, FAM_ATTN, FAM_GEMM_O, FAM_GEMM_UP, FAM_GEMM_DOWN, FAM_GEMM_CLS,
       FAM_SAMPLE, FAM_STATE, FAM_COUNT };
static unsigned long long g_captures = 0;  // generate graphs captured + instantiated so far
static unsigned long long g_launches = 0;  // kernels launched by this library (graph replays add their node count)
void count_launch(unsigned long long n) { g_launches += n; }
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  std::vector<int> fam;  // family of the launch that FOLLOWS event i
  size_t used = 0;
  ~Profiler() { for (auto e : pool) cudaEventDestroy(e); }
  void mark(int family, cudaStream_t st) {
    if (!on) return;
    if (used == pool.size()) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) { on = false; return; }
      pool.push_back(e);
      fam.push_back(-1);
    }
    fam[used] = family;
    cudaEventRecord(pool[used++], st);
  }
};
}  // namespace vnb

struct vnb_model {
  vnb_config cfg;
  vnb_weights w;
  vnb::Profiler prof;
  std::map<std::pair<int, int>, std::unique_ptr<vnb::Workspace>> ws;
  vnb::Workspace* last = nullptr;
  static constexpr int kMaxSteps = 256;
  static constexpr size_t kMaxWorkspaces = 6;
  unsigned long long use_clock = 0;
};

namespace vnb {

static int get_workspace(vnb_model* m, int B, int T, Workspace** out) {
  auto key = std::make_pair(B, T);
  auto it = m->ws.find(key);
  if (it != m->ws.end()) { it->second->last_use = ++m->use_clock; *out = it->second.get(); return 0; }
  // bound the number of live (B, T) workspaces (each holds activations, logits and captured graphs): evict the
  // least recently used one.  cudaFree synchronises the device, so nothing in flight can still touch it.
  while (m->ws.size() >= vnb_model::kMaxWorkspaces) {
    auto lru = m->ws.begin();
    for (auto i = m->ws.begin(); i != m->ws.end(); ++i)
      if (i->second->last_use < lru->second->last_use) lru = i;
    if (m->last == lru->second.get()) m->last = nullptr;
    cudaDeviceSynchronize();
    m->ws.erase(lru);
  }
  const vnb_config& c = m->cfg;
  const int d = c.d_model, L = c.n_layers, Cp = c.n_codebooks - c.n_conditioning_codebooks;
  auto ws = std::make_unique<Workspace>();
  ws->B = B; ws->T = T; ws->M = B * T; ws->Tpad = (T + 7) / 8 * 8;
  const size_t M = ws->M;
  CK(ws->x.alloc(M * d * 4));
  CK(ws->y.alloc(M * d * 2));  // bf16 copy of the residual stream (A operand of QKV / FFN-up / classifier)
  ws->ss_parts = 2 * (d / 256);  // two partial row-sums-of-squares (even / odd chunks) per 256-column tile of the producing GEMM
  CK(ws->ssA.alloc(M * ws->ss_parts * 4, true));
  CK(ws->ssB.alloc(M * ws->ss_parts * 4, true));
  ws->embKp = (c.n_codebooks * 8 + 63) / 64 * 64;                 // gathered latents [hi | hi | lo], each third padded to Kp
  CK(ws->embA.alloc(M * 3 * ws->embKp * 2));
  CK(ws->qk.alloc(M * 2 * d * 2));
  CK(ws->vT.alloc(static_cast<size_t>(B) * d * ws->Tpad * 2, /*zero=*/true));  // padding keys stay 0 forever
  CK(ws->att.alloc(M * d * 2));
  CK(ws->h.alloc(M * 2 * d * 2));
  // ws->logits (M x Cp*V fp32, 0.4-1 GB at the bench shapes) is only needed when generate() samples from a materialised
  // tensor (top-p, fused_sampler = 0): allocated on first use in vnb_generate
  CK(ws->zcur.alloc(M * c.n_codebooks * 4));
  CK(ws->zorig.alloc(M * c.n_codebooks * 4));
  CK(ws->tokens.alloc(M * Cp * 4));
  CK(ws->conf.alloc(M * Cp * 4));
  CK(ws->n0.alloc(4, true));
  CK(ws->dyn.alloc(sizeof(SampleDyn) * vnb_model::kMaxSteps));
  CK(ws->z_in.alloc(M * c.n_codebooks * 8));
  CK(ws->mask_in.alloc(M * c.n_codebooks * 4));
  CK(ws->z_out.alloc(M * c.n_codebooks * 8));
  const __nv_bfloat16* wqkv = reinterpret_cast<const __nv_bfloat16*>(m->w.wqkv);
  const __nv_bfloat16* wo = reinterpret_cast<const __nv_bfloat16*>(m->w.wo);
  const __nv_bfloat16* w1 = reinterpret_cast<const __nv_bfloat16*>(m->w.w1);
  const __nv_bfloat16* w2 = reinterpret_cast<const __nv_bfloat16*>(m->w.w2);
  ws->qkv.resize(L); ws->wo.resize(L); ws->up.resize(L); ws->down.resize(L);
  const size_t dd = static_cast<size_t>(d) * d;
  const float inv_d = 1.0f / static_cast<float>(d), eps = 1e-6f;
  auto consumer = [&](GemmPlan& p, const DevBuf& ss) { p.ss_in = ss.as<float>(); p.ss_parts = ws->ss_parts; p.inv_d = inv_d; p.eps = eps; };
  auto producer = [&](GemmPlan& p, const DevBuf& ss) { return gemm_plan_set_fused_out(&p, ws->y.p, ss.as<float>()); };
  for (int l = 0; l < L; ++l) {
    // residual stream x (fp32) + its bf16 copy y + row sums of squares: ssA feeds QKV, ssB feeds FFN-up
    bool ok = make_gemm_plan(&ws->qkv[l], VNB_EPI_QKV, ws->y.p, wqkv + l * 3 * dd, ws->M, 3 * d, d, ws->qk.p, ws->vT.p,
                             nullptr, T, ws->Tpad, 2 * d) &&
              make_gemm_plan(&ws->wo[l], VNB_EPI_RESID, ws->att.p, wo + l * dd, ws->M, d, d, ws->x.p, nullptr, nullptr, T,
                             ws->Tpad, 0) &&
              make_gemm_plan(&ws->up[l], VNB_EPI_GEGLU, ws->y.p, w1 + l * 4 * dd, ws->M, 4 * d, d, ws->h.p, nullptr,
                             nullptr, T, ws->Tpad, 0) &&
              make_gemm_plan(&ws->down[l], VNB_EPI_RESID, ws->h.p, w2 + l * 2 * dd, ws->M, d, 2 * d, ws->x.p, nullptr,
                             nullptr, T, ws->Tpad, 0);
    if (!ok) return fail("plan layer %d: %s", l, tmap_error());
    consumer(ws->qkv[l], ws->ssA);
    consumer(ws->up[l], ws->ssB);
    if (!producer(ws->wo[l], ws->ssB) || !producer(ws->down[l], ws->ssA)) return fail("plan layer %d: %s", l, tmap_error());
  }
  if (!make_gemm_plan(&ws->cls, VNB_EPI_BIAS_F32, ws->y.p, m->w.wcls, ws->M, Cp * c.vocab_size, d, nullptr, nullptr,
                      m->w.bcls, T, ws->Tpad, 0))
    return fail("plan classifier: %s", tmap_error());
  consumer(ws->cls, ws->ssA);
  // generate loop: the same GEMM with the sampling epilogue; the logits are consumed in the epilogue and never stored
  ws->can_fuse = c.vocab_size % 128 == 0 && c.vocab_size <= 1024;
  if (ws->can_fuse) {
    CK(ws->partials.alloc(M * static_cast<size_t>(Cp) * (c.vocab_size / 128) * 16));
    ws->cls_sample = ws->cls;
    ws->cls_sample.epi = VNB_EPI_SAMPLE;
    ws->cls_sample.out = nullptr;
    ws->cls_sample.zcur = ws->zcur.as<int32_t>();
    ws->cls_sample.partials = ws->partials.p;
    ws->cls_sample.C = c.n_codebooks; ws->cls_sample.ncc = c.n_conditioning_codebooks;
    ws->cls_sample.V = c.vocab_size; ws->cls_sample.mask_token = c.vocab_size;
  }
  // embedding out_proj (layers.py:162) as a split-bf16 tensor-core contraction: x = A . emb_w3^T + bias, which also
  // emits bf16(x) and the row sums of squares the first QKV projection's fused RMSNorm consumes
  if (!make_gemm_plan(&ws->emb, VNB_EPI_BIAS_F32, ws->embA.p, m->w.emb_w3, ws->M, d, 3 * ws->embKp, ws->x.p, nullptr,
                      m->w.emb_b, T, ws->Tpad, 0) ||
      !gemm_plan_set_fused_out(&ws->emb, ws->y.p, ws->ssA.as<float>()))
    return fail("plan embedding: %s", tmap_error());
  if (!make_attn_plan(&ws->attn, ws->qk.p, ws->vT.p, ws->att.p, m->w.rel_bias, m->w.rel_sat, B, T, ws->Tpad, c.n_heads))
    return fail("plan attention: %s", tmap_error());
  ws->last_use = ++m->use_clock;
  *out = ws.get();
  m->ws[key] = std::move(ws);
  return 0;
}

#define LAUNCH(fam_, expr)        \
  do {                            \
    m->prof.mark((fam_), st);     \
    CK(expr);                     \
    ++g_launches;                 \
  } while (0)

// CodebookEmbedding (layers.py:134-162): gather (+ split) the latents, then the out_proj contraction on the tensor cores.
static int run_embed(vnb_model* m, Workspace* ws, const int32_t* codes_btc, const float* latents, cudaStream_t st) {
  const vnb_config& c = m->cfg;
  LAUNCH(FAM_EMBED, launch_embed_gather(codes_btc, latents, m->w.emb_table, ws->embA.p, ws->M, ws->T, c.n_codebooks,
                                        c.vocab_size + 1, c.n_codebooks * 8, ws->embKp, ws->ssA.as<float>(),
                                        c.d_model / 256, ws->ss_parts, st));
  LAUNCH(FAM_EMBED, launch_gemm(ws->emb, st));
  return 0;
}

// x already holds the embedded input; runs the L layers + final norm + classifier into `logits`.
static int run_stack(vnb_model* m, Workspace* ws, float* logits, cudaStream_t st, float* acts = nullptr,
                     const SampleDyn* fused_dyn = nullptr) {
  const vnb_config& c = m->cfg;
  // RMSNorm (transformer.py:43-58) is fused: norm weights are folded into wqkv / w1 / wcls at pack time, the
  // producers of x (embed, attn-out, ffn-down) also emit bf16(x) and per-row sums of squares, and the consumers
  // scale their accumulator rows by rsqrt(mean(x^2) + eps).
  for (int l = 0; l < c.n_layers; ++l) {
    LAUNCH(FAM_GEMM_QKV, launch_gemm(ws->qkv[l], st));
    LAUNCH(FAM_ATTN, launch_attention(ws->attn, st));
    LAUNCH(FAM_GEMM_O, launch_gemm(ws->wo[l], st));
    LAUNCH(FAM_GEMM_UP, launch_gemm(ws->up[l], st));
    LAUNCH(FAM_GEMM_DOWN, launch_gemm(ws->down[l], st));
    if (acts)  // return_activations: the residual stream after this layer (transformer.py:455-456)
      CK(cudaMemcpyAsync(acts + static_cast<size_t>(l) * ws->M * c.d_model, ws->x.p, ws->x.n, cudaMemcpyDeviceToDevice, st));
  }
  if (fused_dyn != nullptr) {  // generate loop: sample in the classifier's epilogue, no logits tensor
    GemmPlan cls = ws->cls_sample;
    cls.dyn = fused_dyn;
    LAUNCH(FAM_GEMM_CLS, launch_gemm(cls, st));
  } else {
    GemmPlan cls = ws->cls;
    cls.out = logits;
    LAUNCH(FAM_GEMM_CLS, launch_gemm(cls, st));
  }
  m->prof.mark(-1, st);
  return 0;
}

}  // namespace vnb

using namespace vnb;

extern "C" {

int32_t vnb_abi_version(void) { return VNB_ABI_VERSION; }
int32_t vnb_set_error_cuda(const char* what, int32_t cuda_error) {
  return fail("%s failed: %s", what, cudaGetErrorString(static_cast<cudaError_t>(cuda_error)));
}
const char* vnb_last_error(void) { return g_err.c_str(); }

int32_t vnb_model_create(const vnb_config* cfg, const vnb_weights* w, vnb_model** out) {
  if (!cfg || !w || !out) return fail("null argument");
  if (cfg->d_model != cfg->n_heads * 64) return fail("d_model must be n_heads*64 (got %d, %d)", cfg->d_model, cfg->n_heads);
  if (cfg->d_model % 256 != 0) return fail("d_model must be a multiple of 256");
  if (cfg->latent_dim != 8) return fail("latent_dim must be 8");
  if (cfg->vocab_size % 256 != 0 || cfg->vocab_size > 1024) return fail("vocab_size must be a multiple of 256, <= 1024");
  if (cfg->n_codebooks * 8 > 128) return fail("n_codebooks too large");
  if (w->rel_sat < 1 || w->rel_sat > 128) return fail("rel_sat out of range");
  int dev = 0, major = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail("vampnet_b200 needs an sm_100 device (got compute capability major %d)", major);
  CK(prepare_gemm());
  auto* m = new vnb_model();
  m->cfg = *cfg;
  m->w = *w;
  *out = m;
  return 0;
}

void vnb_model_destroy(vnb_model* m) { delete m; }

int32_t vnb_forward_codes(vnb_model* m, const int64_t* codes, int32_t B, int32_t T, float* logits, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Workspace* ws;
  if (get_workspace(m, B, T, &ws)) return 1;
  const vnb_config& c = m->cfg;
  // (B,C,T) int64 -> (B,T,C) int32, no masking (mask = zeros)
  LAUNCH(FAM_STATE, launch_gen_init(codes, nullptr, ws->zcur.as<int32_t>(), ws->zorig.as<int32_t>(), ws->n0.as<int32_t>(), B,
                     c.n_codebooks, T, /*ncc=*/c.n_codebooks, c.vocab_size, st));
  if (run_embed(m, ws, ws->zcur.as<int32_t>(), nullptr, st)) return 1;
  m->last = ws;
  return run_stack(m, ws, logits, st);
}

int32_t vnb_forward_latents(vnb_model* m, const float* latents, int32_t B, int32_t T, float* logits, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Workspace* ws;
  if (get_workspace(m, B, T, &ws)) return 1;
  if (run_embed(m, ws, nullptr, latents, st)) return 1;
  m->last = ws;
  return run_stack(m, ws, logits, st);
}

int32_t vnb_forward_latents_acts(vnb_model* m, const float* latents, int32_t B, int32_t T, float* logits, float* acts,
                                 void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Workspace* ws;
  if (get_workspace(m, B, T, &ws)) return 1;
  if (run_embed(m, ws, nullptr, latents, st)) return 1;
  m->last = ws;
  return run_stack(m, ws, logits, st, acts);
}

int32_t vnb_get_hidden(vnb_model* m, float* out, void* stream) {
  if (!m->last) return fail("no forward has run");
  CK(cudaMemcpyAsync(out, m->last->x.p, m->last->x.n, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

// vnb_set_option("fused_sampler", 0|1): sample inside the classifier GEMM's epilogue (default) or from a materialised
// logits tensor.  Nucleus (top-p) filtering needs whole sorted rows and always takes the materialising path.
static int g_fused_sampler = -1;  // -1: not read yet (environment VNB_FUSED_SAMPLER, else on)
static int fused_sampler_enabled() {
  if (g_fused_sampler < 0) {
    const char* e = getenv("VNB_FUSED_SAMPLER");
    g_fused_sampler = e != nullptr ? (e[0] != '0') : 1;
  }
  return g_fused_sampler;
}

static int enqueue_generate(vnb_model* m, Workspace* ws, const int64_t* z, const int32_t* mask, int steps, int64_t* out,
                            cudaStream_t st, bool use_top_p, bool fused) {
  const vnb_config& c = m->cfg;
  const int ncc = c.n_conditioning_codebooks;
  LAUNCH(FAM_STATE, launch_gen_init(z, mask, ws->zcur.as<int32_t>(), ws->zorig.as<int32_t>(), ws->n0.as<int32_t>(), ws->B, c.n_codebooks,
                     ws->T, ncc, c.vocab_size, st));
  SampleArgs sa;
  sa.logits = ws->logits.as<float>();
  sa.zcur = ws->zcur.as<int32_t>();
  sa.zorig = ws->zorig.as<int32_t>();
  sa.tokens = ws->tokens.as<int32_t>();
  sa.conf = ws->conf.as<float>();
  sa.n0 = ws->n0.as<int32_t>();
  sa.B = ws->B; sa.T = ws->T; sa.C = c.n_codebooks; sa.ncc = ncc; sa.V = c.vocab_size; sa.mask_token = c.vocab_size;
  for (int i = 0; i < steps; ++i) {
    if (run_embed(m, ws, ws->zcur.as<int32_t>(), nullptr, st)) return 1;
    const SampleDyn* dyn_i = ws->dyn.as<SampleDyn>() + i;
    if (fused) {
      if (run_stack(m, ws, nullptr, st, nullptr, dyn_i)) return 1;
      LAUNCH(FAM_SAMPLE, launch_sample_combine_dev(sa, ws->partials.p, dyn_i, st));
    } else {
      if (run_stack(m, ws, ws->logits.as<float>(), st)) return 1;
      LAUNCH(FAM_SAMPLE, launch_sample_step_dev(sa, dyn_i, st, use_top_p));
    }
    ++g_launches;  // sample step = two kernels
  }
  LAUNCH(FAM_STATE, launch_gen_finish(ws->tokens.as<int32_t>(), ws->zorig.as<int32_t>(), out, ws->B, c.n_codebooks, ws->T, ncc, st));
  m->prof.mark(-1, st);
  return 0;
}

int32_t vnb_generate(vnb_model* m, const int64_t* z, const int32_t* mask, int32_t B, int32_t T,
                     const vnb_gen_params* p, int64_t* out, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!p || p->sampling_steps < 1 || p->sampling_steps > vnb_model::kMaxSteps) return fail("bad sampling_steps");
  Workspace* ws;
  if (get_workspace(m, B, T, &ws)) return 1;
  m->last = ws;
  const int steps = p->sampling_steps;
  std::vector<SampleDyn> dyn(steps);
  const float inv_t = p->temperature > 0.f ? static_cast<float>(1.0 / static_cast<double>(p->temperature)) : 1.0f;
  for (int i = 0; i < steps; ++i) {
    dyn[i].inv_temp = inv_t;
    dyn[i].gamma = p->gamma[i];
    dyn[i].temp_eff = p->temp_eff[i];
    dyn[i].do_sample = p->do_sample[i];
    dyn[i].is_last = (i == steps - 1);
    dyn[i].step = i;
    dyn[i].seed_lo = p->seed_lo;
    dyn[i].seed_hi = p->seed_hi;
    dyn[i].top_p = p->top_p;
  }
  // pageable source: the runtime stages it before returning, so `dyn` may die at scope exit
  CK(cudaMemcpyAsync(ws->dyn.p, dyn.data(), sizeof(SampleDyn) * steps, cudaMemcpyHostToDevice, st));
  const bool use_top_p = p->top_p > 0.f && p->top_p < 1.f;
  const bool fused = fused_sampler_enabled() != 0 && !use_top_p && ws->can_fuse;
  if (!fused && ws->logits.p == nullptr)  // before any capture: cudaMalloc is not capturable
    CK(ws->logits.alloc(static_cast<size_t>(ws->M) * (m->cfg.n_codebooks - m->cfg.n_conditioning_codebooks) * m->cfg.vocab_size * 4));
  if (!p->use_graph || m->prof.on) return enqueue_generate(m, ws, z, mask, steps, out, st, use_top_p, fused);

  const size_t nz = static_cast<size_t>(B) * m->cfg.n_codebooks * T;
  CK(cudaMemcpyAsync(ws->z_in.p, z, nz * 8, cudaMemcpyDeviceToDevice, st));
  if (mask) CK(cudaMemcpyAsync(ws->mask_in.p, mask, nz * 4, cudaMemcpyDeviceToDevice, st));
  const int64_t* gz = ws->z_in.as<int64_t>();
  const int32_t* gmask = mask ? ws->mask_in.as<int32_t>() : nullptr;
  int64_t* gout = ws->z_out.as<int64_t>();
  GraphKey key{steps, mask != nullptr, use_top_p, get_gemm_pair(), fused};
  auto it = ws->graphs.find(key);
  if (it == ws->graphs.end()) {
    cudaStream_t cap;
    CK(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { cudaStreamDestroy(cap); return fail("begin capture: %s", cudaGetErrorString(e)); }
    const unsigned long long before = g_launches;
    int rc = enqueue_generate(m, ws, gz, gmask, steps, gout, cap, use_top_p, fused);
    const unsigned long long in_graph = g_launches - before;
    g_launches = before;
    e = cudaStreamEndCapture(cap, &graph);
    cudaStreamDestroy(cap);
    if (rc) { if (graph) cudaGraphDestroy(graph); return 1; }
    if (e != cudaSuccess) return fail("end capture: %s", cudaGetErrorString(e));
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail("graph instantiate: %s", cudaGetErrorString(e));
    if (ws->graphs.size() >= 16) {  // bound the cache
      for (auto& kv : ws->graphs) cudaGraphExecDestroy(kv.second);
      ws->graphs.clear();
      ws->graph_kernels.clear();
    }
    it = ws->graphs.emplace(key, exec).first;
    ++g_captures;
    ws->graph_kernels[key] = in_graph;
  }
  CK(cudaGraphLaunch(it->second, st));
  g_launches += ws->graph_kernels[key];
  CK(cudaMemcpyAsync(out, ws->z_out.p, nz * 8, cudaMemcpyDeviceToDevice, st));
  return 0;
}

uint64_t vnb_launch_count(void) { return g_launches; }
uint64_t vnb_graph_capture_count(void) { return g_captures; }

int32_t vnb_set_option(const char* name, int32_t value) {
  if (!name) return fail("null option name");
  if (strcmp(name, "gemm_pair") == 0) {
    set_gemm_pair(value);
    return 0;
  }
  if (strcmp(name, "fused_sampler") == 0) {
    g_fused_sampler = value ? 1 : 0;
    return 0;
  }
  return fail("unknown option '%s'", name);
}
int32_t vnb_get_option(const char* name, int32_t* value) {
  if (!name || !value) return fail("null argument");
  if (strcmp(name, "gemm_pair") == 0) {
    *value = get_gemm_pair();
    return 0;
  }
  if (strcmp(name, "fused_sampler") == 0) {
    *value = fused_sampler_enabled();
    return 0;
  }
  if (strcmp(name, "gemm_pair_max_clusters") == 0) {  // read-only: co-resident CTA pairs on the current device
    *value = get_gemm_max_clusters();
    return 0;
  }
  return fail("unknown option '%s'", name);
}

int32_t vnb_profile_begin(vnb_model* m) {
  m->prof.used = 0;
  m->prof.on = true;
  return 0;
}
int32_t vnb_profile_end(vnb_model* m, float* ms_per_family, int32_t* launches_per_family, int32_t n_families) {
  m->prof.on = false;
  for (int i = 0; i < n_families; ++i) { ms_per_family[i] = 0.f; launches_per_family[i] = 0; }
  if (m->prof.used < 2) return 0;
  CK(cudaEventSynchronize(m->prof.pool[m->prof.used - 1]));
  for (size_t i = 0; i + 1 < m->prof.used; ++i) {
    const int f = m->prof.fam[i];
    if (f < 0 || f >= n_families) continue;
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, m->prof.pool[i], m->prof.pool[i + 1]));
    ms_per_family[f] += ms;
    launches_per_family[f] += 1;
  }
  return 0;
}

int32_t vnb_sample_step(const float* logits, int32_t* zflat, int32_t* tokens_out, float* conf_out, const int32_t* n0,
                        int32_t B, int32_t S, int32_t V, int32_t mask_token, int32_t step, int32_t is_last,
                        int32_t do_sample, float temperature, float gamma, float temp_eff, uint32_t seed_lo,
                        uint32_t seed_hi, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // per-device ring of parameter slots (a process-global one would live on whichever device called first)
  static SampleDyn* scratch_dev[64] = {nullptr};
  static int slot_dev[64] = {0};
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail("device index %d out of range", dev);
  SampleDyn*& scratch = scratch_dev[dev];
  int& slot = slot_dev[dev];
  if (!scratch) CK(cudaMalloc(&scratch, sizeof(SampleDyn) * 64));
  SampleDyn d;
  d.inv_temp = temperature > 0.f ? static_cast<float>(1.0 / static_cast<double>(temperature)) : 1.0f;
  d.gamma = gamma; d.temp_eff = temp_eff; d.do_sample = do_sample; d.is_last = is_last; d.step = step;
  d.seed_lo = seed_lo; d.seed_hi = seed_hi; d.top_p = 0.f;
  SampleDyn* dd = scratch + (slot++ & 63);
  CK(cudaMemcpyAsync(dd, &d, sizeof(d), cudaMemcpyHostToDevice, st));
  SampleArgs sa;
  sa.logits = logits; sa.zcur = zflat; sa.zorig = nullptr; sa.tokens = tokens_out; sa.conf = conf_out; sa.n0 = n0;
  sa.B = B; sa.T = S; sa.C = 1; sa.ncc = 0; sa.V = V; sa.mask_token = mask_token;
  CK(launch_sample_step_dev(sa, dd, st, false));
  return 0;
}

// ------------------------------------------------------------------------------- unit-level ops
int32_t vnb_op_gemm(int32_t epi, const void* A, const void* W, int32_t M, int32_t N, int32_t K, void* out, void* out2,
                    const float* bias, int32_t T, int32_t Tpad, void* stream) {
  GemmPlan p;
  const int d2 = epi == VNB_EPI_QKV ? (N / 3) * 2 : 0;
  if (!make_gemm_plan(&p, epi, A, W, M, N, K, out, out2, bias, T, Tpad, d2)) return fail("gemm plan: %s", tmap_error());
  CK(launch_gemm(p, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}
int32_t vnb_op_attention(const void* qk, const void* vT, void* out, const float* rel_bias, int32_t rel_sat, int32_t B,
                         int32_t T, int32_t Tpad, int32_t H, void* stream) {
  AttnPlan p;
  if (!make_attn_plan(&p, qk, vT, out, rel_bias, rel_sat, B, T, Tpad, H)) return fail("attn plan: %s", tmap_error());
  CK(launch_attention(p, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}
int32_t vnb_dbg_gemm_ref(const void* A, const void* W, int32_t M, int32_t N, int32_t K, float* out, void* stream) {
  CK(launch_gemm_ref(A, W, M, N, K, out, reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}

}  // extern "C"
