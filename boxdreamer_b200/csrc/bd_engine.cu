// C ABI + native engine: owns the workspace and packed weights, enqueues the whole BoxDreamer
// inference path (DINOv2 ViT-B/14-reg -> BETR decoder -> heat maps -> top-20 corners -> PnP) on one stream.
// See include/boxdreamer_b200.h for the contract and the reference call sites each entry replaces.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/boxdreamer_b200.h"
#include "bd_internal.h"

using namespace bd;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CK(expr)                                                                                           \
  do {                                                                                                     \
    cudaError_t _e = (expr);                                                                               \
    if (_e != cudaSuccess) {                                                                               \
      std::string _m = std::string(#expr) + ": " + cudaGetErrorString(_e);                                 \
      const char* _t = tc_last_error();                                                                    \
      if (_t && _t[0]) _m += std::string(" [") + _t + "]";                                                 \
      return fail(BD_ERR_CUDA, _m);                                                                        \
    }                                                                                                      \
  } while (0)

struct Weight {
  float* f32 = nullptr;   // device, fp32
  bf16* b16 = nullptr;    // device, bf16 (tensor path, GEMM operands only)
  std::vector<int64_t> shape;
  size_t numel = 0;
};

struct bd_engine {
  bd_config cfg;
  int device = 0;
  bool tc = false;
  bool finalized = false;
  int S, patch, g, P, d, hd_dec, hd_dino, n_tok, n_prefix, seqpad_dino, kpe;  // kpe: patch-embed K (padded on the tensor path)
  int Lmax, Bmax, Tmax;
  std::map<std::string, Weight> w;
  std::vector<void*> allocs;
  // workspace (element type depends on precision: "act" = bf16 on the tensor path, fp32 on the exact path)
  void* A_pe = nullptr;      // act [L*P, kpe]
  float* X_dino = nullptr;   // f32 [L*n_tok, d]
  float* X_dec = nullptr;    // f32 [L*P, d]
  void* H = nullptr;         // act [Mmax, d]      LayerNorm output
  void* G = nullptr;         // act [Mmax, 4d]     MLP hidden
  void* Q = nullptr; void* K = nullptr; void* V = nullptr;  // act, attention operand layouts
  void* O = nullptr;         // act [Mmax, d]
  float* qkv_scratch = nullptr;  // exact path: f32 [Mmax, 3d]
  float* feats = nullptr;    // f32 [L*P, d]
  void* feats_act = nullptr; // act [L*P, d]
  float* R = nullptr;        // f32 [L*P, d]
  float* PF = nullptr;       // f32 [L*P, d]
  void* A_bb = nullptr;      // act [L*P, patch*patch*8]
  void* Xq = nullptr;        // act [B*P, d]
  float* logits = nullptr;   // f32 [B*P, patch*patch*8]
  float* heat = nullptr;     // f32 [B,8,S,S]
  float* corners_px = nullptr; float* corners_norm = nullptr; float* poses = nullptr;
  float* rec = nullptr;      // f32 [B, 28] packed result record {R|t, 8 normalised corners} (bd_forward_packed)
  float* bbox3d_q = nullptr; float* K_q = nullptr; int64_t* qidx = nullptr;
  // Input staging: two slots, so that bd_forward_host_submit can copy the next batch in while the previous one computes
  // (slot 0 also stages the graph-replayed shapes of bd_forward).  in_bbox_px: projected corners [Lmax,8,2] fp32 (_px entries).
  struct HostSlot {
    void* in_images = nullptr; void* in_bbox = nullptr; void* in_bbox_px = nullptr;
    cudaEvent_t copy_ev[8] = {nullptr};
    cudaEvent_t done = nullptr;
    bool pending = false;          // submitted, not yet waited for
  };
  HostSlot slot[2];
  // Cross-stream ordering of the ONE workspace: device-pointer entries run on the caller's stream, host-buffer entries on
  // host_stream / copy_stream.  Each side waits (on the device, cudaStreamWaitEvent) for the other's last recorded work.
  cudaEvent_t dev_ev = nullptr;      // end of the latest device-pointer entry, recorded on the caller's stream
  bool dev_ev_valid = false;
  float* pos_dec = nullptr;  // f32 [P, d] 2-D sincos table
  cudaStream_t host_stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  // CUDA graphs of the launch chain (graph_run): one executable per (stage, shape, options), captured on `cap_stream`
  struct GraphEntry { cudaGraphExec_t exec = nullptr; int calls = 0; bool unusable = false; long long launches = 0; };
  std::map<std::vector<long long>, GraphEntry> graphs;
  cudaStream_t cap_stream = nullptr;
  int graph_max_views = 24;   // bd_forward (device pointers): largest B*T that is staged and replayed; BOXDREAMER_B200_GRAPH_MAX_VIEWS
  bool graphs_on = true;      // BOXDREAMER_B200_GRAPHS=0 switches every replay off
  // instrumentation: kernel launch counter and optional per-category CUDA-event timing
  long long launches = 0;
  bool profile = false;
  struct Span { int cat; cudaEvent_t a, b; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> ev_pool;
  double cat_ms[BD_PROF_NCAT] = {0};
  long long cat_n[BD_PROF_NCAT] = {0};
};

// Entry points that take a handle run on the handle's device and restore the caller's current device afterwards.
struct DevGuard {
  int prev = -1, want;
  explicit DevGuard(const bd_engine* e) : want(e ? e->device : -1) {
    if (want >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != want) cudaSetDevice(want); else prev = -1;
  }
  ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static cudaEvent_t get_event(bd_engine* e) {
  if (!e->ev_pool.empty()) { cudaEvent_t ev = e->ev_pool.back(); e->ev_pool.pop_back(); return ev; }
  cudaEvent_t ev;
  cudaEventCreate(&ev);
  return ev;
}

static void drop_graphs(bd_engine* e) {
  for (auto& kv : e->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  e->graphs.clear();
}

// launches `expr` (a cudaError_t launcher that enqueues `nk` kernels) under category `cat`
#define LAUNCH(cat, nk, expr)                                             \
  do {                                                                    \
    bd_engine::Span _sp{(cat), nullptr, nullptr};                         \
    if (e->profile) { _sp.a = get_event(e); cudaEventRecord(_sp.a, s); }  \
    CK(expr);                                                             \
    e->launches += (nk) + take_extra_launches();                                                  \
    if (e->profile) { _sp.b = get_event(e); cudaEventRecord(_sp.b, s); e->spans.push_back(_sp); } \
  } while (0)

static size_t act_size(const bd_engine* e) { return e->tc ? 2 : 4; }

static int dalloc(bd_engine* e, void** out, size_t bytes) {
  void* p = nullptr;
  CK(cudaMalloc(&p, bytes ? bytes : 16));
  CK(cudaMemset(p, 0, bytes ? bytes : 16));
  e->allocs.push_back(p);
  *out = p;
  return BD_OK;
}
#define DALLOC(field, bytes)                                               \
  do {                                                                     \
    int _r = dalloc(e, reinterpret_cast<void**>(&(field)), (bytes));       \
    if (_r != BD_OK) return _r;                                            \
  } while (0)

extern "C" const char* bd_last_error(void) { return g_err.c_str(); }
extern "C" int bd_version(void) { return 100; }

// pos_encodiong.py:125-213 as consumed at betr.py:357-364 (see oracle sincos_pos_embed_2d)
static void build_sincos(std::vector<float>& tab, int d, int g) {
  const int half = d / 2, quarter = d / 4;
  tab.assign(static_cast<size_t>(g) * g * d, 0.f);
  for (int y = 0; y < g; ++y)
    for (int x = 0; x < g; ++x) {
      float* row = &tab[(static_cast<size_t>(y) * g + x) * d];
      for (int i = 0; i < quarter; ++i) {
        const double omega = 1.0 / pow(10000.0, static_cast<double>(i) / (half / 2.0));
        const double ax = static_cast<double>(static_cast<float>(x)) * omega;
        const double ay = static_cast<double>(static_cast<float>(y)) * omega;
        row[i] = static_cast<float>(sin(ax));
        row[quarter + i] = static_cast<float>(cos(ax));
        row[half + i] = static_cast<float>(sin(ay));
        row[half + quarter + i] = static_cast<float>(cos(ay));
      }
    }
}

extern "C" int bd_create(bd_handle* out, const bd_config* cfg) {
  if (!out || !cfg) return fail(BD_ERR_INVALID, "bd_create: null argument");
  if (cfg->patch_size <= 0 || cfg->img_size % cfg->patch_size != 0)
    return fail(BD_ERR_INVALID, "bd_create: img_size must be a multiple of patch_size");
  if (cfg->d_model != 768 || cfg->dec_heads != 8 || cfg->dino_heads != 12)
    return fail(BD_ERR_UNSUPPORTED, "bd_create: only d_model=768, decoder 8x96, DINOv2 12x64 are built");
  if (cfg->max_batch <= 0 || cfg->max_views <= 0) return fail(BD_ERR_INVALID, "bd_create: max_batch/max_views must be > 0");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(BD_ERR_CUDA, "bd_create: no CUDA device (this library has no CPU fallback)");
  bd_engine* e = new bd_engine();
  e->cfg = *cfg;
  CK(cudaGetDevice(&e->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, e->device));
  e->tc = cfg->precision == BD_PRECISION_BF16;
  if (const char* ev = getenv("BOXDREAMER_B200_GRAPHS")) e->graphs_on = atoi(ev) != 0;
  if (const char* ev = getenv("BOXDREAMER_B200_GRAPH_MAX_VIEWS")) e->graph_max_views = atoi(ev);
  if (e->tc && prop.major != 10) {
    delete e;
    return fail(BD_ERR_UNSUPPORTED, "bd_create: the bf16 tensor path needs an sm_100 (Blackwell) device");
  }
  e->S = cfg->img_size; e->patch = cfg->patch_size; e->g = e->S / e->patch; e->P = e->g * e->g; e->d = cfg->d_model;
  e->hd_dec = e->d / cfg->dec_heads; e->hd_dino = e->d / cfg->dino_heads;
  e->n_prefix = 1 + cfg->dino_registers; e->n_tok = e->n_prefix + e->P;
  e->seqpad_dino = (e->n_tok + 127) / 128 * 128;
  const int kreal = 3 * e->patch * e->patch;
  e->kpe = e->tc ? (kreal + 63) / 64 * 64 : kreal;
  e->Bmax = cfg->max_batch; e->Tmax = cfg->max_views; e->Lmax = e->Bmax * e->Tmax;
  const size_t as = act_size(e);
  const size_t L = e->Lmax, P = e->P, d = e->d;
  const size_t Md = L * e->n_tok, Mb = L * P, Mmax = Md > Mb ? Md : Mb;
  const size_t seqpad_dec = (static_cast<size_t>(e->Tmax) * P + 127) / 128 * 128;
  const size_t qk_dino = L * cfg->dino_heads * e->seqpad_dino * e->hd_dino;
  const size_t qk_dec = static_cast<size_t>(e->Bmax) * cfg->dec_heads * seqpad_dec * e->hd_dec;
  const size_t qk = qk_dino > qk_dec ? qk_dino : qk_dec;
  const size_t pp8 = static_cast<size_t>(e->patch) * e->patch * 8;
  DALLOC(e->A_pe, Mb * e->kpe * as);
  DALLOC(e->X_dino, Md * d * 4);
  DALLOC(e->X_dec, Mb * d * 4);
  DALLOC(e->H, Mmax * d * as);
  DALLOC(e->G, Mmax * 4 * d * as);
  DALLOC(e->Q, qk * as);
  DALLOC(e->K, qk * as);
  DALLOC(e->V, qk * as);
  DALLOC(e->O, Mmax * d * as);
  if (!e->tc) DALLOC(e->qkv_scratch, Mmax * 3 * d * 4);
  DALLOC(e->feats, Mb * d * 4);
  DALLOC(e->feats_act, Mb * d * as);
  DALLOC(e->R, Mb * d * 4);
  DALLOC(e->PF, Mb * d * 4);
  DALLOC(e->A_bb, Mb * pp8 * as);
  DALLOC(e->Xq, static_cast<size_t>(e->Bmax) * P * d * as);
  DALLOC(e->logits, static_cast<size_t>(e->Bmax) * P * pp8 * 4);
  DALLOC(e->heat, static_cast<size_t>(e->Bmax) * 8 * e->S * e->S * 4);
  DALLOC(e->corners_px, static_cast<size_t>(e->Bmax) * 16 * 4);
  DALLOC(e->corners_norm, static_cast<size_t>(e->Bmax) * 16 * 4);
  DALLOC(e->poses, static_cast<size_t>(e->Bmax) * 16 * 4);
  DALLOC(e->rec, static_cast<size_t>(e->Bmax) * 28 * 4);
  DALLOC(e->bbox3d_q, static_cast<size_t>(e->Bmax) * 24 * 4);
  DALLOC(e->K_q, static_cast<size_t>(e->Bmax) * 9 * 4);
  DALLOC(e->qidx, static_cast<size_t>(e->Bmax) * 8);
  DALLOC(e->pos_dec, P * d * 4);
  std::vector<float> tab;
  build_sincos(tab, e->d, e->g);
  CK(cudaMemcpy(e->pos_dec, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
  *out = e;
  return BD_OK;
}

extern "C" int bd_destroy(bd_handle e) {
  if (!e) return BD_OK;
  DevGuard dev_guard(e);
  cudaDeviceSynchronize();
  for (void* p : e->allocs) cudaFree(p);
  if (e->host_stream) cudaStreamDestroy(e->host_stream);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  for (auto& sl : e->slot) {
    for (int i = 0; i < 8; ++i) if (sl.copy_ev[i]) cudaEventDestroy(sl.copy_ev[i]);
    if (sl.done) cudaEventDestroy(sl.done);
  }
  drop_graphs(e);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  if (e->dev_ev) cudaEventDestroy(e->dev_ev);
  delete e;
  return BD_OK;
}

extern "C" int bd_load_weight(bd_handle e, const char* name, const void* data, const int64_t* shape, int32_t ndim) {
  if (!e || !name || !data || (ndim > 0 && !shape)) return fail(BD_ERR_INVALID, "bd_load_weight: null argument");
  DevGuard dev_guard(e);
  size_t n = 1;
  std::vector<int64_t> shp;
  for (int i = 0; i < ndim; ++i) { n *= static_cast<size_t>(shape[i]); shp.push_back(shape[i]); }
  Weight& w = e->w[name];
  if (w.f32 == nullptr || w.numel != n) {
    void* p = nullptr;
    int r = dalloc(e, &p, n * 4);
    if (r != BD_OK) return r;
    w.f32 = reinterpret_cast<float*>(p);
    w.numel = n;
    w.b16 = nullptr;
  }
  w.shape = shp;
  CK(cudaMemcpy(w.f32, data, n * 4, cudaMemcpyDefault));
  e->finalized = false;
  return BD_OK;
}

static int expect(bd_engine* e, const std::string& name, std::vector<int64_t> shape) {
  auto it = e->w.find(name);
  if (it == e->w.end()) return fail(BD_ERR_STATE, "missing weight: " + name);
  size_t n = 1;
  for (auto s : shape) n *= static_cast<size_t>(s);
  if (it->second.numel != n) return fail(BD_ERR_STATE, "weight has the wrong size: " + name);
  return BD_OK;
}

static int pack_bf16(bd_engine* e, const std::string& name) {
  Weight& w = e->w[name];
  if (!w.b16) {
    void* p = nullptr;
    int r = dalloc(e, &p, w.numel * 2);
    if (r != BD_OK) return r;
    w.b16 = reinterpret_cast<bf16*>(p);
  }
  CK(cast_f32_to_bf16(w.f32, w.b16, w.numel, 0));
  return BD_OK;
}

extern "C" int bd_finalize_weights(bd_handle e) {
  if (!e) return fail(BD_ERR_INVALID, "bd_finalize_weights: null handle");
  DevGuard dev_guard(e);
  drop_graphs(e);   // captured launch chains hold the addresses of the packed weights
  const int64_t d = e->d, pp8 = static_cast<int64_t>(e->patch) * e->patch * 8;
  std::vector<std::string> gemm_w;
  int r;
#define EXPECT(name, ...)                                   \
  do {                                                      \
    r = expect(e, (name), std::vector<int64_t>{__VA_ARGS__}); \
    if (r != BD_OK) return r;                               \
  } while (0)
  EXPECT("decoder.bbox_learnable_query", 1, d);
  for (int i = 0; i < e->cfg.dec_layers; ++i) {
    const std::string p = "decoder.attn." + std::to_string(i) + ".";
    EXPECT(p + "norm1.weight", d); EXPECT(p + "norm1.bias", d);
    EXPECT(p + "attn.qkv.weight", 3 * d, d); EXPECT(p + "attn.qkv.bias", 3 * d);
    EXPECT(p + "attn.q_norm.weight", e->hd_dec); EXPECT(p + "attn.k_norm.weight", e->hd_dec);
    EXPECT(p + "attn.proj.weight", d, d); EXPECT(p + "attn.proj.bias", d);
    EXPECT(p + "norm2.weight", d); EXPECT(p + "norm2.bias", d);
    EXPECT(p + "mlp.fc1.weight", 4 * d, d); EXPECT(p + "mlp.fc1.bias", 4 * d);
    EXPECT(p + "mlp.fc2.weight", d, 4 * d); EXPECT(p + "mlp.fc2.bias", d);
    gemm_w.push_back(p + "attn.qkv.weight"); gemm_w.push_back(p + "attn.proj.weight");
    gemm_w.push_back(p + "mlp.fc1.weight"); gemm_w.push_back(p + "mlp.fc2.weight");
  }
  EXPECT("decoder.bbox_proj.weight", pp8, d); EXPECT("decoder.bbox_proj.bias", pp8);
  EXPECT("decoder.input_transform.fc1.weight", d, d); EXPECT("decoder.input_transform.fc1.bias", d);
  EXPECT("decoder.input_transform.fc2.weight", d, d); EXPECT("decoder.input_transform.fc2.bias", d);
  EXPECT("decoder.bbox_emb.weight", d, pp8); EXPECT("decoder.bbox_emb.bias", d);
  gemm_w.push_back("decoder.bbox_proj.weight"); gemm_w.push_back("decoder.input_transform.fc1.weight");
  gemm_w.push_back("decoder.input_transform.fc2.weight"); gemm_w.push_back("decoder.bbox_emb.weight");
  EXPECT("dino.cls_token", 1, 1, d);
  EXPECT("dino.pos_embed", 1, 1 + e->P, d);
  EXPECT("dino.register_tokens", 1, e->cfg.dino_registers, d);
  EXPECT("dino.patch_embed.proj.weight", d, 3, e->patch, e->patch);
  EXPECT("dino.patch_embed.proj.bias", d);
  for (int i = 0; i < e->cfg.dino_layers; ++i) {
    const std::string p = "dino.blocks." + std::to_string(i) + ".";
    EXPECT(p + "norm1.weight", d); EXPECT(p + "norm1.bias", d);
    EXPECT(p + "attn.qkv.weight", 3 * d, d); EXPECT(p + "attn.qkv.bias", 3 * d);
    EXPECT(p + "attn.proj.weight", d, d); EXPECT(p + "attn.proj.bias", d);
    EXPECT(p + "ls1.gamma", d);
    EXPECT(p + "norm2.weight", d); EXPECT(p + "norm2.bias", d);
    EXPECT(p + "mlp.fc1.weight", 4 * d, d); EXPECT(p + "mlp.fc1.bias", 4 * d);
    EXPECT(p + "mlp.fc2.weight", d, 4 * d); EXPECT(p + "mlp.fc2.bias", d);
    EXPECT(p + "ls2.gamma", d);
    gemm_w.push_back(p + "attn.qkv.weight"); gemm_w.push_back(p + "attn.proj.weight");
    gemm_w.push_back(p + "mlp.fc1.weight"); gemm_w.push_back(p + "mlp.fc2.weight");
  }
  EXPECT("dino.norm.weight", d); EXPECT("dino.norm.bias", d);
#undef EXPECT
  if (e->tc) {
    for (const auto& n : gemm_w) {
      r = pack_bf16(e, n);
      if (r != BD_OK) return r;
    }
    // patch-embed weight [d, 588] -> bf16 [d, kpe] zero padded (K multiple of 64 for the 128B-swizzle TMA boxes)
    Weight& pw = e->w["dino.patch_embed.proj.weight"];
    const int kreal = 3 * e->patch * e->patch;
    std::vector<float> host(pw.numel);
    CK(cudaMemcpy(host.data(), pw.f32, pw.numel * 4, cudaMemcpyDeviceToHost));
    std::vector<float> padded(static_cast<size_t>(d) * e->kpe, 0.f);
    for (int64_t o = 0; o < d; ++o)
      for (int k = 0; k < kreal; ++k) padded[o * e->kpe + k] = host[o * kreal + k];
    float* tmp = nullptr;
    CK(cudaMalloc(&tmp, padded.size() * 4));
    CK(cudaMemcpy(tmp, padded.data(), padded.size() * 4, cudaMemcpyHostToDevice));
    if (!pw.b16) {
      void* p = nullptr;
      r = dalloc(e, &p, padded.size() * 2);
      if (r != BD_OK) { cudaFree(tmp); return r; }
      pw.b16 = reinterpret_cast<bf16*>(p);
    }
    cudaError_t ce = cast_f32_to_bf16(tmp, pw.b16, padded.size(), 0);
    cudaDeviceSynchronize();
    cudaFree(tmp);
    CK(ce);
  }
  CK(cudaDeviceSynchronize());
  e->finalized = true;
  return BD_OK;
}

// ---------------------------------------------------------------------------------------------

static const float* WF(bd_engine* e, const std::string& n) { return e->w[n].f32; }

static cudaError_t linear(bd_engine* e, const void* in, const std::string& wname, int M, int N, int K, int epi, GemmEpi& ep,
                          cudaStream_t s) {
  Weight& w = e->w[wname];
  if (e->tc) return gemm_tc(reinterpret_cast<const bf16*>(in), w.b16, M, N, K, epi, ep, s);
  if (epi == EPI_QKV) ep.out_act = e->qkv_scratch;
  return gemm_f32(reinterpret_cast<const float*>(in), w.f32, M, N, K, epi, ep, s);
}

static cudaError_t attention(bd_engine* e, int L, int heads, int hd, int seq, int seq_pad, cudaStream_t s) {
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  if (e->tc)
    return attention_tc(reinterpret_cast<const bf16*>(e->Q), reinterpret_cast<const bf16*>(e->K), reinterpret_cast<const bf16*>(e->V),
                        reinterpret_cast<bf16*>(e->O), L, heads, hd, seq, seq_pad, scale, s);
  return attention_f32(reinterpret_cast<const float*>(e->Q), reinterpret_cast<const float*>(e->K),
                       reinterpret_cast<const float*>(e->V), reinterpret_cast<float*>(e->O), L, heads, hd, seq, seq_pad, scale, s);
}

static cudaError_t ln_act(bd_engine* e, const float* x, const float* w, const float* b, float eps, int rows, cudaStream_t s) {
  return layernorm(x, w, b, eps, e->tc ? nullptr : reinterpret_cast<float*>(e->H), e->tc ? reinterpret_cast<bf16*>(e->H) : nullptr,
                   rows, e->d, 0, 0, 0, s);
}

// one pre-LN transformer block on the fp32 residual stream X [L*seq, d]
// (the LayerNorms stay stand-alone launches: fusing them behind the residual GEMMs was built and measured in round 2 and gains
// nothing on a power-capped B200, profiles/r02_ln_fusion_measured.txt)
static int run_block(bd_engine* e, float* X, const std::string& p, int L, int seq, int seq_pad, int heads, int hd, float ln_eps,
                     bool qk_norm, const char* g1, const char* g2, int attn_cat, cudaStream_t s) {
  const int M = L * seq, d = e->d;
  LAUNCH(BD_PROF_LAYERNORM, 1, ln_act(e, X, WF(e, p + "norm1.weight"), WF(e, p + "norm1.bias"), ln_eps, M, s));
  GemmEpi q;
  q.bias = WF(e, p + "attn.qkv.bias");
  q.q = e->Q; q.k = e->K; q.v = e->V;
  q.q_norm_w = qk_norm ? WF(e, p + "attn.q_norm.weight") : nullptr;
  q.k_norm_w = qk_norm ? WF(e, p + "attn.k_norm.weight") : nullptr;
  q.seq = seq; q.seq_pad = seq_pad; q.heads = heads; q.head_dim = hd; q.rms_eps = 1e-6f;
  LAUNCH(BD_PROF_GEMM_QKV, e->tc ? 1 : 2, linear(e, e->H, p + "attn.qkv.weight", M, 3 * d, d, EPI_QKV, q, s));
  LAUNCH(attn_cat, 1, attention(e, L, heads, hd, seq, seq_pad, s));
  GemmEpi pr;
  pr.bias = WF(e, p + "attn.proj.bias"); pr.out_f32 = X; pr.ldo = d; pr.gamma = g1 ? WF(e, p + g1) : nullptr;
  LAUNCH(BD_PROF_GEMM_PROJ, 1, linear(e, e->O, p + "attn.proj.weight", M, d, d, EPI_RESID, pr, s));
  LAUNCH(BD_PROF_LAYERNORM, 1, ln_act(e, X, WF(e, p + "norm2.weight"), WF(e, p + "norm2.bias"), ln_eps, M, s));
  GemmEpi f1;
  f1.bias = WF(e, p + "mlp.fc1.bias"); f1.out_act = e->G;
  LAUNCH(BD_PROF_GEMM_FC1, 1, linear(e, e->H, p + "mlp.fc1.weight", M, 4 * d, d, EPI_GELU, f1, s));
  GemmEpi f2;
  f2.bias = WF(e, p + "mlp.fc2.bias"); f2.out_f32 = X; f2.ldo = d; f2.gamma = g2 ? WF(e, p + g2) : nullptr;
  LAUNCH(BD_PROF_GEMM_FC2, 1, linear(e, e->G, p + "mlp.fc2.weight", M, d, 4 * d, EPI_RESID, f2, s));
  return BD_OK;
}

// The decoder's LAST block on the tensor path.  Only the query view's P tokens of each sequence reach the head (betr.py:419-430
// gathers them), so everything behind this block's K/V projection is computed for those rows only: attention from the query
// window over all T*P keys into a compact O, then proj / LayerNorm / MLP on the B*P gathered residual rows `Xc` (fp32, compact).
// Per-row arithmetic is unchanged (same k order in every GEMM, same key-tile order in the attention), so the head sees the
// bit-identical tokens; the block does 1/T of its attention and 1/T of 3 of its 4 GEMMs.  BD_LAST_LAYER_PRUNE=0 disables it.
static int run_block_query_rows(bd_engine* e, float* X, const std::string& p, int B, int T, int P, int seq_pad, int heads, int hd,
                                float ln_eps, const int64_t* query_idx, float* Xc, cudaStream_t s) {
  const int seq = T * P, M = B * seq, Mq = B * P, d = e->d;
  LAUNCH(BD_PROF_LAYERNORM, 1, ln_act(e, X, WF(e, p + "norm1.weight"), WF(e, p + "norm1.bias"), ln_eps, M, s));
  GemmEpi q;
  q.bias = WF(e, p + "attn.qkv.bias");
  q.q = e->Q; q.k = e->K; q.v = e->V;
  q.q_norm_w = WF(e, p + "attn.q_norm.weight");
  q.k_norm_w = WF(e, p + "attn.k_norm.weight");
  q.seq = seq; q.seq_pad = seq_pad; q.heads = heads; q.head_dim = hd; q.rms_eps = 1e-6f;
  LAUNCH(BD_PROF_GEMM_QKV, 1, linear(e, e->H, p + "attn.qkv.weight", M, 3 * d, d, EPI_QKV, q, s));
  LAUNCH(BD_PROF_ATTENTION_WINDOW, 1,
         attention_tc_window(reinterpret_cast<const bf16*>(e->Q), reinterpret_cast<const bf16*>(e->K), reinterpret_cast<const bf16*>(e->V),
                             reinterpret_cast<bf16*>(e->O), B, heads, hd, seq, seq_pad, 1.0f / sqrtf(static_cast<float>(hd)),
                             reinterpret_cast<const long long*>(query_idx), P, s));
  LAUNCH(BD_PROF_GLUE, 1, gather_query(X, query_idx, Xc, nullptr, B, T, P, d, s));
  GemmEpi pr;
  pr.bias = WF(e, p + "attn.proj.bias"); pr.out_f32 = Xc; pr.ldo = d;
  LAUNCH(BD_PROF_GEMM_PROJ, 1, linear(e, e->O, p + "attn.proj.weight", Mq, d, d, EPI_RESID, pr, s));
  LAUNCH(BD_PROF_LAYERNORM, 1, ln_act(e, Xc, WF(e, p + "norm2.weight"), WF(e, p + "norm2.bias"), ln_eps, Mq, s));
  GemmEpi f1;
  f1.bias = WF(e, p + "mlp.fc1.bias"); f1.out_act = e->G;
  LAUNCH(BD_PROF_GEMM_FC1, 1, linear(e, e->H, p + "mlp.fc1.weight", Mq, 4 * d, d, EPI_GELU, f1, s));
  GemmEpi f2;
  f2.bias = WF(e, p + "mlp.fc2.bias"); f2.out_f32 = Xc; f2.ldo = d;
  LAUNCH(BD_PROF_GEMM_FC2, 1, linear(e, e->G, p + "mlp.fc2.weight", Mq, d, 4 * d, EPI_RESID, f2, s));
  return BD_OK;
}

// all `layers` blocks "<prefix><i>." over L sequences of `seq` tokens
static int run_layers(bd_engine* e, float* X, const std::string& prefix, int layers, int L, int seq, int seq_pad, int heads, int hd,
                      float ln_eps, bool qk_norm, const char* g1, const char* g2, int attn_cat, cudaStream_t s) {
  for (int i = 0; i < layers; ++i) {
    int r = run_block(e, X, prefix + std::to_string(i) + ".", L, seq, seq_pad, heads, hd, ln_eps, qk_norm, g1, g2, attn_cat, s);
    if (r != BD_OK) return r;
  }
  return BD_OK;
}

// `out_img_off`: first image slot of feats_out / feats_act this call writes (the host-buffer entry runs the encoder in chunks
// while later images are still in flight); feats_out may be null on the tensor path (only the bf16 copy is consumed then).
static int dino_forward_impl(bd_engine* e, const void* images, int dtype, float* feats_out, int L, cudaStream_t s, int out_img_off = 0) {
  if (!e->finalized) return fail(BD_ERR_STATE, "weights not finalised (call bd_finalize_weights)");
  if (L <= 0 || out_img_off < 0 || out_img_off + L > e->Lmax) return fail(BD_ERR_INVALID, "bd_dino_forward: L exceeds max_batch*max_views");
  if (!feats_out && !e->tc) return fail(BD_ERR_INVALID, "bd_dino_forward: null output");
  const int P = e->P, d = e->d;
  LAUNCH(BD_PROF_GLUE, 1, im2col_patches(images, dtype == BD_BF16, e->A_pe, e->tc, L, e->S, e->patch, e->kpe, s));
  GemmEpi pe;
  pe.bias = WF(e, "dino.patch_embed.proj.bias");
  pe.out_f32 = e->X_dino; pe.ldo = d;
  pe.rp_in = P; pe.rp_out = e->n_tok; pe.rp_off = e->n_prefix;
  pe.addtab = WF(e, "dino.pos_embed") + d;  // rows 1.. of the (already interpolated) table
  LAUNCH(BD_PROF_GEMM_OTHER, 1, linear(e, e->A_pe, "dino.patch_embed.proj.weight", L * P, d, e->kpe, EPI_F32, pe, s));
  LAUNCH(BD_PROF_GLUE, 1, dino_prefix_tokens(e->X_dino, WF(e, "dino.cls_token"), WF(e, "dino.pos_embed"),
                                             WF(e, "dino.register_tokens"), L, e->n_tok, e->cfg.dino_registers, d, s));
  {
    int r = run_layers(e, e->X_dino, "dino.blocks.", e->cfg.dino_layers, L, e->n_tok, e->seqpad_dino, e->cfg.dino_heads, e->hd_dino, 1e-6f,
                       false, "ls1.gamma", "ls2.gamma", BD_PROF_ATTENTION_DINO, s);
    if (r != BD_OK) return r;
  }
  // final LayerNorm, patch tokens only (vision_transformer.py:263-267)
  const size_t out_off = static_cast<size_t>(out_img_off) * P * d;
  LAUNCH(BD_PROF_LAYERNORM, 1, layernorm(e->X_dino, WF(e, "dino.norm.weight"), WF(e, "dino.norm.bias"), 1e-6f,
                                         feats_out ? feats_out + out_off : nullptr,
                                         e->tc ? reinterpret_cast<bf16*>(e->feats_act) + out_off : nullptr, L * P, d, P, e->n_tok,
                                         e->n_prefix, s));
  return BD_OK;
}

static int decoder_forward_impl(bd_engine* e, const void* bbox_feat, int dtype, const float* feats, bool feats_act_valid,
                                const int64_t* query_idx, float* heat_out, float* logits_out, int B, int T, cudaStream_t s) {
  if (!e->finalized) return fail(BD_ERR_STATE, "weights not finalised (call bd_finalize_weights)");
  if (B <= 0 || T <= 0 || B > e->Bmax || T > e->Tmax) return fail(BD_ERR_INVALID, "bd_decoder_forward: B/T exceed the workspace");
  const int P = e->P, d = e->d, L = B * T, M = L * P;
  const int pp8 = e->patch * e->patch * 8;
  const void* fin = feats;
  if (e->tc) {
    if (!feats_act_valid) LAUNCH(BD_PROF_GLUE, 1, cast_f32_to_bf16(feats, reinterpret_cast<bf16*>(e->feats_act), static_cast<size_t>(M) * d, s));
    fin = e->feats_act;
  }
  // rgb branch: input_transform Mlp -> (LayerNorm without affine happens inside the fusion kernel)
  GemmEpi t1;
  t1.bias = WF(e, "decoder.input_transform.fc1.bias"); t1.out_act = e->G;
  LAUNCH(BD_PROF_GEMM_OTHER, 1, linear(e, fin, "decoder.input_transform.fc1.weight", M, d, d, EPI_GELU, t1, s));
  GemmEpi t2;
  t2.bias = WF(e, "decoder.input_transform.fc2.bias"); t2.out_f32 = e->R; t2.ldo = d;
  LAUNCH(BD_PROF_GEMM_OTHER, 1, linear(e, e->G, "decoder.input_transform.fc2.weight", M, d, d, EPI_F32, t2, s));
  // pose branch: patchify(bbox_feat) -> bbox_emb
  LAUNCH(BD_PROF_GLUE, 1, patchify_heat(bbox_feat, dtype == BD_BF16, e->A_bb, e->tc, L, 8, e->S, e->patch, s));
  GemmEpi be;
  be.bias = WF(e, "decoder.bbox_emb.bias"); be.out_f32 = e->PF; be.ldo = d;
  LAUNCH(BD_PROF_GEMM_OTHER, 1, linear(e, e->A_bb, "decoder.bbox_emb.weight", M, d, pp8, EPI_F32, be, s));
  LAUNCH(BD_PROF_GLUE, 1, betr_fuse(e->PF, e->R, WF(e, "decoder.bbox_learnable_query"), e->pos_dec, query_idx, e->X_dec, B, T, P, d, 1e-6f, s));
  const int seq = T * P, seq_pad = (seq + 127) / 128 * 128;
  {
    const char* pe = getenv("BD_LAST_LAYER_PRUNE");   // read per call so that a test can flip it
    const bool prune = e->tc && T > 1 && e->cfg.dec_layers >= 1 && !(pe && atoi(pe) == 0);
    const int full_layers = prune ? e->cfg.dec_layers - 1 : e->cfg.dec_layers;
    int r = run_layers(e, e->X_dec, "decoder.attn.", full_layers, B, seq, seq_pad, e->cfg.dec_heads, e->hd_dec, 1e-5f, true, nullptr,
                       nullptr, BD_PROF_ATTENTION, s);
    if (r != BD_OK) return r;
    if (prune) {   // last block on the query view's rows only; `R` (the adapter branch, consumed by betr_fuse) is free by now
      r = run_block_query_rows(e, e->X_dec, "decoder.attn." + std::to_string(full_layers) + ".", B, T, P, seq_pad, e->cfg.dec_heads, e->hd_dec,
                               1e-5f, query_idx, e->R, s);
      if (r != BD_OK) return r;
      LAUNCH(BD_PROF_GLUE, 1, cast_f32_to_bf16(e->R, reinterpret_cast<bf16*>(e->Xq), static_cast<size_t>(B) * P * d, s));
    } else {
      LAUNCH(BD_PROF_GLUE, 1, gather_query(e->X_dec, query_idx, e->tc ? nullptr : reinterpret_cast<float*>(e->Xq),
                                           e->tc ? reinterpret_cast<bf16*>(e->Xq) : nullptr, B, T, P, d, s));
    }
  }
  float* lg = logits_out ? logits_out : e->logits;
  GemmEpi bp;
  bp.bias = WF(e, "decoder.bbox_proj.bias"); bp.out_f32 = lg; bp.ldo = pp8;
  LAUNCH(BD_PROF_GEMM_OTHER, 1, linear(e, e->Xq, "decoder.bbox_proj.weight", B * P, pp8, d, EPI_F32, bp, s));
  LAUNCH(BD_PROF_GLUE, 1, unpatchify_sigmoid(lg, heat_out, B, 8, e->S, e->patch, s));
  return BD_OK;
}

// A device-pointer entry on stream `s` starts after every submitted host batch has finished with the workspace ...
static int order_after_host_batches(bd_engine* e, cudaStream_t s) {
  for (auto& sl : e->slot)
    if (sl.done && sl.pending) CK(cudaStreamWaitEvent(s, sl.done, 0));
  return BD_OK;
}
// ... and leaves a marker the next host batch waits for.
static int mark_device_entry_end(bd_engine* e, cudaStream_t s) {
  if (!e->dev_ev) CK(cudaEventCreateWithFlags(&e->dev_ev, cudaEventDisableTiming));
  CK(cudaEventRecord(e->dev_ev, s));
  e->dev_ev_valid = true;
  return BD_OK;
}

// ---- CUDA graphs of the launch chain --------------------------------------------------------------------------------------
// A forward is ~220 launches, each with host-side work (tensor-map encoding, weight look-ups): at batch 1 the host, not the GPU,
// sets the latency.  `body(stream)` enqueues a stage that reads and writes engine-owned buffers only, so its launches can be
// captured once per key (stage, shape, options) and replayed with one cudaGraphLaunch.  The first call of a key runs eagerly
// (one-time kernel attributes, lazy allocations); the second is captured on `cap_stream` -- the caller's stream may be the
// legacy default stream, which cannot be captured -- and the executable graph is launched into the caller's stream.  Profiling
// (per-launch events) and BOXDREAMER_B200_GRAPHS=0 take the eager path.
template <typename F>
static int graph_run(bd_engine* e, const std::vector<long long>& key, cudaStream_t s, F&& body) {
  if (!e->graphs_on || e->profile) return body(s);
  bd_engine::GraphEntry& g = e->graphs[key];
  if (g.exec) {
    CK(cudaGraphLaunch(g.exec, s));
    e->launches += g.launches;
    return BD_OK;
  }
  if (g.unusable || g.calls++ == 0) return body(s);
  if (!e->cap_stream) CK(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
  const long long l0 = e->launches;
  CK(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
  const int r = body(e->cap_stream);
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
  const long long n = e->launches - l0;
  e->launches = l0;   // nothing has run yet
  if (r == BD_OK && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&g.exec, graph, 0);
  if (graph) cudaGraphDestroy(graph);
  if (r != BD_OK || ce != cudaSuccess || !g.exec) {   // not capturable on this driver / with these options: stay eager for this key
    cudaGetLastError();
    g.exec = nullptr;
    g.unusable = true;
    return r != BD_OK ? r : body(s);
  }
  g.launches = n;
  CK(cudaGraphLaunch(g.exec, s));
  e->launches += n;
  return BD_OK;
}

static PnpOpts to_opts(const bd_pnp_opts* o) {
  PnpOpts p{0, 0, 1.0f, 0u, 30};
  if (o) { p.mode = o->mode; p.n_hyp = o->n_hyp; p.thr_px = o->thr_px; p.seed = o->seed; p.max_iter = o->max_iter; }
  return p;
}

static int ensure_staging(bd_engine* e, int k) {
  bd_engine::HostSlot& sl = e->slot[k];
  if (sl.in_images) return BD_OK;
  const size_t SS = static_cast<size_t>(e->S) * e->S, Lm = e->Lmax;
  DALLOC(sl.in_images, Lm * 3 * SS * 4);
  DALLOC(sl.in_bbox, Lm * 8 * SS * 4);
  return BD_OK;
}

// The two stages of a forward on the engine-owned staging buffers (in_images / in_bbox / qidx / bbox3d_q / K_q -> heat,
// corners_px, corners_norm, poses): the encoder over images [img0, img0 + L), and decoder + corner extraction + PnP.
static int encoder_body(bd_engine* e, int k, int dtype, int img0, int L, cudaStream_t s) {
  const size_t img_b = static_cast<size_t>(3) * e->S * e->S * (dtype == BD_BF16 ? 2 : 4);
  return dino_forward_impl(e, static_cast<char*>(e->slot[k].in_images) + img0 * img_b, dtype, e->tc ? nullptr : e->feats, L, s, img0);
}
static int decoder_post_body(bd_engine* e, int k, int dtype, int B, int T, const PnpOpts& po, cudaStream_t s) {
  int r = decoder_forward_impl(e, e->slot[k].in_bbox, dtype, e->feats, e->tc, e->qidx, e->heat, nullptr, B, T, s);
  if (r != BD_OK) return r;
  LAUNCH(BD_PROF_TOPK, 1, corners_topk(e->heat, e->corners_px, e->corners_norm, nullptr, B, 8, e->S, s));
  LAUNCH(BD_PROF_PNP, 1, pnp_solve(e->corners_px, e->bbox3d_q, e->K_q, e->poses, po, B, 8, s, po.mode == 0 ? e->rec : nullptr, e->corners_norm));
  return BD_OK;
}
static std::vector<long long> graph_key(int stage, int slot, int dtype, int a, int b, const PnpOpts& po) {
  long long thr_bits = 0;
  memcpy(&thr_bits, &po.thr_px, sizeof(float));
  return {stage, slot, dtype, a, b, po.mode, po.n_hyp, thr_bits, po.seed, po.max_iter};
}

extern "C" int bd_dino_forward(bd_handle e, const void* images, int32_t dtype, float* feats_out, int32_t L, void* stream) {
  if (!e || !images || !feats_out) return fail(BD_ERR_INVALID, "bd_dino_forward: null argument");
  DevGuard dev_guard(e);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int r = order_after_host_batches(e, s);
  if (r == BD_OK) r = dino_forward_impl(e, images, dtype, feats_out, L, s);
  return r != BD_OK ? r : mark_device_entry_end(e, s);
}

extern "C" int bd_decoder_forward(bd_handle e, const void* bbox_feat, int32_t dtype, const float* feats, const int64_t* query_idx,
                                  float* heat_out, float* logits_out, int32_t B, int32_t T, void* stream) {
  if (!e || !bbox_feat || !feats || !query_idx || !heat_out) return fail(BD_ERR_INVALID, "bd_decoder_forward: null argument");
  DevGuard dev_guard(e);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int r = order_after_host_batches(e, s);
  if (r == BD_OK) r = decoder_forward_impl(e, bbox_feat, dtype, feats, false, query_idx, heat_out, logits_out, B, T, s);
  return r != BD_OK ? r : mark_device_entry_end(e, s);
}

extern "C" int bd_corners_topk(bd_handle e, const float* heat, float* corners_px, float* corners_norm, int32_t* idx_out, int32_t B,
                               int32_t S, void* stream) {
  (void)e;
  if (!heat || !corners_px || !corners_norm) return fail(BD_ERR_INVALID, "bd_corners_topk: null argument");
  if (B < 0 || S <= 0) return fail(BD_ERR_INVALID, "bd_corners_topk: bad shape");
  CK(corners_topk(heat, corners_px, corners_norm, idx_out, B, 8, S, reinterpret_cast<cudaStream_t>(stream)));
  return BD_OK;
}

extern "C" int bd_pnp(bd_handle e, const float* corners_px, const float* bbox3d, const float* K, float* poses_out,
                      const bd_pnp_opts* opts, int32_t B, int32_t n_pts, void* stream) {
  (void)e;
  if (!corners_px || !bbox3d || !K || !poses_out) return fail(BD_ERR_INVALID, "bd_pnp: null argument");
  cudaError_t ce = pnp_solve(corners_px, bbox3d, K, poses_out, to_opts(opts), B, n_pts, reinterpret_cast<cudaStream_t>(stream));
  if (ce == cudaErrorNotSupported) return fail(BD_ERR_UNSUPPORTED, "bd_pnp: mode not built");
  if (ce == cudaErrorInvalidValue) return fail(BD_ERR_INVALID, "bd_pnp: n_pts must be in [6,64] (iterative mode) / [6,256] (robust mode)");
  CK(ce);
  return BD_OK;
}

// bd_forward / bd_forward_packed: device pointers in, device results out.  Null result pointers select the engine's own buffers;
// rec_out (mode 0 only) receives the packed [B, 28] record written by the PnP kernel's epilogue.
static int forward_device_inner(bd_handle e, const void* images, const void* bbox_feat, int32_t in_dtype, const int64_t* query_idx,
                                const float* bbox3d_q, const float* K_q, float* heat_out, float* corners_px, float* corners_norm,
                                float* poses_out, float* rec_out, const bd_pnp_opts* opts, int32_t B, int32_t T, cudaStream_t s);
static int forward_device_impl(bd_handle e, const void* images, const void* bbox_feat, int32_t in_dtype, const int64_t* query_idx,
                               const float* bbox3d_q, const float* K_q, float* heat_out, float* corners_px, float* corners_norm,
                               float* poses_out, float* rec_out, const bd_pnp_opts* opts, int32_t B, int32_t T, void* stream) {
  DevGuard dev_guard(e);
  if (B <= 0 || T <= 0 || B > e->Bmax || T > e->Tmax) return fail(BD_ERR_INVALID, "bd_forward: B/T exceed the workspace");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int r = order_after_host_batches(e, s);
  if (r == BD_OK)
    r = forward_device_inner(e, images, bbox_feat, in_dtype, query_idx, bbox3d_q, K_q, heat_out, corners_px, corners_norm, poses_out, rec_out,
                             opts, B, T, s);
  return r != BD_OK ? r : mark_device_entry_end(e, s);
}
static int forward_device_inner(bd_handle e, const void* images, const void* bbox_feat, int32_t in_dtype, const int64_t* query_idx,
                                const float* bbox3d_q, const float* K_q, float* heat_out, float* corners_px, float* corners_norm,
                                float* poses_out, float* rec_out, const bd_pnp_opts* opts, int32_t B, int32_t T, cudaStream_t s) {
  const PnpOpts po = to_opts(opts);
  if (po.mode != 0 && po.mode != 1) return fail(BD_ERR_UNSUPPORTED, "bd_forward: pnp mode not built");
  if (rec_out && po.mode != 0) return fail(BD_ERR_UNSUPPORTED, "bd_forward_packed: the packed record is written by the iterative PnP kernel (mode 0) only");
  if (e->graphs_on && !e->profile && B * T <= e->graph_max_views) {
    // small shapes are launch-bound: stage the inputs into the engine's own buffers (device-to-device, a few MB) and replay
    // the captured launch chain; the results are copied out of the workspace afterwards
    int r = ensure_staging(e, 0);
    if (r != BD_OK) return r;
    const size_t es = in_dtype == BD_BF16 ? 2 : 4, SS = static_cast<size_t>(e->S) * e->S, L = static_cast<size_t>(B) * T;
    CK(cudaMemcpyAsync(e->slot[0].in_images, images, L * 3 * SS * es, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(e->slot[0].in_bbox, bbox_feat, L * 8 * SS * es, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(e->qidx, query_idx, static_cast<size_t>(B) * 8, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(e->bbox3d_q, bbox3d_q, static_cast<size_t>(B) * 24 * 4, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(e->K_q, K_q, static_cast<size_t>(B) * 9 * 4, cudaMemcpyDeviceToDevice, s));
    r = graph_run(e, graph_key(2, 0, in_dtype, B, T, po), s, [&](cudaStream_t st) {
      const int q = encoder_body(e, 0, in_dtype, 0, B * T, st);
      return q != BD_OK ? q : decoder_post_body(e, 0, in_dtype, B, T, po, st);
    });
    if (r != BD_OK) return r;
    if (corners_px) CK(cudaMemcpyAsync(corners_px, e->corners_px, static_cast<size_t>(B) * 16 * 4, cudaMemcpyDeviceToDevice, s));
    if (corners_norm) CK(cudaMemcpyAsync(corners_norm, e->corners_norm, static_cast<size_t>(B) * 16 * 4, cudaMemcpyDeviceToDevice, s));
    if (poses_out) CK(cudaMemcpyAsync(poses_out, e->poses, static_cast<size_t>(B) * 16 * 4, cudaMemcpyDeviceToDevice, s));
    if (rec_out) CK(cudaMemcpyAsync(rec_out, e->rec, static_cast<size_t>(B) * 28 * 4, cudaMemcpyDeviceToDevice, s));
    if (heat_out) CK(cudaMemcpyAsync(heat_out, e->heat, static_cast<size_t>(B) * 8 * SS * 4, cudaMemcpyDeviceToDevice, s));
    return BD_OK;
  }
  float* heat = heat_out ? heat_out : e->heat;
  if (!corners_px) corners_px = e->corners_px;
  if (!corners_norm) corners_norm = e->corners_norm;
  if (!poses_out) poses_out = e->poses;
  int r = dino_forward_impl(e, images, in_dtype, e->tc ? nullptr : e->feats, B * T, s);
  if (r != BD_OK) return r;
  r = decoder_forward_impl(e, bbox_feat, in_dtype, e->feats, e->tc, query_idx, heat, nullptr, B, T, s);
  if (r != BD_OK) return r;
  LAUNCH(BD_PROF_TOPK, 1, corners_topk(heat, corners_px, corners_norm, nullptr, B, 8, e->S, s));
  LAUNCH(BD_PROF_PNP, 1, pnp_solve(corners_px, bbox3d_q, K_q, poses_out, po, B, 8, s, rec_out, corners_norm));
  return BD_OK;
}

extern "C" int bd_forward(bd_handle e, const void* images, const void* bbox_feat, int32_t in_dtype, const int64_t* query_idx,
                          const float* bbox3d_q, const float* K_q, float* heat_out, float* corners_px, float* corners_norm,
                          float* poses_out, const bd_pnp_opts* opts, int32_t B, int32_t T, void* stream) {
  if (!e || !images || !bbox_feat || !query_idx || !bbox3d_q || !K_q || !corners_px || !corners_norm || !poses_out)
    return fail(BD_ERR_INVALID, "bd_forward: null argument");
  return forward_device_impl(e, images, bbox_feat, in_dtype, query_idx, bbox3d_q, K_q, heat_out, corners_px, corners_norm, poses_out,
                             nullptr, opts, B, T, stream);
}

extern "C" int bd_forward_packed(bd_handle e, const void* images, const void* bbox_feat, int32_t in_dtype, const int64_t* query_idx,
                                 const float* bbox3d_q, const float* K_q, float* rec_out, const bd_pnp_opts* opts, int32_t B, int32_t T,
                                 void* stream) {
  if (!e || !images || !bbox_feat || !query_idx || !bbox3d_q || !K_q || !rec_out) return fail(BD_ERR_INVALID, "bd_forward_packed: null argument");
  return forward_device_impl(e, images, bbox_feat, in_dtype, query_idx, bbox3d_q, K_q, nullptr, nullptr, nullptr, nullptr, rec_out, opts,
                             B, T, stream);
}

// Host-buffer entries.  submit: enqueue the H2D copies (copy stream), the forward (compute stream) and the D2H copies of the results
// for staging slot k, return without waiting.  wait: block until slot k's results are in the host buffers given to submit.
static int host_wait_impl(bd_handle e, int k) {
  if (!e || k < 0 || k > 1) return fail(BD_ERR_INVALID, "bd_forward_host_wait: bad handle or slot");
  bd_engine::HostSlot& sl = e->slot[k];
  if (!sl.pending) return BD_OK;
  DevGuard dev_guard(e);
  sl.pending = false;
  CK(cudaEventSynchronize(sl.done));
  return BD_OK;
}

static int host_submit_impl(bd_handle e, int k, const void* images_host, const void* bbox_feat_host, const float* bbox_px_host,
                            int32_t in_dtype, const int64_t* query_idx_host, const float* bbox3d_q_host, const float* K_q_host,
                            float* heat_out_host, float* corners_px_host, float* corners_norm_host, float* poses_out_host,
                            const bd_pnp_opts* opts, int32_t B, int32_t T) {
  if (!e || !images_host || (!bbox_feat_host && !bbox_px_host) || !query_idx_host || !bbox3d_q_host || !K_q_host || !corners_px_host ||
      !corners_norm_host || !poses_out_host)
    return fail(BD_ERR_INVALID, "bd_forward_host: null argument");
  if (k < 0 || k > 1) return fail(BD_ERR_INVALID, "bd_forward_host_submit: slot must be 0 or 1");
  DevGuard dev_guard(e);
  if (B <= 0 || T <= 0 || B > e->Bmax || T > e->Tmax) return fail(BD_ERR_INVALID, "bd_forward_host: B/T exceed the workspace");
  const PnpOpts po = to_opts(opts);
  if (po.mode != 0 && po.mode != 1) return fail(BD_ERR_UNSUPPORTED, "bd_forward_host: pnp mode not built");
  const size_t es = in_dtype == BD_BF16 ? 2 : 4;
  const size_t SS = static_cast<size_t>(e->S) * e->S;
  {
    int r = host_wait_impl(e, k);   // a slot is reused only after its previous batch has been delivered
    if (r != BD_OK) return r;
    r = ensure_staging(e, k);
    if (r != BD_OK) return r;
  }
  bd_engine::HostSlot& sl = e->slot[k];
  if (!e->host_stream) CK(cudaStreamCreateWithFlags(&e->host_stream, cudaStreamNonBlocking));
  if (!e->copy_stream) CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  if (!sl.done) {
    for (int i = 0; i < 8; ++i) CK(cudaEventCreateWithFlags(&sl.copy_ev[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
  }
  cudaStream_t s = e->host_stream;
  if (e->dev_ev_valid) {   // a device-pointer entry may still be using the workspace (and slot 0's staging) on the caller's stream
    CK(cudaStreamWaitEvent(s, e->dev_ev, 0));
    CK(cudaStreamWaitEvent(e->copy_stream, e->dev_ev, 0));
  }
  // Copy order = consumption order: the encoder only needs the images, so they go first; the reference heat maps (73 % of
  // the bytes) follow and land while the encoder runs.  The decoder, the corner extraction and PnP then see the whole batch
  // once.  The images can be split into `nchunk` pieces of whole queries (first one half-sized) so that the encoder starts
  // earlier, but at BASELINE config 2 the smaller encoder GEMMs cost more than the 2 ms of exposed transfer they hide
  // (scripts/bench_e2e_chunks.py on B200: 1 chunk 49.3 ms, 2: 50.7, 3: 51.8, 4: 51.1) -- default 1.  A caller that wants the
  // transfer hidden completely alternates the two slots (submit k+1 before wait k).
  int nchunk = 1;
  if (const char* ev = getenv("BOXDREAMER_B200_HOST_CHUNKS")) nchunk = atoi(ev);
  if (nchunk < 1) nchunk = 1;
  if (nchunk > 7) nchunk = 7;
  if (nchunk > B) nchunk = B;
  CK(cudaMemcpyAsync(e->qidx, query_idx_host, static_cast<size_t>(B) * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(e->bbox3d_q, bbox3d_q_host, static_cast<size_t>(B) * 24 * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(e->K_q, K_q_host, static_cast<size_t>(B) * 9 * 4, cudaMemcpyHostToDevice, s));
  const size_t img_q = static_cast<size_t>(T) * 3 * SS * es, box_q = static_cast<size_t>(T) * 8 * SS * es;  // bytes per query
  int b0s[9];
  b0s[0] = 0;
  for (int c = 1; c <= nchunk; ++c)   // chunk sizes 1 : 2 : 2 : ... (first one half as large as the others)
    b0s[c] = c == nchunk ? B : static_cast<int>((static_cast<long long>(B) * (2 * c - 1) + (2 * nchunk - 2)) / (2 * nchunk - 1));
  for (int c = 0; c < nchunk; ++c) {
    const int b0 = b0s[c], nb = b0s[c + 1] - b0s[c];
    if (nb > 0)
      CK(cudaMemcpyAsync(static_cast<char*>(sl.in_images) + b0 * img_q, static_cast<const char*>(images_host) + b0 * img_q, nb * img_q,
                         cudaMemcpyHostToDevice, e->copy_stream));
    CK(cudaEventRecord(sl.copy_ev[c], e->copy_stream));
  }
  if (bbox_feat_host) {
    CK(cudaMemcpyAsync(sl.in_bbox, bbox_feat_host, B * box_q, cudaMemcpyHostToDevice, e->copy_stream));
  } else {  // 64 bytes per view instead of the maps; rasterised on the device, on the copy stream (overlaps the encoder)
    if (!sl.in_bbox_px) DALLOC(sl.in_bbox_px, static_cast<size_t>(e->Lmax) * 16 * 4);
    CK(cudaMemcpyAsync(sl.in_bbox_px, bbox_px_host, static_cast<size_t>(B) * T * 16 * 4, cudaMemcpyHostToDevice, e->copy_stream));
    CK(bbox_heatmaps(reinterpret_cast<const float*>(sl.in_bbox_px), sl.in_bbox, in_dtype == BD_BF16, B * T, e->S, T, e->copy_stream));
    e->launches += 1;
  }
  CK(cudaEventRecord(sl.copy_ev[7], e->copy_stream));
  for (int c = 0; c < nchunk; ++c) {
    const int b0 = b0s[c], nb = b0s[c + 1] - b0s[c];
    CK(cudaStreamWaitEvent(s, sl.copy_ev[c], 0));
    if (nb <= 0) continue;
    const PnpOpts none{};
    int r = graph_run(e, graph_key(0, k, in_dtype, b0 * T, nb * T, none), s,
                      [&](cudaStream_t st) { return encoder_body(e, k, in_dtype, b0 * T, nb * T, st); });
    if (r != BD_OK) return r;
  }
  CK(cudaStreamWaitEvent(s, sl.copy_ev[7], 0));
  {
    int r = graph_run(e, graph_key(1, k, in_dtype, B, T, po), s, [&](cudaStream_t st) { return decoder_post_body(e, k, in_dtype, B, T, po, st); });
    if (r != BD_OK) return r;
  }
  CK(cudaMemcpyAsync(corners_px_host, e->corners_px, static_cast<size_t>(B) * 16 * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(corners_norm_host, e->corners_norm, static_cast<size_t>(B) * 16 * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(poses_out_host, e->poses, static_cast<size_t>(B) * 16 * 4, cudaMemcpyDeviceToHost, s));
  if (heat_out_host) CK(cudaMemcpyAsync(heat_out_host, e->heat, static_cast<size_t>(B) * 8 * SS * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaEventRecord(sl.done, s));
  sl.pending = true;
  return BD_OK;
}

static int forward_host_impl(bd_handle e, const void* images_host, const void* bbox_feat_host, const float* bbox_px_host,
                             int32_t in_dtype, const int64_t* query_idx_host, const float* bbox3d_q_host, const float* K_q_host,
                             float* heat_out_host, float* corners_px_host, float* corners_norm_host, float* poses_out_host,
                             const bd_pnp_opts* opts, int32_t B, int32_t T) {
  int r = host_submit_impl(e, 0, images_host, bbox_feat_host, bbox_px_host, in_dtype, query_idx_host, bbox3d_q_host, K_q_host, heat_out_host,
                           corners_px_host, corners_norm_host, poses_out_host, opts, B, T);
  return r != BD_OK ? r : host_wait_impl(e, 0);
}

extern "C" int bd_forward_host_submit(bd_handle e, int32_t slot, const void* images_host, const void* bbox_feat_host,
                                      const float* bbox_px_host, int32_t in_dtype, const int64_t* query_idx_host,
                                      const float* bbox3d_q_host, const float* K_q_host, float* heat_out_host, float* corners_px_host,
                                      float* corners_norm_host, float* poses_out_host, const bd_pnp_opts* opts, int32_t B, int32_t T) {
  return host_submit_impl(e, slot, images_host, bbox_feat_host, bbox_px_host, in_dtype, query_idx_host, bbox3d_q_host, K_q_host,
                          heat_out_host, corners_px_host, corners_norm_host, poses_out_host, opts, B, T);
}

extern "C" int bd_forward_host_wait(bd_handle e, int32_t slot) { return host_wait_impl(e, slot); }

extern "C" int bd_forward_host(bd_handle e, const void* images_host, const void* bbox_feat_host, int32_t in_dtype,
                               const int64_t* query_idx_host, const float* bbox3d_q_host, const float* K_q_host,
                               float* heat_out_host, float* corners_px_host, float* corners_norm_host, float* poses_out_host,
                               const bd_pnp_opts* opts, int32_t B, int32_t T) {
  if (!bbox_feat_host) return fail(BD_ERR_INVALID, "bd_forward_host: null argument");
  return forward_host_impl(e, images_host, bbox_feat_host, nullptr, in_dtype, query_idx_host, bbox3d_q_host, K_q_host, heat_out_host,
                           corners_px_host, corners_norm_host, poses_out_host, opts, B, T);
}

extern "C" int bd_forward_host_px(bd_handle e, const void* images_host, const float* bbox_px_host, int32_t in_dtype,
                                  const int64_t* query_idx_host, const float* bbox3d_q_host, const float* K_q_host,
                                  float* heat_out_host, float* corners_px_host, float* corners_norm_host,
                                  float* poses_out_host, const bd_pnp_opts* opts, int32_t B, int32_t T) {
  if (!bbox_px_host) return fail(BD_ERR_INVALID, "bd_forward_host_px: null argument");
  return forward_host_impl(e, images_host, nullptr, bbox_px_host, in_dtype, query_idx_host, bbox3d_q_host, K_q_host, heat_out_host,
                           corners_px_host, corners_norm_host, poses_out_host, opts, B, T);
}

extern "C" int bd_pose_metrics(const float* pose_pred, const float* pose_gt, const float* K, const float* model_pts,
                               int64_t pts_stride, float* out, int32_t B, int32_t N, void* stream) {
  if (!pose_pred || !pose_gt || !K || !model_pts || !out || B <= 0 || N <= 0) return fail(BD_ERR_INVALID, "bd_pose_metrics: bad argument");
  cudaError_t err = pose_metrics(pose_pred, pose_gt, K, model_pts, pts_stride, out, B, N, reinterpret_cast<cudaStream_t>(stream));
  if (err != cudaSuccess) return fail(BD_ERR_CUDA, std::string("bd_pose_metrics: ") + cudaGetErrorString(err));
  return BD_OK;
}

extern "C" int bd_make_bbox_features(const float* bbox_px, void* out, int32_t out_dtype, int32_t L, int32_t S, int32_t group,
                                     void* stream) {
  if (!bbox_px || !out || L <= 0 || S <= 0 || group <= 0 || L % group != 0) return fail(BD_ERR_INVALID, "bd_make_bbox_features: bad argument");
  if (out_dtype != BD_F32 && out_dtype != BD_BF16) return fail(BD_ERR_INVALID, "bd_make_bbox_features: dtype must be BD_F32 or BD_BF16");
  cudaError_t err = bbox_heatmaps(bbox_px, out, out_dtype == BD_BF16, L, S, group, reinterpret_cast<cudaStream_t>(stream));
  if (err != cudaSuccess) return fail(BD_ERR_CUDA, std::string("bd_make_bbox_features: ") + cudaGetErrorString(err));
  return BD_OK;
}

// ---------------------------------------------------------------------------------------------
// kernel-level entry points

extern "C" int bd_gemm(const void* A, const void* W, const float* bias, const float* gamma, void* out, int32_t M, int32_t N,
                       int32_t K, int32_t epilogue, int32_t precision, void* stream) {
  if (!A || !W || !bias || !out) return fail(BD_ERR_INVALID, "bd_gemm: null argument");
  if (epilogue != EPI_F32 && epilogue != EPI_GELU && epilogue != EPI_RESID && epilogue != EPI_ACT)
    return fail(BD_ERR_INVALID, "bd_gemm: epilogue must be 0, 1, 2 or 4");
  GemmEpi e;
  e.bias = bias; e.gamma = gamma; e.ldo = N;
  if (epilogue == EPI_F32 || epilogue == EPI_RESID) e.out_f32 = reinterpret_cast<float*>(out);
  else e.out_act = out;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (precision == BD_PRECISION_BF16)
    CK(gemm_tc(reinterpret_cast<const bf16*>(A), reinterpret_cast<const bf16*>(W), M, N, K, epilogue, e, s));
  else
    CK(gemm_f32(reinterpret_cast<const float*>(A), reinterpret_cast<const float*>(W), M, N, K, epilogue, e, s));
  return BD_OK;
}

extern "C" int bd_qkv_project(const void* x, const void* W, const float* bias, const float* q_norm_w, const float* k_norm_w,
                              void* Q, void* K, void* V, void* scratch, int32_t L, int32_t seq, int32_t seq_pad, int32_t heads,
                              int32_t head_dim, int32_t precision, void* stream) {
  if (!x || !W || !bias || !Q || !K || !V) return fail(BD_ERR_INVALID, "bd_qkv_project: null argument");
  GemmEpi e;
  e.bias = bias; e.q = Q; e.k = K; e.v = V; e.q_norm_w = q_norm_w; e.k_norm_w = k_norm_w;
  e.seq = seq; e.seq_pad = seq_pad; e.heads = heads; e.head_dim = head_dim; e.rms_eps = 1e-6f;
  const int d = heads * head_dim;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (precision == BD_PRECISION_BF16) {
    CK(gemm_tc(reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(W), L * seq, 3 * d, d, EPI_QKV, e, s));
  } else {
    if (!scratch) return fail(BD_ERR_INVALID, "bd_qkv_project: exact path needs scratch");
    e.out_act = scratch;
    CK(gemm_f32(reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(W), L * seq, 3 * d, d, EPI_QKV, e, s));
  }
  return BD_OK;
}

extern "C" int bd_attention(const void* Q, const void* K, const void* V, void* O, int32_t L, int32_t heads, int32_t head_dim,
                            int32_t seq, int32_t seq_pad, float scale, int32_t precision, int32_t variant, void* stream) {
  if (!Q || !K || !V || !O) return fail(BD_ERR_INVALID, "bd_attention: null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (precision == BD_PRECISION_BF16)
    CK(attention_tc(reinterpret_cast<const bf16*>(Q), reinterpret_cast<const bf16*>(K), reinterpret_cast<const bf16*>(V),
                    reinterpret_cast<bf16*>(O), L, heads, head_dim, seq, seq_pad, scale, s));
  else
    CK(attention_f32(reinterpret_cast<const float*>(Q), reinterpret_cast<const float*>(K), reinterpret_cast<const float*>(V),
                     reinterpret_cast<float*>(O), L, heads, head_dim, seq, seq_pad, scale, s));
  return BD_OK;
}

extern "C" int bd_layernorm(const float* x, const float* w, const float* b, float eps, float* out_f32, void* out_bf16, int32_t rows,
                            int32_t d, void* stream) {
  if (!x || (!out_f32 && !out_bf16)) return fail(BD_ERR_INVALID, "bd_layernorm: null argument");
  CK(layernorm(x, w, b, eps, out_f32, reinterpret_cast<bf16*>(out_bf16), rows, d, 0, 0, 0, reinterpret_cast<cudaStream_t>(stream)));
  return BD_OK;
}

// ---------------------------------------------------------------------------------------------
// instrumentation

extern "C" long long bd_launch_count(bd_handle e) { return e ? e->launches : 0; }

extern "C" int bd_profile_enable(bd_handle e, int32_t on) {
  if (!e) return fail(BD_ERR_INVALID, "bd_profile_enable: null handle");
  e->profile = on != 0;
  return BD_OK;
}

extern "C" int bd_profile_read(bd_handle e, double* ms_out, int64_t* count_out, int32_t reset) {
  if (!e || !ms_out || !count_out) return fail(BD_ERR_INVALID, "bd_profile_read: null argument");
  CK(cudaDeviceSynchronize());
  for (auto& sp : e->spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
      e->cat_ms[sp.cat] += ms;
      e->cat_n[sp.cat] += 1;
    }
    e->ev_pool.push_back(sp.a);
    e->ev_pool.push_back(sp.b);
  }
  e->spans.clear();
  for (int i = 0; i < BD_PROF_NCAT; ++i) { ms_out[i] = e->cat_ms[i]; count_out[i] = e->cat_n[i]; }
  if (reset) for (int i = 0; i < BD_PROF_NCAT; ++i) { e->cat_ms[i] = 0; e->cat_n[i] = 0; }
  return BD_OK;
}

// debug aid: per-role clock stamps of the v2 attention kernel (CTA 0); dev_buf must hold 3*512 int64, nullptr disables
namespace bd { void attention_tc2_set_trace(long long* dev_buf); }
extern "C" int bd_debug_attention_trace(void* dev_buf) {
  bd::attention_tc2_set_trace(reinterpret_cast<long long*>(dev_buf));
  return BD_OK;
}
