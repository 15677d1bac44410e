// The step runtime: each chain of the distillation step behind ONE C-ABI call.
//
//   lgd_teacher_forward / lgd_teacher_backward   DynamicTeacher.forward and its backward
//                                                (dynamic_teacher/dynamic_teacher.py:209-301, label_encoder.py:216-276,
//                                                 spatial_transformer.py:30-47)
//   lgd_distill_forward / lgd_distill_backward   BaseDistillator.distill with the SequentialConvs adapter
//                                                (base_distillator.py:34-64, adapters/sequential_convs.py:7-15)
//
// A call enqueues every kernel of its chain (about 60-110 launches) from native code: the host side of the plug-in
// (Python) only allocates three buffers per call -- outputs, a "tape" that lives until the backward, and scratch that
// dies with the call -- and this file carves them up. Nothing is allocated here, nothing synchronises with the host;
// the weight-gradient GEMMs and the label-side backward run on the context's side streams underneath the chain and are
// joined back into the caller's stream before the call returns, so stream-ordered allocators may reuse every buffer as
// soon as the call has returned. The kernels are the library's own entry points (include/lgd_b200.h), issued in the
// same order as the per-kernel orchestration in lgd_b200/engine.py, so both produce bit-identical results
// (tests/test_gpu_chain.py).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "tokenprog.h"

namespace lgd {

constexpr int DESC = LGD_DESC_DIM;
constexpr size_t LIN_WS_BYTES = (size_t)32 << 20;  // split-K scratch of the small-T linears (per stream)

}  // namespace lgd

using namespace lgd;

// ------------------------------------------------------------------------------------------------ context
struct lgd_ctx {
  int device = 0;
  cudaStream_t wgrad_stream = nullptr;  // weight-gradient GEMMs
  cudaStream_t label_stream = nullptr;  // label-side backward (canoni projection + label encoder)
  cudaEvent_t ev_ready = nullptr, ev_label = nullptr, ev_join = nullptr;
  // teacher backward: every parameter gradient except student_proj_2D's is complete once these (and ev_label) fired
  cudaEvent_t ev_early_main = nullptr, ev_early_wgrad = nullptr;
  bool early_label = false;   // the last teacher backward recorded ev_label on the label stream
  bool side_streams = true;
  bool profiling = false;
  bool token_programs = true;   // label encoder forward / label-side backward as one persistent kernel each
  bool fuse_gn_sums = true;     // GroupNorm-backward sums in the epilogue of the dgrad that produces its input gradient
  bool tap_render = true;       // local_inst_proj_2D from per-box tap vectors instead of a convolution (taprender.cu)
  // pinned staging ring for the token programs (op lists travel host -> device asynchronously)
  static constexpr int SLOTS = 8;
  static constexpr size_t SLOT_BYTES = 96 << 10;
  char* pinned = nullptr;
  cudaEvent_t slot_done[SLOTS] = {};
  int next_slot = 0;
  struct Rec {
    const char* name;
    cudaEvent_t e0, e1;
  };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  std::mutex mu;
};

extern "C" lgd_ctx_t* lgd_ctx_create(void) {
  lgd_ctx* c = new lgd_ctx();
  const char* env = getenv("LGD_B200_TOKENPROG");
  if (env != nullptr && env[0] == '0') c->token_programs = false;
  env = getenv("LGD_B200_GN_FUSE");
  if (env != nullptr && env[0] == '0') c->fuse_gn_sums = false;
  env = getenv("LGD_B200_TAP_RENDER");
  if (env != nullptr && env[0] == '0') c->tap_render = false;
  if (cudaGetDevice(&c->device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->wgrad_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->label_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_label, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_early_main, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_early_wgrad, cudaEventDisableTiming) != cudaSuccess ||
      cudaHostAlloc(reinterpret_cast<void**>(&c->pinned), lgd_ctx::SLOTS * lgd_ctx::SLOT_BYTES, cudaHostAllocDefault) !=
          cudaSuccess) {
    set_error("lgd_ctx_create: %s", cudaGetErrorString(cudaGetLastError()));
    delete c;
    return nullptr;
  }
  return c;
}

extern "C" void lgd_ctx_destroy(lgd_ctx_t* c) {
  if (c == nullptr) return;
  for (auto& r : c->recs) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  for (auto e : c->pool) cudaEventDestroy(e);
  if (c->wgrad_stream) cudaStreamDestroy(c->wgrad_stream);
  if (c->label_stream) cudaStreamDestroy(c->label_stream);
  if (c->ev_ready) cudaEventDestroy(c->ev_ready);
  if (c->ev_label) cudaEventDestroy(c->ev_label);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->ev_early_main) cudaEventDestroy(c->ev_early_main);
  if (c->ev_early_wgrad) cudaEventDestroy(c->ev_early_wgrad);
  for (auto e : c->slot_done)
    if (e) cudaEventDestroy(e);
  if (c->pinned) cudaFreeHost(c->pinned);
  delete c;
}

extern "C" int lgd_ctx_wait_early_grads(lgd_ctx_t* c, void* stream) {
  LGD_CHECK_ARG(c != nullptr, "lgd_ctx_wait_early_grads: null context");
  cudaStream_t s = (cudaStream_t)stream;
  LGD_CUDA(cudaStreamWaitEvent(s, c->ev_early_main, 0));
  LGD_CUDA(cudaStreamWaitEvent(s, c->ev_early_wgrad, 0));
  if (c->early_label) LGD_CUDA(cudaStreamWaitEvent(s, c->ev_label, 0));
  return LGD_OK;
}

extern "C" int lgd_ctx_set_side_streams(lgd_ctx_t* c, int enable) {
  LGD_CHECK_ARG(c != nullptr, "lgd_ctx_set_side_streams: null context");
  c->side_streams = enable != 0;
  return LGD_OK;
}

extern "C" int lgd_ctx_set_token_programs(lgd_ctx_t* c, int enable) {
  LGD_CHECK_ARG(c != nullptr, "lgd_ctx_set_token_programs: null context");
  c->token_programs = enable != 0;
  return LGD_OK;
}

extern "C" int lgd_ctx_profile(lgd_ctx_t* c, int enable) {
  LGD_CHECK_ARG(c != nullptr, "lgd_ctx_profile: null context");
  std::lock_guard<std::mutex> lk(c->mu);
  c->profiling = enable != 0;
  return LGD_OK;
}

extern "C" int lgd_ctx_profile_count(lgd_ctx_t* c) { return c ? (int)c->recs.size() : 0; }

extern "C" int lgd_ctx_profile_get(lgd_ctx_t* c, int i, const char** name, float* ms) {
  LGD_CHECK_ARG(c != nullptr && i >= 0 && i < (int)c->recs.size() && name && ms, "lgd_ctx_profile_get: bad arguments");
  *name = c->recs[i].name;
  LGD_CUDA(cudaEventElapsedTime(ms, c->recs[i].e0, c->recs[i].e1));
  return LGD_OK;
}

extern "C" void lgd_ctx_profile_reset(lgd_ctx_t* c) {
  if (c == nullptr) return;
  std::lock_guard<std::mutex> lk(c->mu);
  for (auto& r : c->recs) {
    c->pool.push_back(r.e0);
    c->pool.push_back(r.e1);
  }
  c->recs.clear();
}

namespace lgd {

static cudaEvent_t timing_event(lgd_ctx* c) {
  if (!c->pool.empty()) {
    cudaEvent_t e = c->pool.back();
    c->pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

// RAII bracket of one library call when the context is profiling
struct Prof {
  lgd_ctx* c;
  cudaStream_t s;
  cudaEvent_t e1 = nullptr;
  const char* name;
  Prof(lgd_ctx* c_, const char* n, cudaStream_t s_) : c(c_), s(s_), name(n) {
    if (!c->profiling) return;
    std::lock_guard<std::mutex> lk(c->mu);
    cudaEvent_t e0 = timing_event(c);
    e1 = timing_event(c);
    cudaEventRecord(e0, s);
    c->recs.push_back({name, e0, e1});
  }
  ~Prof() {
    if (e1) cudaEventRecord(e1, s);
  }
};

#define RUN(ctx, stream, fn, ...)                          \
  do {                                                     \
    int _rc;                                               \
    {                                                      \
      Prof _p((ctx), #fn, (cudaStream_t)(stream));         \
      _rc = fn(__VA_ARGS__, (void*)(stream));              \
    }                                                      \
    if (_rc != LGD_OK) return _rc;                         \
  } while (0)

// uploads + launches a token program through the context's pinned staging ring
static int launch_program(lgd_ctx* ctx, TokenProgram& prog, const char* name, void* dev_prog, size_t dev_bytes,
                          cudaStream_t s) {
  LGD_CHECK_ARG(TokenProgram::device_bytes((int)prog.size()) <= lgd_ctx::SLOT_BYTES, "token program too long (%zu ops)",
                prog.size());
  int slot;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    slot = ctx->next_slot;
    ctx->next_slot = (ctx->next_slot + 1) % lgd_ctx::SLOTS;
  }
  if (ctx->slot_done[slot] == nullptr)
    LGD_CUDA(cudaEventCreateWithFlags(&ctx->slot_done[slot], cudaEventDisableTiming));
  else
    LGD_CUDA(cudaEventSynchronize(ctx->slot_done[slot]));   // eight programs ago: long finished
  int rc;
  {
    Prof p(ctx, name, s);
    rc = prog.launch(dev_prog, dev_bytes, ctx->pinned + (size_t)slot * lgd_ctx::SLOT_BYTES, s);
  }
  if (rc != LGD_OK) return rc;
  LGD_CUDA(cudaEventRecord(ctx->slot_done[slot], s));
  return LGD_OK;
}

// ------------------------------------------------------------------------------------------------ arena
struct Arena {
  char* base;
  size_t cap;
  size_t off = 0;
  bool overflow = false;
  Arena(void* b, size_t c) : base(static_cast<char*>(b)), cap(c) {}
  template <typename T>
  T* take(size_t n) {
    const size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += bytes;
    if (base && off > cap) overflow = true;
    return p;
  }
};

// ------------------------------------------------------------------------------------------------ parameters
// Order of the parameter / gradient pointer arrays of the teacher entry points (names relative to "teacher.",
// the reference's state_dict names: SURVEY.md 8(b)).
enum TP {
  SD_C1W, SD_C1B, SD_C2W, SD_C2B, SD_C3W, SD_C3B, SD_F1W, SD_F1B, SD_F2W, SD_F2B, SD_F3W, SD_F3B,   // stn_desc
  SF_C1W, SF_C1B, SF_C2W, SF_C2B, SF_C3W, SF_C3B, SF_F1W, SF_F1B, SF_F2W, SF_F2B, SF_F3W, SF_F3B,   // stn_feat
  LE_C1W, LE_C1B, LE_C2W, LE_C2B, LE_C3W, LE_C3B, LE_C4W, LE_C4B,                                   // label encoder
  CANONI_W, CANONI_B, SPROJ_W, SPROJ_B, LINST2D_W, LINST2D_B, GCTX_W, GCTX_B, LINST1D_W, LINST1D_B,
  REF0_W, REF0_B, REF3_W, REF3_B, REF6_W, REF6_B, MHA_INW, MHA_INB, MHA_OUTW, MHA_OUTB,
  TP_COUNT
};
static const char* const TP_NAMES[TP_COUNT] = {
    "label_encoder_.stn_desc.conv1.weight", "label_encoder_.stn_desc.conv1.bias",
    "label_encoder_.stn_desc.conv2.weight", "label_encoder_.stn_desc.conv2.bias",
    "label_encoder_.stn_desc.conv3.weight", "label_encoder_.stn_desc.conv3.bias",
    "label_encoder_.stn_desc.fc1.weight", "label_encoder_.stn_desc.fc1.bias",
    "label_encoder_.stn_desc.fc2.weight", "label_encoder_.stn_desc.fc2.bias",
    "label_encoder_.stn_desc.fc3.weight", "label_encoder_.stn_desc.fc3.bias",
    "label_encoder_.stn_feat.conv1.weight", "label_encoder_.stn_feat.conv1.bias",
    "label_encoder_.stn_feat.conv2.weight", "label_encoder_.stn_feat.conv2.bias",
    "label_encoder_.stn_feat.conv3.weight", "label_encoder_.stn_feat.conv3.bias",
    "label_encoder_.stn_feat.fc1.weight", "label_encoder_.stn_feat.fc1.bias",
    "label_encoder_.stn_feat.fc2.weight", "label_encoder_.stn_feat.fc2.bias",
    "label_encoder_.stn_feat.fc3.weight", "label_encoder_.stn_feat.fc3.bias",
    "label_encoder_.conv1.weight", "label_encoder_.conv1.bias", "label_encoder_.conv2.weight",
    "label_encoder_.conv2.bias", "label_encoder_.conv3.weight", "label_encoder_.conv3.bias",
    "label_encoder_.conv4.weight", "label_encoder_.conv4.bias",
    "canoni_proj_1D.0.0.weight", "canoni_proj_1D.0.0.bias", "student_proj_2D.0.0.weight", "student_proj_2D.0.0.bias",
    "local_inst_proj_2D.weight", "local_inst_proj_2D.bias", "global_ctx_proj_1D.weight", "global_ctx_proj_1D.bias",
    "local_inst_proj_1D.weight", "local_inst_proj_1D.bias", "refinement_module.0.weight", "refinement_module.0.bias",
    "refinement_module.3.weight", "refinement_module.3.bias", "refinement_module.6.weight", "refinement_module.6.bias",
    "multi_head_attn.in_proj_weight", "multi_head_attn.in_proj_bias", "multi_head_attn.out_proj.weight",
    "multi_head_attn.out_proj.bias"};

enum AP { AD0_W, AD0_B, AD2_W, AD2_B, AD4_W, AD4_B, AP_COUNT };   // relative to "adapter.distill."
static const char* const AP_NAMES[AP_COUNT] = {"adapter.0.weight", "adapter.0.bias", "adapter.2.weight",
                                               "adapter.2.bias", "adapter.4.weight", "adapter.4.bias"};

// ------------------------------------------------------------------------------------------------ dimensions
struct Dims {
  lgd_pyramid_t pyr;
  Pyr p;
  int B, F, T, heads, max_n, img_h, img_w, ctx;
  long long E;        // elements of one pyramid buffer
  int P;              // pixels of one image over the pyramid
  int num_tiles;
  size_t ws_bytes;    // per-stream kernel workspace (max over the kernels' needs)
};

static int make_dims(const lgd_step_desc_t* d, Dims* o) {
  LGD_CHECK_ARG(d != nullptr, "null step descriptor");
  o->pyr = d->pyr;
  int rc = make_pyr(&d->pyr, &o->p);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(d->T > 0 && d->heads > 0 && LGD_CHANNELS % d->heads == 0 && d->max_n > 0 && d->img_h > 0 && d->img_w > 0,
                "bad step descriptor (T %d, heads %d, max_n %d, image %dx%d)", d->T, d->heads, d->max_n, d->img_h,
                d->img_w);
  o->B = d->pyr.batch;
  o->F = d->pyr.num_levels;
  o->T = d->T;
  o->heads = d->heads;
  o->max_n = d->max_n;
  o->img_h = d->img_h;
  o->img_w = d->img_w;
  o->ctx = d->add_context_box;
  o->E = lgd_pyramid_elems(&d->pyr);
  o->P = o->p.pix_start[o->F];
  o->num_tiles = lgd_conv3x3_num_tiles(&d->pyr);
  size_t ws = lgd_conv3x3_wgrad_workspace(&d->pyr);
  auto mx = [&](size_t v) { if (v > ws) ws = v; };
  mx(lgd_in_workspace(&d->pyr));
  mx(lgd_conv3x3_fwd_workspace(&d->pyr));
  mx(lgd_gn_apply_workspace(&d->pyr));
  mx(lgd_gn_bwd_workspace(&d->pyr));
  mx(lgd_channel_sums_workspace(&d->pyr));
  mx(lgd_maskpool_workspace(&d->pyr, d->T));
  o->ws_bytes = ws;
  return LGD_OK;
}

// box table blob (device int32): boxes (T,4) fp32 | labels T | img_of T | img_start B+1 | n_render B | ctx_row B
struct BoxTable {
  const float* boxes;
  const int32_t *labels, *img_of, *img_start, *n_render, *ctx_row;
};
static BoxTable box_table(const int32_t* blob, int T, int B) {
  BoxTable t;
  t.boxes = reinterpret_cast<const float*>(blob);
  t.labels = blob + 4 * T;
  t.img_of = t.labels + T;
  t.img_start = t.img_of + T;
  t.n_render = t.img_start + B + 1;
  t.ctx_row = t.n_render + B;
  return t;
}

// ------------------------------------------------------------------------------------------------ small-T units
// Linear -> LayerNorm(no affine) -> ReLU  (spatial_transformer.py:31-39, label_encoder.py:243-270, layers.py:9-19)
struct Unit {
  int K, N;
  bool norm;
  int wi, bi;                  // parameter indices
  const float* x = nullptr;    // input (M,K), owned by the producer
  float *pre = nullptr, *y = nullptr, *mean = nullptr, *rstd = nullptr;
  void layout(Arena& a, int M) {
    pre = a.take<float>((size_t)M * N);
    if (norm) {
      y = a.take<float>((size_t)M * N);
      mean = a.take<float>(M);
      rstd = a.take<float>(M);
    } else {
      y = pre;
    }
  }
};

struct LabelTape {
  Unit sd[6], sf[6], c1, c2, c3, c4, canoni;
  float *desc, *x1, *x_ft, *cat;
  int32_t* argmax;
};

static void stn_units(Unit* u, int k, int base) {
  const int dims[7] = {k, 64, 128, 1024, 512, 256, k * k};
  for (int i = 0; i < 6; ++i) {
    u[i].K = dims[i];
    u[i].N = dims[i + 1];
    u[i].norm = i < 5;
    u[i].wi = base + 2 * i;
    u[i].bi = base + 2 * i + 1;
  }
}

static void layout_label(Arena& a, const Dims& d, LabelTape* L) {
  const int T = d.T;
  stn_units(L->sd, DESC, SD_C1W);
  stn_units(L->sf, 64, SF_C1W);
  L->c1 = {DESC, 64, true, LE_C1W, LE_C1B};
  L->c2 = {64, 128, true, LE_C2W, LE_C2B};
  L->c3 = {128, 1024, true, LE_C3W, LE_C3B};
  L->c4 = {1088, 256, true, LE_C4W, LE_C4B};
  L->canoni = {256, 256, true, CANONI_W, CANONI_B};
  L->desc = a.take<float>((size_t)T * DESC);
  for (int i = 0; i < 6; ++i) L->sd[i].layout(a, T);
  L->x1 = a.take<float>((size_t)T * DESC);
  L->c1.layout(a, T);
  for (int i = 0; i < 6; ++i) L->sf[i].layout(a, T);
  L->x_ft = a.take<float>((size_t)T * 64);
  L->c2.layout(a, T);
  L->c3.layout(a, T);
  L->cat = a.take<float>((size_t)T * 1088);
  L->argmax = a.take<int32_t>((size_t)d.B * 1024);
  L->c4.layout(a, T);
  L->canoni.layout(a, T);
}

struct Exec {
  lgd_ctx* ctx;
  cudaStream_t s;
  void* lin_ws;
};

static int unit_fwd(const Exec& e, Unit& u, const float* x, int M, const float* const* P) {
  u.x = x;
  RUN(e.ctx, e.s, lgd_linear_fwd, x, u.K, P[u.wi], u.K, P[u.bi], u.pre, u.N, M, u.N, u.K, e.lin_ws, LIN_WS_BYTES);
  if (u.norm) RUN(e.ctx, e.s, lgd_layernorm_fwd, u.pre, u.y, u.mean, u.rstd, M, u.N, 1);
  return LGD_OK;
}

// gy: gradient w.r.t. the unit's output (clobbered when the unit has a norm: the LayerNorm backward runs in place);
// gx (optional): (M,K) gradient w.r.t. the unit's input
static int unit_bwd(const Exec& e, const Unit& u, float* gy, const float* x, int M, const float* const* P,
                    float* const* G, float* gx) {
  if (u.norm) RUN(e.ctx, e.s, lgd_layernorm_bwd, gy, u.pre, u.mean, u.rstd, gy, M, u.N, 1);
  RUN(e.ctx, e.s, lgd_linear_bwd_weight, gy, u.N, x, u.K, G[u.wi], u.K, G[u.bi], M, u.N, u.K, 0, e.lin_ws, LIN_WS_BYTES);
  if (gx != nullptr)
    RUN(e.ctx, e.s, lgd_linear_bwd_input, gy, u.N, P[u.wi], u.K, gx, u.K, M, u.N, u.K, 0, e.lin_ws, LIN_WS_BYTES);
  return LGD_OK;
}

// the same two building blocks as ops of a token program
static void unit_fwd_prog(TokenProgram& pg, Unit& u, const float* x, int M, const float* const* P) {
  u.x = x;
  if (u.norm) {   // GEMM, then ONE stage that reduces the split-K partials and normalises
    pg.linear_layernorm(x, u.K, P[u.wi], u.K, P[u.bi], u.pre, u.y, u.mean, u.rstd, M, u.N, u.K, 1);
    return;
  }
  pg.linear(x, u.K, P[u.wi], u.K, P[u.bi], u.pre, u.N, M, u.N, u.K);
  pg.next_stage();
}
static void unit_bwd_prog(TokenProgram& pg, const Unit& u, float* gy, const float* x, int M, const float* const* P,
                          float* const* G, float* gx) {
  if (u.norm) {
    pg.layernorm_bwd(gy, u.pre, u.mean, u.rstd, gy, M, u.N, 1);
    pg.next_stage();
  }
  pg.linear_bwd_weight(gy, u.N, x, u.K, G[u.wi], u.K, G[u.bi], M, u.N, u.K);
  if (gx != nullptr) pg.linear_bwd_input(gy, u.N, P[u.wi], u.K, gx, u.K, M, u.N, u.K, 0);
  pg.next_stage();
}
static void stn_bwd_prog(TokenProgram& pg, const Unit* u, float* g, int M, const float* const* P, float* const* G,
                         Arena& a, float* gx_out) {
  float* cur = g;
  for (int i = 5; i >= 0; --i) {
    float* gx = (i > 0) ? a.take<float>((size_t)M * u[i].K) : gx_out;
    unit_bwd_prog(pg, u[i], cur, u[i].x, M, P, G, gx);
    cur = gx;
  }
}

static int stn_fwd(const Exec& e, Unit* u, const float* x, int M, const float* const* P) {
  for (int i = 0; i < 6; ++i) {
    int rc = unit_fwd(e, u[i], x, M, P);
    if (rc != LGD_OK) return rc;
    x = u[i].y;
  }
  return LGD_OK;
}

// g: gradient w.r.t. the (M, k*k) transform (clobbered). gx_out (optional): gradient w.r.t. the STN input.
static int stn_bwd(const Exec& e, const Unit* u, float* g, int M, const float* const* P, float* const* G, Arena& a,
                   float* gx_out) {
  float* cur = g;
  for (int i = 5; i >= 0; --i) {
    float* gx = (i > 0) ? a.take<float>((size_t)M * u[i].K) : gx_out;
    int rc = unit_bwd(e, u[i], cur, u[i].x, M, P, G, gx);
    if (rc != LGD_OK) return rc;
    cur = gx;
  }
  return LGD_OK;
}

// ------------------------------------------------------------------------------------------------ teacher tape
// packed fp16 weights of a chain's convolutions (forward and dgrad layouts) + gains: written once per step by the
// forward (one launch for all of them), read by the backward
struct PackedConvs {
  __half* fwd[8];
  __half* dgrad[8];
  float* gains;
  int n;
};
static void layout_packed(Arena& a, int n, PackedConvs* p) {
  p->n = n;
  for (int i = 0; i < n; ++i) p->fwd[i] = a.take<__half>((size_t)9 * C * C);
  for (int i = 0; i < n; ++i) p->dgrad[i] = a.take<__half>((size_t)9 * C * C);
  p->gains = a.take<float>(8);
}
static int pack_all(lgd_ctx* ctx, cudaStream_t s, const PackedConvs& p, const float* const* w, void* ws) {
  void* fwd[8];
  void* dg[8];
  for (int i = 0; i < p.n; ++i) {
    fwd[i] = p.fwd[i];
    dg[i] = p.dgrad[i];
  }
  RUN(ctx, s, lgd_pack_conv_weights_f16_multi, w, p.n, fwd, dg, p.gains, ws, (size_t)p.n * 9 * 64 * sizeof(float));
  return LGD_OK;
}
enum TConv { TC_SPROJ, TC_LINST, TC_REF0, TC_REF3, TC_REF6, TC_COUNT };
enum AConv { AC_0, AC_2, AC_4, AC_COUNT };

struct TeacherTape {
  PackedConvs pk;
  int32_t* ranges;
  LabelTape L;
  float *sp_raw, *sp_stats, *pooled, *q, *k, *v, *att, *probs, *a;
  __half *rend_h, *y0_h, *y1_h, *y2_h;
  float *r0, *st0, *r1, *st1, *r2, *st2;
};
static_assert(sizeof(__half) == 2, "fp16");

static void layout_teacher_tape(Arena& a, const Dims& d, TeacherTape* t) {
  const size_t FT = (size_t)d.F * d.T, E = (size_t)d.E;
  layout_packed(a, TC_COUNT, &t->pk);
  t->ranges = a.take<int32_t>(FT * 4);
  layout_label(a, d, &t->L);
  t->sp_raw = a.take<float>(E);
  t->sp_stats = a.take<float>((size_t)d.F * d.B * 2);
  t->pooled = a.take<float>(FT * C);
  t->q = a.take<float>(FT * C);
  t->k = a.take<float>((size_t)d.T * C);
  t->v = a.take<float>((size_t)d.T * C);
  t->att = a.take<float>(FT * C);
  t->probs = a.take<float>((size_t)d.F * d.heads * d.T * d.max_n);
  t->a = a.take<float>(FT * C);
  t->rend_h = a.take<__half>(E);
  t->y0_h = a.take<__half>(E);
  t->r0 = a.take<float>(E);
  t->st0 = a.take<float>((size_t)d.F * d.B * 2);
  t->y1_h = a.take<__half>(E);
  t->r1 = a.take<float>(E);
  t->st1 = a.take<float>((size_t)d.F * d.B * 2);
  t->y2_h = a.take<__half>(E);
  t->r2 = a.take<float>(E);
  t->st2 = a.take<float>((size_t)d.F * d.B * 2);
}

// conv + GroupNorm statistics
static int conv_stats(const Exec& e, const Dims& d, const __half* x_h, const __half* w_h, const float* bias, float* out,
                      float* tile_stats, float* stats) {
  RUN(e.ctx, e.s, lgd_conv3x3_fwd_f16, &d.pyr, x_h, w_h, bias, 0, 0, out, nullptr, 0, 0, tile_stats);
  RUN(e.ctx, e.s, lgd_gn_finalize, &d.pyr, tile_stats, stats);
  return LGD_OK;
}

struct DgradW {
  __half* w;
  float* gain;
};
// the weight gradient of one convolution on the wgrad stream (or in line when side streams are off)
struct Wgrad {
  lgd_ctx* ctx;
  cudaStream_t main;
  const Dims* d;
  void* ws;        // lgd_conv3x3_wgrad_workspace() bytes, used by the wgrad stream only
  bool used = false;
  int run(const __half* x_h, const __half* g_h, const float* scale3, float* gw, Arena& a) {
    cudaStream_t s = main;
    if (ctx->side_streams && !ctx->profiling) {
      LGD_CUDA(cudaEventRecord(ctx->ev_ready, main));
      LGD_CUDA(cudaStreamWaitEvent(ctx->wgrad_stream, ctx->ev_ready, 0));
      s = ctx->wgrad_stream;
      used = true;
    }
    float* packed = a.take<float>((size_t)9 * C * C);
    RUN(ctx, s, lgd_conv3x3_wgrad_f16, &d->pyr, x_h, g_h, scale3 + 1, packed, ws, lgd_conv3x3_wgrad_workspace(&d->pyr));
    RUN(ctx, s, lgd_unpack_conv_wgrad, packed, gw, 0);
    return LGD_OK;
  }
  int join() {
    if (used) {
      LGD_CUDA(cudaEventRecord(ctx->ev_join, ctx->wgrad_stream));
      LGD_CUDA(cudaStreamWaitEvent(main, ctx->ev_join, 0));
      used = false;
    }
    return LGD_OK;
  }
};

// fp16 dgrad. relu_mask_h: the layer below is conv+ReLU -> masked output + channel sums (its bias gradient).
// want_half: the output feeds another dgrad -> only the scaled fp16 copy (out32 = nullptr) with a measured norm.
struct DgradOut {
  float* out32 = nullptr;
  __half* out_h = nullptr;
  float* scale3 = nullptr;   // {s, 1/s, U_measured} of out_h
  float *sums = nullptr, *total = nullptr;
};
static int dgrad(const Exec& e, const Dims& d, Arena& a, void* ws, const __half* g_h, const float* sc_in,
                 const DgradW& w, const __half* relu_mask_h, bool want_half, float* out32_buf, DgradOut* o) {
  const bool csum = relu_mask_h != nullptr;
  if (csum) {
    o->sums = a.take<float>((size_t)d.F * d.B * C);
    o->total = a.take<float>(C);
  }
  float* tile_stats = nullptr;
  float* sc_out = nullptr;
  if (want_half) {
    o->out_h = a.take<__half>((size_t)d.E);
    sc_out = a.take<float>(3);
    tile_stats = a.take<float>((size_t)d.num_tiles * 2);
    RUN(e.ctx, e.s, lgd_grad_scale, nullptr, 0, 1, w.gain, sc_in + 2, 1.0f, sc_out);
  } else {
    o->out32 = out32_buf;
  }
  RUN(e.ctx, e.s, lgd_conv3x3_dgrad_f16, &d.pyr, g_h, w.w, sc_in + 1, o->out32, 0, nullptr, relu_mask_h,
      o->out_h, sc_out, tile_stats, o->sums, o->total, csum ? ws : nullptr, csum ? d.ws_bytes : 0);
  if (want_half) {
    float* meas = a.take<float>(3);
    RUN(e.ctx, e.s, lgd_grad_scale, tile_stats + 1, d.num_tiles, 2, nullptr, nullptr, 1.0f, meas);
    // {s, 1/s} the copy was written with, U = the measured norm (so that bounds never compound along a chain)
    o->scale3 = a.take<float>(3);
    LGD_CUDA(cudaMemcpyAsync(o->scale3, sc_out, 2 * sizeof(float), cudaMemcpyDeviceToDevice, e.s));
    LGD_CUDA(cudaMemcpyAsync(o->scale3 + 2, meas + 2, sizeof(float), cudaMemcpyDeviceToDevice, e.s));
  }
  return LGD_OK;
}

}  // namespace lgd

// ================================================================================================ queries
extern "C" int lgd_teacher_param_count(void) { return TP_COUNT; }
extern "C" const char* lgd_teacher_param_name(int i) { return (i >= 0 && i < TP_COUNT) ? TP_NAMES[i] : nullptr; }
extern "C" int lgd_adapter_param_count(void) { return AP_COUNT; }
extern "C" const char* lgd_adapter_param_name(int i) { return (i >= 0 && i < AP_COUNT) ? AP_NAMES[i] : nullptr; }

namespace lgd {
struct Field {
  const char* name;
  const void* ptr;
  size_t bytes;
};
static int find_field(const Field* f, int n, const char* name, size_t* offset, size_t* bytes) {
  for (int i = 0; i < n; ++i)
    if (strcmp(f[i].name, name) == 0) {
      *offset = reinterpret_cast<size_t>(f[i].ptr) - 4096;
      *bytes = f[i].bytes;
      return LGD_OK;
    }
  set_error("unknown tape field '%s'", name);
  return LGD_EINVAL;
}
}  // namespace lgd

/* diagnostics: byte offset and size of a named tensor inside the teacher / distillation tape */
extern "C" int lgd_teacher_tape_field(const lgd_step_desc_t* desc, const char* name, size_t* offset, size_t* bytes) {
  Dims d;
  int rc = make_dims(desc, &d);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(name && offset && bytes, "lgd_teacher_tape_field: null pointer");
  Arena a(reinterpret_cast<void*>(4096), ~size_t(0) >> 1);   // fake base: pointers become offsets + 4096
  TeacherTape t;
  layout_teacher_tape(a, d, &t);
  const size_t E = (size_t)d.E, FT = (size_t)d.F * d.T, FB2 = (size_t)d.F * d.B * 2 * 4;
  const Field f[] = {{"ranges", t.ranges, FT * 16}, {"desc", t.L.desc, (size_t)d.T * DESC * 4},
                     {"label_embed", t.L.c4.y, (size_t)d.T * C * 4}, {"canoni", t.L.canoni.y, (size_t)d.T * C * 4},
                     {"sp_raw", t.sp_raw, E * 4}, {"sp_stats", t.sp_stats, FB2}, {"pooled", t.pooled, FT * C * 4},
                     {"a", t.a, FT * C * 4}, {"rend_h", t.rend_h, E * 2}, {"y0_h", t.y0_h, E * 2},
                     {"r0", t.r0, E * 4}, {"st0", t.st0, FB2}, {"y1_h", t.y1_h, E * 2}, {"r1", t.r1, E * 4},
                     {"st1", t.st1, FB2}, {"y2_h", t.y2_h, E * 2}, {"r2", t.r2, E * 4}, {"st2", t.st2, FB2}};
  return find_field(f, (int)(sizeof(f) / sizeof(f[0])), name, offset, bytes);
}

extern "C" size_t lgd_teacher_tape_bytes(const lgd_step_desc_t* desc) {
  Dims d;
  if (make_dims(desc, &d) != LGD_OK) return 0;
  Arena a(nullptr, 0);
  TeacherTape t;
  layout_teacher_tape(a, d, &t);
  return a.off;
}

static size_t teacher_fwd_scratch(const Dims& d) {
  const size_t FT = (size_t)d.F * d.T;
  size_t n = 0;
  auto add = [&](size_t bytes) { n += (bytes + 255) & ~size_t(255); };
  add(LIN_WS_BYTES);
  add(d.ws_bytes);
  add((size_t)d.num_tiles * 2 * 4);
  add(FT * C * 4);              // inst
  add(FT * C * 4);              // ctx vectors
  add((size_t)d.F * d.B * C * 4);   // bias table
  add(4 * LIN_WS_BYTES);            // split-K arenas of the label program (two slices each: K / V projections)
  add(lgd_ctx::SLOT_BYTES);         // its op list
  add(4 * LIN_WS_BYTES);            // split-K arenas of the relation program
  add(lgd_ctx::SLOT_BYTES);
  if (d.max_n <= LGD_TAP_MAX_ROWS) add(lgd_tap_render_workspace(&d.pyr, d.T, 0));   // tap rendering (when enabled)
  return n + 4096;
}

static size_t teacher_bwd_scratch(const Dims& d) {
  const size_t FT = (size_t)d.F * d.T, E = (size_t)d.E, T = d.T;
  size_t n = 0;
  auto add = [&](size_t bytes) { n += (bytes + 255) & ~size_t(255); };
  add(2 * LIN_WS_BYTES);
  add(2 * d.ws_bytes);
  add(2 * E * 4);                       // fp32 ping-pong gradient pyramids
  add(5 * E * 2 + 5 * 256);             // scaled fp16 gradient pyramids (alive until the wgrads are joined)
  add(5 * ((size_t)9 * C * C * 2 + 256 + 9 * C * 4 + 256));   // dgrad weights + gains
  add(5 * (size_t)9 * C * C * 4);       // packed weight gradients
  add(8 * ((size_t)d.F * d.B * C * 4 + C * 4 + 512));          // channel sums, totals
  add(4 * ((size_t)d.num_tiles * 2 * 4 + 1024));
  add(2 * ((size_t)d.num_tiles * 4 * 4 + 256));   // per-tile GroupNorm-backward sums
  add(12 * FT * C * 4);                 // token-sized gradients of the relation / rendering block
  add((size_t)d.F * d.heads * T * d.max_n * 4);   // attention score gradients
  // label side: every activation gradient once (sum of the unit widths) + the two transform gradients
  add(T * (size_t)(2 * (DESC * DESC + 64 * 64) + 4 * (1088 + 1024 + 512 + 256 + 128 + 64 + DESC) + 8 * 1024) * 4);
  add(64 * 512);
  add(4 * LIN_WS_BYTES);            // split-K arenas of the label-side program (two slices each)
  add(lgd_ctx::SLOT_BYTES);
  add(2 * (4 * LIN_WS_BYTES + lgd_ctx::SLOT_BYTES));   // the two relation programs of the backward
  if (d.max_n <= LGD_TAP_MAX_ROWS) add(lgd_tap_render_workspace(&d.pyr, d.T, 1) + FT * C * 4);   // tap rendering
  return n + 8192;
}

extern "C" size_t lgd_teacher_scratch_bytes(const lgd_step_desc_t* desc, int backward) {
  Dims d;
  if (make_dims(desc, &d) != LGD_OK) return 0;
  return backward ? teacher_bwd_scratch(d) : teacher_fwd_scratch(d);
}

// ================================================================================================ teacher forward
extern "C" int lgd_teacher_forward(lgd_ctx_t* ctx, const lgd_step_desc_t* desc, const int32_t* box_blob,
                                   const void* stu_half, const float* const* params_host, float* tea, float* masks,
                                   void* tape, size_t tape_bytes, void* scratch, size_t scratch_bytes, void* stream) {
  LGD_CHECK_ARG(ctx && box_blob && stu_half && params_host && tea && tape && scratch, "lgd_teacher_forward: null pointer");
  Dims d;
  int rc = make_dims(desc, &d);
  if (rc != LGD_OK) return rc;
  for (int i = 0; i < TP_COUNT; ++i)
    LGD_CHECK_ARG(params_host[i] != nullptr || (!d.ctx && (i == GCTX_W || i == GCTX_B)),
                  "lgd_teacher_forward: parameter %s is missing", TP_NAMES[i]);
  const float* const* P = params_host;
  Arena ta(tape, tape_bytes), sa(scratch, scratch_bytes);
  TeacherTape t;
  layout_teacher_tape(ta, d, &t);
  LGD_CHECK_ARG(!ta.overflow, "lgd_teacher_forward: tape too small (%zu < %zu)", tape_bytes, ta.off);
  LGD_CHECK_ARG(scratch_bytes >= teacher_fwd_scratch(d), "lgd_teacher_forward: scratch too small");
  cudaStream_t s = (cudaStream_t)stream;
  Exec e{ctx, s, sa.take<char>(LIN_WS_BYTES)};
  void* ws = sa.take<char>(d.ws_bytes);
  float* tile_stats = sa.take<float>((size_t)d.num_tiles * 2);
  const BoxTable tb = box_table(box_blob, d.T, d.B);
  const int T = d.T, F = d.F, B = d.B;
  const __half* stu_h = static_cast<const __half*>(stu_half);

  // every convolution weight of the chain, both layouts, one launch (kept in the tape for the backward)
  {
    const float* w5[TC_COUNT] = {P[SPROJ_W], P[LINST2D_W], P[REF0_W], P[REF3_W], P[REF6_W]};
    if ((rc = pack_all(ctx, s, t.pk, w5, ws)) != LGD_OK) return rc;
  }
  // a4: exact membership intervals (+ the reference's float masks when the caller wants them)
  RUN(ctx, s, lgd_box_ranges, tb.boxes, T, d.img_h, d.img_w, &d.pyr, t.ranges);
  if (masks != nullptr) RUN(ctx, s, lgd_masks_from_ranges, t.ranges, T, &d.pyr, masks);

  // a1 + a2: descriptors, label encoder, canonical projection
  LabelTape& L = t.L;
  RUN(ctx, s, lgd_encode_descriptors, tb.boxes, tb.labels, T, d.img_h, d.img_w, L.desc);
  if (ctx->token_programs) {
    // label encoder + canonical projection: ONE persistent kernel (tokenprog.cu) instead of ~50 dependent launches
    TokenProgram pg(sa.take<char>(2 * LIN_WS_BYTES), sa.take<char>(2 * LIN_WS_BYTES), 2 * LIN_WS_BYTES, LIN_WS_BYTES);
    const float* x = L.desc;
    for (int i = 0; i < 6; ++i) { unit_fwd_prog(pg, L.sd[i], x, T, P); x = L.sd[i].y; }
    pg.rowvec_fwd(L.desc, L.sd[5].y, L.x1, T, DESC);
    pg.next_stage();
    unit_fwd_prog(pg, L.c1, L.x1, T, P);
    x = L.c1.y;
    for (int i = 0; i < 6; ++i) { unit_fwd_prog(pg, L.sf[i], x, T, P); x = L.sf[i].y; }
    pg.rowvec_fwd(L.c1.y, L.sf[5].y, L.x_ft, T, 64);
    pg.next_stage();
    unit_fwd_prog(pg, L.c2, L.x_ft, T, P);
    unit_fwd_prog(pg, L.c3, L.c2.y, T, P);
    pg.segmax_fwd(L.x_ft, 64, L.c3.y, 1024, tb.img_start, B, L.cat, L.argmax);
    pg.next_stage();
    unit_fwd_prog(pg, L.c4, L.cat, T, P);
    unit_fwd_prog(pg, L.canoni, L.c4.y, T, P);
    // a6: the key / value projections of the relation block only need the canonical label embeddings
    pg.linear(L.canoni.y, C, P[MHA_INW] + C * C, C, P[MHA_INB] + C, t.k, C, T, C, C);
    pg.linear(L.canoni.y, C, P[MHA_INW] + 2 * C * C, C, P[MHA_INB] + 2 * C, t.v, C, T, C, C);
    pg.next_stage();
    const size_t pbytes = TokenProgram::device_bytes((int)pg.size() + 8);
    if ((rc = launch_program(ctx, pg, "lgd_label_program_fwd", sa.take<char>(pbytes), pbytes, s)) != LGD_OK) return rc;
  } else {
    if ((rc = stn_fwd(e, L.sd, L.desc, T, P)) != LGD_OK) return rc;
    RUN(ctx, s, lgd_rowvec_matmul_fwd, L.desc, L.sd[5].y, L.x1, T, DESC);
    if ((rc = unit_fwd(e, L.c1, L.x1, T, P)) != LGD_OK) return rc;
    if ((rc = stn_fwd(e, L.sf, L.c1.y, T, P)) != LGD_OK) return rc;
    RUN(ctx, s, lgd_rowvec_matmul_fwd, L.c1.y, L.sf[5].y, L.x_ft, T, 64);
    if ((rc = unit_fwd(e, L.c2, L.x_ft, T, P)) != LGD_OK) return rc;
    if ((rc = unit_fwd(e, L.c3, L.c2.y, T, P)) != LGD_OK) return rc;
    RUN(ctx, s, lgd_segmax_concat_fwd, L.x_ft, 64, L.c3.y, 1024, tb.img_start, B, L.cat, L.argmax);
    if ((rc = unit_fwd(e, L.c4, L.cat, T, P)) != LGD_OK) return rc;
    if ((rc = unit_fwd(e, L.canoni, L.c4.y, T, P)) != LGD_OK) return rc;
  }
  const float* canoni = L.canoni.y;

  // a3: student_proj_2D = conv + GN(1) + ReLU; the normalised map is applied inside the pooling, never written
  if ((rc = conv_stats(e, d, stu_h, t.pk.fwd[TC_SPROJ], P[SPROJ_B], t.sp_raw, tile_stats, t.sp_stats)) != LGD_OK) return rc;
  // a5: mask average pooling -> appearance embeddings (F,T,256)
  RUN(ctx, s, lgd_maskpool_fwd, &d.pyr, t.sp_raw, t.sp_stats, t.ranges, tb.img_of, T, t.pooled, ws, d.ws_bytes);

  // a6: inter-object relation adaptation (stuGuided: queries = appearance embeddings, keys = values = label side)
  const float *Wi = P[MHA_INW], *bi = P[MHA_INB];
  RUN(ctx, s, lgd_linear_fwd, t.pooled, C, Wi, C, bi, t.q, C, F * T, C, C, e.lin_ws, LIN_WS_BYTES);
  if (!ctx->token_programs) {   // (with token programs the label program has produced K and V)
    RUN(ctx, s, lgd_linear_fwd, canoni, C, Wi + C * C, C, bi + C, t.k, C, T, C, C, e.lin_ws, LIN_WS_BYTES);
    RUN(ctx, s, lgd_linear_fwd, canoni, C, Wi + 2 * C * C, C, bi + 2 * C, t.v, C, T, C, C, e.lin_ws, LIN_WS_BYTES);
  }
  RUN(ctx, s, lgd_attention_fwd, t.q, F, t.k, t.v, 1, F, T, d.heads, C, tb.img_of, tb.img_start, d.max_n, t.att, t.probs);

  // a6 out projection + a7 1-D projections (intra-object knowledge mapping)
  float* inst = sa.take<float>((size_t)F * T * C);
  float* ctxv = d.ctx ? sa.take<float>((size_t)F * T * C) : nullptr;
  if (ctx->token_programs) {   // one persistent kernel: out projection, then the instance and context projections side by side
    TokenProgram pg(sa.take<char>(2 * LIN_WS_BYTES), sa.take<char>(2 * LIN_WS_BYTES), 2 * LIN_WS_BYTES, LIN_WS_BYTES);
    pg.linear(t.att, C, P[MHA_OUTW], C, P[MHA_OUTB], t.a, C, F * T, C, C);
    pg.next_stage();
    pg.linear(t.a, C, P[LINST1D_W], C, P[LINST1D_B], inst, C, F * T, C, C);
    if (d.ctx) pg.linear(t.a, C, P[GCTX_W], C, P[GCTX_B], ctxv, C, F * T, C, C);
    pg.next_stage();
    const size_t pbytes = TokenProgram::device_bytes((int)pg.size() + 8);
    if ((rc = launch_program(ctx, pg, "lgd_relation_program_fwd", sa.take<char>(pbytes), pbytes, s)) != LGD_OK) return rc;
  } else {
    RUN(ctx, s, lgd_linear_fwd, t.att, C, P[MHA_OUTW], C, P[MHA_OUTB], t.a, C, F * T, C, C, e.lin_ws, LIN_WS_BYTES);
    RUN(ctx, s, lgd_linear_fwd, t.a, C, P[LINST1D_W], C, P[LINST1D_B], inst, C, F * T, C, C, e.lin_ws, LIN_WS_BYTES);
    if (d.ctx) RUN(ctx, s, lgd_linear_fwd, t.a, C, P[GCTX_W], C, P[GCTX_B], ctxv, C, F * T, C, C, e.lin_ws, LIN_WS_BYTES);
  }
  const float* bias0 = P[LINST2D_B];
  int bias0_sl = 0, bias0_si = 0;
  if (d.ctx) {
    float* table = sa.take<float>((size_t)F * B * C);
    RUN(ctx, s, lgd_ctx_bias_table, ctxv, tb.ctx_row, P[LINST2D_B], F, B, T, table);
    bias0 = table;
    bias0_sl = B * C;
    bias0_si = C;
  }
  if (ctx->tap_render && d.max_n <= LGD_TAP_MAX_ROWS) {
    // the rendered map is piecewise constant over box rectangles: its 3x3 convolution comes from per-box tap vectors
    // (taprender.cu); neither the rendered map nor a convolution launch exists
    const size_t tbytes = lgd_tap_render_workspace(&d.pyr, T, 0);
    void* tws = sa.take<char>(tbytes);
    RUN(ctx, s, lgd_tap_render_fwd, &d.pyr, inst, P[LINST2D_W], t.ranges, tb.img_start, tb.n_render, T, d.max_n, bias0,
        bias0_sl, bias0_si, t.y0_h, nullptr, tws, tbytes);
  } else {
    RUN(ctx, s, lgd_render_fwd, &d.pyr, inst, t.ranges, tb.img_start, tb.n_render, T, nullptr, 1, t.rend_h);
    RUN(ctx, s, lgd_conv3x3_fwd_f16, &d.pyr, t.rend_h, t.pk.fwd[TC_LINST], bias0, bias0_sl, bias0_si, nullptr, t.y0_h, 1, 1,
        nullptr);
  }

  // a8: refinement module
  if ((rc = conv_stats(e, d, t.y0_h, t.pk.fwd[TC_REF0], P[REF0_B], t.r0, tile_stats, t.st0)) != LGD_OK) return rc;
  RUN(ctx, s, lgd_gn_apply, &d.pyr, t.r0, t.st0, nullptr, 1, 0, t.y1_h, nullptr, nullptr, 0);
  if ((rc = conv_stats(e, d, t.y1_h, t.pk.fwd[TC_REF3], P[REF3_B], t.r1, tile_stats, t.st1)) != LGD_OK) return rc;
  RUN(ctx, s, lgd_gn_apply, &d.pyr, t.r1, t.st1, nullptr, 1, 0, t.y2_h, nullptr, nullptr, 0);
  if ((rc = conv_stats(e, d, t.y2_h, t.pk.fwd[TC_REF6], P[REF6_B], t.r2, tile_stats, t.st2)) != LGD_OK) return rc;
  RUN(ctx, s, lgd_gn_apply, &d.pyr, t.r2, t.st2, tea, 0, 0, nullptr, nullptr, nullptr, 0);
  LGD_CHECK_ARG(!sa.overflow, "lgd_teacher_forward: scratch overflow");
  return LGD_OK;
}

// ================================================================================================ teacher backward
namespace lgd {

// GroupNorm(1)(+ReLU) backward producing the scaled fp16 operand of the dgrad in front of it and that convolution's
// bias gradient
static int gn_bwd_half(const Exec& e, const Dims& d, Arena& a, void* ws, const float* gy, const float* x,
                       const float* st, int relu, float* gbias, __half** gh, float** sc) {
  *gh = a.take<__half>((size_t)d.E);
  *sc = a.take<float>(3);
  RUN(e.ctx, e.s, lgd_gn_bwd, &d.pyr, gy, x, st, relu, nullptr, 1, *gh, *sc, nullptr, gbias, ws, d.ws_bytes);
  return LGD_OK;
}

// plain fp16 dgrad whose epilogue emits the sums of the GroupNorm backward that consumes its output, followed by that
// GroupNorm backward without its sums pass (2.5 F1 instead of 4.5 F1 of traffic)
static int dgrad_then_gn_bwd(const Exec& e, const Dims& d, Arena& a, void* ws, const __half* g_h, const float* sc_in,
                             const DgradW& w, float* out32, const float* gn_x, const float* gn_stats, int relu,
                             const __half* gn_y_h, float* gbias, __half** gh, float** sc) {
  float* tile_gn = a.take<float>((size_t)d.num_tiles * 4);
  // gn_y_h = relu(GroupNorm(gn_x)) as the fp16 operand copy of the convolution being differentiated: the sums come from it
  // (half the epilogue bytes of the fp32 gn_x, no statistics)
  if (relu && gn_y_h != nullptr)
    RUN(e.ctx, e.s, lgd_conv3x3_dgrad_f16_gnsums_y, &d.pyr, g_h, w.w, sc_in + 1, out32, gn_y_h, tile_gn);
  else
    RUN(e.ctx, e.s, lgd_conv3x3_dgrad_f16_gnsums, &d.pyr, g_h, w.w, sc_in + 1, out32, gn_x, gn_stats, relu, tile_gn);
  *gh = a.take<__half>((size_t)d.E);
  *sc = a.take<float>(3);
  RUN(e.ctx, e.s, lgd_gn_bwd_tile_sums, &d.pyr, out32, gn_x, gn_stats, relu, tile_gn, nullptr, 1, *gh, *sc, nullptr, gbias,
      ws, d.ws_bytes);
  return LGD_OK;
}

}  // namespace lgd

extern "C" int lgd_teacher_backward(lgd_ctx_t* ctx, const lgd_step_desc_t* desc, const int32_t* box_blob,
                                    const void* stu_half, const float* const* params_host,
                                    const float* const* gtea_levels_host, const float* gtea_pyramid,
                                    const void* tape, size_t tape_bytes, float* const* grads_host,
                                    float* const* gstu_levels_host, float* gstu_pyramid, int gstu_accumulate,
                                    void* wgrad_workspace, void* scratch, size_t scratch_bytes, void* stream) {
  LGD_CHECK_ARG(ctx && box_blob && stu_half && params_host && (gtea_levels_host || gtea_pyramid) && tape && grads_host &&
                    wgrad_workspace && scratch, "lgd_teacher_backward: null pointer");
  Dims d;
  int rc = make_dims(desc, &d);
  if (rc != LGD_OK) return rc;
  const float* const* P = params_host;
  float* const* G = grads_host;
  for (int i = 0; i < TP_COUNT; ++i)
    LGD_CHECK_ARG((P[i] && G[i]) || (!d.ctx && (i == GCTX_W || i == GCTX_B)),
                  "lgd_teacher_backward: parameter / gradient %s is missing", TP_NAMES[i]);
  Arena ta(const_cast<void*>(tape), tape_bytes), sa(scratch, scratch_bytes);
  TeacherTape t;
  layout_teacher_tape(ta, d, &t);
  LGD_CHECK_ARG(!ta.overflow, "lgd_teacher_backward: tape too small");
  LGD_CHECK_ARG(scratch_bytes >= teacher_bwd_scratch(d), "lgd_teacher_backward: scratch too small");
  // the label side's unit inputs are not part of the tape layout: restore them
  LabelTape& L = t.L;
  {
    const float* x = L.desc;
    for (int i = 0; i < 6; ++i) { L.sd[i].x = x; x = L.sd[i].y; }
    x = L.c1.y;
    for (int i = 0; i < 6; ++i) { L.sf[i].x = x; x = L.sf[i].y; }
    L.c1.x = L.x1; L.c2.x = L.x_ft; L.c3.x = L.c2.y; L.c4.x = L.cat; L.canoni.x = L.c4.y;
  }
  cudaStream_t s = (cudaStream_t)stream;
  Exec e{ctx, s, sa.take<char>(LIN_WS_BYTES)};
  void* lin_ws_label = sa.take<char>(LIN_WS_BYTES);
  void* ws = sa.take<char>(d.ws_bytes);
  const BoxTable tb = box_table(box_blob, d.T, d.B);
  const int T = d.T, F = d.F, B = d.B;
  const __half* stu_h = static_cast<const __half*>(stu_half);
  const size_t E = (size_t)d.E;
  float* ping[2] = {sa.take<float>(E), sa.take<float>(E)};
  Wgrad wg{ctx, s, &d, wgrad_workspace};
  const bool need_feat = gstu_levels_host != nullptr || gstu_pyramid != nullptr;

  // cotangents arrive as NCHW maps (transposed here) or already as an NHWC pyramid buffer (read in place)
  const float* g_tea = gtea_pyramid;
  if (g_tea == nullptr) {
    RUN(ctx, s, lgd_nchw_to_pyramid, gtea_levels_host, &d.pyr, ping[0], 0, nullptr);
    g_tea = ping[0];
  }

  // a8 backward
  __half* gh;
  float* sc;
  DgradW w;
  DgradOut o;
  if ((rc = gn_bwd_half(e, d, sa, ws, g_tea, t.r2, t.st2, 0, G[REF6_B], &gh, &sc)) != LGD_OK) return rc;
  if ((rc = wg.run(t.y2_h, gh, sc, G[REF6_W], sa)) != LGD_OK) return rc;
  w = DgradW{t.pk.dgrad[TC_REF6], t.pk.gains + TC_REF6};
  if (ctx->fuse_gn_sums) {
    if ((rc = dgrad_then_gn_bwd(e, d, sa, ws, gh, sc, w, ping[1], t.r1, t.st1, 1, t.y2_h, G[REF3_B], &gh, &sc)) != LGD_OK) return rc;
  } else {
    if ((rc = dgrad(e, d, sa, ws, gh, sc, w, nullptr, false, ping[1], &(o = DgradOut()))) != LGD_OK) return rc;
    if ((rc = gn_bwd_half(e, d, sa, ws, o.out32, t.r1, t.st1, 1, G[REF3_B], &gh, &sc)) != LGD_OK) return rc;
  }
  if ((rc = wg.run(t.y1_h, gh, sc, G[REF3_W], sa)) != LGD_OK) return rc;
  w = DgradW{t.pk.dgrad[TC_REF3], t.pk.gains + TC_REF3};
  if (ctx->fuse_gn_sums) {
    if ((rc = dgrad_then_gn_bwd(e, d, sa, ws, gh, sc, w, ping[0], t.r0, t.st0, 1, t.y1_h, G[REF0_B], &gh, &sc)) != LGD_OK) return rc;
  } else {
    if ((rc = dgrad(e, d, sa, ws, gh, sc, w, nullptr, false, ping[0], &(o = DgradOut()))) != LGD_OK) return rc;
    if ((rc = gn_bwd_half(e, d, sa, ws, o.out32, t.r0, t.st0, 1, G[REF0_B], &gh, &sc)) != LGD_OK) return rc;
  }
  if ((rc = wg.run(t.y0_h, gh, sc, G[REF0_W], sa)) != LGD_OK) return rc;
  w = DgradW{t.pk.dgrad[TC_REF0], t.pk.gains + TC_REF0};
  // y0 = relu(conv(rendered) + bias/ctx): the dgrad epilogue masks by y0 > 0 and yields the per-(level,image) channel
  // sums = gradient of the bias / context vector of local_inst_proj_2D
  const bool tap = ctx->tap_render && d.max_n <= LGD_TAP_MAX_ROWS;
  DgradOut r0;
  // tap rendering reads the masked gradient as fp32 (ping[1]); the convolution path as the scaled fp16 operand
  if ((rc = dgrad(e, d, sa, ws, gh, sc, w, t.y0_h, !tap, tap ? ping[1] : nullptr, &r0)) != LGD_OK) return rc;

  // a7 backward
  LGD_CUDA(cudaMemcpyAsync(G[LINST2D_B], r0.total, C * sizeof(float), cudaMemcpyDeviceToDevice, s));
  float* g_inst = sa.take<float>((size_t)F * T * C);
  if (tap) {
    // the instance embeddings are recomputed (one token-sized GEMM) instead of being kept in the tape
    float* inst = sa.take<float>((size_t)F * T * C);
    RUN(ctx, s, lgd_linear_fwd, t.a, C, P[LINST1D_W], C, P[LINST1D_B], inst, C, F * T, C, C, e.lin_ws, LIN_WS_BYTES);
    const size_t tbytes = lgd_tap_render_workspace(&d.pyr, T, 1);
    void* tws = sa.take<char>(tbytes);
    RUN(ctx, s, lgd_tap_render_bwd, &d.pyr, r0.out32, inst, P[LINST2D_W], t.ranges, tb.img_of, tb.img_start, tb.n_render,
        T, g_inst, G[LINST2D_W], tws, tbytes);
  } else {
    if ((rc = wg.run(t.rend_h, r0.out_h, r0.scale3, G[LINST2D_W], sa)) != LGD_OK) return rc;
    w = DgradW{t.pk.dgrad[TC_LINST], t.pk.gains + TC_LINST};
    if ((rc = dgrad(e, d, sa, ws, r0.out_h, r0.scale3, w, nullptr, false, ping[1], &(o = DgradOut()))) != LGD_OK) return rc;
    RUN(ctx, s, lgd_render_bwd, &d.pyr, o.out32, t.ranges, tb.img_of, tb.img_start, tb.n_render, T, g_inst, ws, d.ws_bytes);
  }
  float* g_a = sa.take<float>((size_t)F * T * C);
  float* g_ctxv = d.ctx ? sa.take<float>((size_t)F * T * C) : nullptr;
  float* g_att = sa.take<float>((size_t)F * T * C);
  float* gq = sa.take<float>((size_t)F * T * C);
  float* gk = sa.take<float>((size_t)T * C);
  float* gv = sa.take<float>((size_t)T * C);
  float* gs = sa.take<float>((size_t)F * d.heads * T * d.max_n);
  float* g_pooled = sa.take<float>((size_t)F * T * C);
  float* g_kin = sa.take<float>((size_t)T * C);
  const float* Wi = P[MHA_INW];
  float* gWi = G[MHA_INW];
  float* gbi = G[MHA_INB];
  const float* canoni = L.canoni.y;
  if (d.ctx) RUN(ctx, s, lgd_ctx_bias_table_bwd, r0.sums, tb.ctx_row, tb.img_of, F, B, T, g_ctxv);
  if (ctx->token_programs) {
    // Only the INPUT gradients of the relation block are on the chain's critical path: two persistent kernels around the
    // attention backward compute them; the six weight-gradient GEMMs (+ bias column sums) join the label-side program,
    // which runs on the label stream underneath the student-side backward.
    {
      TokenProgram pg(sa.take<char>(2 * LIN_WS_BYTES), sa.take<char>(2 * LIN_WS_BYTES), 2 * LIN_WS_BYTES, LIN_WS_BYTES);
      pg.linear_bwd_input(g_inst, C, P[LINST1D_W], C, g_a, C, F * T, C, C, 0);
      pg.next_stage();
      if (d.ctx) {
        pg.linear_bwd_input(g_ctxv, C, P[GCTX_W], C, g_a, C, F * T, C, C, 1);
        pg.next_stage();
      }
      pg.linear_bwd_input(g_a, C, P[MHA_OUTW], C, g_att, C, F * T, C, C, 0);
      pg.next_stage();
      const size_t pbytes = TokenProgram::device_bytes((int)pg.size() + 8);
      if ((rc = launch_program(ctx, pg, "lgd_relation_program_bwd", sa.take<char>(pbytes), pbytes, s)) != LGD_OK) return rc;
    }
    RUN(ctx, s, lgd_attention_bwd, g_att, t.q, F, t.k, t.v, 1, F, T, d.heads, C, tb.img_of, tb.img_start, d.max_n, t.probs,
        gs, gq, gk, gv);
    {
      TokenProgram pg(sa.take<char>(2 * LIN_WS_BYTES), sa.take<char>(2 * LIN_WS_BYTES), 2 * LIN_WS_BYTES, LIN_WS_BYTES);
      pg.linear_bwd_input(gq, C, Wi, C, g_pooled, C, F * T, C, C, 0);
      pg.linear_bwd_input(gk, C, Wi + C * C, C, g_kin, C, T, C, C, 0);
      pg.next_stage();
      pg.linear_bwd_input(gv, C, Wi + 2 * C * C, C, g_kin, C, T, C, C, 1);
      pg.next_stage();
      const size_t pbytes = TokenProgram::device_bytes((int)pg.size() + 8);
      if ((rc = launch_program(ctx, pg, "lgd_relation_program_bwd", sa.take<char>(pbytes), pbytes, s)) != LGD_OK) return rc;
    }
  } else {
    RUN(ctx, s, lgd_linear_bwd_weight, g_inst, C, t.a, C, G[LINST1D_W], C, G[LINST1D_B], F * T, C, C, 0, e.lin_ws, LIN_WS_BYTES);
    RUN(ctx, s, lgd_linear_bwd_input, g_inst, C, P[LINST1D_W], C, g_a, C, F * T, C, C, 0, e.lin_ws, LIN_WS_BYTES);
    if (d.ctx) {
      RUN(ctx, s, lgd_linear_bwd_input, g_ctxv, C, P[GCTX_W], C, g_a, C, F * T, C, C, 1, e.lin_ws, LIN_WS_BYTES);
      RUN(ctx, s, lgd_linear_bwd_weight, g_ctxv, C, t.a, C, G[GCTX_W], C, G[GCTX_B], F * T, C, C, 0, e.lin_ws, LIN_WS_BYTES);
    }
    // a6 backward
    RUN(ctx, s, lgd_linear_bwd_weight, g_a, C, t.att, C, G[MHA_OUTW], C, G[MHA_OUTB], F * T, C, C, 0, e.lin_ws, LIN_WS_BYTES);
    RUN(ctx, s, lgd_linear_bwd_input, g_a, C, P[MHA_OUTW], C, g_att, C, F * T, C, C, 0, e.lin_ws, LIN_WS_BYTES);
    RUN(ctx, s, lgd_attention_bwd, g_att, t.q, F, t.k, t.v, 1, F, T, d.heads, C, tb.img_of, tb.img_start, d.max_n, t.probs,
        gs, gq, gk, gv);
    RUN(ctx, s, lgd_linear_bwd_weight, gq, C, t.pooled, C, gWi, C, gbi, F * T, C, C, 0, e.lin_ws, LIN_WS_BYTES);
    RUN(ctx, s, lgd_linear_bwd_input, gq, C, Wi, C, g_pooled, C, F * T, C, C, 0, e.lin_ws, LIN_WS_BYTES);
    RUN(ctx, s, lgd_linear_bwd_weight, gk, C, canoni, C, gWi + C * C, C, gbi + C, T, C, C, 0, e.lin_ws, LIN_WS_BYTES);
    RUN(ctx, s, lgd_linear_bwd_input, gk, C, Wi + C * C, C, g_kin, C, T, C, C, 0, e.lin_ws, LIN_WS_BYTES);
    RUN(ctx, s, lgd_linear_bwd_weight, gv, C, canoni, C, gWi + 2 * C * C, C, gbi + 2 * C, T, C, C, 0, e.lin_ws, LIN_WS_BYTES);
    RUN(ctx, s, lgd_linear_bwd_input, gv, C, Wi + 2 * C * C, C, g_kin, C, T, C, C, 1, e.lin_ws, LIN_WS_BYTES);
  }
  float* g_canoni = g_kin;   // key and value inputs are the same tensor: gradients accumulated above

  // label side (canoni_proj_1D, label encoder): ~100 latency-bound launches that only end in parameter gradients, on
  // the label stream underneath the student-side backward below, which does not depend on them
  {
    cudaStream_t ls = s;
    const bool side = ctx->side_streams && !ctx->profiling;
    if (side) {
      LGD_CUDA(cudaEventRecord(ctx->ev_label, s));
      LGD_CUDA(cudaStreamWaitEvent(ctx->label_stream, ctx->ev_label, 0));
      ls = ctx->label_stream;
    }
    Exec le{ctx, ls, side ? lin_ws_label : e.lin_ws};
    float* g_le = sa.take<float>((size_t)T * C);
    float* gcat = sa.take<float>((size_t)T * 1088);
    float* g_xft = sa.take<float>((size_t)T * 64);
    float* g_a3 = sa.take<float>((size_t)T * 1024);
    float* g_a2 = sa.take<float>((size_t)T * 128);
    float* g_xft2 = sa.take<float>((size_t)T * 64);
    float* g_a1 = sa.take<float>((size_t)T * 64);
    float* g_tfeat = sa.take<float>((size_t)T * 64 * 64);
    float* g_a1b = sa.take<float>((size_t)T * 64);
    float* g_x1 = sa.take<float>((size_t)T * DESC);
    float* g_desc = sa.take<float>((size_t)T * DESC);
    float* g_tdesc = sa.take<float>((size_t)T * DESC * DESC);
    if (ctx->token_programs) {
      // two GEMMs per stage (weight and input gradient of a unit) -> two workspace slices per arena
      TokenProgram pg(sa.take<char>(2 * LIN_WS_BYTES), sa.take<char>(2 * LIN_WS_BYTES), 2 * LIN_WS_BYTES, LIN_WS_BYTES);
      // weight gradients of the relation block (their inputs are complete on the main stream at this point)
      pg.linear_bwd_weight(g_inst, C, t.a, C, G[LINST1D_W], C, G[LINST1D_B], F * T, C, C);
      if (d.ctx) pg.linear_bwd_weight(g_ctxv, C, t.a, C, G[GCTX_W], C, G[GCTX_B], F * T, C, C);
      pg.next_stage();
      pg.linear_bwd_weight(g_a, C, t.att, C, G[MHA_OUTW], C, G[MHA_OUTB], F * T, C, C);
      pg.linear_bwd_weight(gq, C, t.pooled, C, gWi, C, gbi, F * T, C, C);
      pg.next_stage();
      pg.linear_bwd_weight(gk, C, canoni, C, gWi + C * C, C, gbi + C, T, C, C);
      pg.linear_bwd_weight(gv, C, canoni, C, gWi + 2 * C * C, C, gbi + 2 * C, T, C, C);
      pg.next_stage();
      unit_bwd_prog(pg, L.canoni, g_canoni, L.canoni.x, T, P, G, g_le);
      unit_bwd_prog(pg, L.c4, g_le, L.c4.x, T, P, G, gcat);
      pg.segmax_bwd(gcat, 64, 1024, tb.img_start, B, L.argmax, g_xft, g_a3);
      pg.next_stage();
      unit_bwd_prog(pg, L.c3, g_a3, L.c3.x, T, P, G, g_a2);
      unit_bwd_prog(pg, L.c2, g_a2, L.c2.x, T, P, G, g_xft2);
      pg.axpy(g_xft2, g_xft, (long long)T * 64);
      pg.next_stage();
      pg.rowvec_bwd(g_xft, L.c1.y, L.sf[5].y, g_a1, g_tfeat, T, 64);
      pg.next_stage();
      stn_bwd_prog(pg, L.sf, g_tfeat, T, P, G, sa, g_a1b);
      pg.axpy(g_a1b, g_a1, (long long)T * 64);
      pg.next_stage();
      unit_bwd_prog(pg, L.c1, g_a1, L.c1.x, T, P, G, g_x1);
      pg.rowvec_bwd(g_x1, L.desc, L.sd[5].y, g_desc, g_tdesc, T, DESC);
      pg.next_stage();
      stn_bwd_prog(pg, L.sd, g_tdesc, T, P, G, sa, nullptr);   // descriptors are data
      const size_t pbytes = TokenProgram::device_bytes((int)pg.size() + 8);
      if ((rc = launch_program(ctx, pg, "lgd_label_program_bwd", sa.take<char>(pbytes), pbytes, ls)) != LGD_OK) return rc;
    } else {
      if ((rc = unit_bwd(le, L.canoni, g_canoni, L.canoni.x, T, P, G, g_le)) != LGD_OK) return rc;
      if ((rc = unit_bwd(le, L.c4, g_le, L.c4.x, T, P, G, gcat)) != LGD_OK) return rc;
      RUN(ctx, ls, lgd_segmax_concat_bwd, gcat, 64, 1024, tb.img_start, B, L.argmax, g_xft, g_a3);
      if ((rc = unit_bwd(le, L.c3, g_a3, L.c3.x, T, P, G, g_a2)) != LGD_OK) return rc;
      if ((rc = unit_bwd(le, L.c2, g_a2, L.c2.x, T, P, G, g_xft2)) != LGD_OK) return rc;
      RUN(ctx, ls, lgd_axpy, g_xft2, g_xft, (int64_t)T * 64);
      RUN(ctx, ls, lgd_rowvec_matmul_bwd, g_xft, L.c1.y, L.sf[5].y, g_a1, g_tfeat, T, 64);
      if ((rc = stn_bwd(le, L.sf, g_tfeat, T, P, G, sa, g_a1b)) != LGD_OK) return rc;
      RUN(ctx, ls, lgd_axpy, g_a1b, g_a1, (int64_t)T * 64);
      if ((rc = unit_bwd(le, L.c1, g_a1, L.c1.x, T, P, G, g_x1)) != LGD_OK) return rc;
      RUN(ctx, ls, lgd_rowvec_matmul_bwd, g_x1, L.desc, L.sd[5].y, g_desc, g_tdesc, T, DESC);
      if ((rc = stn_bwd(le, L.sd, g_tdesc, T, P, G, sa, nullptr)) != LGD_OK) return rc;   // descriptors are data
    }
    if (side) LGD_CUDA(cudaEventRecord(ctx->ev_label, ls));
  }

  // Everything enqueued so far ends in every parameter gradient of the teacher except student_proj_2D's: mark it on the
  // three streams, so that the gradient exchange of that part (lgd_ctx_wait_early_grads) can run underneath the
  // student-side backward below instead of after the chain.
  {
    const bool side = ctx->side_streams && !ctx->profiling;
    LGD_CUDA(cudaEventRecord(ctx->ev_early_main, s));
    LGD_CUDA(cudaEventRecord(ctx->ev_early_wgrad, side ? ctx->wgrad_stream : s));
    ctx->early_label = side;
  }

  // a5 + a3 backward (appearance embeddings -> student_proj_2D)
  {
    float* g_y = ping[0];
    RUN(ctx, s, lgd_maskpool_bwd, &d.pyr, g_pooled, t.ranges, tb.img_start, T, g_y);
    if ((rc = gn_bwd_half(e, d, sa, ws, g_y, t.sp_raw, t.sp_stats, 1, G[SPROJ_B], &gh, &sc)) != LGD_OK) return rc;
    if ((rc = wg.run(stu_h, gh, sc, G[SPROJ_W], sa)) != LGD_OK) return rc;
    if (need_feat) {
      w = DgradW{t.pk.dgrad[TC_SPROJ], t.pk.gains + TC_SPROJ};
      // channels-last callers take the dgrad output as it is (their gradient tensors are views of gstu_pyramid)
      float* dst = gstu_levels_host ? ping[1] : gstu_pyramid;
      if ((rc = dgrad(e, d, sa, ws, gh, sc, w, nullptr, false, dst, &(o = DgradOut()))) != LGD_OK) return rc;
      if (gstu_levels_host) RUN(ctx, s, lgd_pyramid_to_nchw, o.out32, &d.pyr, gstu_levels_host, gstu_accumulate);
    }
  }
  if (ctx->side_streams && !ctx->profiling) LGD_CUDA(cudaStreamWaitEvent(s, ctx->ev_label, 0));
  if ((rc = wg.join()) != LGD_OK) return rc;
  LGD_CHECK_ARG(!sa.overflow, "lgd_teacher_backward: scratch overflow (%zu > %zu)", sa.off, scratch_bytes);
  return LGD_OK;
}

// ================================================================================================ distillation loss
namespace lgd {
struct DistillTape {
  PackedConvs pk;
  __half *a1_h, *a2_h;
  float *s, *st_s, *st_t, *bwd_sums, *gs_terms;
};
static void layout_distill_tape(Arena& a, const Dims& d, DistillTape* t) {
  const size_t E = (size_t)d.E, FB = (size_t)d.F * d.B;
  layout_packed(a, AC_COUNT, &t->pk);
  t->a1_h = a.take<__half>(E);
  t->a2_h = a.take<__half>(E);
  t->s = a.take<float>(E);
  t->st_s = a.take<float>(FB * C * 2);
  t->st_t = a.take<float>(FB * C * 2);
  t->bwd_sums = a.take<float>(FB * 2 * C);
  t->gs_terms = a.take<float>(FB);
}
static size_t distill_fwd_scratch(const Dims& d) { return d.ws_bytes + 3 * ((size_t)9 * C * C * 2 + 256) + 4096; }
static size_t distill_bwd_scratch(const Dims& d) {
  const size_t E = (size_t)d.E;
  size_t n = d.ws_bytes + 4096;
  n += 3 * (E * 2 + 256) + E * 4 + 256;     // scaled fp16 gradients of the three stages, fp32 input gradient
  n += 3 * ((size_t)9 * C * C * 2 + 256 + 9 * C * 4 + 256 + 256) + 3 * ((size_t)9 * C * C * 4);
  n += 4 * ((size_t)d.F * d.B * C * 4 + C * 4 + 512) + 4 * ((size_t)d.num_tiles * 2 * 4 + 1024) + 64 * 256;
  return n;
}
}  // namespace lgd

extern "C" int lgd_distill_tape_field(const lgd_step_desc_t* desc, const char* name, size_t* offset, size_t* bytes) {
  Dims d;
  int rc = make_dims(desc, &d);
  if (rc != LGD_OK) return rc;
  LGD_CHECK_ARG(name && offset && bytes, "lgd_distill_tape_field: null pointer");
  Arena a(reinterpret_cast<void*>(4096), ~size_t(0) >> 1);
  DistillTape t;
  layout_distill_tape(a, d, &t);
  const size_t E = (size_t)d.E;
  const Field f[] = {{"a1_h", t.a1_h, E * 2}, {"a2_h", t.a2_h, E * 2}, {"s", t.s, E * 4}};
  return find_field(f, 3, name, offset, bytes);
}

extern "C" size_t lgd_distill_tape_bytes(const lgd_step_desc_t* desc) {
  Dims d;
  if (make_dims(desc, &d) != LGD_OK) return 0;
  Arena a(nullptr, 0);
  DistillTape t;
  layout_distill_tape(a, d, &t);
  return a.off;
}
extern "C" size_t lgd_distill_scratch_bytes(const lgd_step_desc_t* desc, int backward) {
  Dims d;
  if (make_dims(desc, &d) != LGD_OK) return 0;
  return backward ? distill_bwd_scratch(d) : distill_fwd_scratch(d);
}

extern "C" int lgd_distill_forward(lgd_ctx_t* ctx, const lgd_step_desc_t* desc, const void* stu_half,
                                   const float* tea, const float* const* params_host, float coef, void* tea_ready_event,
                                   float* loss, void* tape, size_t tape_bytes, void* scratch, size_t scratch_bytes,
                                   void* stream) {
  LGD_CHECK_ARG(ctx && stu_half && tea && params_host && loss && tape && scratch, "lgd_distill_forward: null pointer");
  Dims d;
  int rc = make_dims(desc, &d);
  if (rc != LGD_OK) return rc;
  const float* const* P = params_host;
  for (int i = 0; i < AP_COUNT; ++i) LGD_CHECK_ARG(P[i], "lgd_distill_forward: parameter %s is missing", AP_NAMES[i]);
  Arena ta(tape, tape_bytes), sa(scratch, scratch_bytes);
  DistillTape t;
  layout_distill_tape(ta, d, &t);
  LGD_CHECK_ARG(!ta.overflow, "lgd_distill_forward: tape too small");
  LGD_CHECK_ARG(scratch_bytes >= distill_fwd_scratch(d), "lgd_distill_forward: scratch too small");
  cudaStream_t s = (cudaStream_t)stream;
  Exec e{ctx, s, nullptr};
  void* ws = sa.take<char>(d.ws_bytes);
  const __half* stu_h = static_cast<const __half*>(stu_half);
  {
    const float* w3[AC_COUNT] = {P[AD0_W], P[AD2_W], P[AD4_W]};
    if ((rc = pack_all(ctx, s, t.pk, w3, ws)) != LGD_OK) return rc;
  }
  RUN(ctx, s, lgd_conv3x3_fwd_f16, &d.pyr, stu_h, t.pk.fwd[AC_0], P[AD0_B], 0, 0, nullptr, t.a1_h, 1, 1, nullptr);
  RUN(ctx, s, lgd_conv3x3_fwd_f16, &d.pyr, t.a1_h, t.pk.fwd[AC_2], P[AD2_B], 0, 0, nullptr, t.a2_h, 1, 1, nullptr);
  RUN(ctx, s, lgd_conv3x3_fwd_f16, &d.pyr, t.a2_h, t.pk.fwd[AC_4], P[AD4_B], 0, 0, t.s, nullptr, 0, 0, nullptr);
  // running next to the teacher chain: the loss is the first consumer of the teacher pyramid
  if (tea_ready_event != nullptr) LGD_CUDA(cudaStreamWaitEvent(s, (cudaEvent_t)tea_ready_event, 0));
  RUN(ctx, s, lgd_in_mse_moments_fwd, &d.pyr, t.s, tea, coef, t.st_s, t.st_t, t.bwd_sums, t.gs_terms, loss, ws,
      d.ws_bytes);
  return LGD_OK;
}

extern "C" int lgd_distill_backward(lgd_ctx_t* ctx, const lgd_step_desc_t* desc, const void* stu_half,
                                    const float* tea, const float* const* params_host, float coef, const float* gloss,
                                    const void* tape, size_t tape_bytes, float* const* grads_host,
                                    float* const* gstu_levels_host, float* gstu_pyramid, int gstu_accumulate,
                                    void* wgrad_workspace, void* scratch, size_t scratch_bytes, void* stream) {
  LGD_CHECK_ARG(ctx && stu_half && tea && params_host && gloss && tape && grads_host && wgrad_workspace && scratch,
                "lgd_distill_backward: null pointer");
  Dims d;
  int rc = make_dims(desc, &d);
  if (rc != LGD_OK) return rc;
  const float* const* P = params_host;
  float* const* G = grads_host;
  for (int i = 0; i < AP_COUNT; ++i)
    LGD_CHECK_ARG(P[i] && G[i], "lgd_distill_backward: parameter / gradient %s is missing", AP_NAMES[i]);
  Arena ta(const_cast<void*>(tape), tape_bytes), sa(scratch, scratch_bytes);
  DistillTape t;
  layout_distill_tape(ta, d, &t);
  LGD_CHECK_ARG(!ta.overflow, "lgd_distill_backward: tape too small");
  LGD_CHECK_ARG(scratch_bytes >= distill_bwd_scratch(d), "lgd_distill_backward: scratch too small");
  cudaStream_t s = (cudaStream_t)stream;
  Exec e{ctx, s, nullptr};
  void* ws = sa.take<char>(d.ws_bytes);
  const __half* stu_h = static_cast<const __half*>(stu_half);
  Wgrad wg{ctx, s, &d, wgrad_workspace};
  const bool need_feat = gstu_levels_host != nullptr || gstu_pyramid != nullptr;

  __half* gh = sa.take<__half>((size_t)d.E);
  float* sc = sa.take<float>(3);
  RUN(ctx, s, lgd_in_mse_bwd, &d.pyr, t.s, tea, t.st_s, t.st_t, t.bwd_sums, coef, gloss, nullptr, 1, t.gs_terms, gh, sc,
      nullptr, G[AD4_B], ws, d.ws_bytes);
  DgradW w;
  DgradOut r2, r1, r0;
  if ((rc = wg.run(t.a2_h, gh, sc, G[AD4_W], sa)) != LGD_OK) return rc;
  w = DgradW{t.pk.dgrad[AC_4], t.pk.gains + AC_4};
  if ((rc = dgrad(e, d, sa, ws, gh, sc, w, t.a2_h, true, nullptr, &r2)) != LGD_OK) return rc;
  LGD_CUDA(cudaMemcpyAsync(G[AD2_B], r2.total, C * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if ((rc = wg.run(t.a1_h, r2.out_h, r2.scale3, G[AD2_W], sa)) != LGD_OK) return rc;
  w = DgradW{t.pk.dgrad[AC_2], t.pk.gains + AC_2};
  if ((rc = dgrad(e, d, sa, ws, r2.out_h, r2.scale3, w, t.a1_h, true, nullptr, &r1)) != LGD_OK) return rc;
  LGD_CUDA(cudaMemcpyAsync(G[AD0_B], r1.total, C * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if ((rc = wg.run(stu_h, r1.out_h, r1.scale3, G[AD0_W], sa)) != LGD_OK) return rc;
  if (need_feat) {
    w = DgradW{t.pk.dgrad[AC_0], t.pk.gains + AC_0};
    float* dst = gstu_levels_host ? sa.take<float>((size_t)d.E) : gstu_pyramid;
    if ((rc = dgrad(e, d, sa, ws, r1.out_h, r1.scale3, w, nullptr, false, dst, &r0)) != LGD_OK) return rc;
    if (gstu_levels_host) RUN(ctx, s, lgd_pyramid_to_nchw, r0.out32, &d.pyr, gstu_levels_host, gstu_accumulate);
  }
  if ((rc = wg.join()) != LGD_OK) return rc;
  LGD_CHECK_ARG(!sa.overflow, "lgd_distill_backward: scratch overflow (%zu > %zu)", sa.off, scratch_bytes);
  return LGD_OK;
}
