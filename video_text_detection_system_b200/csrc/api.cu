// libvtd_b200.so: context, weight loader and the C-ABI entry points declared in include/vtd.h.
//
// The host runtime is deliberately small: a context owns one stream, one device arena laid out at
// vtd_create for the configured maximum batch, and two "programs" (detector, recogniser) -- flat lists
// of conv / pool launches with every pointer, shape and tensor map bound when the weights are loaded, so
// a batch is a fixed sequence of launches with no allocation, no shape logic and no host<->device sync: in the
// speed tier even the recogniser is launched against the crop count in device memory (recognize_locked); the
// host looks at the count where it synchronises anyway and only then runs chunks beyond the first.
#include <nvtx3/nvToolsExt.h>
#include "common.cuh"
#include "resize_tab.h"
#include "../../include/vtd.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace vtd;

namespace {

std::string g_create_error;
std::mutex g_create_mutex;

struct DebugEntry { const void* p; int C, H, W; OutLayout lay; bool f32; bool per_crop; };

struct Op {
  enum Kind { CONV, POOL } kind = CONV;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // profiling (vtd_set_profiling): last launch of this op
  double prof_ms = 0.0; long long prof_n = 0; bool prof_pending = false;
  ConvDesc d{};
  TcPlan* plan = nullptr;
  StemPoolPlan* sp = nullptr;   // the op is the fused DBNet stem + max-pool (stem_pool_tcgen05)
  bool fused_head = false;      // plan is the one-pass DB head (dbhead_fused_tcgen05): 3x3 convolutions + both tails
  // pool
  const void* pin = nullptr; void* pout = nullptr;
  int H = 0, W = 0, C = 0, kh = 0, kw = 0, sh = 0, sw = 0, ph = 0, pw = 0;
};

struct Act { void* p = nullptr; int H = 0, W = 0, C = 0; };

// stage-level device timers (vtd_set_profiling): the launches of the path that are not conv/pool ops
enum { ST_PREPROCESS = 0, ST_HEAD_TAIL, ST_BOXES, ST_CROP, ST_LSTM0, ST_LSTM1, ST_CTC, ST_COUNT };
struct StageProf {
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double ms = 0.0; long long n = 0; bool pending = false;
};

struct HostConv {              // folded, repacked host weights [Cout][KH][KW][Cin_pad]
  std::vector<float> w, b;
  int Cout = 0, Cin = 0, Cin_pad = 0, KH = 0, KW = 0;
};

}  // namespace

struct vtd_ctx {
  vtd_config cfg{};
  std::mutex mu;
  std::string err;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  LaunchCounter lc;
  bool profiling = false;
  StageProf stage_prof[ST_COUNT];
  std::vector<void*> allocs;
  // VTD_FLAG_GUARD_ALLOCS: every arena allocation sits between two canary pages, checked by vtd_check_guards()
  struct Guarded { uint8_t* user; size_t bytes; };
  std::vector<Guarded> guarded;
  size_t esz = 4;                       // activation element size
  bool bf16_mode = false;
  int T = 0;                            // CRNN sequence length
  int rc = 0;                           // recogniser chunk capacity (crops)

  // detector state
  bool det_loaded = false, rec_loaded = false;
  std::vector<Op> det_prog, rec_prog;
  std::map<std::string, DebugEntry> dbg;
  bool use_stem_pool = false;           // DBNet stem and its max-pool run as one kernel (input buffer has a 6-px left border)
  bool use_win = false, use_tchead = false, use_tclstm = false;   // tcgen05 stems / head tail / LSTM (bf16 tier)
  OutLayout pre_lay{}, crops_lay{};
  TcPlan* head_plan = nullptr;
  bool head_fused = false;              // the DB head runs as ONE kernel (no feature map in HBM)
  float cur_thr = 0.5f; const float* cur_bias = nullptr;   // arguments of the detect call in flight (read by the fused head op)
  TcPlan* lstm_plan[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [layer][step parity] (step-wise variant)
  LstmPlan* plstm[2] = {nullptr, nullptr};   // persistent clustered variant (default)
  bool use_plstm = false;
  bf16* h16 = nullptr;                  // [2 ping-pong][2 dirs][rc][256]
  void* pre = nullptr;                  // [B,dh,dw,4]  (bf16 + use_win: zero-bordered [B,dh+6,dw+8,4])
  void* head_feat = nullptr;            // [B,dh/4,dw/4,128]
  HeadTailWeights htw{};
  float *prob = nullptr, *thresh = nullptr; uint8_t* mask = nullptr;
  float* stage_f32 = nullptr; size_t stage_f32_elems = 0;   // NCHW fp32 staging for the forward drop-ins

  // frames
  uint8_t* frames_store = nullptr; size_t frame_bytes_cap = 0;
  // pageable host frames (NumPy arrays of a Python caller) are staged through this pinned buffer by a few copy threads,
  // so that the host->device copies are real asynchronous DMA instead of the driver's serialised bounce copies
  uint8_t* host_stage = nullptr; size_t host_stage_bytes = 0; cudaEvent_t host_stage_event = nullptr;
  const uint8_t** store_ptrs_dev = nullptr;    // constant: the staging slots
  const uint8_t** ext_ptrs_dev = nullptr;      // caller-owned device frames of the current batch
  const uint8_t** frame_ptrs_dev = nullptr;    // whichever of the two the current batch uses
  const uint8_t** frame_ptrs_pinned = nullptr;
  cudaEvent_t ptrs_event = nullptr;
  OverlayItem* ov_items = nullptr; int* ov_end = nullptr; bool ov_tables = false;   // annotated-frame overlay (lazy)
  // speed tier: the recogniser's first chunk is launched against the crop count ON THE DEVICE (no host round trip in the step);
  // the host reads the total later, where it synchronises anyway, and only then runs the chunks beyond the first (rare)
  bool rec_dyn_ok = false;                     // every recogniser launch can take the device-side count
  const int* n_dyn = nullptr; int n_first = 0; // set while such a chunk is being launched (run_op passes them on)
  cudaEvent_t total_event = nullptr;           // the crop total has reached pinned_int[0]
  bool rec_pending = false; int pending_n = 0; // a batch whose total has not been looked at yet
  int cur_h = 0, cur_w = 0, cur_pitch = 0, cur_n = 0, cur_pix = 0;
  ResizeTab tx, ty; int tab_h = -1, tab_w = -1;
  float* norm_lut = nullptr;            // [3][256] u8 -> normalised fp32

  // boxes
  uint8_t* box_work = nullptr; BoxWorkLayout box_lay{};
  vtd_record* records = nullptr; int* counts = nullptr; int* offsets = nullptr;
  int* pinned_int = nullptr;            // [max_batch+2]
  int overflow_cached = -1;             // overflow flag fetched with the last record read-back (-1: not fetched since the last extraction)
  // arbitrary-size post-process drop-in
  uint8_t* pp_work = nullptr; BoxWorkLayout pp_lay{}; int pp_h = 0, pp_w = 0;
  float* pp_prob = nullptr; uint8_t* pp_mask = nullptr; vtd_record* pp_records = nullptr; int* pp_counts = nullptr;

  // recogniser state
  void* crops = nullptr;                // [rc,32,cw,4]
  void* seq = nullptr;                  // [rc,T,512] conv features
  float* xproj = nullptr;               // [rc,T,2048]
  void* rnn_out[2] = {nullptr, nullptr};
  void* whh[2] = {nullptr, nullptr};
  float *hbuf = nullptr, *cbuf = nullptr;
  float* logits = nullptr;              // [rc,T,logits_ld]  (97, or 128 when the classifier runs on tcgen05)
  int logits_ld = 97;
  Op xproj_op[2], fc_op;
  uint8_t* ids_dev = nullptr; int* len_dev = nullptr; float* conf_dev = nullptr;
  // crop-list drop-in staging
  uint8_t* list_store = nullptr; size_t list_cap = 0;
  const uint8_t** list_ptrs = nullptr; int* list_meta = nullptr;   // device: [rc] ptrs, [3*rc] h,w,pitch
  void* trocr = nullptr;                // TrocrState (trocr_host.inc): the transformer recogniser, when loaded
};

namespace {

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      char buf__[512];                                                                             \
      snprintf(buf__, sizeof(buf__), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      c->err = buf__;                                                                              \
      return VTD_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

#define FAIL(code, ...)                                   \
  do {                                                    \
    char buf__[512];                                      \
    snprintf(buf__, sizeof(buf__), __VA_ARGS__);          \
    c->err = buf__;                                       \
    return code;                                          \
  } while (0)

constexpr size_t GUARD_BYTES = 4096;
constexpr int GUARD_PATTERN = 0xA5;

int dev_alloc(vtd_ctx* c, void** p, size_t bytes) {
  if (bytes == 0) bytes = 256;
  if (c->cfg.flags & VTD_FLAG_GUARD_ALLOCS) {
    const size_t padded = (bytes + 255) & ~(size_t)255;          // the tail canary starts right after the (rounded) buffer
    uint8_t* base = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&base), padded + 2 * GUARD_BYTES));
    c->allocs.push_back(base);
    CK(cudaMemset(base, GUARD_PATTERN, GUARD_BYTES));
    CK(cudaMemset(base + GUARD_BYTES + bytes, GUARD_PATTERN, padded - bytes + GUARD_BYTES));
    *p = base + GUARD_BYTES;
    c->guarded.push_back({base + GUARD_BYTES, bytes});
    return VTD_OK;
  }
  CK(cudaMalloc(p, bytes));
  c->allocs.push_back(*p);
  return VTD_OK;
}
template <typename P> int dalloc(vtd_ctx* c, P** p, size_t bytes) { return dev_alloc(c, reinterpret_cast<void**>(p), bytes); }

// ---- state-dict access -----------------------------------------------------------------------------
struct SD {
  std::map<std::string, const vtd_tensor*> m;
  const vtd_tensor* get(const std::string& k) const { auto it = m.find(k); return it == m.end() ? nullptr : it->second; }
};

long long numel(const vtd_tensor* t) { long long n = 1; for (int i = 0; i < t->ndim; ++i) n *= t->shape[i]; return n; }

// conv (+ optional bias) (+ optional BN) -> folded weights in [Cout][KH][KW][Cin_pad]
int fold_conv(vtd_ctx* c, const SD& sd, const std::string& conv, const std::string& bn, int cin_pad_to, HostConv* out) {
  const vtd_tensor* w = sd.get(conv + ".weight");
  if (!w || w->ndim != 4) FAIL(VTD_ERR_WEIGHT, "missing or non-4D weight '%s.weight'", conv.c_str());
  const int Cout = (int)w->shape[0], Cin = (int)w->shape[1], KH = (int)w->shape[2], KW = (int)w->shape[3];
  const vtd_tensor* b = sd.get(conv + ".bias");
  if (b && numel(b) != Cout) FAIL(VTD_ERR_WEIGHT, "bias '%s.bias' has wrong size", conv.c_str());
  std::vector<float> scale(Cout, 1.f), shift(Cout, 0.f);
  for (int o = 0; o < Cout; ++o) shift[o] = b ? b->data[o] : 0.f;
  if (!bn.empty()) {
    const vtd_tensor *g = sd.get(bn + ".weight"), *be = sd.get(bn + ".bias"), *mu = sd.get(bn + ".running_mean"),
                     *var = sd.get(bn + ".running_var");
    if (!g || !be || !mu || !var || numel(g) != Cout || numel(be) != Cout || numel(mu) != Cout || numel(var) != Cout)
      FAIL(VTD_ERR_WEIGHT, "missing or mis-shaped BatchNorm '%s'", bn.c_str());
    for (int o = 0; o < Cout; ++o) {
      float s = g->data[o] / sqrtf(var->data[o] + 1e-5f);
      scale[o] = s;
      shift[o] = (shift[o] - mu->data[o]) * s + be->data[o];
    }
  }
  const int Cp = cin_pad_to > Cin ? cin_pad_to : Cin;
  out->Cout = Cout; out->Cin = Cin; out->Cin_pad = Cp; out->KH = KH; out->KW = KW;
  out->w.assign((size_t)Cout * KH * KW * Cp, 0.f);
  out->b = shift;
  for (int o = 0; o < Cout; ++o)
    for (int i = 0; i < Cin; ++i)
      for (int r = 0; r < KH; ++r)
        for (int s = 0; s < KW; ++s)
          out->w[(((size_t)o * KH + r) * KW + s) * Cp + i] = w->data[(((size_t)o * Cin + i) * KH + r) * KW + s] * scale[o];
  return VTD_OK;
}

int upload_act_type(vtd_ctx* c, const std::vector<float>& h, void** dev) {
  if (c->bf16_mode) {
    std::vector<bf16> t(h.size());
    for (size_t i = 0; i < h.size(); ++i) t[i] = f32_to_16(h[i]);
    int r = dev_alloc(c, dev, t.size() * 2); if (r) return r;
    CK(cudaMemcpy(*dev, t.data(), t.size() * 2, cudaMemcpyHostToDevice));
  } else {
    int r = dev_alloc(c, dev, h.size() * 4); if (r) return r;
    CK(cudaMemcpy(*dev, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  }
  return VTD_OK;
}
int upload_f32(vtd_ctx* c, const std::vector<float>& h, float** dev) {
  int r = dalloc(c, dev, h.size() * 4); if (r) return r;
  CK(cudaMemcpy(*dev, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  return VTD_OK;
}

// Append a conv op whose input/output live in the arena; N is the capacity the buffers were sized for.
int add_conv(vtd_ctx* c, std::vector<Op>* prog, const HostConv& hc, int N, const Act& in, int stride, int pad,
             bool relu, const Act* res, int res_mode, bool out_f32, Act* out, Op* standalone = nullptr, int pool = 0,
             bool* fused = nullptr) {
  if (in.C != hc.Cin_pad) FAIL(VTD_ERR_WEIGHT, "layer expects %d input channels, activation has %d", hc.Cin_pad, in.C);
  Op op;
  op.kind = Op::CONV;
  ConvDesc& d = op.d;
  d.in = in.p;
  d.N = N; d.H = in.H; d.W = in.W; d.Cin = in.C;
  d.KH = hc.KH; d.KW = hc.KW; d.stride = stride; d.pad = pad;
  d.Ho = (in.H + 2 * pad - hc.KH) / stride + 1;
  d.Wo = (in.W + 2 * pad - hc.KW) / stride + 1;
  d.Cout = hc.Cout;
  d.relu = relu ? 1 : 0;
  d.res = res ? res->p : nullptr;
  d.res_mode = res ? res_mode : RES_NONE;
  d.out_mode = OUT_NHWC;
  d.out_f32 = out_f32 ? 1 : 0;
  void* wdev = nullptr;
  int r = upload_act_type(c, hc.w, &wdev); if (r) return r;
  d.w = wdev;
  float* bdev = nullptr;
  r = upload_f32(c, hc.b, &bdev); if (r) return r;
  d.bias = bdev;
  size_t oes = out_f32 ? 4 : c->esz;
  void* o = nullptr;
  // pool != 0 (tcgen05 tier only): the max-pool that follows is fused into the epilogue and `out` is the pooled map
  // (the plan refuses when the staging would cost a ring slot: the caller then keeps the separate pooling kernel)
  if (pool && !(c->bf16_mode && tc_supported(d))) pool = 0;
  r = dev_alloc(c, &o, (size_t)N * d.Ho * d.Wo * d.Cout * oes); if (r) return r;
  d.out = o;
  if (c->bf16_mode && tc_supported(d)) {
    std::string e;
    d.pool = pool;
    op.plan = tc_plan_create(d, &e);
    if (!op.plan && pool) { pool = 0; d.pool = 0; op.plan = tc_plan_create(d, &e); }
    if (!op.plan) FAIL(VTD_ERR_CUDA, "tcgen05 plan: %s", e.c_str());
  }
  const int pw = pool == 1 ? 2 : 1, ph = pool ? 2 : 1;
  if (fused) *fused = pool != 0;
  out->p = o; out->H = d.Ho / ph; out->W = d.Wo / pw; out->C = d.Cout;
  if (standalone) *standalone = op; else prog->push_back(op);
  return VTD_OK;
}

int add_pool(vtd_ctx* c, std::vector<Op>* prog, int N, const Act& in, int kh, int kw, int sh, int sw, int ph, int pw,
             Act* out) {
  Op op;
  op.kind = Op::POOL;
  op.pin = in.p; op.H = in.H; op.W = in.W; op.C = in.C;
  op.kh = kh; op.kw = kw; op.sh = sh; op.sw = sw; op.ph = ph; op.pw = pw;
  int Ho = (in.H + 2 * ph - kh) / sh + 1, Wo = (in.W + 2 * pw - kw) / sw + 1;
  void* o = nullptr;
  int r = dev_alloc(c, &o, (size_t)N * Ho * Wo * in.C * c->esz); if (r) return r;
  op.pout = o;
  out->p = o; out->H = Ho; out->W = Wo; out->C = in.C;
  prog->push_back(op);
  return VTD_OK;
}

cudaError_t run_op(vtd_ctx* c, const Op& op, int n) {
  if (op.kind == Op::POOL) {
    if (c->bf16_mode)
      return maxpool_nhwc<bf16>((const bf16*)op.pin, (bf16*)op.pout, n, op.H, op.W, op.C, op.kh, op.kw, op.sh, op.sw,
                                op.ph, op.pw, c->stream, &c->lc);
    return maxpool_nhwc<float>((const float*)op.pin, (float*)op.pout, n, op.H, op.W, op.C, op.kh, op.kw, op.sh, op.sw,
                               op.ph, op.pw, c->stream, &c->lc);
  }
  if (op.sp) return stem_pool_tcgen05(op.sp, n, c->stream, &c->lc, c->n_dyn, c->n_first);
  if (op.fused_head) return dbhead_fused_tcgen05(op.plan, n, c->cur_thr, c->cur_bias, c->stream, &c->lc);
  if (op.plan) return conv_tcgen05(op.plan, n, c->stream, &c->lc, c->n_dyn, c->n_first);
  ConvDesc d = op.d;
  d.N = n;
  if (c->bf16_mode) return conv_generic<bf16>(d, c->stream, &c->lc);
  return conv_generic<float>(d, c->stream, &c->lc);
}

// With profiling on, every op of a program is bracketed by CUDA events on the launching stream; the pair of
// the previous launch is harvested (it has long completed) before it is re-recorded, so there is no sync.
void prof_harvest(Op& op) {
  if (!op.prof_pending) return;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, op.ev0, op.ev1) == cudaSuccess) { op.prof_ms += ms; op.prof_n++; op.prof_pending = false; }
  else cudaGetLastError();   // not ready yet: keep pending, clear the sticky-free error
}

void stage_harvest(StageProf& sp) {
  float ms = 0.f;
  if (sp.pending && cudaEventElapsedTime(&ms, sp.ev0, sp.ev1) == cudaSuccess) { sp.ms += ms; sp.n++; }
  sp.pending = false;
}
// NVTX ranges (header-only NVTX v3; no-ops unless a tool is attached): one per API call and per stage of the path, so a
// timeline shows preprocess / detector / boxes / crop / crnn / lstm / ctc per batch and context
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
static const char* const kStageNames[] = {"vtd.preprocess", "vtd.head_tail", "vtd.boxes", "vtd.crop", "vtd.lstm0", "vtd.lstm1", "vtd.ctc"};

struct StageTimer {                      // brackets a stage with CUDA events on the launching stream when profiling is on
  vtd_ctx* c; StageProf* sp; NvtxRange range;
  StageTimer(vtd_ctx* ctx, int id) : c(ctx), sp(ctx->profiling ? &ctx->stage_prof[id] : nullptr), range(kStageNames[id]) {
    if (!sp) return;
    if (!sp->ev0) { cudaEventCreate(&sp->ev0); cudaEventCreate(&sp->ev1); }
    if (sp->pending) { cudaEventSynchronize(sp->ev1); stage_harvest(*sp); }
    cudaEventRecord(sp->ev0, c->stream);
  }
  ~StageTimer() { if (sp) { cudaEventRecord(sp->ev1, c->stream); sp->pending = true; } }
};

int run_op_prof(vtd_ctx* c, Op& op, int n) {
  if (!c->profiling) { CK(run_op(c, op, n)); return VTD_OK; }
  if (!op.ev0) { CK(cudaEventCreate(&op.ev0)); CK(cudaEventCreate(&op.ev1)); }
  if (op.prof_pending) { cudaEventSynchronize(op.ev1); prof_harvest(op); }
  CK(cudaEventRecord(op.ev0, c->stream));
  CK(run_op(c, op, n));
  CK(cudaEventRecord(op.ev1, c->stream));
  op.prof_pending = true;
  return VTD_OK;
}

int run_prog(vtd_ctx* c, std::vector<Op>& prog, int n) {
  for (Op& op : prog) { int r = run_op_prof(c, op, n); if (r) return r; }
  return VTD_OK;
}

void reg_dbg(vtd_ctx* c, const char* name, const Act& a, bool f32 = false, bool per_crop = false) {
  c->dbg[name] = DebugEntry{a.p, a.C, a.H, a.W, dense_layout(a.H, a.W, a.C), f32, per_crop};
}
void reg_dbg_lay(vtd_ctx* c, const char* name, const void* p, int C, int H, int W, OutLayout lay, bool per_crop) {
  c->dbg[name] = DebugEntry{p, C, H, W, lay, false, per_crop};
}

// ---- Pillow resize coefficient tables: resize_tab.h (host-only, also compiled into the CPU test harness) ----
int build_tab(vtd_ctx* c, int in_size, int out_size, ResizeTab* t) {
  std::vector<int> lo, cnt, kk;
  int ksize = 0, maxcnt = 0;
  compute_resize_tab(in_size, out_size, &lo, &cnt, &kk, &ksize, &maxcnt);
  if (t->lo) { cudaFree(t->lo); cudaFree(t->cnt); cudaFree(t->kk); t->lo = t->cnt = t->kk = nullptr; }
  CK(cudaMalloc(&t->lo, out_size * 4)); CK(cudaMalloc(&t->cnt, out_size * 4)); CK(cudaMalloc(&t->kk, kk.size() * 4));
  CK(cudaMemcpyAsync(t->lo, lo.data(), out_size * 4, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(t->cnt, cnt.data(), out_size * 4, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(t->kk, kk.data(), kk.size() * 4, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));   // the host vectors die here
  t->in_size = in_size; t->out_size = out_size; t->ksize = ksize; t->maxcnt = maxcnt;
  return VTD_OK;
}

// ---- detector program ----------------------------------------------------------------------------------
int build_detector(vtd_ctx* c, const SD& sd) {
  const int B = c->cfg.max_batch, dh = c->cfg.det_h, dw = c->cfg.det_w;
  std::vector<Op>& P = c->det_prog;
  int r;
  Act x; x.p = c->pre; x.H = dh; x.W = dw; x.C = 4;
  HostConv hc;
  if ((r = fold_conv(c, sd, "backbone.0", "backbone.1", 4, &hc))) return r;
  Act a;
  if (c->use_win) {
    if (hc.KH != 7 || hc.KW != 7 || hc.Cout != 64) FAIL(VTD_ERR_WEIGHT, "stem must be 7x7, 64 channels");
    // window weights [64][7 rows][8 taps][4 ch]: tap 0 is the extra left pixel of the 16-byte aligned window
    std::vector<float> ww((size_t)64 * 7 * 32, 0.f);
    for (int o = 0; o < 64; ++o)
      for (int rr = 0; rr < 7; ++rr)
        for (int ss = 0; ss < 7; ++ss)
          for (int ch = 0; ch < 4; ++ch)
            ww[(((size_t)o * 7 + rr) * 8 + ss + 1) * 4 + ch] = hc.w[(((size_t)o * 7 + rr) * 7 + ss) * 4 + ch];
    void* wdev = nullptr; float* bdev = nullptr; void* o = nullptr;
    if ((r = upload_act_type(c, ww, &wdev)) || (r = upload_f32(c, hc.b, &bdev)) ||
        (!c->use_stem_pool && (r = dev_alloc(c, &o, (size_t)B * (dh / 2) * (dw / 2) * 64 * c->esz))))
      return r;
    Op op;
    op.kind = Op::CONV;
    op.d.H = dh; op.d.W = dw; op.d.Cin = 4; op.d.Ho = dh / 2; op.d.Wo = dw / 2; op.d.Cout = 64; op.d.KH = op.d.KW = 7;
    op.d.stride = 2; op.d.pad = 3; op.d.N = B;
    std::string e;
    if (c->use_stem_pool) {
      // conv1 + BN + ReLU + MaxPool2d(3, 2, 1) in one kernel: `o` (allocated above for the stem map) is not used
      void* po = nullptr;
      if ((r = dev_alloc(c, &po, (size_t)B * (dh / 4) * (dw / 4) * 64 * c->esz))) return r;
      op.sp = stem_pool_plan_create(c->pre, B, dh, dw, wdev, bdev, po, &e);
      if (!op.sp) FAIL(VTD_ERR_CUDA, "fused stem plan: %s", e.c_str());
      P.push_back(op);
      a.p = po; a.H = dh / 4; a.W = dw / 4; a.C = 64;
    } else {
      op.plan = tc_plan_create_win(c->pre, B, dh + 6, dw + 8, 4, 2, 7, dh / 2, dw / 2, wdev, bdev, o, 1, &e);
      if (!op.plan) FAIL(VTD_ERR_CUDA, "tcgen05 stem plan: %s", e.c_str());
      P.push_back(op);
      a.p = o; a.H = dh / 2; a.W = dw / 2; a.C = 64;
    }
  } else {
    if ((r = add_conv(c, &P, hc, B, x, 2, 3, true, nullptr, RES_NONE, false, &a))) return r;
  }
  if (!c->use_stem_pool && (r = add_pool(c, &P, B, a, 3, 3, 2, 2, 1, 1, &a))) return r;
  const bool r50 = c->cfg.backbone == 50;
  const int nblocks18[4] = {2, 2, 2, 2}, nblocks50[4] = {3, 4, 6, 3};
  Act feats[4];
  for (int li = 0; li < 4; ++li) {
    const int nb = r50 ? nblocks50[li] : nblocks18[li];
    for (int bi = 0; bi < nb; ++bi) {
      const std::string pre = "backbone." + std::to_string(4 + li) + "." + std::to_string(bi);
      const int stride = (bi == 0 && li > 0) ? 2 : 1;
      Act idn = a;
      const bool has_ds = sd.get(pre + ".downsample.0.weight") != nullptr;
      if (has_ds) {
        if ((r = fold_conv(c, sd, pre + ".downsample.0", pre + ".downsample.1", 0, &hc))) return r;
        if ((r = add_conv(c, &P, hc, B, a, stride, 0, false, nullptr, RES_NONE, false, &idn))) return r;
      }
      Act o1, o2, o3;
      if (!r50) {
        if ((r = fold_conv(c, sd, pre + ".conv1", pre + ".bn1", 0, &hc))) return r;
        if ((r = add_conv(c, &P, hc, B, a, stride, 1, true, nullptr, RES_NONE, false, &o1))) return r;
        if ((r = fold_conv(c, sd, pre + ".conv2", pre + ".bn2", 0, &hc))) return r;
        if ((r = add_conv(c, &P, hc, B, o1, 1, 1, true, &idn, RES_SAME, false, &o2))) return r;
        a = o2;
      } else {
        if ((r = fold_conv(c, sd, pre + ".conv1", pre + ".bn1", 0, &hc))) return r;
        if ((r = add_conv(c, &P, hc, B, a, 1, 0, true, nullptr, RES_NONE, false, &o1))) return r;
        if ((r = fold_conv(c, sd, pre + ".conv2", pre + ".bn2", 0, &hc))) return r;
        if ((r = add_conv(c, &P, hc, B, o1, stride, 1, true, nullptr, RES_NONE, false, &o2))) return r;
        if ((r = fold_conv(c, sd, pre + ".conv3", pre + ".bn3", 0, &hc))) return r;
        if ((r = add_conv(c, &P, hc, B, o2, 1, 0, true, &idn, RES_SAME, false, &o3))) return r;
        a = o3;
      }
    }
    feats[li] = a;
  }
  reg_dbg(c, "c2", feats[0]); reg_dbg(c, "c3", feats[1]); reg_dbg(c, "c4", feats[2]); reg_dbg(c, "c5", feats[3]);
  // FPN (text_detector.py:42-56, repaired): lateral i takes C(5-i); only layer_blocks[3] is live
  Act last;
  if ((r = fold_conv(c, sd, "fpn.inner_blocks.0", "", 0, &hc))) return r;
  if ((r = add_conv(c, &P, hc, B, feats[3], 1, 0, false, nullptr, RES_NONE, false, &last))) return r;
  for (int i = 1; i < 4; ++i) {
    if ((r = fold_conv(c, sd, "fpn.inner_blocks." + std::to_string(i), "", 0, &hc))) return r;
    Act nxt;
    if ((r = add_conv(c, &P, hc, B, feats[3 - i], 1, 0, false, &last, RES_UP2, false, &nxt))) return r;
    last = nxt;
  }
  reg_dbg(c, "p2_in", last);
  Act p2;
  if ((r = fold_conv(c, sd, "fpn.layer_blocks.3", "", 0, &hc))) return r;
  if ((r = add_conv(c, &P, hc, B, last, 1, 1, false, nullptr, RES_NONE, false, &p2))) return r;
  reg_dbg(c, "p2", p2);
  // DB head: the two 3x3 convs as one 256->128 conv (BN folded, ReLU)
  HostConv hp, ht;
  if ((r = fold_conv(c, sd, "head.probability_head.0", "head.probability_head.1", 0, &hp))) return r;
  if ((r = fold_conv(c, sd, "head.threshold_head.0", "head.threshold_head.1", 0, &ht))) return r;
  HostConv hm = hp;
  hm.Cout = hp.Cout + ht.Cout;
  hm.w.insert(hm.w.end(), ht.w.begin(), ht.w.end());
  hm.b.insert(hm.b.end(), ht.b.begin(), ht.b.end());
  if (hm.Cout != 128) FAIL(VTD_ERR_WEIGHT, "DB head must be 2 x 64 channels");
  // tail weights: ConvT(64->64,k2,s2)+BN, ConvT(64->1,k2,s2)
  std::vector<float> w1(2 * 256 * 64), b1(2 * 256), w2(2 * 64 * 4), b2(2);
  const char* heads[2] = {"head.probability_head", "head.threshold_head"};
  for (int h = 0; h < 2; ++h) {
    const std::string hn = heads[h];
    const vtd_tensor *w = sd.get(hn + ".3.weight"), *b = sd.get(hn + ".3.bias");
    const vtd_tensor *g = sd.get(hn + ".4.weight"), *be = sd.get(hn + ".4.bias"), *mu = sd.get(hn + ".4.running_mean"),
                     *var = sd.get(hn + ".4.running_var");
    const vtd_tensor *wl = sd.get(hn + ".6.weight"), *bl = sd.get(hn + ".6.bias");
    if (!w || !b || !g || !be || !mu || !var || !wl || !bl || numel(w) != 64 * 64 * 4 || numel(wl) != 64 * 4)
      FAIL(VTD_ERR_WEIGHT, "missing or mis-shaped tensors under '%s'", hn.c_str());
    for (int co = 0; co < 64; ++co) {
      float s = g->data[co] / sqrtf(var->data[co] + 1e-5f);
      float sh = (b->data[co] - mu->data[co]) * s + be->data[co];
      for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx) {
          int row = (dy * 2 + dx) * 64 + co;
          b1[h * 256 + row] = sh;
          for (int ci = 0; ci < 64; ++ci)   // ConvTranspose2d weight is [Cin][Cout][kH][kW]
            w1[((size_t)h * 256 + row) * 64 + ci] = w->data[(((size_t)ci * 64 + co) * 2 + dy) * 2 + dx] * s;
        }
      for (int k = 0; k < 4; ++k) w2[((size_t)h * 64 + co) * 4 + k] = wl->data[(size_t)co * 4 + k];   // [64][1][2][2]
    }
    b2[h] = bl->data[0];
  }
  float *dw1, *db1, *dw2, *db2;
  if ((r = upload_f32(c, w1, &dw1)) || (r = upload_f32(c, b1, &db1)) || (r = upload_f32(c, w2, &dw2)) ||
      (r = upload_f32(c, b2, &db2)))
    return r;
  c->htw.w1 = dw1; c->htw.b1 = db1; c->htw.w2 = dw2; c->htw.b2 = db2;
  void* w1b_fused = nullptr;
  if (c->use_tchead && !(c->cfg.flags & VTD_FLAG_UNFUSED_HEAD)) {
    // one-pass head: the merged 3x3 convolution, both transposed convolutions, sigmoid and mask in one kernel
    Op op;
    op.kind = Op::CONV;
    ConvDesc& d = op.d;
    d.in = p2.p; d.N = B; d.H = p2.H; d.W = p2.W; d.Cin = p2.C; d.KH = d.KW = 3; d.stride = 1; d.pad = 1;
    d.Ho = p2.H; d.Wo = p2.W; d.Cout = 128; d.relu = 1; d.res_mode = RES_NONE; d.out_mode = OUT_NHWC;
    void* wdev = nullptr; float* bdev = nullptr;
    if ((r = upload_act_type(c, hm.w, &wdev)) || (r = upload_f32(c, hm.b, &bdev)) || (r = upload_act_type(c, w1, &w1b_fused))) return r;
    d.w = wdev; d.bias = bdev;
    std::string e;
    op.plan = tc_plan_create_headfused(d, w1b_fused, b1.data(), w2.data(), b2.data(), c->prob, c->thresh, c->mask, &e);
    if (op.plan) { op.fused_head = true; P.push_back(op); c->head_fused = true; return VTD_OK; }
  }
  Act hf;
  if ((r = add_conv(c, &P, hm, B, p2, 1, 1, true, nullptr, RES_NONE, false, &hf))) return r;
  reg_dbg(c, "head", hf);
  c->head_feat = hf.p;
  if (c->use_tchead) {
    void* w1b = w1b_fused;
    if (!w1b && (r = upload_act_type(c, w1, &w1b))) return r;
    std::string e;
    c->head_plan = tc_plan_create_dbhead(c->head_feat, B, dh / 4, dw / 4, w1b, b1.data(), w2.data(), b2.data(), c->prob,
                                         c->thresh, c->mask, &e);
    if (!c->head_plan) FAIL(VTD_ERR_CUDA, "tcgen05 head plan: %s", e.c_str());
  }
  return VTD_OK;
}

// ---- recogniser program -----------------------------------------------------------------------------------
int build_recognizer(vtd_ctx* c, const SD& sd) {
  const int B = c->rc, cw = c->cfg.crop_w;
  std::vector<Op>& P = c->rec_prog;
  int r;
  Act a; a.p = c->crops; a.H = 32; a.W = cw; a.C = 4;
  const bool fuse_pools = c->bf16_mode && !dev_env("VTD_NO_POOL_FUSION");
  struct L { int conv, bn, k, pad; int pool; };   // pool: 0 none, 1 = 2x2 s2, 2 = (2,1) s(2,1)
  const L layers[7] = {{0, 1, 3, 1, 1}, {4, 5, 3, 1, 1}, {8, 9, 3, 1, 0}, {11, 12, 3, 1, 2},
                       {15, 16, 3, 1, 0}, {18, 19, 3, 1, 2}, {22, 23, 2, 0, 0}};
  HostConv hc;
  for (int i = 0; i < 7; ++i) {
    if ((r = fold_conv(c, sd, "cnn." + std::to_string(layers[i].conv), "cnn." + std::to_string(layers[i].bn),
                       i == 0 ? 4 : 0, &hc)))
      return r;
    if (hc.KH != layers[i].k) FAIL(VTD_ERR_WEIGHT, "cnn.%d has an unexpected kernel size", layers[i].conv);
    if (i == 0 && c->use_win) {
      if (hc.Cout != 64) FAIL(VTD_ERR_WEIGHT, "cnn.0 must have 64 output channels");
      // window weights [64][3 rows][4 taps][8 ch]: taps 0..2 are the filter columns, tap 3 and channels 3..7 are zero
      std::vector<float> ww((size_t)64 * 3 * 32, 0.f);
      for (int o = 0; o < 64; ++o)
        for (int rr = 0; rr < 3; ++rr)
          for (int ss = 0; ss < 3; ++ss)
            for (int ch = 0; ch < 3; ++ch)
              ww[(((size_t)o * 3 + rr) * 4 + ss) * 8 + ch] = hc.w[(((size_t)o * 3 + rr) * 3 + ss) * 4 + ch];
      void* wdev = nullptr; float* bdev = nullptr; void* o = nullptr;
      if ((r = upload_act_type(c, ww, &wdev)) || (r = upload_f32(c, hc.b, &bdev)) ||
          (r = dev_alloc(c, &o, (size_t)B * 32 * cw * 64 * c->esz)))
        return r;
      const int fuse = (fuse_pools && layers[i].pool && cw % 2 == 0) ? layers[i].pool : 0;
      Op op;
      op.kind = Op::CONV;
      op.d.H = 32; op.d.W = cw; op.d.Cin = 4; op.d.Ho = 32; op.d.Wo = cw; op.d.Cout = 64; op.d.KH = op.d.KW = 3;
      op.d.stride = 1; op.d.pad = 1; op.d.N = B;
      std::string e;
      op.d.pool = fuse;
      if (fuse == 1 && !(c->cfg.flags & VTD_FLAG_UNFUSED_STEM))      // direct windows + pooling ring: the DBNet stem's kernel
        op.sp = stem_pool_plan_create_crnn(c->crops, B, cw, wdev, bdev, o, &e);
      if (!op.sp) {
        op.plan = tc_plan_create_win(c->crops, B, 34, cw + 4, 8, 1, 3, 32, cw, wdev, bdev, o, 1, &e, fuse);
        if (!op.plan) FAIL(VTD_ERR_CUDA, "tcgen05 CRNN stem plan: %s", e.c_str());
      }
      P.push_back(op);
      a.p = o; a.H = fuse ? 16 : 32; a.W = fuse == 1 ? cw / 2 : cw; a.C = 64;
      if (fuse) continue;
    } else {
      // nn.MaxPool2d after conv+BN+ReLU (text_recognizer.py:17-23): fused into the conv epilogue in the tcgen05 tier
      const int fuse = (fuse_pools && layers[i].pool && a.H % 2 == 0 && (layers[i].pool == 2 || a.W % 2 == 0) &&
                        hc.Cout % 64 == 0 && a.C % 64 == 0) ? layers[i].pool : 0;
      bool fused = false;
      if ((r = add_conv(c, &P, hc, B, a, 1, layers[i].pad, true, nullptr, RES_NONE, false, &a, nullptr, fuse, &fused))) return r;
      if (fused) continue;
    }
    if (layers[i].pool == 1) { if ((r = add_pool(c, &P, B, a, 2, 2, 2, 2, 0, 0, &a))) return r; }
    else if (layers[i].pool == 2) { if ((r = add_pool(c, &P, B, a, 2, 1, 2, 1, 0, 0, &a))) return r; }
  }
  if (a.H != 1 || a.W != c->T || a.C != 512) FAIL(VTD_ERR_WEIGHT, "CRNN conv stack ends at %dx%dx%d", a.H, a.W, a.C);
  reg_dbg(c, "cnn", a, false, true);
  c->seq = a.p;
  const int H = 256;
  Act in = a;
  for (int l = 0; l < 2; ++l) {
    const std::string sfx[2] = {"_l" + std::to_string(l), "_l" + std::to_string(l) + "_reverse"};
    HostConv xp;
    xp.Cout = 8 * H; xp.Cin = 512; xp.Cin_pad = 512; xp.KH = xp.KW = 1;
    xp.w.assign((size_t)8 * H * 512, 0.f); xp.b.assign(8 * H, 0.f);
    std::vector<float> whh((size_t)2 * 4 * H * H);
    for (int d = 0; d < 2; ++d) {
      const vtd_tensor *wih = sd.get("rnn.weight_ih" + sfx[d]), *wh = sd.get("rnn.weight_hh" + sfx[d]),
                       *bi = sd.get("rnn.bias_ih" + sfx[d]), *bh = sd.get("rnn.bias_hh" + sfx[d]);
      if (!wih || !wh || !bi || !bh || numel(wih) != 4 * H * 512 || numel(wh) != 4 * H * H || numel(bi) != 4 * H ||
          numel(bh) != 4 * H)
        FAIL(VTD_ERR_WEIGHT, "missing or mis-shaped LSTM tensors '%s'", sfx[d].c_str());
      // row order of the gate dimension: PyTorch's (gate, unit), or -- for the tcgen05 recurrence -- (unit tile of
      // 64, gate, unit in tile), so that one 256-column MMA tile holds i,f,g,o of the same 64 hidden units
      for (int g = 0; g < 4; ++g)
        for (int u = 0; u < H; ++u) {
          const int src = g * H + u;
          const int dst = c->use_tclstm ? (u / 64) * 256 + g * 64 + (u % 64) : src;
          memcpy(&xp.w[((size_t)d * 4 * H + dst) * 512], wih->data + (size_t)src * 512, sizeof(float) * 512);
          xp.b[d * 4 * H + dst] = bi->data[src] + bh->data[src];
          memcpy(&whh[((size_t)d * 4 * H + dst) * H], wh->data + (size_t)src * H, sizeof(float) * H);
        }
    }
    Act xo;
    if ((r = add_conv(c, nullptr, xp, B, in, 1, 0, false, nullptr, RES_NONE, /*out_f32=*/!c->use_plstm, &xo, &c->xproj_op[l])))
      return r;
    if (l == 0) c->xproj = (float*)xo.p;
    if ((r = upload_act_type(c, whh, &c->whh[l]))) return r;
    if ((r = dev_alloc(c, &c->rnn_out[l], (size_t)B * c->T * 2 * H * c->esz))) return r;
    in.p = c->rnn_out[l]; in.H = 1; in.W = c->T; in.C = 2 * H;
    reg_dbg(c, l == 0 ? "rnn0" : "rnn1", in, false, true);
  }
  if ((r = dalloc(c, &c->hbuf, (size_t)2 * 2 * B * H * 4)) || (r = dalloc(c, &c->cbuf, (size_t)2 * B * H * 4))) return r;
  if (c->use_tclstm && c->use_plstm) {
    for (int l = 0; l < 2; ++l) {
      std::string e;
      c->plstm[l] = lstm_plan_create(c->whh[l], &e);
      if (!c->plstm[l]) FAIL(VTD_ERR_CUDA, "persistent LSTM plan: %s", e.c_str());
    }
  } else if (c->use_tclstm) {
    if ((r = dalloc(c, &c->h16, (size_t)2 * 2 * B * H * 2))) return r;
    for (int l = 0; l < 2; ++l)
      for (int pp = 0; pp < 2; ++pp) {
        std::string e;
        c->lstm_plan[l][pp] = tc_plan_create_lstm(c->h16 + (size_t)pp * 2 * B * H, c->h16 + (size_t)(pp ^ 1) * 2 * B * H, B,
                                                  c->whh[l], (const float*)c->xproj_op[l].d.out, c->cbuf, c->rnn_out[l],
                                                  c->T, &e);
        if (!c->lstm_plan[l][pp]) FAIL(VTD_ERR_CUDA, "tcgen05 LSTM plan: %s", e.c_str());
      }
  }
  HostConv fc;
  {
    const vtd_tensor *w = sd.get("classifier.weight"), *b = sd.get("classifier.bias");
    if (!w || !b || w->ndim != 2 || w->shape[1] != 2 * H || numel(b) != w->shape[0])
      FAIL(VTD_ERR_WEIGHT, "missing or mis-shaped classifier");
    if (w->shape[0] != 97) FAIL(VTD_ERR_WEIGHT, "classifier must have 97 outputs (text_recognizer.py:81)");
    // bf16 tier: pad the 97 classes to 128 zero rows so the GEMM runs on the tcgen05 path; the decode reads 97
    c->logits_ld = c->bf16_mode ? 128 : 97;
    fc.Cout = c->logits_ld; fc.Cin = fc.Cin_pad = 2 * H; fc.KH = fc.KW = 1;
    fc.w.assign((size_t)fc.Cout * 2 * H, 0.f); fc.b.assign(fc.Cout, 0.f);
    memcpy(fc.w.data(), w->data, sizeof(float) * 97 * 2 * H);
    memcpy(fc.b.data(), b->data, sizeof(float) * 97);
  }
  Act lo;
  if ((r = add_conv(c, nullptr, fc, B, in, 1, 0, false, nullptr, RES_NONE, true, &lo, &c->fc_op))) return r;
  c->logits = (float*)lo.p;
  c->dbg["logits"] = DebugEntry{lo.p, 97, 1, c->T, dense_layout(1, c->T, c->logits_ld), true, true};
  // every launch of the recogniser on the tcgen05 / persistent-LSTM kernels: they take the crop count from device memory
  bool dyn = c->bf16_mode && c->use_tclstm && c->use_plstm && c->xproj_op[0].plan && c->xproj_op[1].plan && c->fc_op.plan;
  for (const Op& o : c->rec_prog) dyn = dyn && o.kind == Op::CONV && (o.plan || o.sp);
  c->rec_dyn_ok = dyn;
  return VTD_OK;
}

// crops [0,nc) already in c->crops -> logits in c->logits
int run_crnn(vtd_ctx* c, int nc) {
  NvtxRange range("vtd.crnn");
  int r = run_prog(c, c->rec_prog, nc); if (r) return r;
  for (int l = 0; l < 2; ++l) {
    // the second layer's xproj reuses its own buffer (allocated by add_conv)
    { int r2 = run_op_prof(c, c->xproj_op[l], nc); if (r2) return r2; }
    const float* xp = (const float*)c->xproj_op[l].d.out;
    StageTimer st(c, l == 0 ? ST_LSTM0 : ST_LSTM1);
    if (c->use_tclstm && c->use_plstm) {
      CK(bilstm_layer_tcgen05(c->plstm[l], c->xproj_op[l].d.out, c->rnn_out[l], nc, c->T, c->stream, &c->lc, c->n_dyn, c->n_first));
    } else if (c->use_tclstm) {
      CK(cudaMemsetAsync(c->h16, 0, (size_t)2 * c->rc * 256 * 2, c->stream));          // h_0 = 0 (parity 0)
      CK(cudaMemsetAsync(c->cbuf, 0, (size_t)2 * c->rc * 256 * 4, c->stream));         // c_0 = 0
      for (int step = 0; step < c->T; ++step) CK(lstm_step_tcgen05(c->lstm_plan[l][step & 1], nc, step, c->stream, &c->lc));
    } else if (c->bf16_mode)
      CK((bilstm_layer<bf16, bf16>(xp, (const bf16*)c->whh[l], (bf16*)c->rnn_out[l], c->hbuf, c->cbuf, nc, c->T, 256,
                                   c->stream, &c->lc)));
    else
      CK((bilstm_layer<float, float>(xp, (const float*)c->whh[l], (float*)c->rnn_out[l], c->hbuf, c->cbuf, nc, c->T, 256,
                                     c->stream, &c->lc)));
  }
  { int r2 = run_op_prof(c, c->fc_op, nc); if (r2) return r2; }
  return VTD_OK;
}

int detect_maps_locked(vtd_ctx* c, int n, float thr, const float* logit_bias) {
  c->cur_thr = thr; c->cur_bias = logit_bias;
  NvtxRange range("vtd.detector");
  int r = run_prog(c, c->det_prog, n); if (r) return r;
  if (c->head_fused) return VTD_OK;       // the last op of the program was the whole head
  const int dh = c->cfg.det_h, dw = c->cfg.det_w;
  StageTimer st(c, ST_HEAD_TAIL);
  if (c->head_plan)
    CK(dbhead_tcgen05(c->head_plan, n, thr, logit_bias, c->stream, &c->lc));
  else if (c->bf16_mode)
    CK(db_head_tail<bf16>((const bf16*)c->head_feat, c->htw, n, dh / 4, dw / 4, logit_bias, thr, c->prob, c->thresh,
                          c->mask, c->stream, &c->lc));
  else
    CK(db_head_tail<float>((const float*)c->head_feat, c->htw, n, dh / 4, dw / 4, logit_bias, thr, c->prob, c->thresh,
                           c->mask, c->stream, &c->lc));
  return VTD_OK;
}

int preprocess_locked(vtd_ctx* c, const uint8_t* const* frames, int n, int h, int w, int pitch, int pixfmt,
                      int on_device) {
  if (n <= 0 || n > c->cfg.max_batch) FAIL(VTD_ERR_CAPACITY, "n=%d outside 1..max_batch=%d", n, c->cfg.max_batch);
  if (h <= 0 || w <= 0 || h > c->cfg.max_src_h || w > c->cfg.max_src_w)
    FAIL(VTD_ERR_CAPACITY, "frame %dx%d exceeds max_src %dx%d", h, w, c->cfg.max_src_h, c->cfg.max_src_w);
  if (pixfmt != VTD_PIX_BGR && pixfmt != VTD_PIX_NV12) FAIL(VTD_ERR_ARG, "unknown pixel format %d", pixfmt);
  if (pixfmt == VTD_PIX_NV12 && ((h | w) & 1)) FAIL(VTD_ERR_ARG, "NV12 frames need even dimensions");
  const int row_bytes = pixfmt == VTD_PIX_BGR ? w * 3 : w;
  if (pitch < row_bytes) FAIL(VTD_ERR_ARG, "pitch %d smaller than a row (%d bytes)", pitch, row_bytes);
  const int rows = pixfmt == VTD_PIX_BGR ? h : h + h / 2;
  int dev_pitch = pitch;
  if (!on_device) {
    dev_pitch = row_bytes;
    const size_t fb = (size_t)rows * row_bytes;
    if (fb > c->frame_bytes_cap) FAIL(VTD_ERR_CAPACITY, "frame of %zu bytes exceeds the staging slot", fb);
    for (int i = 0; i < n; ++i) if (!frames[i]) FAIL(VTD_ERR_ARG, "frame %d is a null pointer", i);
    // page-locked frames (cudaHostAlloc / cudaHostRegister / torch pin_memory) go straight to the copy engine
    cudaPointerAttributes pa;
    const bool pinned = cudaPointerGetAttributes(&pa, frames[0]) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) {
      for (int i = 0; i < n; ++i) {
        uint8_t* dst = c->frames_store + (size_t)i * c->frame_bytes_cap;
        if (pitch == row_bytes) CK(cudaMemcpyAsync(dst, frames[i], fb, cudaMemcpyHostToDevice, c->stream));
        else CK(cudaMemcpy2DAsync(dst, row_bytes, frames[i], pitch, row_bytes, rows, cudaMemcpyHostToDevice, c->stream));
      }
    } else {
      // pageable frames: copy threads pack frame i into its pinned slot and issue its DMA at once, so the host copy of
      // frame i+1 overlaps the transfer of frame i
      const size_t need = fb * (size_t)n;
      if (c->host_stage_bytes < need) {
        if (c->host_stage) { cudaFreeHost(c->host_stage); c->host_stage = nullptr; c->host_stage_bytes = 0; }
        CK(cudaHostAlloc(reinterpret_cast<void**>(&c->host_stage), fb * (size_t)c->cfg.max_batch, cudaHostAllocDefault));
        c->host_stage_bytes = fb * (size_t)c->cfg.max_batch;
      }
      if (!c->host_stage_event) CK(cudaEventCreateWithFlags(&c->host_stage_event, cudaEventDisableTiming));
      else CK(cudaEventSynchronize(c->host_stage_event));       // the previous batch's DMAs have read the slots
      const int nthreads = n < 8 ? n : 8;     // ~8-10 GB/s per copying core; a 16-frame 1080p batch is 100 MB
      std::atomic<int> next{0};
      std::atomic<int> failed{0};
      auto work = [&]() {
        cudaSetDevice(c->cfg.device);
        for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) {
          uint8_t* slot = c->host_stage + (size_t)i * fb;
          if (pitch == row_bytes) memcpy(slot, frames[i], fb);
          else for (int y = 0; y < rows; ++y) memcpy(slot + (size_t)y * row_bytes, frames[i] + (size_t)y * pitch, row_bytes);
          if (cudaMemcpyAsync(c->frames_store + (size_t)i * c->frame_bytes_cap, slot, fb, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
            failed.store(1);
        }
      };
      std::vector<std::thread> pool;
      for (int t = 1; t < nthreads; ++t) pool.emplace_back(work);
      work();
      for (std::thread& t : pool) t.join();
      if (failed.load()) { cudaGetLastError(); FAIL(VTD_ERR_CUDA, "host->device copy of a staged frame failed"); }
      CK(cudaEventRecord(c->host_stage_event, c->stream));
    }
    c->frame_ptrs_dev = c->store_ptrs_dev;
  } else {
    CK(cudaEventSynchronize(c->ptrs_event));      // the previous upload has consumed the pinned array
    for (int i = 0; i < n; ++i) c->frame_ptrs_pinned[i] = frames[i];
    CK(cudaMemcpyAsync(c->ext_ptrs_dev, c->frame_ptrs_pinned, sizeof(void*) * n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaEventRecord(c->ptrs_event, c->stream));
    c->frame_ptrs_dev = c->ext_ptrs_dev;
  }
  if (c->tab_h != h || c->tab_w != w) {
    int r = build_tab(c, w, c->cfg.det_w, &c->tx); if (r) return r;
    r = build_tab(c, h, c->cfg.det_h, &c->ty); if (r) return r;
    c->tab_h = h; c->tab_w = w;
  }
  StageTimer st(c, ST_PREPROCESS);
  if (c->bf16_mode)
    CK(preprocess_frames<bf16>(c->frame_ptrs_dev, n, h, w, dev_pitch, pixfmt, c->tx, c->ty, c->norm_lut, (bf16*)c->pre,
                               c->pre_lay, c->stream, &c->lc));
  else
    CK(preprocess_frames<float>(c->frame_ptrs_dev, n, h, w, dev_pitch, pixfmt, c->tx, c->ty, c->norm_lut, (float*)c->pre,
                                c->pre_lay, c->stream, &c->lc));
  c->cur_h = h; c->cur_w = w; c->cur_pitch = dev_pitch; c->cur_n = n; c->cur_pix = pixfmt;
  return VTD_OK;
}

int extract_locked(vtd_ctx* c, int n, int orig_h, int orig_w) {
  BoxParams bp;
  bp.n = n; bp.n_alloc = c->cfg.max_batch; bp.mh = c->cfg.det_h; bp.mw = c->cfg.det_w;
  bp.clip_h = c->cfg.det_h; bp.clip_w = c->cfg.det_w; bp.orig_h = orig_h; bp.orig_w = orig_w;
  bp.kmax = c->cfg.max_boxes; bp.unclip = c->cfg.unclip_ratio;
  c->overflow_cached = -1;
  StageTimer st(c, ST_BOXES);
  CK(extract_boxes(c->prob, c->mask, bp, c->box_work, c->box_lay, c->records, c->counts, c->stream, &c->lc));
  return VTD_OK;
}

// crops [first, first + nc) of the current batch -> crop gather, CRNN, CTC into the records
int recognize_chunk(vtd_ctx* c, int n, int first, int nc) {
  {
    StageTimer st(c, ST_CROP);
    if (c->bf16_mode)
      CK(crop_resize_records<bf16>(c->frame_ptrs_dev, c->cur_h, c->cur_w, c->cur_pitch, c->records, c->offsets, n,
                                   c->cfg.max_boxes, first, nc, c->cfg.crop_w, c->cur_pix == VTD_PIX_NV12, (bf16*)c->crops, c->crops_lay,
                                   c->stream, &c->lc));
    else
      CK(crop_resize_records<float>(c->frame_ptrs_dev, c->cur_h, c->cur_w, c->cur_pitch, c->records, c->offsets, n,
                                    c->cfg.max_boxes, first, nc, c->cfg.crop_w, c->cur_pix == VTD_PIX_NV12, (float*)c->crops, c->crops_lay,
                                    c->stream, &c->lc));
  }
  int r = run_crnn(c, nc); if (r) return r;
  StageTimer st(c, ST_CTC);
  CK(ctc_into_records(c->logits, nc, first, c->T, 97, c->logits_ld, c->cfg.canonical_ctc, c->offsets, n, c->cfg.max_boxes,
                      c->records, c->stream, &c->lc));
  return VTD_OK;
}

// The batch's crop total is looked at here, at a point where the host waits for the stream anyway: whatever lies beyond the
// first chunk (more than `rc` crops in the batch) is recognised now.  Every entry point that reads results or reuses the
// batch's frames calls this first.
int finish_pending(vtd_ctx* c) {
  if (!c->rec_pending) return VTD_OK;
  c->rec_pending = false;
  CK(cudaEventSynchronize(c->total_event));
  const int total = c->pinned_int[0], n = c->pending_n;
  for (int first = c->rc; first < total; first += c->rc) {
    const int nc = total - first < c->rc ? total - first : c->rc;
    int r = recognize_chunk(c, n, first, nc); if (r) return r;
  }
  return VTD_OK;
}

int recognize_locked(vtd_ctx* c, int n) {
  CK(scan_counts(c->counts, n, c->offsets, c->stream, &c->lc));
  CK(cudaMemcpyAsync(c->pinned_int, c->offsets + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (c->rec_dyn_ok) {
    // first chunk against the device-side total: launches sized for what the batch can hold at most, the kernels process
    // what it does hold; no synchronisation here
    if (!c->total_event) CK(cudaEventCreateWithFlags(&c->total_event, cudaEventDisableTiming));
    CK(cudaEventRecord(c->total_event, c->stream));
    const long long most = (long long)n * c->cfg.max_boxes;
    const int nc = most < c->rc ? (int)most : c->rc;
    c->n_dyn = c->offsets + n; c->n_first = 0;
    const int r = recognize_chunk(c, n, 0, nc);
    c->n_dyn = nullptr;
    if (r) return r;
    c->rec_pending = most > c->rc;                      // only then can there be a second chunk
    c->pending_n = n;
    return VTD_OK;
  }
  CK(cudaStreamSynchronize(c->stream));
  const int total = c->pinned_int[0];
  for (int first = 0; first < total; first += c->rc) {
    const int nc = total - first < c->rc ? total - first : c->rc;
    int r = recognize_chunk(c, n, first, nc); if (r) return r;
  }
  return VTD_OK;
}

int read_records_locked(vtd_ctx* c, int n, vtd_record* rh, int* ch) {
  { int fr = finish_pending(c); if (fr) return fr; }
  // the overflow flag rides along with the read-back (same synchronisation): vtd_overflow_flag() then costs nothing
  CK(cudaMemcpyAsync(c->pinned_int + 1, c->box_work + c->box_lay.overflow, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (ch) CK(cudaMemcpyAsync(ch, c->counts, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
  if (rh) CK(cudaMemcpyAsync(rh, c->records, sizeof(vtd_record) * (size_t)n * c->cfg.max_boxes, cudaMemcpyDeviceToHost,
                             c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->overflow_cached = c->pinned_int[1];
  return VTD_OK;
}

// Serialises the calls on a context and makes its device current for their duration; the caller's current device
// (torch's, for a Python caller) is put back on the way out.
#include "trocr_host.inc"

TrocrState* trocr_of(vtd_ctx* c) {
  if (!c->trocr) c->trocr = new TrocrState();
  return static_cast<TrocrState*>(c->trocr);
}

void trocr_free(vtd_ctx* c) {
  if (!c->trocr) return;
  TrocrState* t = static_cast<TrocrState*>(c->trocr);
  auto kill = [](Op& o) { if (o.plan) tc_plan_destroy(o.plan); o.plan = nullptr; };
  kill(t->patch_op); kill(t->lm);
  for (auto& e : t->el) { kill(e.qkv); kill(e.proj); kill(e.fc1); kill(e.fc2); }
  for (auto& d : t->dl) { kill(d.qkv); kill(d.so); kill(d.cq); kill(d.co); kill(d.fc1); kill(d.fc2); kill(d.ckv); }
  for (auto& kv : t->step_graphs) if (kv.second) cudaGraphExecDestroy(kv.second);
  if (t->pinned) cudaFreeHost(t->pinned);
  if (t->crop_store) cudaFree(t->crop_store);
  if (t->tmp) cudaFree(t->tmp);
  if (t->tab_dev) cudaFree(t->tab_dev);
  if (t->px_stage) cudaFree(t->px_stage);
  delete t;
  c->trocr = nullptr;
}

struct Guard {
  vtd_ctx* c; std::lock_guard<std::mutex> lk; int prev = -1;
  explicit Guard(vtd_ctx* ctx) : c(ctx), lk(ctx->mu) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != ctx->cfg.device) cudaSetDevice(ctx->cfg.device); else prev = -1;
  }
  ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};

int make_sd(vtd_ctx* c, const vtd_tensor* t, int n, SD* sd) {
  if (!t || n <= 0) FAIL(VTD_ERR_ARG, "empty state dict");
  for (int i = 0; i < n; ++i) {
    if (!t[i].name || !t[i].data) FAIL(VTD_ERR_ARG, "state dict entry %d has a null name or data pointer", i);
    sd->m[t[i].name] = &t[i];
  }
  return VTD_OK;
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int vtd_abi_version(void) { return 1; }

const char* vtd_last_error(vtd_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int vtd_create(vtd_ctx** out, const vtd_config* cfg) {
  std::lock_guard<std::mutex> lk(g_create_mutex);
  auto bad = [&](int code, const std::string& m) { g_create_error = m; if (out) *out = nullptr; return code; };
  if (!out || !cfg) return bad(VTD_ERR_ARG, "null argument");
  if (cfg->backbone != 18 && cfg->backbone != 50) return bad(VTD_ERR_ARG, "backbone must be 18 or 50");
  if (cfg->dtype != VTD_FP32 && cfg->dtype != VTD_16BIT) return bad(VTD_ERR_ARG, "dtype must be VTD_FP32 or VTD_16BIT");
  if (cfg->det_h <= 0 || cfg->det_w <= 0 || cfg->det_h % 32 || cfg->det_w % 32)
    return bad(VTD_ERR_ARG, "det_h/det_w must be positive multiples of 32");
  if (cfg->crop_w < 16 || cfg->crop_w % 4 || cfg->crop_w / 4 - 1 > 36) return bad(VTD_ERR_ARG, "crop_w must be a multiple of 4 in [16,148]");
  if (cfg->max_batch <= 0 || cfg->max_batch > 256) return bad(VTD_ERR_ARG, "max_batch must be in 1..256");
  if (cfg->max_boxes <= 0 || cfg->max_boxes > 1024) return bad(VTD_ERR_ARG, "max_boxes must be in 1..1024");
  if (cfg->max_src_h <= 0 || cfg->max_src_w <= 0) return bad(VTD_ERR_ARG, "max_src_h/max_src_w must be positive");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return bad(VTD_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return bad(VTD_ERR_ARG, "device ordinal out of range");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) return bad(VTD_ERR_CUDA, cudaGetErrorString(e));
  if (prop.major != 10)
    return bad(VTD_ERR_CUDA, std::string("device '") + prop.name + "' is not sm_100 (this library is built for sm_100a only)");
  int prev_dev = -1;
  cudaGetDevice(&prev_dev);
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev == cfg->device ? -1 : prev_dev};
  if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return bad(VTD_ERR_CUDA, cudaGetErrorString(e));

  vtd_ctx* c = new vtd_ctx();
  c->cfg = *cfg;
  if (!(c->cfg.unclip_ratio > 0.f)) c->cfg.unclip_ratio = 1.0f;
  c->bf16_mode = cfg->dtype == VTD_16BIT;      // the 16-bit speed tier (half or bfloat16 storage, common.cuh)
  c->esz = c->bf16_mode ? 2 : 4;
  c->T = cfg->crop_w / 4 - 1;
  c->use_win = c->bf16_mode && !dev_env("VTD_NO_WIN");
  c->use_stem_pool = c->use_win && cfg->det_w / 2 >= 128 && !(cfg->flags & VTD_FLAG_UNFUSED_STEM);
  c->use_tchead = c->bf16_mode && !dev_env("VTD_NO_TCHEAD");
  c->use_tclstm = c->bf16_mode && !dev_env("VTD_NO_TCLSTM");
  c->use_plstm = c->use_tclstm && !dev_env("VTD_NO_PLSTM");
  long long want = (long long)cfg->max_batch * cfg->max_boxes;
  c->rc = (int)(want < 1024 ? want : 1024);
  auto fail = [&](int code) { g_create_error = c->err; vtd_destroy(c); *out = nullptr; return code; };
  if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
    c->err = cudaGetErrorString(e); return fail(VTD_ERR_CUDA);
  }
  c->stream = c->own_stream;
  const int B = cfg->max_batch, dh = cfg->det_h, dw = cfg->det_w;
  const size_t px = (size_t)dh * dw;
  int r;
  c->frame_bytes_cap = (((size_t)cfg->max_src_h * cfg->max_src_w * 3) + 255) & ~(size_t)255;
  if ((r = dalloc(c, &c->frames_store, c->frame_bytes_cap * B)) || (r = dalloc(c, &c->store_ptrs_dev, sizeof(void*) * B)) ||
      (r = dalloc(c, &c->ext_ptrs_dev, sizeof(void*) * B)) ||
      (r = dev_alloc(c, &c->pre, (size_t)B * (c->use_win ? (size_t)(dh + 6) * (dw + (c->use_stem_pool ? 10 : 8)) : px) * 4 * c->esz + 4096 /* the stem's row copies overhang */)) || (r = dalloc(c, &c->prob, (size_t)B * px * 4)) ||
      (r = dalloc(c, &c->thresh, (size_t)B * px * 4)) || (r = dalloc(c, &c->mask, (size_t)B * px)) ||
      // records and counts are ONE block (counts directly after the records): a rank's results travel in one collective
      (r = dalloc(c, &c->records, sizeof(vtd_record) * (size_t)B * cfg->max_boxes + sizeof(int) * B)) ||
      (r = dalloc(c, &c->offsets, sizeof(int) * (B + 1))) ||
      (r = dev_alloc(c, &c->crops, (size_t)c->rc * (c->use_win ? (size_t)34 * (cfg->crop_w + 4) * 8 : (size_t)32 * cfg->crop_w * 4) * c->esz + 4096 /* row copies overhang */)) ||
      (r = dalloc(c, &c->ids_dev, (size_t)c->rc * VTD_IDS_STRIDE)) || (r = dalloc(c, &c->len_dev, (size_t)c->rc * 4)) ||
      (r = dalloc(c, &c->conf_dev, (size_t)c->rc * 4)) || (r = dalloc(c, &c->list_ptrs, sizeof(void*) * c->rc)) ||
      (r = dalloc(c, &c->list_meta, sizeof(int) * 3 * c->rc)))
    return fail(r);
  {
    const int kc = cfg->max_boxes * 4 < 1024 ? 1024 : cfg->max_boxes * 4;
    size_t bytes = box_work_bytes(B, dh, dw, kc, &c->box_lay);
    if ((r = dalloc(c, &c->box_work, bytes))) return fail(r);
    cudaMemset(c->box_work, 0, bytes);
  }
  if (cudaMallocHost(&c->frame_ptrs_pinned, sizeof(void*) * B) != cudaSuccess ||
      cudaMallocHost(&c->pinned_int, sizeof(int) * (B + 2)) != cudaSuccess) {
    c->err = "cudaMallocHost failed"; return fail(VTD_ERR_CUDA);
  }
  if (cudaEventCreateWithFlags(&c->ptrs_event, cudaEventDisableTiming) != cudaSuccess) {
    c->err = "cudaEventCreate failed"; return fail(VTD_ERR_CUDA);
  }
  for (int i = 0; i < B; ++i) c->frame_ptrs_pinned[i] = c->frames_store + (size_t)i * c->frame_bytes_cap;
  cudaMemcpy(c->store_ptrs_dev, c->frame_ptrs_pinned, sizeof(void*) * B, cudaMemcpyHostToDevice);
  c->frame_ptrs_dev = c->store_ptrs_dev;
  c->counts = reinterpret_cast<int*>(c->records + (size_t)B * cfg->max_boxes);
  cudaMemset(c->counts, 0, sizeof(int) * B);
  if ((r = dalloc(c, &c->norm_lut, 768 * sizeof(float)))) return fail(r);
  if (build_normalize_lut(c->norm_lut, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) {
    c->err = "normalisation table kernel failed"; return fail(VTD_ERR_CUDA);
  }
  if (c->use_win) {   // zero-bordered stem inputs: 4 px left/right + 3 rows top/bottom (7x7 s2), 1 px + 1 row (3x3)
    c->pre_lay = padded_layout(dh, dw, 4, 3, 3, c->use_stem_pool ? 6 : 4, 4);
    c->crops_lay = padded_layout(32, cfg->crop_w, 8, 1, 1, 1, 3);
    cudaMemset(c->pre, 0, (size_t)B * (dh + 6) * (dw + (c->use_stem_pool ? 10 : 8)) * 4 * c->esz);
    cudaMemset(c->crops, 0, (size_t)c->rc * 34 * (cfg->crop_w + 4) * 8 * c->esz);
  } else {
    c->pre_lay = dense_layout(dh, dw, 4);
    c->crops_lay = dense_layout(32, cfg->crop_w, 4);
  }
  reg_dbg_lay(c, "input", c->pre, 4, dh, dw, c->pre_lay, false);
  reg_dbg_lay(c, "crops", c->crops, 4, 32, cfg->crop_w, c->crops_lay, true);
  *out = c;
  return VTD_OK;
}

void vtd_destroy(vtd_ctx* c) {
  if (!c) return;
  int prev_dev = -1;
  cudaGetDevice(&prev_dev);
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev == c->cfg.device ? -1 : prev_dev};
  cudaSetDevice(c->cfg.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (Op& op : c->det_prog) { if (op.plan) tc_plan_destroy(op.plan); if (op.sp) stem_pool_plan_destroy(op.sp); }
  for (Op& op : c->rec_prog) { if (op.plan) tc_plan_destroy(op.plan); if (op.sp) stem_pool_plan_destroy(op.sp); }
  for (int l = 0; l < 2; ++l) if (c->xproj_op[l].plan) tc_plan_destroy(c->xproj_op[l].plan);
  if (c->fc_op.plan) tc_plan_destroy(c->fc_op.plan);
  if (c->head_plan) tc_plan_destroy(c->head_plan);
  trocr_free(c);
  for (int l = 0; l < 2; ++l) if (c->plstm[l]) lstm_plan_destroy(c->plstm[l]);
  for (int l = 0; l < 2; ++l) for (int pp = 0; pp < 2; ++pp) if (c->lstm_plan[l][pp]) tc_plan_destroy(c->lstm_plan[l][pp]);
  for (void* p : c->allocs) cudaFree(p);
  for (ResizeTab* t : {&c->tx, &c->ty}) if (t->lo) { cudaFree(t->lo); cudaFree(t->cnt); cudaFree(t->kk); }
  if (c->pp_work) { cudaFree(c->pp_work); cudaFree(c->pp_prob); cudaFree(c->pp_mask); cudaFree(c->pp_records); cudaFree(c->pp_counts); }
  if (c->stage_f32) cudaFree(c->stage_f32);
  if (c->list_store) cudaFree(c->list_store);
  if (c->host_stage) cudaFreeHost(c->host_stage);
  if (c->host_stage_event) cudaEventDestroy(c->host_stage_event);
  if (c->frame_ptrs_pinned) cudaFreeHost(c->frame_ptrs_pinned);
  if (c->pinned_int) cudaFreeHost(c->pinned_int);
  if (c->ptrs_event) cudaEventDestroy(c->ptrs_event);
  if (c->total_event) cudaEventDestroy(c->total_event);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

int vtd_set_stream(vtd_ctx* c, void* s) {
  if (!c) return VTD_ERR_ARG;
  Guard g(c);
  CK(cudaStreamSynchronize(c->stream));
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return VTD_OK;
}
void* vtd_stream(vtd_ctx* c) { return c ? (void*)c->stream : nullptr; }
int vtd_sync(vtd_ctx* c) {
  if (!c) return VTD_ERR_ARG;
  Guard g(c);
  { int fr = finish_pending(c); if (fr) return fr; }
  CK(cudaStreamSynchronize(c->stream));
  return VTD_OK;
}
int64_t vtd_launch_count(vtd_ctx* c) { return c ? c->lc.n : 0; }
int vtd_time_T(vtd_ctx* c) { return c ? c->T : 0; }
int vtd_overflow_flag(vtd_ctx* c) {
  if (!c) return 0;
  Guard g(c);
  if (c->overflow_cached >= 0) return c->overflow_cached;
  int v = 0;
  cudaMemcpyAsync(&v, c->box_work + c->box_lay.overflow, 4, cudaMemcpyDeviceToHost, c->stream);
  cudaStreamSynchronize(c->stream);
  return v;
}

int vtd_check_guards(vtd_ctx* c, int64_t* bad_bytes) {
  if (!c || !bad_bytes) return VTD_ERR_ARG;
  Guard g(c);
  *bad_bytes = 0;
  if (!(c->cfg.flags & VTD_FLAG_GUARD_ALLOCS)) FAIL(VTD_ERR_STATE, "context was not created with VTD_FLAG_GUARD_ALLOCS");
  CK(cudaStreamSynchronize(c->stream));
  std::vector<uint8_t> h(2 * GUARD_BYTES + 256);
  int first = -1;
  for (size_t i = 0; i < c->guarded.size(); ++i) {
    const vtd_ctx::Guarded& a = c->guarded[i];
    const size_t padded = (a.bytes + 255) & ~(size_t)255;
    const size_t tail = padded - a.bytes + GUARD_BYTES;
    CK(cudaMemcpy(h.data(), a.user - GUARD_BYTES, GUARD_BYTES, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h.data() + GUARD_BYTES, a.user + a.bytes, tail, cudaMemcpyDeviceToHost));
    int64_t bad = 0;
    for (size_t k = 0; k < GUARD_BYTES + tail; ++k) bad += h[k] != GUARD_PATTERN;
    if (bad && first < 0) first = (int)i;
    *bad_bytes += bad;
  }
  if (*bad_bytes) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%lld canary bytes overwritten; first damaged allocation: #%d of %zu (%zu bytes)", (long long)*bad_bytes,
             first, c->guarded.size(), c->guarded[first].bytes);
    c->err = buf;
  }
  return VTD_OK;
}

int vtd_load_detector(vtd_ctx* c, const vtd_tensor* t, int n) {
  if (!c) return VTD_ERR_ARG;
  Guard g(c);
  if (c->det_loaded) FAIL(VTD_ERR_STATE, "detector weights already loaded (create a new context to reload)");
  SD sd; int r = make_sd(c, t, n, &sd); if (r) return r;
  r = build_detector(c, sd); if (r) return r;
  CK(cudaDeviceSynchronize());
  c->det_loaded = true;
  return VTD_OK;
}

int vtd_load_recognizer(vtd_ctx* c, const vtd_tensor* t, int n) {
  if (!c) return VTD_ERR_ARG;
  Guard g(c);
  if (c->rec_loaded) FAIL(VTD_ERR_STATE, "recogniser weights already loaded (create a new context to reload)");
  SD sd; int r = make_sd(c, t, n, &sd); if (r) return r;
  r = build_recognizer(c, sd); if (r) return r;
  CK(cudaDeviceSynchronize());
  c->rec_loaded = true;
  return VTD_OK;
}

int vtd_preprocess(vtd_ctx* c, const uint8_t* const* frames, int n, int h, int w, int pitch, int pixfmt, int on_dev) {
  if (!c || !frames) return VTD_ERR_ARG;
  Guard g(c);
  { int fr = finish_pending(c); if (fr) return fr; }
  return preprocess_locked(c, frames, n, h, w, pitch, pixfmt, on_dev);
}

int vtd_detect_maps(vtd_ctx* c, int n, float thr, const float* logit_bias) {
  if (!c) return VTD_ERR_ARG;
  Guard g(c);
  { int fr = finish_pending(c); if (fr) return fr; }
  if (!c->det_loaded) FAIL(VTD_ERR_STATE, "vtd_load_detector has not been called");
  if (n <= 0 || n > c->cfg.max_batch) FAIL(VTD_ERR_CAPACITY, "n=%d outside 1..max_batch=%d", n, c->cfg.max_batch);
  return detect_maps_locked(c, n, thr, logit_bias);
}

int vtd_get_maps(vtd_ctx* c, float** p, float** t, uint8_t** m) {
  if (!c) return VTD_ERR_ARG;
  if (p) *p = c->prob; if (t) *t = c->thresh; if (m) *m = c->mask;
  return VTD_OK;
}

int vtd_read_maps(vtd_ctx* c, int n, float* ph, float* th, uint8_t* mh) {
  if (!c) return VTD_ERR_ARG;
  Guard g(c);
  if (n <= 0 || n > c->cfg.max_batch) FAIL(VTD_ERR_CAPACITY, "n=%d outside 1..max_batch", n);
  const size_t px = (size_t)c->cfg.det_h * c->cfg.det_w * n;
  if (ph) CK(cudaMemcpyAsync(ph, c->prob, px * 4, cudaMemcpyDeviceToHost, c->stream));
  if (th) CK(cudaMemcpyAsync(th, c->thresh, px * 4, cudaMemcpyDeviceToHost, c->stream));
  if (mh) CK(cudaMemcpyAsync(mh, c->mask, px, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return VTD_OK;
}

int vtd_dbnet_forward(vtd_ctx* c, const float* x, int n, float* ph, float* th) {
  if (!c || !x) return VTD_ERR_ARG;
  Guard g(c);
  { int fr = finish_pending(c); if (fr) return fr; }
  if (!c->det_loaded) FAIL(VTD_ERR_STATE, "vtd_load_detector has not been called");
  if (n <= 0) FAIL(VTD_ERR_ARG, "n must be positive");
  const int B = c->cfg.max_batch, dh = c->cfg.det_h, dw = c->cfg.det_w;
  const size_t px = (size_t)dh * dw;
  const size_t need = (size_t)B * 3 * px;
  if (c->stage_f32_elems < need) {
    if (c->stage_f32) cudaFree(c->stage_f32);
    c->stage_f32 = nullptr; c->stage_f32_elems = 0;
    CK(cudaMalloc(&c->stage_f32, need * 4));
    c->stage_f32_elems = need;
  }
  for (int i0 = 0; i0 < n; i0 += B) {
    const int m = n - i0 < B ? n - i0 : B;
    CK(cudaMemcpyAsync(c->stage_f32, x + (size_t)i0 * 3 * px, (size_t)m * 3 * px * 4, cudaMemcpyHostToDevice, c->stream));
    if (c->bf16_mode) CK(nchw_f32_to_nhwc<bf16>(c->stage_f32, (bf16*)c->pre, m, 3, dh, dw, c->pre_lay, c->stream, &c->lc));
    else CK(nchw_f32_to_nhwc<float>(c->stage_f32, (float*)c->pre, m, 3, dh, dw, c->pre_lay, c->stream, &c->lc));
    int r = detect_maps_locked(c, m, 0.5f, nullptr); if (r) return r;
    if (ph) CK(cudaMemcpyAsync(ph + (size_t)i0 * px, c->prob, (size_t)m * px * 4, cudaMemcpyDeviceToHost, c->stream));
    if (th) CK(cudaMemcpyAsync(th + (size_t)i0 * px, c->thresh, (size_t)m * px * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return VTD_OK;
}

int vtd_extract_boxes(vtd_ctx* c, int n, int orig_h, int orig_w) {
  if (!c) return VTD_ERR_ARG;
  Guard g(c);
  { int fr = finish_pending(c); if (fr) return fr; }
  if (n <= 0 || n > c->cfg.max_batch) FAIL(VTD_ERR_CAPACITY, "n=%d outside 1..max_batch", n);
  if (orig_h <= 0 || orig_w <= 0) FAIL(VTD_ERR_ARG, "orig size must be positive");
  return extract_locked(c, n, orig_h, orig_w);
}

int vtd_postprocess_map(vtd_ctx* c, const float* prob_host, int mh, int mw, int clip_h, int clip_w, int orig_w,
                        int orig_h, float thr, vtd_record* out, int cap, int* n_out) {
  if (!c || !prob_host || !n_out) return VTD_ERR_ARG;
  Guard g(c);
  { int fr = finish_pending(c); if (fr) return fr; }
  if (mh <= 0 || mw <= 0 || clip_h <= 0 || clip_w <= 0 || orig_w <= 0 || orig_h <= 0) FAIL(VTD_ERR_ARG, "sizes must be positive");
  if ((long long)mh * mw > (1LL << 28)) FAIL(VTD_ERR_CAPACITY, "map too large");
  const int kmax = c->cfg.max_boxes;
  if (c->pp_h != mh || c->pp_w != mw) {
    if (c->pp_work) { cudaFree(c->pp_work); cudaFree(c->pp_prob); cudaFree(c->pp_mask); cudaFree(c->pp_records); cudaFree(c->pp_counts); }
    c->pp_work = nullptr; c->pp_h = c->pp_w = 0;
    const int kc = kmax * 4 < 1024 ? 1024 : kmax * 4;
    size_t bytes = box_work_bytes(1, mh, mw, kc, &c->pp_lay);
    CK(cudaMalloc(&c->pp_work, bytes));
    CK(cudaMemset(c->pp_work, 0, bytes));
    CK(cudaMalloc(&c->pp_prob, (size_t)mh * mw * 4));
    CK(cudaMalloc(&c->pp_mask, (size_t)mh * mw));
    CK(cudaMalloc(&c->pp_records, sizeof(vtd_record) * kmax));
    CK(cudaMalloc(&c->pp_counts, sizeof(int)));
    c->pp_h = mh; c->pp_w = mw;
  }
  CK(cudaMemcpyAsync(c->pp_prob, prob_host, (size_t)mh * mw * 4, cudaMemcpyHostToDevice, c->stream));
  CK(threshold_mask(c->pp_prob, c->pp_mask, (long long)mh * mw, thr, c->stream, &c->lc));
  BoxParams bp;
  bp.n = 1; bp.n_alloc = 1; bp.mh = mh; bp.mw = mw; bp.clip_h = clip_h; bp.clip_w = clip_w;
  bp.orig_h = orig_h; bp.orig_w = orig_w; bp.kmax = kmax; bp.unclip = c->cfg.unclip_ratio;
  CK(extract_boxes(c->pp_prob, c->pp_mask, bp, c->pp_work, c->pp_lay, c->pp_records, c->pp_counts, c->stream, &c->lc));
  CK(cudaMemcpyAsync(c->pinned_int, c->pp_counts, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  int cnt = c->pinned_int[0];
  *n_out = cnt;
  int m = cnt < cap ? cnt : cap;
  if (out && m > 0) {
    CK(cudaMemcpyAsync(out, c->pp_records, sizeof(vtd_record) * m, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return VTD_OK;
}

int vtd_recognize_boxes(vtd_ctx* c, int n) {
  if (!c) return VTD_ERR_ARG;
  Guard g(c);
  { int fr = finish_pending(c); if (fr) return fr; }
  if (!c->rec_loaded) FAIL(VTD_ERR_STATE, "vtd_load_recognizer has not been called");
  if (n <= 0 || n > c->cfg.max_batch || n > c->cur_n) FAIL(VTD_ERR_STATE, "n=%d does not match the preprocessed batch (%d)", n, c->cur_n);
  return recognize_locked(c, n);
}

int vtd_recognize_crops(vtd_ctx* c, const uint8_t* const* crops, const int* h, const int* w, const int* pitch, int n_crops,
                        uint8_t* ids_out, int* len_out, float* conf_out, float* logits_out) {
  if (!c || !crops || !h || !w || !pitch) return VTD_ERR_ARG;
  Guard g(c);
  { int fr = finish_pending(c); if (fr) return fr; }
  if (!c->rec_loaded) FAIL(VTD_ERR_STATE, "vtd_load_recognizer has not been called");
  if (n_crops <= 0) return VTD_OK;
  const int T = c->T;
  std::vector<const uint8_t*> ptrs(c->rc);
  std::vector<int> meta(3 * (size_t)c->rc);
  for (int first = 0; first < n_crops; first += c->rc) {
    const int nc = n_crops - first < c->rc ? n_crops - first : c->rc;
    size_t total = 0;
    for (int i = 0; i < nc; ++i) {
      if (h[first + i] <= 0 || w[first + i] <= 0 || pitch[first + i] < 3 * w[first + i] || !crops[first + i])
        FAIL(VTD_ERR_ARG, "crop %d is empty or has a bad pitch", first + i);
      total += ((size_t)h[first + i] * w[first + i] * 3 + 15) & ~(size_t)15;
    }
    if (total > c->list_cap) {
      if (c->list_store) cudaFree(c->list_store);
      c->list_store = nullptr; c->list_cap = 0;
      CK(cudaMalloc(&c->list_store, total * 2));
      c->list_cap = total * 2;
    }
    size_t off = 0;
    for (int i = 0; i < nc; ++i) {
      const int hh = h[first + i], ww = w[first + i];
      CK(cudaMemcpy2DAsync(c->list_store + off, (size_t)ww * 3, crops[first + i], pitch[first + i], (size_t)ww * 3, hh,
                           cudaMemcpyHostToDevice, c->stream));
      ptrs[i] = c->list_store + off;
      meta[i] = hh; meta[c->rc + i] = ww; meta[2 * c->rc + i] = ww * 3;
      off += ((size_t)hh * ww * 3 + 15) & ~(size_t)15;
    }
    CK(cudaMemcpyAsync(c->list_ptrs, ptrs.data(), sizeof(void*) * nc, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->list_meta, meta.data(), sizeof(int) * 3 * c->rc, cudaMemcpyHostToDevice, c->stream));
    if (c->bf16_mode)
      CK(crop_resize_list<bf16>(c->list_ptrs, c->list_meta, c->list_meta + c->rc, c->list_meta + 2 * c->rc, nc,
                                c->cfg.crop_w, (bf16*)c->crops, c->crops_lay, c->stream, &c->lc));
    else
      CK(crop_resize_list<float>(c->list_ptrs, c->list_meta, c->list_meta + c->rc, c->list_meta + 2 * c->rc, nc,
                                 c->cfg.crop_w, (float*)c->crops, c->crops_lay, c->stream, &c->lc));
    int r = run_crnn(c, nc); if (r) return r;
    CK(ctc_greedy(c->logits, nc, T, 97, c->logits_ld, 0, c->cfg.canonical_ctc, c->ids_dev, VTD_IDS_STRIDE, c->len_dev, c->conf_dev,
                  c->stream, &c->lc));
    if (ids_out) CK(cudaMemcpyAsync(ids_out + (size_t)first * VTD_IDS_STRIDE, c->ids_dev, (size_t)nc * VTD_IDS_STRIDE, cudaMemcpyDeviceToHost, c->stream));
    if (len_out) CK(cudaMemcpyAsync(len_out + first, c->len_dev, (size_t)nc * 4, cudaMemcpyDeviceToHost, c->stream));
    if (conf_out) CK(cudaMemcpyAsync(conf_out + first, c->conf_dev, (size_t)nc * 4, cudaMemcpyDeviceToHost, c->stream));
    if (logits_out) CK(cudaMemcpy2DAsync(logits_out + (size_t)first * T * 97, 97 * 4, c->logits, (size_t)c->logits_ld * 4, 97 * 4, (size_t)nc * T, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return VTD_OK;
}

int vtd_crnn_forward(vtd_ctx* c, const float* x, int n, float* logits_host) {
  if (!c || !x || !logits_host) return VTD_ERR_ARG;
  Guard g(c);
  { int fr = finish_pending(c); if (fr) return fr; }
  if (!c->rec_loaded) FAIL(VTD_ERR_STATE, "vtd_load_recognizer has not been called");
  const int cw = c->cfg.crop_w, T = c->T;
  const size_t per = (size_t)3 * 32 * cw;
  const size_t need = (size_t)c->rc * per;
  if (c->stage_f32_elems < need) {
    if (c->stage_f32) cudaFree(c->stage_f32);
    c->stage_f32 = nullptr; c->stage_f32_elems = 0;
    CK(cudaMalloc(&c->stage_f32, need * 4));
    c->stage_f32_elems = need;
  }
  for (int first = 0; first < n; first += c->rc) {
    const int nc = n - first < c->rc ? n - first : c->rc;
    CK(cudaMemcpyAsync(c->stage_f32, x + (size_t)first * per, (size_t)nc * per * 4, cudaMemcpyHostToDevice, c->stream));
    if (c->bf16_mode) CK(nchw_f32_to_nhwc<bf16>(c->stage_f32, (bf16*)c->crops, nc, 3, 32, cw, c->crops_lay, c->stream, &c->lc));
    else CK(nchw_f32_to_nhwc<float>(c->stage_f32, (float*)c->crops, nc, 3, 32, cw, c->crops_lay, c->stream, &c->lc));
    int r = run_crnn(c, nc); if (r) return r;
    CK(cudaMemcpy2DAsync(logits_host + (size_t)first * T * 97, 97 * 4, c->logits, (size_t)c->logits_ld * 4, 97 * 4, (size_t)nc * T, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return VTD_OK;
}

int vtd_ctc_decode(vtd_ctx* c, const float* x, int B, int T, int V, int is_prob, uint8_t* ids_out, int* len_out,
                   float* conf_out) {
  if (!c || !x) return VTD_ERR_ARG;
  Guard g(c);
  if (B <= 0) return VTD_OK;
  if (T <= 0 || T > VTD_IDS_STRIDE || V < 2) FAIL(VTD_ERR_ARG, "T must be in 1..%d and V >= 2", VTD_IDS_STRIDE);
  float* dx = nullptr; uint8_t* dids = nullptr; int* dlen = nullptr; float* dconf = nullptr;
  const size_t ne = (size_t)B * T * V;
  CK(cudaMalloc(&dx, ne * 4));
  cudaError_t e1 = cudaMalloc(&dids, (size_t)B * VTD_IDS_STRIDE), e2 = cudaMalloc(&dlen, (size_t)B * 4),
              e3 = cudaMalloc(&dconf, (size_t)B * 4);
  int rc = VTD_OK;
  auto cleanup = [&]() { cudaFree(dx); cudaFree(dids); cudaFree(dlen); cudaFree(dconf); };
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { cleanup(); FAIL(VTD_ERR_CUDA, "cudaMalloc failed in vtd_ctc_decode"); }
  cudaError_t e = cudaMemcpyAsync(dx, x, ne * 4, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = ctc_greedy(dx, B, T, V, V, is_prob, c->cfg.canonical_ctc, dids, VTD_IDS_STRIDE, dlen, dconf, c->stream, &c->lc);
  if (e == cudaSuccess && ids_out) e = cudaMemcpyAsync(ids_out, dids, (size_t)B * VTD_IDS_STRIDE, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess && len_out) e = cudaMemcpyAsync(len_out, dlen, (size_t)B * 4, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess && conf_out) e = cudaMemcpyAsync(conf_out, dconf, (size_t)B * 4, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cleanup();
  if (e != cudaSuccess) { c->err = std::string("vtd_ctc_decode: ") + cudaGetErrorString(e); rc = VTD_ERR_CUDA; }
  return rc;
}

int vtd_load_trocr(vtd_ctx* c, const vtd_tensor* t, int n, int crops_per_chunk) {
  if (!c) return VTD_ERR_ARG;
  Guard g(c);
  if (c->trocr && trocr_of(c)->loaded) FAIL(VTD_ERR_STATE, "transformer recogniser already loaded (create a new context to reload)");
  SD sd; int r = make_sd(c, t, n, &sd); if (r) return r;
  r = build_trocr(c, sd, crops_per_chunk > 0 ? crops_per_chunk : 32); if (r) return r;
  CK(cudaDeviceSynchronize());
  return VTD_OK;
}

int vtd_trocr_info(vtd_ctx* c, int32_t* info8) {
  if (!c || !info8) return VTD_ERR_ARG;
  Guard g(c);
  if (!c->trocr || !trocr_of(c)->loaded) FAIL(VTD_ERR_STATE, "vtd_load_trocr has not been called");
  TrocrState& t = *trocr_of(c);
  info8[0] = t.S; info8[1] = t.T; info8[2] = t.De; info8[3] = t.Dd; info8[4] = t.V; info8[5] = t.Lcap; info8[6] = t.cap; info8[7] = t.Ld;
  return VTD_OK;
}

int vtd_trocr_forward(vtd_ctx* c, const float* pixel_values, int n, const int32_t* decoder_ids, int L, int max_length,
                      float* enc_out, float* logits_out, int32_t* ids_out, int* len_out) {
  if (!c || !pixel_values) return VTD_ERR_ARG;
  Guard g(c);
  if (!c->trocr || !trocr_of(c)->loaded) FAIL(VTD_ERR_STATE, "vtd_load_trocr has not been called");
  TrocrState& t = *trocr_of(c);
  if (n <= 0) return VTD_OK;
  const size_t per = (size_t)3 * t.S * t.S;
  if (!t.px_stage) CK(cudaMalloc(&t.px_stage, (size_t)t.cap * per * 4));
  for (int first = 0; first < n; first += t.cap) {
    const int nc = n - first < t.cap ? n - first : t.cap;
    CK(cudaMemcpyAsync(t.px_stage, pixel_values + (size_t)first * per, (size_t)nc * per * 4, cudaMemcpyHostToDevice, c->stream));
    CK(nchw_to_patches(t.px_stage, t.patches, nc, t.S, t.P, c->stream, &c->lc));
    int r = trocr_encode(c, nc); if (r) return r;
    if (enc_out) {
      std::vector<bf16> h((size_t)nc * t.T * t.De);
      CK(cudaMemcpyAsync(h.data(), t.enc, h.size() * 2, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      float* o = enc_out + (size_t)first * t.T * t.De;
      for (size_t i = 0; i < h.size(); ++i) o[i] = (float)h[i];
    }
    if (decoder_ids && logits_out) {                      // teacher forcing: the logits of every given position
      if (L <= 0 || L > t.Lcap) FAIL(VTD_ERR_ARG, "L must be in 1..%d", t.Lcap);
      std::vector<int> ids((size_t)nc * t.Lcap, t.pad_id);
      for (int b = 0; b < nc; ++b) for (int i = 0; i < L; ++i) ids[(size_t)b * t.Lcap + i] = decoder_ids[(size_t)(first + b) * L + i];
      CK(cudaMemcpyAsync(t.ids_dev, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      for (int tp = 0; tp < L; ++tp) {
        r = trocr_decode_step(c, nc, tp); if (r) return r;
        CK(cudaMemcpy2DAsync(logits_out + ((size_t)first * L + tp) * t.V, (size_t)L * t.V * 4, t.logits, (size_t)t.Vp * 4, (size_t)t.V * 4, nc,
                             cudaMemcpyDeviceToHost, c->stream));
      }
      CK(cudaStreamSynchronize(c->stream));
    }
    if (ids_out) {
      r = trocr_generate(c, nc, max_length, ids_out + (size_t)first * max_length, len_out ? len_out + first : nullptr);
      if (r) return r;
    }
  }
  return VTD_OK;
}

int vtd_trocr_generate_crops(vtd_ctx* c, const uint8_t* const* crops, const int* h, const int* w, const int* pitch, int n_crops,
                             int max_length, int32_t* ids_out, int* len_out) {
  if (!c || !crops || !h || !w || !pitch || !ids_out) return VTD_ERR_ARG;
  Guard g(c);
  if (!c->trocr || !trocr_of(c)->loaded) FAIL(VTD_ERR_STATE, "vtd_load_trocr has not been called");
  TrocrState& t = *trocr_of(c);
  for (int first = 0; first < n_crops; first += t.cap) {
    const int nc = n_crops - first < t.cap ? n_crops - first : t.cap;
    // stage the crops, their Pillow coefficient tables (csrc/resize_tab.h, per crop and axis) and the pass-1 intermediate
    size_t bytes = 0, tmp_bytes = 0;
    int max_h = 1;
    for (int i = 0; i < nc; ++i) {
      const int hh = h[first + i], ww = w[first + i];
      if (hh <= 0 || ww <= 0 || pitch[first + i] < 3 * ww || !crops[first + i]) FAIL(VTD_ERR_ARG, "crop %d is empty or has a bad pitch", first + i);
      bytes += ((size_t)hh * ww * 3 + 15) & ~(size_t)15;
      tmp_bytes += ((size_t)hh * t.S * 3 + 15) & ~(size_t)15;
      if (hh > max_h) max_h = hh;
    }
    if (bytes > t.crop_cap) { if (t.crop_store) cudaFree(t.crop_store); t.crop_store = nullptr; t.crop_cap = 0; CK(cudaMalloc(&t.crop_store, bytes * 2)); t.crop_cap = bytes * 2; }
    if (tmp_bytes > t.tmp_cap) { if (t.tmp) cudaFree(t.tmp); t.tmp = nullptr; t.tmp_cap = 0; CK(cudaMalloc(&t.tmp, tmp_bytes * 2)); t.tmp_cap = tmp_bytes * 2; }
    std::vector<const uint8_t*> ptrs(nc);
    std::vector<TrocrCropMeta> meta(nc);
    std::vector<int> tab;
    size_t off = 0; long long toff = 0;
    for (int i = 0; i < nc; ++i) {
      const int hh = h[first + i], ww = w[first + i];
      CK(cudaMemcpy2DAsync(t.crop_store + off, (size_t)ww * 3, crops[first + i], pitch[first + i], (size_t)ww * 3, hh, cudaMemcpyHostToDevice, c->stream));
      ptrs[i] = t.crop_store + off;
      off += ((size_t)hh * ww * 3 + 15) & ~(size_t)15;
      TrocrCropMeta m;
      m.h = hh; m.w = ww; m.pitch = ww * 3; m.tmp_off = toff;
      toff += (long long)(((size_t)hh * t.S * 3 + 15) & ~(size_t)15);
      for (int axis = 0; axis < 2; ++axis) {
        std::vector<int> lo, cnt, kk; int ks = 0, mc = 0;
        compute_resize_tab(axis == 0 ? ww : hh, t.S, &lo, &cnt, &kk, &ks, &mc);
        (axis == 0 ? m.offx : m.offy) = (int)tab.size();
        (axis == 0 ? m.ksx : m.ksy) = ks;
        tab.insert(tab.end(), lo.begin(), lo.end()); tab.insert(tab.end(), cnt.begin(), cnt.end()); tab.insert(tab.end(), kk.begin(), kk.end());
      }
      meta[i] = m;
    }
    if (tab.size() * 4 > t.tab_cap) { if (t.tab_dev) cudaFree(t.tab_dev); t.tab_dev = nullptr; t.tab_cap = 0; CK(cudaMalloc(&t.tab_dev, tab.size() * 8)); t.tab_cap = tab.size() * 8; }
    CK(cudaMemcpyAsync(t.tab_dev, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(t.crop_ptrs, ptrs.data(), sizeof(void*) * nc, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(t.meta_dev, meta.data(), sizeof(TrocrCropMeta) * nc, cudaMemcpyHostToDevice, c->stream));
    CK(trocr_resize_patches(t.crop_ptrs, t.meta_dev, t.tab_dev, t.tmp, t.patches, nc, t.S, t.P, max_h, c->stream, &c->lc));
    CK(cudaStreamSynchronize(c->stream));                  // the host vectors above die with this iteration
    int r = trocr_encode(c, nc); if (r) return r;
    r = trocr_generate(c, nc, max_length, ids_out + (size_t)first * max_length, len_out ? len_out + first : nullptr);
    if (r) return r;
  }
  return VTD_OK;
}

int vtd_run_batch(vtd_ctx* c, const uint8_t* const* frames, int n, int h, int w, int pitch, int pixfmt, int on_dev,
                  float thr, const float* logit_bias, int recognize, vtd_record* rh, int* ch) {
  if (!c || !frames) return VTD_ERR_ARG;
  Guard g(c);
  { int fr = finish_pending(c); if (fr) return fr; }
  NvtxRange range("vtd_run_batch");
  if (!c->det_loaded) FAIL(VTD_ERR_STATE, "vtd_load_detector has not been called");
  if (recognize && !c->rec_loaded) FAIL(VTD_ERR_STATE, "vtd_load_recognizer has not been called");
  int r = preprocess_locked(c, frames, n, h, w, pitch, pixfmt, on_dev); if (r) return r;
  if ((r = detect_maps_locked(c, n, thr, logit_bias))) return r;
  if ((r = extract_locked(c, n, h, w))) return r;
  if (recognize && (r = recognize_locked(c, n))) return r;
  if (rh || ch) return read_records_locked(c, n, rh, ch);
  return VTD_OK;
}

// ProcessingService._draw_detections (processing_service.py:188-218) on the device: csrc/overlay.cu
int vtd_draw_detections(vtd_ctx* c, uint8_t* const* frames, int n, int h, int w, int pitch, int on_dev,
                        const vtd_overlay_item* items, int n_items) {
  static_assert(sizeof(vtd_overlay_item) == sizeof(OverlayItem) && sizeof(OverlayItem) == 256, "overlay item layout");
  static_assert(offsetof(vtd_overlay_item, label) == offsetof(OverlayItem, label), "overlay item layout");
  static_assert((int)VTD_OVERLAY_LABEL_MAX == OV_LABEL_MAX, "overlay label capacity");
  if (!c || !frames || (n_items > 0 && !items)) return VTD_ERR_ARG;
  Guard g(c);
  { int fr = finish_pending(c); if (fr) return fr; }
  constexpr int PER_FRAME = 256;               // the kernel's list of later, overlapping detections (overlay.cu)
  if (n <= 0 || n > c->cfg.max_batch) FAIL(VTD_ERR_CAPACITY, "n=%d outside 1..max_batch=%d", n, c->cfg.max_batch);
  if (h <= 0 || w <= 0 || pitch < w * 3) FAIL(VTD_ERR_ARG, "frame %dx%d with pitch %d", h, w, pitch);
  for (int i = 0; i < n; ++i) if (!frames[i]) FAIL(VTD_ERR_ARG, "frame %d is a null pointer", i);
  if (n_items < 0) FAIL(VTD_ERR_ARG, "n_items=%d", n_items);
  if (n_items == 0) return VTD_OK;
  std::vector<int> per(n, 0);
  for (int i = 0; i < n_items; ++i) {
    const vtd_overlay_item& it = items[i];
    if (it.frame < 0 || it.frame >= n) FAIL(VTD_ERR_ARG, "item %d names frame %d of %d", i, it.frame, n);
    if (it.label_len < 0 || it.label_len > VTD_OVERLAY_LABEL_MAX) FAIL(VTD_ERR_ARG, "item %d: label of %d bytes", i, it.label_len);
    for (int k = 0; k < 4; ++k)
      if (it.bbox[k] > (1 << 24) || it.bbox[k] < -(1 << 24)) FAIL(VTD_ERR_ARG, "item %d: coordinate %d", i, it.bbox[k]);
    if (++per[it.frame] > PER_FRAME) FAIL(VTD_ERR_CAPACITY, "more than %d items on frame %d", PER_FRAME, it.frame);
  }
  // group by frame, keeping the draw order inside each frame
  std::vector<int> start(n + 1, 0);
  for (int f = 0; f < n; ++f) start[f + 1] = start[f] + per[f];
  std::vector<OverlayItem> sorted(n_items);
  std::vector<int> fend(n_items), fill(start.begin(), start.end() - 1);
  for (int i = 0; i < n_items; ++i) {
    const int at = fill[items[i].frame]++;
    memcpy(&sorted[at], &items[i], sizeof(OverlayItem));
    fend[at] = start[items[i].frame + 1];
  }
  int r;
  if (!c->ov_items) {
    const size_t cap = (size_t)c->cfg.max_batch * PER_FRAME;
    if ((r = dalloc(c, &c->ov_items, cap * sizeof(OverlayItem))) || (r = dalloc(c, &c->ov_end, cap * sizeof(int)))) return r;
  }
  if (!c->ov_tables) { CK(overlay_upload_tables(c->stream)); c->ov_tables = true; }
  const size_t rowb = (size_t)w * 3;
  uint8_t* const* ptrs_dev;
  int dev_pitch = pitch;
  if (!on_dev) {
    const size_t fb = rowb * h;
    if (fb > c->frame_bytes_cap) FAIL(VTD_ERR_CAPACITY, "frame of %zu bytes exceeds the staging slot", fb);
    dev_pitch = (int)rowb;
    for (int i = 0; i < n; ++i)
      CK(cudaMemcpy2DAsync(c->frames_store + (size_t)i * c->frame_bytes_cap, rowb, frames[i], pitch, rowb, h, cudaMemcpyHostToDevice,
                           c->stream));
    ptrs_dev = const_cast<uint8_t* const*>(reinterpret_cast<const uint8_t* const*>(c->store_ptrs_dev));
  } else {
    CK(cudaEventSynchronize(c->ptrs_event));
    for (int i = 0; i < n; ++i) c->frame_ptrs_pinned[i] = frames[i];
    CK(cudaMemcpyAsync(c->ext_ptrs_dev, c->frame_ptrs_pinned, sizeof(void*) * n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaEventRecord(c->ptrs_event, c->stream));
    ptrs_dev = const_cast<uint8_t* const*>(reinterpret_cast<const uint8_t* const*>(c->ext_ptrs_dev));
  }
  CK(cudaMemcpyAsync(c->ov_items, sorted.data(), sizeof(OverlayItem) * n_items, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->ov_end, fend.data(), sizeof(int) * n_items, cudaMemcpyHostToDevice, c->stream));
  CK(draw_overlay(ptrs_dev, h, w, dev_pitch, c->ov_items, c->ov_end, n_items, c->stream, &c->lc));
  if (!on_dev)
    for (int i = 0; i < n; ++i)
      CK(cudaMemcpy2DAsync(frames[i], pitch, c->frames_store + (size_t)i * c->frame_bytes_cap, rowb, rowb, h, cudaMemcpyDeviceToHost,
                           c->stream));
  CK(cudaStreamSynchronize(c->stream));          // host frames are complete; `sorted`/`fend` die here
  return VTD_OK;
}

int vtd_read_records(vtd_ctx* c, int n, vtd_record* rh, int* ch) {
  if (!c) return VTD_ERR_ARG;
  Guard g(c);
  if (n <= 0 || n > c->cfg.max_batch) FAIL(VTD_ERR_CAPACITY, "n=%d outside 1..max_batch", n);
  return read_records_locked(c, n, rh, ch);
}

int vtd_get_records(vtd_ctx* c, vtd_record** r, int** cnt) {
  if (!c) return VTD_ERR_ARG;
  if (r) *r = c->records; if (cnt) *cnt = c->counts;
  return VTD_OK;
}

static std::vector<Op*> prof_ops(vtd_ctx* c, int which) {
  std::vector<Op*> v;
  if (which == 0) for (Op& o : c->det_prog) v.push_back(&o);
  else {
    for (Op& o : c->rec_prog) v.push_back(&o);
    if (c->rec_loaded) { v.push_back(&c->xproj_op[0]); v.push_back(&c->xproj_op[1]); v.push_back(&c->fc_op); }
  }
  return v;
}

int vtd_set_profiling(vtd_ctx* c, int on) {
  if (!c) return VTD_ERR_ARG;
  Guard g(c);
  CK(cudaStreamSynchronize(c->stream));
  for (int w = 0; w < 2; ++w)
    for (Op* o : prof_ops(c, w)) { prof_harvest(*o); if (on) { o->prof_ms = 0.0; o->prof_n = 0; o->prof_pending = false; } }
  for (StageProf& sp : c->stage_prof) { stage_harvest(sp); if (on) { sp.ms = 0.0; sp.n = 0; } }
  c->profiling = on != 0;
  return VTD_OK;
}

int vtd_op_count(vtd_ctx* c, int which) {
  if (!c) return 0;
  Guard g(c);
  return which == 2 ? (int)ST_COUNT : (int)prof_ops(c, which).size();
}

/* info[16]: kind(0 conv,1 pool), tensor_core(0/1), H, W, Cin, Ho, Wo, Cout, KH, KW, stride, launches, 0... ; ms = summed
 * device time of those launches */
int vtd_op_info(vtd_ctx* c, int which, int idx, int64_t* info, double* ms) {
  if (!c || !info) return VTD_ERR_ARG;
  Guard g(c);
  if (which == 2) {                       // stages: info[0] = 2, info[12] = stage id (see vtd.h), info[11] = launches
    if (idx < 0 || idx >= ST_COUNT) FAIL(VTD_ERR_ARG, "stage index out of range");
    StageProf& sp = c->stage_prof[idx];
    if (sp.pending) { cudaEventSynchronize(sp.ev1); stage_harvest(sp); }
    for (int i = 0; i < 16; ++i) info[i] = 0;
    info[0] = 2; info[11] = sp.n; info[12] = idx;
    if (ms) *ms = sp.ms;
    return VTD_OK;
  }
  std::vector<Op*> v = prof_ops(c, which);
  if (idx < 0 || idx >= (int)v.size()) FAIL(VTD_ERR_ARG, "op index out of range");
  Op& o = *v[idx];
  if (o.prof_pending) { cudaEventSynchronize(o.ev1); prof_harvest(o); }
  for (int i = 0; i < 16; ++i) info[i] = 0;
  if (o.kind == Op::CONV) {
    info[0] = 0; info[1] = (o.plan || o.sp) ? 1 : 0; info[2] = o.d.H; info[3] = o.d.W; info[4] = o.d.Cin; info[5] = o.d.Ho;
    info[6] = o.d.Wo; info[7] = o.d.Cout; info[8] = o.d.KH; info[9] = o.d.KW; info[10] = o.d.stride;
  } else {
    info[0] = 1; info[2] = o.H; info[3] = o.W; info[4] = o.C; info[8] = o.kh; info[9] = o.kw; info[10] = o.sh;
  }
  info[11] = o.prof_n;
  if (ms) *ms = o.prof_ms;
  return VTD_OK;
}

int vtd_debug_tensor(vtd_ctx* c, const char* name, int n, float* host_out, int64_t capacity, int64_t* shape4) {
  if (!c || !name) return VTD_ERR_ARG;
  Guard g(c);
  { int fr = finish_pending(c); if (fr) return fr; }
  auto it = c->dbg.find(name);
  if (it == c->dbg.end()) FAIL(VTD_ERR_ARG, "unknown debug tensor '%s'", name);
  const DebugEntry& d = it->second;
  const int cap_n = d.per_crop ? c->rc : c->cfg.max_batch;
  if (n <= 0 || n > cap_n) FAIL(VTD_ERR_CAPACITY, "n=%d outside 1..%d", n, cap_n);
  const int C = (strcmp(name, "input") == 0 || strcmp(name, "crops") == 0) ? 3 : d.C;
  if (shape4) { shape4[0] = n; shape4[1] = C; shape4[2] = d.H; shape4[3] = d.W; }
  const int64_t ne = (int64_t)n * C * d.H * d.W;
  if (!host_out) return VTD_OK;
  if (capacity < ne) FAIL(VTD_ERR_CAPACITY, "host buffer holds %lld elements, need %lld", (long long)capacity, (long long)ne);
  float* tmp = nullptr;
  CK(cudaMalloc(&tmp, ne * 4));
  cudaError_t e;
  if (d.f32 || !c->bf16_mode) e = nhwc_to_nchw_f32<float>((const float*)d.p, tmp, n, C, d.H, d.W, d.lay, c->stream, &c->lc);
  else e = nhwc_to_nchw_f32<bf16>((const bf16*)d.p, tmp, n, C, d.H, d.W, d.lay, c->stream, &c->lc);
  if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, tmp, ne * 4, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(tmp);
  if (e != cudaSuccess) { c->err = std::string("vtd_debug_tensor: ") + cudaGetErrorString(e); return VTD_ERR_CUDA; }
  return VTD_OK;
}

}  // extern "C"
#pragma GCC visibility pop
