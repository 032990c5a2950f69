// engine.cu -- handle, weight packing, static execution plans and the C ABI of libysp.so (see include/ysp.h).
//
// The network topology of the hot path is fixed (YOLOv12n detector 4-ch nc=1 + YOLO-Seg++ head, SURVEY App. A.3 /
// 3.2), so it is written here once as a graph builder that emits a static list of kernel launches ("plan") per
// (batch, H, W).  Activations live in a caller-provided workspace as NHWC views; concatenations are formed by
// construction (producers write into channel slices of the consumer's buffer), buffers are packed by lifetime so
// the working set of consecutive layers stays L2-resident.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/ysp.h"
#include "kernels.h"
#include "kernels_train.h"

using namespace ysp;

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
  return code;
}
namespace ysp { void set_error(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); } }   // train.cu
#define CUDA_OK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(YSP_ECUDA, "%s: %s", #x, cudaGetErrorString(e_)); } while (0)

namespace {

struct HostT { std::vector<float> d; std::vector<int64_t> shape; };

struct DevConv {           // packed weights of one conv (BN folded)
  float* w = nullptr;      // dense: [K][wld] fp32; depthwise: [k*k][C]
  float* bias = nullptr;   // [Cout]
  void* w_tc = nullptr;    // bf16 [Cout_pad][K_pad] K-major (tensor-core path), or NULL
  int Cout = 0, Cin = 0, kh = 0, kw = 0, wld = 0, K = 0, Ktc = 0;
  float w_unscale = 1.f;   // TC32 pack: weights are stored times a power of two (keeps the fp16 lo parts normal); epilogue undoes it
  bool dw = false;
};

struct TRef { int buf = -1; int N = 0, H = 0, W = 0, C = 0, cs = 0, co = 0, dt = 0; bool zpad = false; int pw = 0, ph = 0; };
struct BufInfo { size_t bytes = 0, off = 0; int first = 1 << 30, last = -1; std::vector<int> regions; };
struct RunCtx { char* ws; void* ext[16]; cudaStream_t s; };

static inline size_t esize(int dt) { return dt == DT_F32 ? 4 : 2; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct StepInfo { std::string name; std::string kind; double bytes = 0, flops = 0; int launches = 1; };
struct ProfAgg { std::string kind; double ms = 0, bytes = 0, flops = 0; long calls = 0; long launches = 0; };

struct Plan {
  std::vector<std::function<void(RunCtx&)>> steps;
  std::vector<StepInfo> infos;
  std::vector<int> lane, region;             // per step: stream lane (0 = caller's stream) and fork-join region (0 = none)
  std::vector<std::pair<int, int>> region_span = {{0, 0}};   // [first step, last step] per region id
  int split = -1;                            // seg plan: first step that needs the detector's bottleneck
  int share_step = -1, share_buf = -1;       // det plan: steps [0, share_step) produce layer 1 (buffer share_buf), kept alive
  bool shared_stem = false;                  // seg plan: encoder.0/1 skipped, encoder.2 reads the detector's layer 1 (X_E1)
  std::vector<BufInfo> bufs;
  std::map<std::string, TRef> named;
  std::vector<TcConvPlan*> tc_plans;
  std::vector<Tc32ConvPlan*> tc32_plans;
  size_t ws_bytes = 0;
  int launches = 0;
  ~Plan() { for (auto* t : tc_plans) tc_conv_plan_destroy(t); for (auto* t : tc32_plans) tc32_conv_plan_destroy(t); }
  void* ptr(const RunCtx& c, const TRef& t) const {
    char* base = t.buf >= 0 ? c.ws + bufs[t.buf].off : reinterpret_cast<char*>(c.ext[-1 - t.buf]);
    return base ? base + (size_t)t.co * esize(t.dt) : nullptr;
  }
};

}  // namespace

struct ysp_handle {
  int device = 0, mode = YSP_MODE_FP32;
  std::map<std::string, HostT> host;
  std::map<std::string, DevConv> convs;
  std::map<std::string, float*> vecs;        // small raw fp32 vectors (ECA conv1d weights)
  bool det_ready = false, seg_ready = false;
  bool keep_all = false;                     // disable buffer reuse so ysp_debug_tensor sees every intermediate
  bool no_share = false;                     // YSP_NO_SHARE at ysp_create: seg head recomputes encoder layers 0-1 (A/B testing)
  std::map<std::string, std::unique_ptr<Plan>> plans;
  Plan* last_plan = nullptr;
  int last_launches = 0;
  bool profiling = false;
  std::map<std::string, ProfAgg> prof;
  std::vector<cudaEvent_t> prof_events;
  std::vector<cudaStream_t> lanes;           // extra streams for independent branches (lane k -> lanes[k-1])
  std::vector<cudaEvent_t> sync_events;      // fork/join events (timing disabled)
  cudaStream_t aux = nullptr;                // pipeline: seg encoder / NMS run here, concurrently with the detector
  std::string build_err;
  int act_dt() const { return mode == YSP_MODE_BF16 ? DT_BF16 : DT_F32; }
  bool tc32() const { return mode == YSP_MODE_TC32; }   // fp32 storage, convs on tcgen05 with fp16 hi/lo operand splits
};

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// weight packing
// ---------------------------------------------------------------------------------------------------------------------
static const HostT* find(ysp_handle* h, const std::string& k) {
  auto it = h->host.find(k);
  return it == h->host.end() ? nullptr : &it->second;
}

// One conv with BN folded, still in PyTorch layout [Cout][Cin/g][kh][kw] (double-rounded to fp32 at upload).
struct Folded { int Cout = 0, Cin = 0, kh = 0, kw = 0; bool dw = false; std::vector<double> w; std::vector<double> b; };

// `prefix` names either an ultralytics Conv module (prefix.conv.weight [+ prefix.conv.bias] [+ prefix.bn.*]) or a
// plain nn.Conv2d (prefix.weight, prefix.bias).  BN folded in double precision with the module's eps.
static int fold_conv(ysp_handle* h, const std::string& prefix, double bn_eps, Folded& f) {
  const HostT* w = find(h, prefix + ".conv.weight");
  const HostT* b = nullptr;
  const HostT *g = nullptr, *be = nullptr, *mu = nullptr, *var = nullptr;
  if (w) {
    b = find(h, prefix + ".conv.bias");
    g = find(h, prefix + ".bn.weight");
    if (g) {
      be = find(h, prefix + ".bn.bias"); mu = find(h, prefix + ".bn.running_mean"); var = find(h, prefix + ".bn.running_var");
      if (!be || !mu || !var) return fail(YSP_ENOWEIGHT, "incomplete BatchNorm state for %s", prefix.c_str());
    }
  } else {
    w = find(h, prefix + ".weight");
    b = find(h, prefix + ".bias");
  }
  if (!w) return fail(YSP_ENOWEIGHT, "missing weight %s(.conv).weight", prefix.c_str());
  if (w->shape.size() != 4) return fail(YSP_EINVAL, "%s: conv weight must be 4-d", prefix.c_str());
  f.Cout = (int)w->shape[0]; f.Cin = (int)w->shape[1]; f.kh = (int)w->shape[2]; f.kw = (int)w->shape[3];
  f.dw = (f.Cin == 1 && f.Cout > 1 && f.kh > 1);
  const size_t per = (size_t)f.Cin * f.kh * f.kw;
  f.w.resize((size_t)f.Cout * per); f.b.resize(f.Cout);
  for (int co = 0; co < f.Cout; ++co) {
    double bb = b ? b->d[co] : 0.0, sc = 1.0, sh = bb;
    if (g) {
      sc = (double)g->d[co] / std::sqrt((double)var->d[co] + bn_eps);
      sh = (double)be->d[co] + (bb - (double)mu->d[co]) * sc;
    }
    f.b[co] = sh;
    for (size_t i = 0; i < per; ++i) f.w[co * per + i] = (double)w->d[co * per + i] * sc;
  }
  return 0;
}

// Pack one conv, or several convs that read the same input concatenated along Cout (one GEMM instead of several).
static int pack_conv_cat(ysp_handle* h, const std::string& key, const std::vector<std::string>& prefixes, double bn_eps,
                         DevConv** out) {
  auto it = h->convs.find(key);
  if (it != h->convs.end()) { *out = &it->second; return 0; }
  Folded f;
  for (size_t i = 0; i < prefixes.size(); ++i) {
    Folded g;
    int rc = fold_conv(h, prefixes[i], bn_eps, g);
    if (rc) return rc;
    if (i == 0) { f = std::move(g); continue; }
    if (g.Cin != f.Cin || g.kh != f.kh || g.kw != f.kw || g.dw || f.dw)
      return fail(YSP_EINVAL, "cannot concatenate %s with %s", prefixes[i].c_str(), prefixes[0].c_str());
    f.w.insert(f.w.end(), g.w.begin(), g.w.end());
    f.b.insert(f.b.end(), g.b.begin(), g.b.end());
    f.Cout += g.Cout;
  }
  DevConv dc;
  dc.Cout = f.Cout; dc.kh = f.kh; dc.kw = f.kw; dc.dw = f.dw;
  std::vector<float> hw, hb(dc.Cout);
  for (int co = 0; co < dc.Cout; ++co) hb[co] = (float)f.b[co];
  const int taps = dc.kh * dc.kw;
  if (dc.dw) {
    dc.Cin = dc.Cout; dc.K = taps; dc.wld = dc.Cout;
    hw.assign((size_t)taps * dc.Cout, 0.f);
    for (int c = 0; c < dc.Cout; ++c)
      for (int t = 0; t < taps; ++t) hw[(size_t)t * dc.Cout + c] = (float)f.w[(size_t)c * taps + t];
  } else {
    dc.Cin = f.Cin; dc.K = taps * dc.Cin; dc.wld = (dc.Cout + 3) / 4 * 4;
    hw.assign((size_t)dc.K * dc.wld, 0.f);
    for (int co = 0; co < dc.Cout; ++co)
      for (int ci = 0; ci < dc.Cin; ++ci)
        for (int t = 0; t < taps; ++t)
          hw[(size_t)(t * dc.Cin + ci) * dc.wld + co] = (float)f.w[((size_t)co * dc.Cin + ci) * taps + t];
  }
  CUDA_OK(cudaMalloc(&dc.w, hw.size() * 4));
  CUDA_OK(cudaMemcpy(dc.w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMalloc(&dc.bias, hb.size() * 4));
  CUDA_OK(cudaMemcpy(dc.bias, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
  if (!dc.dw && h->mode == YSP_MODE_BF16) {
    // tensor-core layout: [Cout_pad16][Ktc] bf16, K-major, k = tap*Cin_pad + ci with Cin padded to a multiple of 16
    int cin_pad = (dc.Cin + 15) / 16 * 16;
    int cout_pad = (dc.Cout + 15) / 16 * 16;
    dc.Ktc = taps * cin_pad;
    std::vector<uint16_t> hbf((size_t)cout_pad * dc.Ktc, 0);
    for (int co = 0; co < dc.Cout; ++co)
      for (int ci = 0; ci < dc.Cin; ++ci)
        for (int t = 0; t < taps; ++t) {
          float v = (float)f.w[((size_t)co * dc.Cin + ci) * taps + t];
          uint32_t u; memcpy(&u, &v, 4);
          uint32_t r = u + 0x7fffu + ((u >> 16) & 1u);          // round-to-nearest-even to bf16
          hbf[(size_t)co * dc.Ktc + (size_t)t * cin_pad + ci] = (uint16_t)(r >> 16);
        }
    CUDA_OK(cudaMalloc(&dc.w_tc, hbf.size() * 2));
    CUDA_OK(cudaMemcpy(dc.w_tc, hbf.data(), hbf.size() * 2, cudaMemcpyHostToDevice));
  }
  if (!dc.dw && h->mode == YSP_MODE_TC32) {
    // parity-mode tensor-core pack (conv_tc32.cu): per (n-tile, tap, Cin chunk) block [hi|lo][8-channel group][N_tile][8] fp16,
    // w = hi + lo with hi = rn16(w), lo = rn16(w - hi)
    const Tc32Tiling tl = tc32_tiling(dc.Cin, dc.Cout, taps);
    dc.Ktc = taps * tl.cin_pad;
    // Scale by 2^e so that max|w| lands in [2^13, 2^14): hi stays far from the fp16 overflow, and lo = rn16(w - hi) is a
    // NORMAL fp16 number for every weight above 2^-17 of the largest one (unscaled, lo of a typical 0.05 weight would be
    // subnormal and carry an absolute error of 2^-25, four times the 2^-22 relative error of the split itself).
    double wmax = 0.0;
    for (double v : f.w) wmax = std::max(wmax, std::fabs(v));
    int e2 = 0;
    if (wmax > 0.0) { int ex; std::frexp(wmax, &ex); e2 = 14 - ex; }      // wmax = m * 2^ex, m in [0.5, 1)
    const float wsc = std::ldexp(1.0f, e2);
    dc.w_unscale = std::ldexp(1.0f, -e2);
    std::vector<__half> hp(tl.total_bytes / 2, __float2half_rn(0.f));
    const int ngrp = tl.Kc / 8;
    for (int tn = 0; tn < tl.n_tiles_n; ++tn)
      for (int t = 0; t < taps; ++t)
        for (int kc = 0; kc < tl.kchunks; ++kc) {
          __half* blk = hp.data() + ((size_t)(tn * taps + t) * tl.kchunks + kc) * (tl.b_bytes / 2);
          for (int g = 0; g < ngrp; ++g)
            for (int n = 0; n < tl.N_tile; ++n)
              for (int j = 0; j < 8; ++j) {
                const int co = tn * tl.N_tile + n, ci = kc * tl.Kc + g * 8 + j;
                if (co >= dc.Cout || ci >= dc.Cin) continue;
                const float v = (float)f.w[((size_t)co * dc.Cin + ci) * taps + t] * wsc;
                const __half hi = __float2half_rn(v);
                const __half lo = __float2half_rn(v - __half2float(hi));
                blk[((size_t)g * tl.N_tile + n) * 8 + j] = hi;
                blk[((size_t)(ngrp + g) * tl.N_tile + n) * 8 + j] = lo;
              }
        }
    CUDA_OK(cudaMalloc(&dc.w_tc, tl.total_bytes));
    CUDA_OK(cudaMemcpy(dc.w_tc, hp.data(), tl.total_bytes, cudaMemcpyHostToDevice));
  }
  auto res = h->convs.emplace(key, dc);
  *out = &res.first->second;
  return 0;
}

static int pack_conv(ysp_handle* h, const std::string& prefix, double bn_eps, DevConv** out) {
  return pack_conv_cat(h, prefix, {prefix}, bn_eps, out);
}

static int pack_vec(ysp_handle* h, const std::string& key, float** out) {
  auto it = h->vecs.find(key);
  if (it != h->vecs.end()) { *out = it->second; return 0; }
  const HostT* t = find(h, key);
  if (!t) return fail(YSP_ENOWEIGHT, "missing weight %s", key.c_str());
  float* d = nullptr;
  CUDA_OK(cudaMalloc(&d, t->d.size() * 4));
  CUDA_OK(cudaMemcpy(d, t->d.data(), t->d.size() * 4, cudaMemcpyHostToDevice));
  h->vecs[key] = d;
  *out = d;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// graph builder
// ---------------------------------------------------------------------------------------------------------------------
// ext slots
enum { X_IMG = 0, X_Y = 1, X_P3 = 2, X_P4 = 3, X_P5 = 4, X_LOGITS = 5, X_OUT = 6, X_BOTT = 7, X_IMG_U8 = 8, X_E1 = 9 };

struct Builder {
  std::map<int, TRef> pending_gate;          // buffer id -> ECA gate [N][C] not yet applied: the consumer conv folds it into its operand load
  int folding_buf = -1;                      // buffer whose pending gate the op being emitted applies
  ysp_handle* h; Plan* plan; int dt; std::string ns; double bn_eps; int rc = 0;
  int cur_lane = 0, cur_region = 0;
  // independent branches: fork(); lane(k); ...; lane(j); ...; join();  -- steps in different lanes may run concurrently
  void fork() { plan->region_span.push_back({(int)plan->steps.size(), (int)plan->steps.size()}); cur_region = (int)plan->region_span.size() - 1; }
  void lane(int k) { cur_lane = k; }
  void join() { cur_region = 0; cur_lane = 0; }
  Builder(ysp_handle* h_, Plan* p_, const std::string& ns_, double eps) : h(h_), plan(p_), dt(h_->act_dt()), ns(ns_), bn_eps(eps) {}

  TRef alloc(int N, int H, int W, int C, int dtype = -1, int cs = 0) {
    if (dtype < 0) dtype = dt;
    int align = dtype == DT_F32 ? 4 : 8;                       // 16-byte pixel rows
    if (cs == 0) cs = (C + align - 1) / align * align;
    BufInfo b; b.bytes = align_up((size_t)N * H * W * cs * esize(dtype), 256);
    plan->bufs.push_back(b);
    TRef t; t.buf = (int)plan->bufs.size() - 1; t.N = N; t.H = H; t.W = W; t.C = C; t.cs = cs; t.co = 0; t.dt = dtype;
    return t;
  }
  static TRef ext(int slot, int N, int H, int W, int C, int cs, int dtype) {
    TRef t; t.buf = -1 - slot; t.N = N; t.H = H; t.W = W; t.C = C; t.cs = cs; t.co = 0; t.dt = dtype; return t;
  }
  static TRef slice(TRef t, int c0, int c) { t.co += c0; t.C = c; return t; }
  void name(const std::string& n, const TRef& t) { plan->named[ns + ":" + n] = t; }

  void touch(const TRef& t) {
    if (t.buf < 0) return;
    BufInfo& b = plan->bufs[t.buf];
    int step = (int)plan->steps.size();
    b.first = std::min(b.first, step); b.last = std::max(b.last, step);
    if (cur_region && (b.regions.empty() || b.regions.back() != cur_region)) b.regions.push_back(cur_region);
  }
  static double tbytes(const TRef& t) { return (double)t.N * t.H * t.W * t.C * esize(t.dt); }
  void emit(std::function<void(RunCtx&)> f, std::initializer_list<const TRef*> uses, int nlaunch = 1,
            StepInfo info = StepInfo()) {
    for (auto* t : uses) {
      if (!t) continue;
      if (t->buf >= 0 && pending_gate.count(t->buf) && t->buf != folding_buf && !rc)
        rc = fail(YSP_EINVAL, "step %s reads a tensor whose ECA gate is still pending (only a flat tc32 1x1 conv can fold it)", info.name.c_str());
      touch(*t);
    }
    plan->steps.push_back(std::move(f));
    plan->lane.push_back(cur_region ? cur_lane : 0);
    plan->region.push_back(cur_region);
    if (cur_region) plan->region_span[cur_region].second = (int)plan->steps.size() - 1;
    plan->launches += nlaunch;
    info.launches = nlaunch;
    if (info.name.empty()) info.name = "misc";
    if (info.kind.empty()) info.kind = "misc";
    info.name = ns + ":" + info.name;
    plan->infos.push_back(info);
  }

  // 4-channel stem conv reading the caller's tensor directly (fp32 NCHW or u8 HWC4): kernels_stem_attn.cu
  void stem(const std::string& prefix, TRef out, int B, int H, int W) {
    if (rc) return;
    DevConv* dc = nullptr;
    if ((rc = pack_conv(h, ns + "." + prefix, bn_eps, &dc))) return;
    if (dc->dw || dc->Cin != 4 || dc->Cout != 16 || dc->kh != 3) { rc = fail(YSP_EINVAL, "stem %s: expected a 3x3 4->16 conv", prefix.c_str()); return; }
    Plan* pl = plan; int d = dt; const float* w = dc->w; const float* bias = dc->bias; int wld = dc->wld;
    emit([=](RunCtx& c) {
      const bool u8 = c.ext[X_IMG_U8] != nullptr;
      launch_stem_conv(u8 ? c.ext[X_IMG_U8] : c.ext[X_IMG], u8, pl->ptr(c, out), w, bias, B, H, W, out.H, out.W, out.cs, wld, d, c.s);
    }, {&out}, 1,
    StepInfo{prefix, "stem_conv", (double)B * H * W * 16 + tbytes(out), 2.0 * B * out.H * out.W * 36.0 * 16.0, 1});
  }

  // dense conv (ultralytics Conv / nn.Conv2d); OH/OW default to 'same'/stride arithmetic on (padH, padW)
  void conv(const std::string& prefix, TRef in, TRef out, int k, int s, int act, const TRef* res = nullptr,
            int padH = 0, int padW = 0) {
    if (rc) return;
    DevConv* dc = nullptr;
    {   // "a+b" = convs a and b (same input) concatenated along Cout into one GEMM
      std::vector<std::string> parts;
      size_t st = 0, pos;
      while ((pos = prefix.find('+', st)) != std::string::npos) { parts.push_back(ns + "." + prefix.substr(st, pos - st)); st = pos + 1; }
      parts.push_back(ns + "." + prefix.substr(st));
      if ((rc = pack_conv_cat(h, ns + "." + prefix, parts, bn_eps, &dc))) return;
    }
    if (dc->dw || dc->Cin != in.C || dc->Cout != out.C || dc->kh != k) {
      rc = fail(YSP_EINVAL, "conv %s: weight [%d,%d,%d,%d] does not match in C=%d out C=%d k=%d", prefix.c_str(), dc->Cout,
                dc->Cin, dc->kh, dc->kw, in.C, out.C, k);
      return;
    }
    ConvP p = {};
    p.w = dc->w; p.bias = dc->bias;
    p.N = in.N; p.H = in.H; p.W = in.W; p.Cin = in.C; p.in_cs = in.cs;
    p.OH = out.H; p.OW = out.W; p.Cout = out.C; p.out_cs = out.cs; p.res_cs = res ? res->cs : 0;
    p.kh = k; p.kw = k; p.stride = s; p.pad = k / 2; p.act = act;
    p.M = out.N * out.H * out.W; p.K = dc->K; p.wld = dc->wld;
    p.cout_store = out.zpad ? std::min((out.C + 15) / 16 * 16, out.cs) : out.C;
    p.in_zpad = in.zpad ? 1 : 0;
    p.in_pw = in.pw; p.in_ph = in.ph;
    (void)padH; (void)padW;
    Plan* pl = plan;
    TRef rres = res ? *res : TRef();
    bool has_res = res != nullptr;
    int in_dt = in.dt, out_dt = out.dt;
    TcConvPlan* tcp = nullptr;
    bool halo = false;
    const void* w_tc = dc->w_tc; const int Ktc = dc->Ktc;
    if (h->mode == YSP_MODE_BF16 && dc->w_tc && in_dt == DT_BF16 && getenv("YSP_NO_TC") == nullptr) {
      ConvP q = p; q.K = dc->Ktc;
      const char* only = getenv("YSP_TC_ONLY");       // debugging aid: tensor-core path only for matching layers
      halo = getenv("YSP_NO_HALO") == nullptr && conv_halo_supported(q, in_dt, out_dt);
      if (halo) {
      } else if ((!only || prefix.find(only) != std::string::npos) && tc_conv_supported(q)) {
        tcp = tc_conv_plan_create(q, dc->w_tc, out_dt);
        if (tcp) pl->tc_plans.push_back(tcp);
      }
    }
    Tc32ConvPlan* tcp32 = nullptr;
    bool halo32 = false;
    const float w_unscale = dc->w_unscale;
    if (h->tc32() && dc->w_tc && in_dt == DT_F32 && out_dt == DT_F32 && getenv("YSP_NO_TC") == nullptr) {
      ConvP q = p; q.K = dc->Ktc;
      const char* only = getenv("YSP_TC_ONLY");
      halo32 = getenv("YSP_NO_HALO") == nullptr && conv_halo32_supported(q);
      if (halo32) {
      } else if ((!only || prefix.find(only) != std::string::npos) && tc32_conv_supported(q)) {
        tcp32 = tc32_conv_plan_create(q, dc->w_tc, dc->w_unscale);
        if (tcp32) pl->tc32_plans.push_back(tcp32);
      }
    }
    TRef gate; bool has_gate = false;
    if (in.buf >= 0 && pending_gate.count(in.buf)) {
      if (!tcp32 || !tc32_conv_plan_flat(tcp32) || in.co != 0 || in.C != pending_gate[in.buf].C) {
        rc = fail(YSP_EINVAL, "conv %s: its input carries a pending ECA gate but the conv is not a flat tc32 1x1 conv", prefix.c_str());
        return;
      }
      gate = pending_gate[in.buf]; has_gate = true;
      folding_buf = in.buf;
    }
    if ((in.pw || in.ph) && !tcp && !halo && !tcp32) { rc = fail(YSP_EINVAL, "conv %s: pitched input needs the tensor-core path", prefix.c_str()); return; }
    emit([=](RunCtx& c) {
      ConvP q = p;
      q.in = pl->ptr(c, in); q.out = pl->ptr(c, out); q.res = has_res ? pl->ptr(c, rres) : nullptr;
      q.in_scale = has_gate ? (const float*)pl->ptr(c, gate) : nullptr;
      if (halo) launch_conv_halo(q, w_tc, Ktc, c.s);
      else if (halo32) launch_conv_halo32(q, w_tc, w_unscale, c.s);
      else if (tcp32) launch_conv_tc32(tcp32, q, c.s);
      else if (tcp) launch_conv_tc(tcp, q, c.s);
      else launch_conv_dense(q, in_dt, out_dt, c.s);
    }, {&in, &out, res, has_gate ? &gate : nullptr}, 1,
    StepInfo{prefix, std::string(halo ? "halo_conv" : halo32 ? "halo32_conv" : tcp32 ? "tc32_conv" : tcp ? "tc_conv" : "conv") + std::to_string(k) + "x" + std::to_string(k) + (s == 2 ? "s2" : ""),
             tbytes(in) + tbytes(out) + (res ? tbytes(*res) : 0.0) + (double)dc->K * dc->Cout * ((tcp || halo) ? 2 : 4),
             2.0 * p.M * (double)dc->K * dc->Cout, 1});
    if (has_gate) { pending_gate.erase(in.buf); folding_buf = -1; }      // applied: later readers would see the UNscaled tensor,
  }                                                                        // so there must be none (checked in emit for pending ones)

  void dw(const std::string& prefix, TRef in, TRef out, int k, int act, const TRef* res = nullptr, int grp = 0,
          int grp_stride = 0) {
    if (rc) return;
    DevConv* dc = nullptr;
    if ((rc = pack_conv(h, ns + "." + prefix, bn_eps, &dc))) return;
    if (!dc->dw || dc->Cout != out.C || dc->kh != k || (out.C & 3)) {
      rc = fail(YSP_EINVAL, "dwconv %s: weight [%d,%d,%d,%d] does not match C=%d k=%d", prefix.c_str(), dc->Cout, dc->Cin,
                dc->kh, dc->kw, out.C, k);
      return;
    }
    DwP p = {};
    p.w = dc->w; p.bias = dc->bias;
    p.N = out.N; p.H = out.H; p.W = out.W; p.C = out.C; p.in_cs = in.cs; p.out_cs = out.cs; p.res_cs = res ? res->cs : 0;
    p.k = k; p.pad = k / 2; p.act = act;
    p.grp = grp ? grp : out.C; p.grp_stride = grp ? grp_stride : out.C;
    Plan* pl = plan; TRef rres = res ? *res : TRef(); bool has_res = res != nullptr; int d = dt;
    emit([=](RunCtx& c) {
      DwP q = p;
      q.in = pl->ptr(c, in); q.out = pl->ptr(c, out); q.res = has_res ? pl->ptr(c, rres) : nullptr;
      launch_conv_dw(q, d, c.s);
    }, {&in, &out, res}, 1,
    StepInfo{prefix, "dwconv" + std::to_string(k) + "x" + std::to_string(k),
             2.0 * tbytes(out) + (res ? tbytes(*res) : 0.0) + (double)k * k * out.C * 4,
             2.0 * out.N * out.H * out.W * (double)out.C * k * k, 1});
  }

  void ew(int mode, TRef a, const TRef* b, TRef out) {
    if (rc) return;
    EwP p = {};
    p.N = out.N; p.H = a.H; p.W = a.W; p.C = out.C; p.a_cs = a.cs; p.b_cs = b ? b->cs : 0; p.out_cs = out.cs;
    p.OH = out.H; p.OW = out.W;
    Plan* pl = plan; TRef rb = b ? *b : TRef(); bool has_b = b != nullptr; int d = dt;
    emit([=](RunCtx& c) {
      EwP q = p;
      q.a = pl->ptr(c, a); q.b = has_b ? pl->ptr(c, rb) : nullptr; q.out = pl->ptr(c, out);
      if (mode == 0) launch_add(q, d, c.s);
      else if (mode == 1) launch_up_nearest2(q, d, c.s);
      else launch_up_bilinear2(q, d, c.s);
    }, {&a, b, &out}, 1,
    StepInfo{mode == 0 ? "add" : (mode == 1 ? "up_nearest" : "up_bilinear"), mode == 0 ? "add" : (mode == 1 ? "up_nearest" : "up_bilinear"),
             tbytes(a) + (b ? tbytes(*b) : 0.0) + tbytes(out), mode == 2 ? 8.0 * out.N * out.H * out.W * out.C : 0.0, 1});
  }

  void eca(const std::string& key, TRef x) {
    if (rc) return;
    float* w3 = nullptr;
    if ((rc = pack_vec(h, ns + "." + key, &w3))) return;
    TRef mean = alloc(x.N, 1, 1, x.C, DT_F32);
    Plan* pl = plan; int d = dt;
    // Parity mode with the fused decoder: only the gate is computed here; the one consumer of the scaled tensor (the low-res
    // [conv1 | residual] GEMM of the next DoubleLightConv) multiplies by it while converting its operands, so the
    // read-modify-write pass over the tensor disappears.  YSP_NO_ECA_FOLD=1: the two-pass form.
    if (h->tc32() && x.buf >= 0 && x.co == 0 && eca_gate_f32_supported(x.C, x.cs) && getenv("YSP_NO_ECA_FOLD") == nullptr &&
        getenv("YSP_NO_FUSE") == nullptr && getenv("YSP_NO_TC") == nullptr) {
      folding_buf = x.buf;
      emit([=](RunCtx& c) {
        launch_eca_gate_f32(pl->ptr(c, x), x.N, x.H * x.W, x.C, x.cs, w3, (float*)pl->ptr(c, mean), c.s);
      }, {&x, &mean}, 1, StepInfo{key, "eca_gate", tbytes(x), 1.0 * x.N * x.H * x.W * x.C, 1});
      folding_buf = -1;
      pending_gate[x.buf] = mean;
      return;
    }
    emit([=](RunCtx& c) {
      launch_eca(pl->ptr(c, x), x.N, x.H * x.W, x.C, x.cs, w3, (float*)pl->ptr(c, mean), d, c.s);
    }, {&x, &mean}, 2, StepInfo{key, "eca", 3.0 * tbytes(x), 2.0 * x.N * x.H * x.W * x.C, 2});
  }

  void attention(TRef qkv, TRef out, int heads, int area) {
    if (rc) return;
    Plan* pl = plan; int d = dt;
    // tcgen05 kernel (attention_tc.cu) for 64-token areas in both tensor-core modes; the FFMA mode and other shapes use the
    // SIMT kernel, YSP_NO_ATTN_TC=1 switches back to the mma.sync kernel (bf16) / the SIMT kernel for A/B timing
    const bool tcmode = (h->mode == YSP_MODE_BF16 || h->tc32()) && getenv("YSP_NO_ATTN_TC") == nullptr && getenv("YSP_NO_TC") == nullptr;
    const bool use_tc = tcmode && attention_tc_supported(qkv.H * qkv.W, heads, area > 0 ? area : 1, qkv.cs, out.cs, dt);
    emit([=](RunCtx& c) {
      const int ntok = qkv.H * qkv.W, ar = area > 0 ? area : 1;
      if (use_tc)
        launch_attention_tc(pl->ptr(c, qkv), pl->ptr(c, out), qkv.N, ntok, heads, ar, qkv.cs, out.cs, d, c.s);
      else if (d == DT_BF16 && ntok % ar == 0 && ntok / ar == 64)
        launch_attention64_bf16(pl->ptr(c, qkv), pl->ptr(c, out), qkv.N, ntok, heads, ar, qkv.cs, out.cs, c.s);
      else
        launch_attention(pl->ptr(c, qkv), pl->ptr(c, out), qkv.N, ntok, out.C, heads, ar, qkv.cs, out.cs, d, c.s);
    }, {&qkv, &out}, 1,
    StepInfo{"attention", use_tc ? "attention_tc" : "attention", tbytes(qkv) + tbytes(out),
             4.0 * qkv.N * (double)(qkv.H * qkv.W) * (qkv.H * qkv.W / (area > 0 ? area : 1)) * out.C, 1});
  }

  // ---- ultralytics blocks, concat-by-construction -------------------------------------------
  void bottleneck(const std::string& p, TRef x, TRef out, bool shortcut, int k0, int k1, double e) {
    int c_ = (int)(out.C * e);
    TRef t = alloc(x.N, x.H, x.W, c_);
    if ((dt == DT_BF16 || h->tc32()) && c_ % 16 != 0) {       // zero-pad the hidden channels to 16 so cv2 is a tensor-core conv too
      t = alloc(x.N, x.H, x.W, c_, -1, (c_ + 15) / 16 * 16);
      t.zpad = true;
    }
    conv(p + ".cv1", x, t, k0, 1, ACT_SILU);
    bool add = shortcut && x.C == out.C;
    conv(p + ".cv2", t, out, k1, 1, ACT_SILU, add ? &x : nullptr);
  }
  void c3k(const std::string& p, TRef x, TRef out, int n, bool shortcut) {
    int c_ = (int)(out.C * 0.5);
    // buffer [m(cv1(x)) | cv2(x) | cv1(x)]: cv1 and cv2 read the same input, so they run as ONE GEMM (N = 2c_) writing
    // channels [c_, 3c_); cv3 reads the first 2c_ channels (the concat of the reference)
    TRef buf = alloc(x.N, x.H, x.W, 3 * c_);
    TRef cat = slice(buf, 0, 2 * c_), a = slice(buf, 2 * c_, c_);
    conv(p + ".cv2+" + p + ".cv1", x, slice(buf, c_, 2 * c_), 1, 1, ACT_SILU);
    TRef cur = a;
    for (int i = 0; i < n; ++i) {
      TRef dst = (i == n - 1) ? slice(buf, 0, c_) : alloc(x.N, x.H, x.W, c_);
      bottleneck(p + ".m." + std::to_string(i), cur, dst, shortcut, 3, 3, 1.0);
      cur = dst;
    }
    conv(p + ".cv3", cat, out, 1, 1, ACT_SILU);
  }
  void c3k2(const std::string& p, TRef x, TRef out, bool is_c3k, double e, bool shortcut) {
    int c = (int)(out.C * e);
    TRef cat = alloc(x.N, x.H, x.W, 3 * c);
    conv(p + ".cv1", x, slice(cat, 0, 2 * c), 1, 1, ACT_SILU);
    if (is_c3k) c3k(p + ".m.0", slice(cat, c, c), slice(cat, 2 * c, c), 2, shortcut);
    else bottleneck(p + ".m.0", slice(cat, c, c), slice(cat, 2 * c, c), shortcut, 3, 3, 0.5);
    conv(p + ".cv2", cat, out, 1, 1, ACT_SILU);
  }
  void ablock(const std::string& p, TRef x, TRef out, int area) {
    const int dim = x.C, heads = dim / 32;
    TRef qkv = alloc(x.N, x.H, x.W, 3 * dim);
    conv(p + ".attn.qkv", x, qkv, 1, 1, ACT_NONE);
    TRef xa = alloc(x.N, x.H, x.W, dim);
    attention(qkv, xa, heads, area);
    TRef xpe = alloc(x.N, x.H, x.W, dim);
    dw(p + ".attn.pe", slice(qkv, 64, dim), xpe, 7, ACT_NONE, &xa, 32, 96);   // v channels: head*96 + 64 + d
    TRef x1 = alloc(x.N, x.H, x.W, dim);
    conv(p + ".attn.proj", xpe, x1, 1, 1, ACT_NONE, &x);
    TRef hid = alloc(x.N, x.H, x.W, 2 * dim);
    conv(p + ".mlp.0", x1, hid, 1, 1, ACT_SILU);
    conv(p + ".mlp.1", hid, out, 1, 1, ACT_NONE, &x1);
  }
  void a2c2f(const std::string& p, TRef x, TRef out, int n, bool a2, int area) {
    int c_ = (int)(out.C * 0.5);
    TRef cat = alloc(x.N, x.H, x.W, (1 + n) * c_);
    conv(p + ".cv1", x, slice(cat, 0, c_), 1, 1, ACT_SILU);
    for (int i = 0; i < n; ++i) {
      TRef src = slice(cat, i * c_, c_), dst = slice(cat, (i + 1) * c_, c_);
      std::string m = p + ".m." + std::to_string(i);
      if (a2) {
        TRef tmp = alloc(x.N, x.H, x.W, c_);
        ablock(m + ".0", src, tmp, area);
        ablock(m + ".1", tmp, dst, area);
      } else {
        c3k(m, src, dst, 2, true);
      }
    }
    conv(p + ".cv2", cat, out, 1, 1, ACT_SILU);
  }
  // GhostConv: y = Conv(c1, c_/2, 1); out = cat(y, DW5x5(y)); optional residual (GhostBottleneck identity shortcut)
  void ghostconv(const std::string& p, TRef x, TRef out, int act, const TRef* res) {
    int hc = out.C / 2;
    if (!res) {
      conv(p + ".cv1", x, slice(out, 0, hc), 1, 1, act);
      dw(p + ".cv2", slice(out, 0, hc), slice(out, hc, hc), 5, act);
    } else {
      TRef y = alloc(x.N, x.H, x.W, hc);
      conv(p + ".cv1", x, y, 1, 1, act);
      TRef r0 = slice(*res, 0, hc), r1 = slice(*res, hc, hc);
      dw(p + ".cv2", y, slice(out, hc, hc), 5, act, &r1);
      ew(0, y, &r0, slice(out, 0, hc));
    }
  }
  void c3ghost(const std::string& p, TRef x, TRef out) {
    int c_ = (int)(out.C * 0.5);
    // [m(cv1(x)) | cv2(x) | cv1(x)] with cv1/cv2 as one GEMM, like c3k()
    TRef buf = alloc(x.N, x.H, x.W, 3 * c_);
    TRef cat = slice(buf, 0, 2 * c_), a = slice(buf, 2 * c_, c_);
    conv(p + ".cv2+" + p + ".cv1", x, slice(buf, c_, 2 * c_), 1, 1, ACT_SILU);
    // GhostBottleneck(c_, c_), stride 1: GhostConv(c_, c_/2) -> Identity -> GhostConv(c_/2, c_, act=False); + x
    TRef dst = slice(buf, 0, c_);
    if (dt == DT_BF16 && getenv("YSP_NO_GHOSTFUSE") == nullptr && ghost_fused_supported(c_, a.cs, dst.cs)) {
      DevConv *k1 = nullptr, *d1 = nullptr, *k3 = nullptr, *d2 = nullptr;   // the four convs as ONE kernel (kernels_ghost.cu)
      if ((rc = pack_conv(h, ns + "." + p + ".m.0.conv.0.cv1", bn_eps, &k1))) return;
      if ((rc = pack_conv(h, ns + "." + p + ".m.0.conv.0.cv2", bn_eps, &d1))) return;
      if ((rc = pack_conv(h, ns + "." + p + ".m.0.conv.2.cv1", bn_eps, &k3))) return;
      if ((rc = pack_conv(h, ns + "." + p + ".m.0.conv.2.cv2", bn_eps, &d2))) return;
      if (k1->dw || k3->dw || !d1->dw || !d2->dw || k1->Cin != c_ || k1->Cout != c_ / 4 || d1->Cout != c_ / 4 || d1->kh != 5 ||
          k3->Cin != c_ / 2 || k3->Cout != c_ / 2 || d2->Cout != c_ / 2 || d2->kh != 5 || k1->kh != 1 || k3->kh != 1) {
        rc = fail(YSP_EINVAL, "c3ghost %s: unexpected GhostBottleneck weight shapes", p.c_str());
        return;
      }
      GhostP q = {};
      q.w1 = k1->w; q.b1 = k1->bias; q.w1ld = k1->wld; q.dw1 = d1->w; q.bd1 = d1->bias;
      q.w3 = k3->w; q.b3 = k3->bias; q.w3ld = k3->wld; q.dw2 = d2->w; q.bd2 = d2->bias;
      q.N = x.N; q.H = x.H; q.W = x.W; q.a_cs = a.cs; q.out_cs = dst.cs;
      Plan* pl = plan;
      const double px = (double)x.N * x.H * x.W;
      emit([=](RunCtx& c) {
        GhostP r = q;
        r.a = (const bf16*)pl->ptr(c, a); r.out = (bf16*)pl->ptr(c, dst);
        launch_ghost_fused(r, c_, c.s);
      }, {&a, &dst}, 1,
      StepInfo{p + ".m.0.fused", "ghost_fused", 2.0 * tbytes(a), 2.0 * px * (c_ * c_ / 4.0 + 25.0 * c_ / 4 + c_ * c_ / 4.0 + 25.0 * c_ / 2), 1});
    } else {
      TRef g1 = alloc(x.N, x.H, x.W, c_ / 2);
      ghostconv(p + ".m.0.conv.0", a, g1, ACT_SILU, nullptr);
      ghostconv(p + ".m.0.conv.2", g1, dst, ACT_NONE, &a);
    }
    conv(p + ".cv3", cat, out, 1, 1, ACT_SILU);
  }
  // bilinear x2 + DoubleLightConv [+ output head] fused (kernels_fused.cu): the two 1x1 convs that read the upsampled
  // tensor are linear, so they run at LOW resolution as one GEMM (P = [conv.0.conv1 | residual_conv]); one kernel does
  // the rest per hi-res tile.  `out` is the hi-res NHWC result, or (head_prefix != "") the fp32 [N,1,2h,2w] logits.
  void doublelight_fused(const std::string& p, TRef xlow, TRef out, const std::string& head_prefix) {
    if (rc) return;
    const int C = head_prefix.empty() ? out.C : 16;
    TRef P = alloc(xlow.N, xlow.H, xlow.W, 2 * C);
    conv(p + ".conv.0.conv1+" + p + ".residual_conv", xlow, P, 1, 1, ACT_NONE);
    DevConv *d1 = nullptr, *c2 = nullptr, *d2 = nullptr, *hd = nullptr;
    if ((rc = pack_conv(h, ns + "." + p + ".conv.0.conv2", bn_eps, &d1))) return;
    if ((rc = pack_conv(h, ns + "." + p + ".conv.1.conv1", bn_eps, &c2))) return;
    if ((rc = pack_conv(h, ns + "." + p + ".conv.1.conv2", bn_eps, &d2))) return;
    if (!head_prefix.empty() && (rc = pack_conv(h, ns + "." + head_prefix, bn_eps, &hd))) return;
    if (!d1->dw || !d2->dw || d1->Cout != C || d2->Cout != C || c2->Cin != C || c2->Cout != C || c2->kh != 1 || d1->kh != 3 ||
        d2->kh != 3 || (hd && (hd->Cin != C || hd->Cout != 1))) {
      rc = fail(YSP_EINVAL, "doublelight %s: unexpected weight shapes", p.c_str());
      return;
    }
    const double hi = (double)xlow.N * xlow.H * xlow.W * 4;
    if (h->tc32() && c2->w_tc && dlc32_supported(C, hd != nullptr) && P.cs % 4 == 0 && (hd || out.cs % 4 == 0) &&
        getenv("YSP_NO_DLC32") == nullptr) {
      // parity mode: composite up2 o depthwise on CUDA cores + the pointwise GEMM on tcgen05 (kernels_dlc32.cu)
      Dlc32P q = {};
      q.dw1 = d1->w; q.b1 = d1->bias; q.wpack = c2->w_tc; q.w_unscale = c2->w_unscale; q.b2 = c2->bias; q.dw2 = d2->w; q.b3 = d2->bias;
      q.wo = hd ? hd->w : nullptr; q.bo = hd ? hd->bias : nullptr; q.wo_ld = hd ? hd->wld : 0;
      q.N = xlow.N; q.h = xlow.H; q.w = xlow.W; q.C = C; q.p_cs = P.cs; q.out_cs = out.cs;
      Plan* pl = plan;
      emit([=](RunCtx& c) {
        Dlc32P r = q;
        r.P = pl->ptr(c, P); r.out = pl->ptr(c, out);
        launch_dlc32(r, c.s);
      }, {&P, &out}, 1,
      StepInfo{p + ".fused", "dlc32", tbytes(P) + (hd ? hi * 4 : tbytes(out)), 2.0 * hi * C * (9 + C + 9 + (hd ? 1 : 0)) + 16.0 * hi * C, 1});
      return;
    }
    DlcP q = {};
    q.dw1 = d1->w; q.b1 = d1->bias; q.w2 = c2->w; q.b2 = c2->bias; q.w2ld = c2->wld; q.dw2 = d2->w; q.b3 = d2->bias;
    q.wo = hd ? hd->w : nullptr; q.bo = hd ? hd->bias : nullptr; q.wo_ld = hd ? hd->wld : 0;
    q.N = xlow.N; q.h = xlow.H; q.w = xlow.W; q.C = C; q.p_cs = P.cs; q.out_cs = out.cs;
    Plan* pl = plan; int d = dt;
    emit([=](RunCtx& c) {
      DlcP r = q;
      r.P = pl->ptr(c, P); r.out = pl->ptr(c, out);
      launch_dlc_fused(r, d, c.s);
    }, {&P, &out}, 1,
    StepInfo{p + ".fused", "dlc_fused", tbytes(P) + (hd ? hi * 4 : tbytes(out)), 2.0 * hi * C * (9 + C + 9 + (hd ? 1 : 0)) + 16.0 * hi * C, 1});
  }

  // bf16 mode, (Cin, C) in {(32,16)+head, (64,32)}: the whole stage on tcgen05 (kernels_dlc_tc.cu); no low-res GEMM needed.
  bool doublelight_tc(const std::string& p, TRef xlow, TRef out, const std::string& head_prefix) {
    if (rc) return true;
    const bool head = !head_prefix.empty();
    const int C = head ? 16 : out.C, Cin = xlow.C;
    if (dt != DT_BF16 || !dlc_tc_supported(Cin, C, head) || getenv("YSP_NO_DLC_TC") != nullptr) return false;
    DevConv *c1 = nullptr, *d1 = nullptr, *c2 = nullptr, *d2 = nullptr, *rr = nullptr, *hd = nullptr;
    if ((rc = pack_conv(h, ns + "." + p + ".conv.0.conv1", bn_eps, &c1))) return true;
    if ((rc = pack_conv(h, ns + "." + p + ".conv.0.conv2", bn_eps, &d1))) return true;
    if ((rc = pack_conv(h, ns + "." + p + ".conv.1.conv1", bn_eps, &c2))) return true;
    if ((rc = pack_conv(h, ns + "." + p + ".conv.1.conv2", bn_eps, &d2))) return true;
    if ((rc = pack_conv(h, ns + "." + p + ".residual_conv", bn_eps, &rr))) return true;
    if (head && (rc = pack_conv(h, ns + "." + head_prefix, bn_eps, &hd))) return true;
    if (c1->Cin != Cin || c1->Cout != C || !d1->dw || d1->Cout != C || c2->Cin != C || c2->Cout != C || !d2->dw || d2->Cout != C ||
        rr->Cin != Cin || rr->Cout != C || (hd && (hd->Cin != C || hd->Cout != 1)) || (xlow.cs & 7)) {
      rc = fail(YSP_EINVAL, "doublelight_tc %s: unexpected weight shapes", p.c_str());
      return true;
    }
    const std::string key = ns + "." + p + ".tcpack";
    auto it = h->vecs.find(key);
    if (it == h->vecs.end()) {
      float* buf = nullptr;
      if (cudaMalloc(&buf, dlc_tc_pack_bytes(Cin, C)) != cudaSuccess) { rc = fail(YSP_ECUDA, "cudaMalloc(tcpack)"); return true; }
      DlcTcPrep q = {};
      q.w1 = c1->w; q.c1 = c1->bias; q.w1ld = c1->wld; q.dw1 = d1->w; q.b1 = d1->bias;
      q.w2 = c2->w; q.c2 = c2->bias; q.w2ld = c2->wld; q.dw2 = d2->w; q.b3 = d2->bias;
      q.wr = rr->w; q.cr = rr->bias; q.wrld = rr->wld; q.wo = hd ? hd->w : nullptr; q.wold = hd ? hd->wld : 0;
      q.Cin = Cin; q.C = C;
      launch_dlc_tc_prepare(q, buf, nullptr);
      if (cudaDeviceSynchronize() != cudaSuccess) { rc = fail(YSP_ECUDA, "dlc_tc_prepare"); return true; }
      it = h->vecs.emplace(key, buf).first;
    }
    DlcTcP q = {};
    q.wpack = reinterpret_cast<const uint8_t*>(it->second);
    q.bo = hd ? hd->bias : nullptr;
    q.N = xlow.N; q.h = xlow.H; q.w = xlow.W; q.Cin = Cin; q.C = C; q.x_cs = xlow.cs; q.out_cs = out.cs;
    Plan* pl = plan;
    const double hi = (double)xlow.N * xlow.H * xlow.W * 4;
    emit([=](RunCtx& c) {
      DlcTcP r = q;
      r.x = pl->ptr(c, xlow); r.out = pl->ptr(c, out);
      launch_dlc_tc(r, c.s);
    }, {&xlow, &out}, 1,
    // ALGORITHMIC flops of the reference stage (two hi-res 1x1 convs on up(x), two depthwise 3x3, one 1x1, head), not the
    // 9x larger dense-equivalent count the composite tensor-core convs execute
    StepInfo{p + ".tc", "dlc_tc", tbytes(xlow) + (head ? hi * 4 : tbytes(out)),
             2.0 * hi * (2.0 * Cin * C + 18.0 * C + (double)C * C + (head ? C : 0)), 1});
    return true;
  }

  // DoubleLightConv (YOLOSegPlusPlus.py:33-58) on an already-upsampled input
  void doublelight(const std::string& p, TRef x, TRef out) {
    TRef r = alloc(x.N, x.H, x.W, out.C);
    conv(p + ".residual_conv", x, r, 1, 1, ACT_NONE);
    TRef a = alloc(x.N, x.H, x.W, out.C);
    conv(p + ".conv.0.conv1", x, a, 1, 1, ACT_NONE);
    TRef b = alloc(x.N, x.H, x.W, out.C);
    dw(p + ".conv.0.conv2", a, b, 3, ACT_SILU);
    TRef c = alloc(x.N, x.H, x.W, out.C);
    conv(p + ".conv.1.conv1", b, c, 1, 1, ACT_NONE);
    dw(p + ".conv.1.conv2", c, out, 3, ACT_SILU, &r);
  }
};

// pack buffers by lifetime (greedy first-fit over a timeline); keep_all => no reuse
static void assign_offsets(Plan* plan, bool keep_all) {
  struct Live { size_t off, bytes; int last; };
  // steps of a fork-join region may execute in any order: a buffer touched inside one is live for the whole region
  for (auto& b : plan->bufs)
    for (int r : b.regions) { b.first = std::min(b.first, plan->region_span[r].first); b.last = std::max(b.last, plan->region_span[r].second); }
  std::vector<int> order(plan->bufs.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return plan->bufs[a].first < plan->bufs[b].first; });
  std::vector<Live> live;
  size_t top = 0;
  for (int id : order) {
    BufInfo& b = plan->bufs[id];
    if (b.last < 0) { b.off = 0; continue; }   // never used
    if (!keep_all)
      live.erase(std::remove_if(live.begin(), live.end(), [&](const Live& l) { return l.last < b.first; }), live.end());
    std::sort(live.begin(), live.end(), [](const Live& x, const Live& y) { return x.off < y.off; });
    size_t off = 0;
    for (const Live& l : live) {
      if (off + b.bytes <= l.off) break;
      off = std::max(off, l.off + l.bytes);
    }
    b.off = off;
    live.push_back({off, b.bytes, b.last});
    top = std::max(top, off + b.bytes);
  }
  plan->ws_bytes = top;
}

static cudaEvent_t sync_event(ysp_handle* h, size_t i);
enum { EV_BOTT = 53 };   // recorded by the detector plan when the bottleneck logits are written (see ysp_pipeline)

static void input_step(Builder& g, TRef x, int B, int H, int W) {
  Plan* pl = g.plan; int dt = g.dt;
  g.emit([=](RunCtx& c) {
    if (c.ext[X_IMG_U8]) launch_u8_to_nhwc((const uint8_t*)c.ext[X_IMG_U8], pl->ptr(c, x), B, H, W, x.cs, dt, c.s);
    else launch_nchw_to_nhwc((const float*)c.ext[X_IMG], pl->ptr(c, x), B, 4, H, W, x.cs, dt, c.s);
  }, {&x}, 1, StepInfo{"input", "input_layout", (double)B * H * W * 4 * 4 + Builder::tbytes(x), 0.0, 1});
}

// Detector: YOLOv12n, 4-ch, nc=1.  Input HxW is zero-padded (virtually) to S = ceil32.
static int build_detector(ysp_handle* h, Plan* plan, int B, int H, int W) {
  Builder g(h, plan, "det", 1e-3);
  const int SH = (H + 31) / 32 * 32, SW = (W + 31) / 32 * 32;
  auto dims = [&](int s, int& oh, int& ow) { oh = SH / s; ow = SW / s; };
  int h2, w2, h4, w4, h8, w8, h16, w16, h32, w32;
  dims(2, h2, w2); dims(4, h4, w4); dims(8, h8, w8); dims(16, h16, w16); dims(32, h32, w32);
  // backbone
  TRef t0 = g.alloc(B, h2, w2, 16);  g.stem("model.0", t0, B, H, W);                        g.name("model.0", t0);
  TRef t1 = g.alloc(B, h4, w4, 32);  g.conv("model.1", t0, t1, 3, 2, ACT_SILU);             g.name("model.1", t1);
  plan->share_step = (int)plan->steps.size(); plan->share_buf = t1.buf;
  plan->bufs[t1.buf].last = 1 << 29;          // kept for the whole plan: the seg head may read it (shared stem, see build_seg)
  TRef t2 = g.alloc(B, h4, w4, 64);  g.c3k2("model.2", t1, t2, false, 0.25, true);          g.name("model.2", t2);
  TRef t3 = g.alloc(B, h8, w8, 64);  g.conv("model.3", t2, t3, 3, 2, ACT_SILU);             g.name("model.3", t3);
  TRef cat13 = g.alloc(B, h8, w8, 256);        // [up(L11) 128 | L4 128]
  TRef t4 = Builder::slice(cat13, 128, 128);   g.c3k2("model.4", t3, t4, false, 0.25, true);          g.name("model.4", t4);
  TRef t5 = g.alloc(B, h16, w16, 128); g.conv("model.5", t4, t5, 3, 2, ACT_SILU);           g.name("model.5", t5);
  TRef cat10 = g.alloc(B, h16, w16, 384);      // [up(L8) 256 | L6 128]
  TRef t6 = Builder::slice(cat10, 256, 128);   g.a2c2f("model.6", t5, t6, 2, true, 4);      g.name("model.6", t6);
  TRef t7 = g.alloc(B, h32, w32, 256); g.conv("model.7", t6, t7, 3, 2, ACT_SILU);           g.name("model.7", t7);
  TRef cat19 = g.alloc(B, h32, w32, 384);      // [L18 128 | L8 256]
  TRef t8 = Builder::slice(cat19, 128, 256);   g.a2c2f("model.8", t7, t8, 2, true, 1);      g.name("model.8", t8);
  // neck
  g.ew(1, t8, nullptr, Builder::slice(cat10, 0, 256));                                       // 9, 10
  TRef cat16 = g.alloc(B, h16, w16, 192);      // [L15 64 | L11 128]
  TRef t11 = Builder::slice(cat16, 64, 128);   g.a2c2f("model.11", cat10, t11, 1, false, -1); g.name("model.11", t11);
  g.ew(1, t11, nullptr, Builder::slice(cat13, 0, 128));                                      // 12, 13
  TRef t14 = g.alloc(B, h8, w8, 64);   g.a2c2f("model.14", cat13, t14, 1, false, -1);       g.name("model.14", t14);
  // Detect (nc=1): raw maps NHWC fp32 [.., 65] (row stride 68).  The P3 class branch only needs layer 14, and its last
  // channel is the seg head's bottleneck (evaluate_model.py:142-144): it runs on a side lane NEXT TO layers 15-20 and
  // publishes the bottleneck (event EV_BOTT) so that the decoder can start a millisecond before the detector is done.
  TRef raws[3];
  raws[0] = g.alloc(B, t14.H, t14.W, 65, DT_F32, 68);
  auto cls_branch = [&](int i, const TRef& f) {
    std::string s = std::to_string(i);
    TRef d1 = g.alloc(B, f.H, f.W, f.C), p1 = g.alloc(B, f.H, f.W, 64), d2 = g.alloc(B, f.H, f.W, 64),
         p2 = g.alloc(B, f.H, f.W, 64);
    g.dw("model.21.cv3." + s + ".0.0", f, d1, 3, ACT_SILU);
    g.conv("model.21.cv3." + s + ".0.1", d1, p1, 1, 1, ACT_SILU);
    g.dw("model.21.cv3." + s + ".1.0", p1, d2, 3, ACT_SILU);
    g.conv("model.21.cv3." + s + ".1.1", d2, p2, 1, 1, ACT_SILU);
    g.conv("model.21.cv3." + s + ".2", p2, Builder::slice(raws[i], 64, 1), 1, 1, ACT_NONE);
  };
  g.fork();
  g.lane(1);
  cls_branch(0, t14);
  {
    Plan* pl0 = plan; TRef r0 = raws[0]; ysp_handle* hh = h;
    const int bh = H / 8, bw = W / 8;
    g.emit([=](RunCtx& c) {
      if (c.ext[X_BOTT]) launch_bottleneck_nhwc((const float*)pl0->ptr(c, r0), 68, 64, B, r0.H, r0.W, (float*)c.ext[X_BOTT], bh, bw, c.s);
      cudaEventRecord(sync_event(hh, EV_BOTT), c.s);
    }, {&r0}, 1, StepInfo{"bottleneck", "bottleneck", (double)B * bh * bw * 8, 0.0, 1});
  }
  g.lane(0);
  g.conv("model.15", t14, Builder::slice(cat16, 0, 64), 3, 2, ACT_SILU);                     // 15, 16
  TRef t17 = g.alloc(B, h16, w16, 128); g.a2c2f("model.17", cat16, t17, 1, false, -1);      g.name("model.17", t17);
  g.conv("model.18", t17, Builder::slice(cat19, 0, 128), 3, 2, ACT_SILU);                    // 18, 19
  TRef t20 = g.alloc(B, h32, w32, 256); g.c3k2("model.20", cat19, t20, true, 0.5, true);    g.name("model.20", t20);
  g.join();
  TRef feats[3] = {t14, t17, t20};
  g.fork();                                   // the remaining 5 chains: 3 box branches + class branches of P4, P5
  for (int i = 0; i < 3; ++i) {
    TRef f = feats[i];
    std::string s = std::to_string(i);
    if (i > 0) raws[i] = g.alloc(B, f.H, f.W, 65, DT_F32, 68);
    TRef raw = raws[i];
    TRef a = g.alloc(B, f.H, f.W, 64), b = g.alloc(B, f.H, f.W, 64);
    g.lane(2 * i);
    g.conv("model.21.cv2." + s + ".0", f, a, 3, 1, ACT_SILU);
    g.conv("model.21.cv2." + s + ".1", a, b, 3, 1, ACT_SILU);
    g.conv("model.21.cv2." + s + ".2", b, Builder::slice(raw, 0, 64), 1, 1, ACT_NONE);
    if (i > 0) { g.lane(2 * i + 1); cls_branch(i, f); }
  }
  g.join();
  if (g.rc) return g.rc;
  DecodeP dp = {};
  dp.B = B; dp.nc = 1; dp.cs = 68;
  dp.A = 0;
  for (int i = 0; i < 3; ++i) { dp.h[i] = raws[i].H; dp.w[i] = raws[i].W; dp.A += raws[i].H * raws[i].W; }
  dp.stride[0] = 8.f; dp.stride[1] = 16.f; dp.stride[2] = 32.f;
  dp.bh = H / 8; dp.bw = W / 8;
  Plan* pl = plan;
  TRef r0 = raws[0], r1 = raws[1], r2 = raws[2];
  g.emit([=](RunCtx& c) {
    DecodeP q = dp;
    q.raw[0] = (const float*)pl->ptr(c, r0); q.raw[1] = (const float*)pl->ptr(c, r1); q.raw[2] = (const float*)pl->ptr(c, r2);
    q.y = (float*)c.ext[X_Y]; q.p[0] = (float*)c.ext[X_P3]; q.p[1] = (float*)c.ext[X_P4]; q.p[2] = (float*)c.ext[X_P5];
    q.bott = nullptr;                             // already published by the "bottleneck" step above
    launch_detect_decode(q, c.s);
  }, {&r0, &r1, &r2}, 1, StepInfo{"detect_decode", "detect_decode", (double)B * dp.A * (65 * 4 * 2 + 5 * 4), (double)B * dp.A * 200.0, 1});
  assign_offsets(plan, h->keep_all);
  return 0;
}

// YOLO-Seg++ head (YOLOSegPlusPlus.py:150-178, :242-272)
// `shared_stem` (pipeline, both tensor-core modes): encoder layers 0 and 1 are NOT recomputed.  They are stride-2 3x3 convs whose
// outputs inside the H/4 x W/4 window depend only on input pixels inside the H x W image, and the detector computes the
// very same layers (shared weights, YOLOSegPlusPlus.py:150) on the image zero-padded to %32 -- so the detector's layer-1
// output, read through a pitched H/4 x W/4 view, IS encoder.1's output, exactly.  (From layer 2 on the 3x3 stride-1
// convs see real values instead of zero padding at the window border, so sharing stops there: decision D1.)
static int build_seg(ysp_handle* h, Plan* plan, int B, int H, int W, bool shared_stem = false) {
  Builder g(h, plan, "seg", 1e-5);
  const int h2 = (H + 1) / 2, w2 = (W + 1) / 2, h4 = (h2 + 1) / 2, w4 = (w2 + 1) / 2, h8 = (h4 + 1) / 2, w8 = (w4 + 1) / 2;
  // encoder = detector layers 0..4 (frozen, BN already folded; falls back to eps 1e-3 when it arrives unfused)
  g.bn_eps = 1e-3;
  TRef e1;
  if (shared_stem) {
    const int SH = (H + 31) / 32 * 32, SW = (W + 31) / 32 * 32;
    e1 = Builder::ext(X_E1, B, h4, w4, 32, 32, g.dt);
    e1.pw = SW / 4; e1.ph = SH / 4;
    plan->shared_stem = true;
  } else {
    TRef e0 = g.alloc(B, h2, w2, 16);  g.stem("encoder.0", e0, B, H, W);                    g.name("encoder.0", e0);
    e1 = g.alloc(B, h4, w4, 32);       g.conv("encoder.1", e0, e1, 3, 2, ACT_SILU);         g.name("encoder.1", e1);
  }
  TRef cat2 = g.alloc(B, h4, w4, 128);          // dec2 input: [dec1 out 64 | skipA 64]
  TRef skipA = Builder::slice(cat2, 64, 64);    g.c3k2("encoder.2", e1, skipA, false, 0.25, true);   g.name("encoder.2", skipA);
  TRef e3 = g.alloc(B, h8, w8, 64);  g.conv("encoder.3", skipA, e3, 3, 2, ACT_SILU);        g.name("encoder.3", e3);
  // The reference's ablation variant `_YOLOSegPlusPlus.py` (:157, :264-268) has no bottleneck channel: decoder.0 is
  // C3Ghost(128, 96) and its input is the skip alone.  Detected from the checkpoint (Cin of decoder.0.0.cv1); the
  // `logits` argument is then ignored.
  bool use_logits = true;
  {
    auto it = h->host.find("seg.decoder.0.0.cv1.conv.weight");
    if (it != h->host.end() && it->second.shape.size() >= 2 && it->second.shape[1] == 128) use_logits = false;
  }
  const int c0 = use_logits ? 129 : 128, c0s = use_logits ? ((g.dt == DT_F32 && !h->tc32()) ? 132 : 144) : 128;
  TRef cat0 = g.alloc(B, h8, w8, c0, -1, c0s);  // dec0 input: [skipB 128 | logits 1 | zero pad]
  cat0.zpad = use_logits && (g.dt == DT_BF16 || h->tc32());
  TRef skipB = Builder::slice(cat0, 0, 128);    g.c3k2("encoder.4", e3, skipB, false, 0.25, true);   g.name("encoder.4", skipB);
  g.bn_eps = 1e-5;
  plan->split = (int)plan->steps.size();       // everything above is independent of the detector
  if (use_logits) {
    TRef lg = Builder::slice(cat0, 128, 1);
    Plan* pl = plan; int dt = g.dt; int zp = c0s - 129;
    g.emit([=](RunCtx& c) {
      launch_logits_to_nhwc((const float*)c.ext[X_LOGITS], pl->ptr(c, lg), B, h8, w8, lg.cs, zp, dt, c.s);
    }, {&lg}, 1, StepInfo{"logits_in", "input_layout", (double)B * h8 * w8 * 8, 0.0, 1});
  }
  // decoder
  TRef d0 = g.alloc(B, h8, w8, 96);
  g.c3ghost("decoder.0.0", cat0, d0);
  g.eca("decoder.0.1.conv.weight", d0);                                                       g.name("decoder.0", d0);
  const bool fuse = getenv("YSP_NO_FUSE") == nullptr;
  TRef d1 = Builder::slice(cat2, 0, 64);
  if (fuse) g.doublelight_fused("decoder.1.1", d0, d1, "");
  else { TRef u1 = g.alloc(B, h4, w4, 96); g.ew(2, d0, nullptr, u1); g.doublelight("decoder.1.1", u1, d1); }
  g.name("decoder.1", d1);
  TRef d2 = g.alloc(B, h4, w4, 64);
  g.c3ghost("decoder.2.0", cat2, d2);
  g.eca("decoder.2.1.conv.weight", d2);                                                       g.name("decoder.2", d2);
  TRef d3 = g.alloc(B, h2, w2, 32);
  TRef out = Builder::ext(X_OUT, B, H, W, 1, 1, DT_F32);
  if (fuse) {
    if (!g.doublelight_tc("decoder.3.1", d2, d3, "")) g.doublelight_fused("decoder.3.1", d2, d3, "");
    g.name("decoder.3", d3);
    if (!g.doublelight_tc("decoder.4.1", d3, out, "output")) g.doublelight_fused("decoder.4.1", d3, out, "output");
  } else {
    TRef u3 = g.alloc(B, h2, w2, 64);   g.ew(2, d2, nullptr, u3);
    g.doublelight("decoder.3.1", u3, d3);                                                     g.name("decoder.3", d3);
    TRef u4 = g.alloc(B, H, W, 32);     g.ew(2, d3, nullptr, u4);
    TRef d4 = g.alloc(B, H, W, 16);     g.doublelight("decoder.4.1", u4, d4);                g.name("decoder.4", d4);
    g.conv("output", d4, out, 1, 1, ACT_NONE);
  }
  if (g.rc) return g.rc;
  assign_offsets(plan, h->keep_all);
  return 0;
}

static int get_plan(ysp_handle* h, const char* kind, int B, int H, int W, Plan** out) {
  char key[96];
  snprintf(key, sizeof(key), "%s:%d:%d:%d:%d", kind, B, H, W, (int)h->keep_all);
  auto it = h->plans.find(key);
  if (it != h->plans.end()) { *out = it->second.get(); return 0; }
  std::unique_ptr<Plan> p(new Plan());
  int rc = (kind[0] == 'd') ? build_detector(h, p.get(), B, H, W) : build_seg(h, p.get(), B, H, W, strcmp(kind, "segs") == 0);
  if (rc) return rc;
  *out = p.get();
  h->plans[key] = std::move(p);
  return 0;
}

static void prof_add(ysp_handle* h, const StepInfo& in, float ms) {
  ProfAgg& a = h->prof[in.name + "|" + in.kind];
  a.kind = in.kind; a.ms += ms; a.bytes += in.bytes; a.flops += in.flops; a.calls += 1; a.launches += in.launches;
}

static cudaEvent_t sync_event(ysp_handle* h, size_t i) {
  while (h->sync_events.size() <= i) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); h->sync_events.push_back(e); }
  return h->sync_events[i];
}
static cudaStream_t lane_stream(ysp_handle* h, int lane, cudaStream_t main) {
  if (lane == 0) return main;
  while ((int)h->lanes.size() < lane) { cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking); h->lanes.push_back(s); }
  return h->lanes[lane - 1];
}

// Run steps [lo, hi) of a plan on c.s.  Fork-join regions spread their lanes over extra streams (event fork/join);
// while profiling everything runs serially on c.s with an event pair around each step.
static int run_plan(ysp_handle* h, Plan* p, RunCtx& c, int lo = 0, int hi = -1, size_t ev_base = 0) {
  if (hi < 0) hi = (int)p->steps.size();
  static const bool no_lanes = getenv("YSP_NO_LANES") != nullptr;
  if (h->profiling) {
    size_t need = 2 * p->steps.size();
    while (h->prof_events.size() < need) { cudaEvent_t e; cudaEventCreate(&e); h->prof_events.push_back(e); }
    for (int i = lo; i < hi; ++i) {
      cudaEventRecord(h->prof_events[2 * i], c.s);
      p->steps[i](c);
      cudaEventRecord(h->prof_events[2 * i + 1], c.s);
    }
    cudaStreamSynchronize(c.s);
    for (int i = lo; i < hi; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, h->prof_events[2 * i], h->prof_events[2 * i + 1]);
      prof_add(h, p->infos[i], ms);
    }
  } else if (no_lanes) {
    for (int i = lo; i < hi; ++i) p->steps[i](c);
  } else {
    const cudaStream_t main = c.s;
    size_t ev = ev_base;
    int cur = 0;
    unsigned used = 0;                                     // lanes (bit k) used in the current region
    auto join = [&]() {
      for (int k = 1; k < 32; ++k)
        if (used & (1u << k)) {
          cudaEvent_t e = sync_event(h, ev++);
          cudaEventRecord(e, lane_stream(h, k, main));
          cudaStreamWaitEvent(main, e, 0);
        }
      used = 0;
    };
    for (int i = lo; i < hi; ++i) {
      const int r = p->region[i];
      if (r != cur) {
        if (cur) join();
        if (r) {
          cudaEvent_t e = sync_event(h, ev++);
          cudaEventRecord(e, main);
          for (int j = p->region_span[r].first; j <= p->region_span[r].second; ++j) {
            int k = p->lane[j];
            if (k && !(used & (1u << k))) { cudaStreamWaitEvent(lane_stream(h, k, main), e, 0); used |= 1u << k; }
          }
        }
        cur = r;
      }
      RunCtx cc = c;
      cc.s = lane_stream(h, r ? p->lane[i] : 0, main);
      p->steps[i](cc);
    }
    if (cur) join();
  }
  h->last_plan = p;
  for (int i = lo; i < hi; ++i) h->last_launches += p->infos[i].launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(YSP_ECUDA, "kernel launch: %s", cudaGetErrorString(e));
  return 0;
}

static int check_device(ysp_handle* h) {
  if (!h) return fail(YSP_EINVAL, "null handle");
  CUDA_OK(cudaSetDevice(h->device));
  return 0;
}

}  // namespace

// =====================================================================================================================
// C ABI
// =====================================================================================================================
extern "C" {

int ysp_version(void) { return 100; }
const char* ysp_last_error(void) { return g_err; }

int ysp_create(ysp_handle** out, int device, int mode) {
  if (!out || (mode != YSP_MODE_FP32 && mode != YSP_MODE_BF16 && mode != YSP_MODE_TC32)) return fail(YSP_EINVAL, "ysp_create: bad arguments");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return fail(YSP_ECUDA, "no CUDA device (%s); libysp has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(YSP_EINVAL, "device %d out of range (%d devices)", device, n);
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(YSP_ECUDA, "device %d is sm_%d%d; libysp is built for sm_100a only", device, prop.major, prop.minor);
  ysp_handle* h = new ysp_handle();
  h->device = device; h->mode = mode;
  h->keep_all = getenv("YSP_KEEP_INTERMEDIATES") != nullptr;
  h->no_share = getenv("YSP_NO_SHARE") != nullptr;
  *out = h;
  return 0;
}

void ysp_destroy(ysp_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  h->plans.clear();
  for (auto& kv : h->convs) { cudaFree(kv.second.w); cudaFree(kv.second.bias); if (kv.second.w_tc) cudaFree(kv.second.w_tc); }
  for (auto& kv : h->vecs) cudaFree(kv.second);
  for (auto st : h->lanes) cudaStreamDestroy(st);
  if (h->aux) cudaStreamDestroy(h->aux);
  for (auto e : h->sync_events) cudaEventDestroy(e);
  for (auto e : h->prof_events) cudaEventDestroy(e);
  delete h;
}

int ysp_set_keep_intermediates(ysp_handle* h, int on) {
  if (!h) return fail(YSP_EINVAL, "null handle");
  h->keep_all = on != 0;
  return 0;
}

int ysp_load_weight(ysp_handle* h, const char* name, const float* h_data, int ndim, const int64_t* shape) {
  if (!h || !name || !h_data || ndim < 0 || ndim > 8) return fail(YSP_EINVAL, "ysp_load_weight: bad arguments");
  HostT t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= (size_t)shape[i]; }
  t.d.assign(h_data, h_data + n);
  h->host[name] = std::move(t);
  return 0;
}

int ysp_finalize(ysp_handle* h, int which) {
  int rc = check_device(h);
  if (rc) return rc;
  // dry-run both graphs at a nominal shape: touches (packs + uploads) every weight the topology needs
  if (which & 1) {
    Plan p;
    if ((rc = build_detector(h, &p, 1, 64, 64))) return rc;
    h->det_ready = true;
  }
  if (which & 2) {
    Plan p;
    if ((rc = build_seg(h, &p, 1, 64, 64))) return rc;
    h->seg_ready = true;
  }
  CUDA_OK(cudaDeviceSynchronize());
  return 0;
}

size_t ysp_pipeline_workspace_bytes(ysp_handle* h, int B, int H, int W, int max_det) {
  if (!h || B <= 0 || H <= 0 || W <= 0) return 0;
  if (cudaSetDevice(h->device) != cudaSuccess) return 0;
  if (max_det <= 0) max_det = 300;
  size_t total = 0;
  Plan* p = nullptr;
  if (h->det_ready && get_plan(h, "det", B, H, W, &p) == 0) total += align_up(p->ws_bytes, 256);
  if (h->seg_ready && (H % 8 == 0) && (W % 8 == 0)) {
    // ysp_pipeline runs the "segs" plan (shared stem) in the tensor-core modes and "seg" otherwise: size for the larger of the two
    size_t sb = 0;
    if (get_plan(h, "seg", B, H, W, &p) == 0) sb = p->ws_bytes;
    if ((h->mode == YSP_MODE_BF16 || h->tc32()) && !h->no_share && get_plan(h, "segs", B, H, W, &p) == 0) sb = std::max(sb, p->ws_bytes);
    total += align_up(sb, 256);
  }
  const int SH = (H + 31) / 32 * 32, SW = (W + 31) / 32 * 32;
  const int A = (SH / 8) * (SW / 8) + (SH / 16) * (SW / 16) + (SH / 32) * (SW / 32);
  total += align_up((size_t)B * 5 * A * 4, 256);                 // y when the caller does not want it
  total += align_up((size_t)B * (H / 8) * (W / 8) * 4, 256);     // bottleneck logits
  total += nms_workspace_bytes(B, 5, A, max_det) + 256;
  return total;
}

size_t ysp_workspace_bytes(ysp_handle* h, int B, int H, int W) { return ysp_pipeline_workspace_bytes(h, B, H, W, 300); }

int ysp_normalize_u8(const uint8_t* d_u8, float* d_out, int B, int H, int W, void* stream) {
  if (!d_u8 || !d_out || B < 0 || H <= 0 || W <= 0) return fail(YSP_EINVAL, "ysp_normalize_u8: bad arguments");
  if (B == 0) return 0;
  launch_normalize_u8(d_u8, d_out, B, H, W, (cudaStream_t)stream);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int ysp_detector_forward(ysp_handle* h, const float* d_img, int B, int H, int W, float* d_y, float* d_p3, float* d_p4,
                         float* d_p5, void* d_ws, size_t ws_bytes, void* stream) {
  int rc = check_device(h);
  if (rc) return rc;
  if (!h->det_ready) return fail(YSP_ESTATE, "detector weights not finalized");
  if (!d_img || B <= 0 || H <= 0 || W <= 0) return fail(YSP_EINVAL, "ysp_detector_forward: bad arguments");
  Plan* p = nullptr;
  if ((rc = get_plan(h, "det", B, H, W, &p))) return rc;
  if (ws_bytes < p->ws_bytes || !d_ws) return fail(YSP_ESTATE, "workspace too small: need %zu bytes, got %zu", p->ws_bytes, ws_bytes);
  RunCtx c = {};
  c.ws = (char*)d_ws; c.s = (cudaStream_t)stream;
  c.ext[X_IMG] = (void*)d_img; c.ext[X_Y] = d_y; c.ext[X_P3] = d_p3; c.ext[X_P4] = d_p4; c.ext[X_P5] = d_p5;
  h->last_launches = 0;
  return run_plan(h, p, c);
}

int ysp_bottleneck(const float* d_p3, int B, int C, int Hs, int Ws, float* d_logits, int hh, int ww, void* stream) {
  if (!d_p3 || !d_logits || B < 0 || C <= 0 || hh > Hs || ww > Ws) return fail(YSP_EINVAL, "ysp_bottleneck: bad arguments");
  if (B == 0) return 0;
  launch_bottleneck(d_p3, B, C, Hs, Ws, d_logits, hh, ww, (cudaStream_t)stream);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int ysp_segpp_forward(ysp_handle* h, const float* d_x, const float* d_logits, float* d_out, int B, int H, int W,
                      void* d_ws, size_t ws_bytes, void* stream) {
  int rc = check_device(h);
  if (rc) return rc;
  if (!h->seg_ready) return fail(YSP_ESTATE, "seg head weights not finalized");
  if (!d_x || !d_logits || !d_out || B <= 0) return fail(YSP_EINVAL, "ysp_segpp_forward: bad arguments");
  if (H % 8 || W % 8 || H <= 0 || W <= 0) return fail(YSP_EINVAL, "H and W must be positive multiples of 8 (got %dx%d)", H, W);
  Plan* p = nullptr;
  if ((rc = get_plan(h, "seg", B, H, W, &p))) return rc;
  if (ws_bytes < p->ws_bytes || !d_ws) return fail(YSP_ESTATE, "workspace too small: need %zu bytes, got %zu", p->ws_bytes, ws_bytes);
  RunCtx c = {};
  c.ws = (char*)d_ws; c.s = (cudaStream_t)stream;
  c.ext[X_IMG] = (void*)d_x; c.ext[X_LOGITS] = (void*)d_logits; c.ext[X_OUT] = d_out;
  h->last_launches = 0;
  return run_plan(h, p, c);
}

// Frozen encoder only (YOLOSegPlusPlus.py:255-259): runs the seg plan up to the first decoder step and hands the two
// skips to the caller as dense NHWC fp32 -- the input of the training step (train.cu).
int ysp_encoder_forward(ysp_handle* h, const float* d_x, float* d_skipA, float* d_skipB, int B, int H, int W, void* d_ws,
                        size_t ws_bytes, void* stream) {
  int rc = check_device(h);
  if (rc) return rc;
  if (!h->seg_ready) return fail(YSP_ESTATE, "seg head weights not finalized");
  if (!d_x || !d_skipA || !d_skipB || B <= 0) return fail(YSP_EINVAL, "ysp_encoder_forward: bad arguments");
  if (H % 8 || W % 8 || H <= 0 || W <= 0) return fail(YSP_EINVAL, "H and W must be positive multiples of 8 (got %dx%d)", H, W);
  Plan* p = nullptr;
  if ((rc = get_plan(h, "seg", B, H, W, &p))) return rc;
  if (ws_bytes < p->ws_bytes || !d_ws) return fail(YSP_ESTATE, "workspace too small: need %zu bytes, got %zu", p->ws_bytes, ws_bytes);
  RunCtx c = {};
  c.ws = (char*)d_ws; c.s = (cudaStream_t)stream;
  c.ext[X_IMG] = (void*)d_x;
  h->last_launches = 0;
  if ((rc = run_plan(h, p, c, 0, p->split, 0))) return rc;
  auto a = p->named.find("seg:encoder.2"), b = p->named.find("seg:encoder.4");
  if (a == p->named.end() || b == p->named.end()) return fail(YSP_ESTATE, "ysp_encoder_forward: skips not found in the plan");
  const TRef &ta = a->second, &tb = b->second;
  launch_export_view(p->ptr(c, ta), ta.cs, ta.dt, d_skipA, ta.C, (long long)ta.N * ta.H * ta.W, c.s);
  launch_export_view(p->ptr(c, tb), tb.cs, tb.dt, d_skipB, tb.C, (long long)tb.N * tb.H * tb.W, c.s);
  h->last_launches += 2;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int ysp_resize_u8(const uint8_t* d_src, int B, int h, int w, int C, int dh, int dw, int interp, uint8_t* d_dst_u8,
                  float* d_dst_f32, void* stream) {
  if (!d_src || (!d_dst_u8 && !d_dst_f32) || B < 0 || h <= 0 || w <= 0 || dh <= 0 || dw <= 0)
    return fail(YSP_EINVAL, "ysp_resize_u8: bad arguments");
  if (C != 1 && C != 4) return fail(YSP_EINVAL, "ysp_resize_u8: C must be 1 (mask) or 4 (slice), got %d", C);
  if (interp != 0 && interp != 1) return fail(YSP_EINVAL, "ysp_resize_u8: interp must be 0 (INTER_NEAREST) or 1 (INTER_LINEAR)");
  if (B == 0) return 0;
  launch_resize_u8(d_src, B, h, w, C, dh, dw, interp, d_dst_u8, d_dst_f32, (cudaStream_t)stream);
  CUDA_OK(cudaGetLastError());
  return 0;
}

size_t ysp_nms_workspace_bytes(int B, int C, int A, int max_det) { return nms_workspace_bytes(B, C, A, max_det); }

int ysp_nms(const float* d_pred, int B, int C, int A, int nc, float conf_thres, float iou_thres, int max_det,
            int max_nms, float max_wh, int agnostic, const int32_t* d_classes, int n_classes, float* d_out_boxes,
            int64_t* d_out_idx, int32_t* d_out_count, void* d_ws, size_t ws_bytes, void* stream) {
  if (!(conf_thres >= 0.f && conf_thres <= 1.f)) return fail(YSP_EINVAL, "Invalid Confidence threshold %g, valid values are between 0.0 and 1.0", conf_thres);
  if (!(iou_thres >= 0.f && iou_thres <= 1.f)) return fail(YSP_EINVAL, "Invalid IoU %g, valid values are between 0.0 and 1.0", iou_thres);
  if (nc <= 0) nc = C - 4;
  if (B < 0 || A < 0 || C < 5 || nc < 1 || 4 + nc > C || max_det < 0 || max_nms < 0) return fail(YSP_EINVAL, "ysp_nms: bad shape B=%d C=%d A=%d nc=%d", B, C, A, nc);
  if (B == 0) return 0;
  if (!d_pred || !d_out_idx || !d_out_count || !d_out_boxes) return fail(YSP_EINVAL, "ysp_nms: null pointer");
  if (max_det == 0 || A == 0) { CUDA_OK(cudaMemsetAsync(d_out_count, 0, 4 * (size_t)B, (cudaStream_t)stream)); return 0; }
  int rc = launch_nms(d_pred, B, C, A, nc, conf_thres, iou_thres, max_det, max_nms, max_wh, agnostic, d_classes, n_classes,
                      d_out_boxes, d_out_idx, d_out_count, d_ws, ws_bytes, (cudaStream_t)stream);
  if (rc) return fail(YSP_ESTATE, "ysp_nms: workspace too small (need %zu bytes)", nms_workspace_bytes(B, C, A, max_det));
  CUDA_OK(cudaGetLastError());
  return 0;
}

int ysp_nms_core(const float* d_boxes, const float* d_scores, int N, float iou_thres, int64_t* d_keep, int32_t* d_count,
                 void* d_ws, size_t ws_bytes, void* stream) {
  if (N < 0 || !d_count) return fail(YSP_EINVAL, "ysp_nms_core: bad arguments");
  int rc = launch_nms_core(d_boxes, d_scores, N, iou_thres, d_keep, d_count, d_ws, ws_bytes, (cudaStream_t)stream);
  if (rc) return fail(YSP_ESTATE, "ysp_nms_core: workspace too small (need %zu bytes)", nms_workspace_bytes(1, 5, N, N));
  CUDA_OK(cudaGetLastError());
  return 0;
}

int ysp_xywh2xyxy_inplace(float* d_pred, int B, int C, int A, void* stream) {
  if (!d_pred || C < 4) return fail(YSP_EINVAL, "ysp_xywh2xyxy_inplace: bad arguments");
  launch_xywh2xyxy_inplace(d_pred, B, C, A, (cudaStream_t)stream);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int ysp_mask_dice(const float* d_logits, const float* d_target, int B, int HW, int32_t* d_counts, uint8_t* d_mask,
                  void* stream) {
  if (!d_logits || !d_counts || B < 0 || HW < 0) return fail(YSP_EINVAL, "ysp_mask_dice: bad arguments");
  if (B == 0) return 0;
  launch_mask_dice(d_logits, d_target, nullptr, B, HW, d_counts, d_mask, nullptr, (cudaStream_t)stream);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int ysp_conf_gate(const float* d_det_boxes, const int32_t* d_det_count, int B, int max_det, int row, float conf_gate,
                  int32_t* d_counts, uint8_t* d_mask, int HW, uint8_t* d_gated, void* stream) {
  if (!d_det_boxes || !d_det_count || !d_counts || B < 0 || max_det <= 0 || row < 5 || HW < 0)
    return fail(YSP_EINVAL, "ysp_conf_gate: bad arguments");
  if (B == 0) return 0;
  launch_conf_gate(d_det_boxes, d_det_count, B, max_det, row, conf_gate, d_counts, d_mask, HW, d_gated, (cudaStream_t)stream);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int ysp_pipeline(ysp_handle* h, const ysp_pipeline_io* io, int B, int H, int W, void* d_ws, size_t ws_bytes,
                 void* stream) {
  int rc = check_device(h);
  if (rc) return rc;
  if (!h->det_ready || !h->seg_ready) return fail(YSP_ESTATE, "weights not finalized");
  if (!io || B <= 0 || (!io->d_img && !io->d_img_u8) || !io->d_mask_logits || !io->d_det_boxes || !io->d_det_idx ||
      !io->d_det_count || !io->d_counts)
    return fail(YSP_EINVAL, "ysp_pipeline: bad arguments");
  if (H % 8 || W % 8) return fail(YSP_EINVAL, "H and W must be multiples of 8 (got %dx%d)", H, W);
  if (io->d_mask_bits && (H * W) % 128) return fail(YSP_EINVAL, "d_mask_bits needs H*W %% 128 == 0 (got %d)", H * W);
  if (!(io->conf_thres >= 0.f && io->conf_thres <= 1.f) || !(io->iou_thres >= 0.f && io->iou_thres <= 1.f))
    return fail(YSP_EINVAL, "invalid thresholds");
  Plan *pd = nullptr, *ps = nullptr;
  if ((rc = get_plan(h, "det", B, H, W, &pd))) return rc;
  const bool share = (h->mode == YSP_MODE_BF16 || h->tc32()) && !h->no_share && (long long)B * (H / 4) * (W / 4) >= 128;
  if ((rc = get_plan(h, share ? "segs" : "seg", B, H, W, &ps))) return rc;
  const int SH = (H + 31) / 32 * 32, SW = (W + 31) / 32 * 32;
  const int A = (SH / 8) * (SW / 8) + (SH / 16) * (SW / 16) + (SH / 32) * (SW / 32);
  const int max_det = io->max_det > 0 ? io->max_det : 300;
  // workspace carve-up: [det arena | seg arena | y | bottleneck | nms].  The two arenas are disjoint because the seg
  // encoder (which does not depend on the detector) runs concurrently with the detector on an auxiliary stream.
  const size_t det_b = align_up(pd->ws_bytes, 256), seg_b = align_up(ps->ws_bytes, 256);
  size_t off_y = det_b + seg_b;
  size_t off_b = off_y + align_up((size_t)B * 5 * A * 4, 256);
  size_t off_n = off_b + align_up((size_t)B * (H / 8) * (W / 8) * 4, 256);
  size_t need = off_n + nms_workspace_bytes(B, 5, A, max_det);
  if (!d_ws || ws_bytes < need) return fail(YSP_ESTATE, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
  char* ws = (char*)d_ws;
  float* y = io->d_y ? io->d_y : (float*)(ws + off_y);
  float* bott = io->d_bottleneck ? io->d_bottleneck : (float*)(ws + off_b);
  cudaStream_t s = (cudaStream_t)stream;
  h->last_launches = 0;
  static const bool no_overlap = getenv("YSP_NO_OVERLAP") != nullptr;
  const bool overlap = !h->profiling && !no_overlap && ps->split > 0;
  if (overlap && !h->aux) CUDA_OK(cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking));
  cudaStream_t s2 = overlap ? h->aux : s;
  enum { EV_FORK = 48, EV_ENC, EV_DET, EV_NMS, EV_E1 };          // sync-event slots above those used inside the plans
  RunCtx c = {};
  c.ws = ws; c.s = s;
  c.ext[X_IMG] = (void*)io->d_img; c.ext[X_IMG_U8] = (void*)io->d_img_u8; c.ext[X_Y] = y; c.ext[X_BOTT] = bott;
  RunCtx c2 = {};
  c2.ws = ws + det_b; c2.s = s2;
  c2.ext[X_IMG] = (void*)io->d_img; c2.ext[X_IMG_U8] = (void*)io->d_img_u8; c2.ext[X_LOGITS] = bott;
  c2.ext[X_OUT] = io->d_mask_logits;
  c2.ext[X_E1] = ws + pd->bufs[pd->share_buf].off;                                       // detector layer-1 output (shared stem)
  int det_lo = 0;
  if (overlap) {
    cudaEventRecord(sync_event(h, EV_FORK), s);
    cudaStreamWaitEvent(s2, sync_event(h, EV_FORK), 0);
    if (ps->shared_stem) {                       // the seg encoder starts once the detector's layers 0-1 are done
      if ((rc = run_plan(h, pd, c, 0, pd->share_step))) return rc;
      det_lo = pd->share_step;
      cudaEventRecord(sync_event(h, EV_E1), s);
      cudaStreamWaitEvent(s2, sync_event(h, EV_E1), 0);
    }
    if ((rc = run_plan(h, ps, c2, 0, ps->split, 24))) return rc;                         // seg encoder on the aux stream
  }
  if ((rc = run_plan(h, pd, c, det_lo))) return rc;                                     // evaluate_model.py:141-144
  cudaEvent_t pe[4] = {nullptr, nullptr, nullptr, nullptr};
  if (h->profiling) { for (auto& e : pe) cudaEventCreate(&e); cudaEventRecord(pe[0], s); }
  if (overlap) {
    // aux stream: the decoder starts as soon as the detector plan has published the bottleneck (EV_BOTT, recorded on the
    // P3 class-branch lane), i.e. concurrently with detector layers 15-20, the rest of the Detect head, decode and NMS
    cudaStreamWaitEvent(s2, sync_event(h, EV_BOTT), 0);
    if ((rc = run_plan(h, ps, c2, ps->split, -1, 24))) return rc;                        // :156
    launch_mask_dice(io->d_mask_logits, io->d_target, io->d_target_u8, B, H * W, io->d_counts, io->d_mask, io->d_mask_bits, s2);   // :157-174
    cudaEventRecord(sync_event(h, EV_ENC), s2);
    if (launch_nms(y, B, 5, A, 1, io->conf_thres, io->iou_thres, max_det, 30000, 7680.f, 0, nullptr, 0, io->d_det_boxes,
                   io->d_det_idx, io->d_det_count, ws + off_n, ws_bytes - off_n, s))     // :147
      return fail(YSP_ESTATE, "nms workspace");
    cudaStreamWaitEvent(s, sync_event(h, EV_ENC), 0);
    h->last_launches += 3;
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (launch_nms(y, B, 5, A, 1, io->conf_thres, io->iou_thres, max_det, 30000, 7680.f, 0, nullptr, 0, io->d_det_boxes,
                 io->d_det_idx, io->d_det_count, ws + off_n, ws_bytes - off_n, s))       // :147
    return fail(YSP_ESTATE, "nms workspace");
  if (h->profiling) cudaEventRecord(pe[1], s);
  h->last_launches += 2;
  c2.s = s;
  if ((rc = run_plan(h, ps, c2, 0, -1, 24))) return rc;                                  // :156
  if (h->profiling) cudaEventRecord(pe[2], s);
  launch_mask_dice(io->d_mask_logits, io->d_target, io->d_target_u8, B, H * W, io->d_counts, io->d_mask, io->d_mask_bits, s);   // :157-174
  h->last_launches += 1;
  if (h->profiling) {
    cudaEventRecord(pe[3], s);
    cudaStreamSynchronize(s);
    float m0 = 0.f, m1 = 0.f;
    cudaEventElapsedTime(&m0, pe[0], pe[1]);
    cudaEventElapsedTime(&m1, pe[2], pe[3]);
    StepInfo a{"pipe:nms", "nms", (double)B * A * 5 * 4 + (double)B * max_det * 32, 0.0, 2};
    StepInfo b{"pipe:mask_dice", "mask_dice", (double)B * H * W * (io->d_target ? 8 : 4) + (io->d_mask ? (double)B * H * W : 0.0), 0.0, 1};
    prof_add(h, a, m0); prof_add(h, b, m1);
    for (auto& e : pe) cudaEventDestroy(e);
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

int ysp_objectmap_transform(const float* d_maps, float* d_out, int B, int n, void* stream) {
  if (!d_maps || !d_out || B < 0 || n <= 0) return fail(YSP_EINVAL, "ysp_objectmap_transform: bad arguments");
  launch_objectmap_transform(d_maps, d_out, B, n, (cudaStream_t)stream);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int ysp_scale_boxes(float* d_boxes, long long n, int row, float gain, float pad_x, float pad_y, float w0, float h0, void* stream) {
  if (!d_boxes || n < 0 || row < 4 || !(gain > 0.f)) return fail(YSP_EINVAL, "ysp_scale_boxes: bad arguments");
  launch_scale_boxes(d_boxes, n, row, gain, pad_x, pad_y, w0, h0, (cudaStream_t)stream);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int ysp_last_launch_count(ysp_handle* h) { return h ? h->last_launches : 0; }

int ysp_profile(ysp_handle* h, int enable) {
  if (!h) return fail(YSP_EINVAL, "null handle");
  h->profiling = enable != 0;
  if (enable == 2) h->prof.clear();
  return 0;
}

int ysp_profile_report(ysp_handle* h, char* buf, size_t cap) {
  if (!h || !buf || cap < 3) return fail(YSP_EINVAL, "ysp_profile_report: bad arguments");
  std::string o = "[";
  bool first = true;
  for (auto& kv : h->prof) {
    char line[512];
    std::string nm = kv.first.substr(0, kv.first.find('|'));
    snprintf(line, sizeof(line), "%s{\"name\":\"%s\",\"kind\":\"%s\",\"ms\":%.6f,\"calls\":%ld,\"launches\":%ld,\"bytes\":%.1f,\"flops\":%.1f}",
             first ? "" : ",", nm.c_str(), kv.second.kind.c_str(), kv.second.ms, kv.second.calls, kv.second.launches,
             kv.second.bytes, kv.second.flops);
    o += line;
    first = false;
  }
  o += "]";
  if (o.size() + 1 > cap) return fail(YSP_EINVAL, "ysp_profile_report: buffer too small (need %zu)", o.size() + 1);
  memcpy(buf, o.c_str(), o.size() + 1);
  return (int)o.size();
}

int ysp_debug_tensor(ysp_handle* h, const char* name, void* d_ws, float* d_out, int64_t* shape, void* stream) {
  if (!h || !name || !shape) return fail(YSP_EINVAL, "ysp_debug_tensor: bad arguments");
  for (auto& kv : h->plans) {
    Plan* p = kv.second.get();
    if (p != h->last_plan) continue;
    auto it = p->named.find(name);
    if (it == p->named.end()) continue;
    const TRef& t = it->second;
    shape[0] = t.N; shape[1] = t.C; shape[2] = t.H; shape[3] = t.W;
    if (d_out) {
      RunCtx c = {}; c.ws = (char*)d_ws;
      launch_nhwc_to_nchw_f32(p->ptr(c, t), d_out, t.N, t.H, t.W, t.C, t.cs, t.dt, (cudaStream_t)stream);
      CUDA_OK(cudaGetLastError());
    }
    return 0;
  }
  return fail(YSP_EINVAL, "no tensor named %s in the last plan", name);
}

}  // extern "C"
