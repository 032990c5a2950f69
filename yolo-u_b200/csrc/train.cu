// train.cu -- the seg-head training step behind the C ABI (include/ysp.h: ysp_train_*, ysp_adamw).
//
// Restates, as explicit forward + hand-derived backward passes, what autograd does for the reference's training loop
// (train.py:302-331, non-AMP branch) on the trainable part of YOLOSegPlusPlus (YOLOSegPlusPlus.py:150-178, :242-272):
//   decoder.0  C3Ghost(129->96) + ECA      decoder.1  bilinear x2 + DoubleLightConv(96->64)
//   decoder.2  C3Ghost(128->64) + ECA      decoder.3/4 bilinear x2 + DoubleLightConv(64->32 / 32->16)     output 1x1
// in train() mode (BatchNorm with batch statistics, running statistics updated with momentum 0.1, eps 1e-5), the
// monai Dice loss of train.py:98-104, and leaves gradients in a flat buffer laid out like the parameters so the host
// can all-reduce it (one NCCL call) and apply AdamW (ysp_adamw).  The frozen encoder runs through the inference
// engine (ysp_encoder_forward); its skips arrive here as dense NHWC fp32.
//
// No autograd tape: the topology is fixed, so the forward saves exactly what its backward needs (pre-BN conv outputs,
// batch statistics, unit inputs) in a bump-allocated workspace; a dry run at creation sizes it.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/ysp.h"
#include "kernels_train.h"

namespace ysp { void set_error(const char* msg); }

using namespace ysp;

namespace {

struct TInfo { std::string name; int kind; int64_t off, numel; };   // kind 0: parameter, 1: BN running statistic

struct ConvBN {               // ultralytics Conv / DWConv: conv(bias=False) -> BatchNorm2d -> SiLU | identity
  int dw = 0, k = 1, Cin = 0, Cout = 0, act = 0;
  int64_t w = 0, g = 0, b = 0, rm = 0, rv = 0;
  // saved by the forward for the backward
  const float* x = nullptr; int ldx = 0; float* z = nullptr; float *mean = nullptr, *invstd = nullptr;
  int N = 0, H = 0, W = 0;
  const ConvBN* src = nullptr;   // non-null: x is src->z and src's BN + activation is applied as x is loaded (InTf)
};
struct Lin { int Cin = 0, Cout = 0; int64_t w = 0, b = 0; };   // nn.Conv2d(k=1, bias=True)

struct Ghost {                // C3Ghost(n=1) + ECA
  int Cin, Cout;
  ConvBN cv1, cv2, cv3, g1, g2, h1, h2;
  int64_t eca_w;
  float *a, *cat, *gc1, *hb, *c, *mean, *gate;
};
struct Dlc {                  // Upsample(bilinear x2) + DoubleLightConv
  int Cin, C;
  ConvBN p, q, p2, q2;
  Lin r;
  const float* xl; int ldxl, h, w;      // the low-resolution input (read again by the weight gradients)
};

struct Ctx {
  bool dry = true;
  char* ws = nullptr;
  size_t top = 0, sums_top = 0, dz_max = 0;
  size_t sums_bytes = 0;       // fixed after the dry run
  cudaStream_t s = 0;
  const float* P = nullptr; float* G = nullptr; float* S = nullptr;
  float momentum = 0.1f;
  float* dz = nullptr;
  int launches = 0;
  double bytes = 0;            // ALGORITHMIC HBM bytes of the launched kernels (compulsory reads + writes, fp32)
  void acct(double floats) { bytes += 4.0 * floats; }
  float* alloc(size_t nfloats) {
    size_t o = (top + 255) & ~(size_t)255;
    top = o + nfloats * 4;
    return dry ? nullptr : reinterpret_cast<float*>(ws + o);
  }
  double* sums(size_t n) {     // zero-initialised double accumulators (one memset per step covers the whole arena)
    size_t o = sums_top;
    sums_top += n * 8;
    return dry ? nullptr : reinterpret_cast<double*>(ws + o);
  }
  void need_dz(size_t nfloats) { if (nfloats > dz_max) dz_max = nfloats; }
};

}  // namespace

struct ysp_trainer {
  int device = 0, B = 0, H = 0, W = 0;
  std::vector<TInfo> tensors;
  int64_t n_params = 0, n_stats = 0;
  Ghost gh0, gh2;
  Dlc dl1, dl3, dl4;
  Lin out;
  size_t ws_bytes = 0, sums_bytes = 0, dz_floats = 0;
  int last_launches = 0;
  double last_bytes = 0;
};

namespace {

int tfail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  ysp::set_error(buf);
  return code;
}

// ---- parameter registry (names = the reference module's state_dict keys) --------------------------------------------
int64_t add_param(ysp_trainer* t, const std::string& name, int64_t n) {
  int64_t off = t->n_params;
  t->tensors.push_back({name, 0, off, n});
  t->n_params += (n + 3) & ~(int64_t)3;          // keep every tensor 16-byte aligned inside the flat buffers
  return off;
}
int64_t add_stat(ysp_trainer* t, const std::string& name, int64_t n) {
  int64_t off = t->n_stats;
  t->tensors.push_back({name, 1, off, n});
  t->n_stats += (n + 3) & ~(int64_t)3;
  return off;
}
ConvBN make_convbn(ysp_trainer* t, const std::string& pre, int Cin, int Cout, int k, int dw, int act) {
  ConvBN u;
  u.dw = dw; u.k = k; u.Cin = Cin; u.Cout = Cout; u.act = act;
  u.w = add_param(t, pre + ".conv.weight", dw ? (int64_t)Cout * k * k : (int64_t)Cout * Cin);
  u.g = add_param(t, pre + ".bn.weight", Cout);
  u.b = add_param(t, pre + ".bn.bias", Cout);
  u.rm = add_stat(t, pre + ".bn.running_mean", Cout);
  u.rv = add_stat(t, pre + ".bn.running_var", Cout);
  return u;
}
Ghost make_ghost(ysp_trainer* t, const std::string& pre, int Cin, int Cout) {
  Ghost g = {};
  g.Cin = Cin; g.Cout = Cout;
  const int c_ = Cout / 2, gg = c_ / 4;
  g.cv1 = make_convbn(t, pre + ".0.cv1", Cin, c_, 1, 0, 1);
  g.cv2 = make_convbn(t, pre + ".0.cv2", Cin, c_, 1, 0, 1);
  g.cv3 = make_convbn(t, pre + ".0.cv3", 2 * c_, Cout, 1, 0, 1);
  g.g1 = make_convbn(t, pre + ".0.m.0.conv.0.cv1", c_, gg, 1, 0, 1);
  g.g2 = make_convbn(t, pre + ".0.m.0.conv.0.cv2", gg, gg, 5, 1, 1);
  g.h1 = make_convbn(t, pre + ".0.m.0.conv.2.cv1", 2 * gg, c_ / 2, 1, 0, 0);
  g.h2 = make_convbn(t, pre + ".0.m.0.conv.2.cv2", c_ / 2, c_ / 2, 5, 1, 0);
  g.eca_w = add_param(t, pre + ".1.conv.weight", 3);
  return g;
}
Dlc make_dlc(ysp_trainer* t, const std::string& pre, int Cin, int C) {
  Dlc d = {};
  d.Cin = Cin; d.C = C;
  d.p = make_convbn(t, pre + ".1.conv.0.conv1", Cin, C, 1, 0, 0);
  d.q = make_convbn(t, pre + ".1.conv.0.conv2", C, C, 3, 1, 1);
  d.p2 = make_convbn(t, pre + ".1.conv.1.conv1", C, C, 1, 0, 0);
  d.q2 = make_convbn(t, pre + ".1.conv.1.conv2", C, C, 3, 1, 1);
  d.r.Cin = Cin; d.r.Cout = C;
  d.r.w = add_param(t, pre + ".1.residual_conv.weight", (int64_t)C * Cin);
  d.r.b = add_param(t, pre + ".1.residual_conv.bias", C);
  return d;
}

// ---- unit forward / backward -------------------------------------------------------------------------------------------
constexpr float kBnEps = 1e-5f;    // decoder BNs keep PyTorch's default eps

InTf tf_of(const Ctx& c, const ConvBN* src) {
  InTf t;
  if (src) { t.gamma = c.P + src->g; t.beta = c.P + src->b; t.mean = src->mean; t.invstd = src->invstd; t.act = src->act; }
  return t;
}

// y == nullptr: "lazy" unit -- only z and the batch statistics are produced; the consumer passes this unit as `src` and
// applies BN + activation while loading (saves writing and re-reading the normalised tensor).
void convbn_fwd(Ctx& c, ConvBN& u, const float* x, int ldx, int N, int H, int W, float* y, int ldy, const float* res,
                int ldr, const ConvBN* src = nullptr, const Lin* rider = nullptr, float* rider_out = nullptr) {
  const long long M = (long long)N * H * W;
  u.x = src ? src->z : x; u.ldx = src ? src->Cout : ldx; u.N = N; u.H = H; u.W = W; u.src = src;
  u.z = c.alloc((size_t)M * u.Cout);
  u.mean = c.alloc(u.Cout);
  u.invstd = c.alloc(u.Cout);
  double* sums = c.sums(2 * (size_t)u.Cout);
  c.need_dz((size_t)M * u.Cout);
  if (c.dry) return;
  BnRef bn = {c.P + u.g, c.P + u.b, u.mean, u.invstd};
  const InTf tf = tf_of(c, src);
  bool stats_done = false;
  if (u.dw) {
    stats_done = launch_dw_fwd_stats(u.x, u.ldx, c.P + u.w, u.z, u.Cout, sums, N, H, W, u.Cout, u.k, c.s, tf);   // conv + BN statistics
    if (!stats_done) launch_dw_conv(u.x, u.ldx, c.P + u.w, u.z, u.Cout, N, H, W, u.Cout, u.k, 0, 0, c.s);
  } else {
    PwDual du;
    if (rider) {     // a biased 1x1 conv on the same input (DoubleLightConv.residual_conv) rides in the same GEMM
      du.Jsplit = u.Cout; du.Wb = c.P + rider->w; du.ldwb = rider->Cin; du.biasb = c.P + rider->b; du.Cb = rider_out; du.ldcb = rider->Cout;
    }
    launch_pw_gemm(u.x, u.ldx, c.P + u.w, u.Cin, 0, nullptr, u.z, u.Cout, M, u.Cin, u.Cout + (rider ? rider->Cout : 0), 0, c.s, sums,
                   tf, du);                                                                                         // conv + BN statistics
    stats_done = true;
    if (rider) c.acct((double)M * rider->Cout);
  }
  if (!stats_done) { launch_col_reduce(0, u.z, u.Cout, nullptr, 0, bn, 0, sums, u.Cout, 1, M, c.s); c.launches += 1; }
  launch_bn_finalize(sums, u.Cout, M, kBnEps, c.momentum, u.mean, u.invstd, c.S ? c.S + u.rm : nullptr,
                     c.S ? c.S + u.rv : nullptr, c.s);
  c.launches += 2;
  c.acct((double)M * ((u.dw ? u.Cout : u.Cin) + u.Cout));                          // conv read / write
  if (y) {
    launch_bn_apply(u.z, u.Cout, bn, u.act, res, ldr, y, ldy, u.Cout, M, c.s);
    c.launches += 1;
    c.acct((double)M * u.Cout * (2 + (res ? 1 : 0)));
  }
}

// dy -> parameter gradients (+ input gradient into dx[:, :dx_ch], accumulated when beta).
// ready: this unit's BatchNorm-backward sums were already produced (by the fused backward of its consumer);
// psums: where the fused depthwise backward leaves the sums of the unit that produced ITS input (u.src).
void convbn_bwd(Ctx& c, ConvBN& u, const float* dy, int ldd, float* dx, int lddx, int beta, int dx_ch,
                const Lin* rider = nullptr, const float* rider_dy = nullptr, int rider_ldd = 0, double* ready = nullptr,
                double** psums = nullptr) {
  const long long M = (long long)u.N * u.H * u.W;
  double* sums = ready ? ready : c.sums(2 * (size_t)u.Cout);
  double* ps = (psums && u.src) ? c.sums(2 * (size_t)u.src->Cout) : nullptr;
  if (psums) *psums = nullptr;
  if (c.dry) return;
  BnRef bn = {c.P + u.g, c.P + u.b, u.mean, u.invstd};
  if (!ready) { launch_col_reduce(1, dy, ldd, u.z, u.Cout, bn, u.act, sums, u.Cout, 1, M, c.s); c.launches += 1; c.acct((double)M * u.Cout * 2); }
  if (u.dw && dx && !beta && dx_ch == u.Cout &&
      launch_dw_bwd_fused(dy, ldd, u.z, u.Cout, bn, u.act, sums, u.x, u.ldx, tf_of(c, u.src), c.P + u.w, dx, lddx, c.G + u.w,
                          c.G + u.g, c.G + u.b, ps, u.N, u.H, u.W, u.Cout, u.k, c.s)) {
    if (psums) *psums = ps;
    c.launches += 1;
    c.acct((double)M * u.Cout * 4);          // dy, z, x in; dx out
    return;
  }
  launch_bn_bwd_apply(dy, ldd, u.z, u.Cout, bn, u.act, sums, c.dz, u.Cout, c.G + u.g, c.G + u.b, u.Cout, M, c.s);
  if (u.dw) {
    launch_dw_wgrad(c.dz, u.Cout, u.x, u.ldx, c.G + u.w, u.N, u.H, u.W, u.Cout, u.k, c.s, tf_of(c, u.src));
    if (dx) launch_dw_conv(c.dz, u.Cout, c.P + u.w, dx, lddx, u.N, u.H, u.W, u.Cout, u.k, 1, beta, c.s);
  } else {
    launch_pw_wgrad(c.dz, u.Cout, u.x, u.ldx, c.G + u.w, u.Cin, M, u.Cin, u.Cout, c.s, tf_of(c, u.src));
    if (dx) {
      PwDual du;
      if (rider) { du.A2 = rider_dy; du.lda2 = rider_ldd; du.W2 = c.P + rider->w; du.ldw2 = rider->Cin; du.I2 = rider->Cout; }
      launch_pw_gemm(c.dz, u.Cout, c.P + u.w, u.Cin, 1, nullptr, dx, lddx, M, u.Cout, dx_ch, beta, c.s, nullptr, InTf(), du);
      if (rider) c.acct((double)M * rider->Cout);
    }
  }
  c.launches += dx ? 3 : 2;
  // apply (dy, z -> dz), wgrad (dz, x), dgrad (dz -> dx [+ dx])
  c.acct((double)M * (u.Cout * 5 + (u.dw ? u.Cout : u.Cin) + (dx ? dx_ch * (1 + beta) : 0)));
}

// ---- C3Ghost + ECA -------------------------------------------------------------------------------------------------------
void ghost_fwd(Ctx& c, Ghost& g, const float* xin, int ldin, int N, int H, int W, float* out, int ldo) {
  const long long M = (long long)N * H * W, HW = (long long)H * W;
  const int c_ = g.Cout / 2, gg = c_ / 4, hh = c_ / 2;
  g.a = c.alloc((size_t)M * c_);
  g.cat = c.alloc((size_t)M * 2 * c_);
  g.gc1 = c.alloc((size_t)M * 2 * gg);
  g.hb = c.alloc((size_t)M * c_);
  g.c = c.alloc((size_t)M * g.Cout);
  g.mean = c.alloc((size_t)N * g.Cout);
  g.gate = c.alloc((size_t)N * g.Cout);
  convbn_fwd(c, g.cv1, xin, ldin, N, H, W, g.a, c_, nullptr, 0);                       // C3.cv1
  convbn_fwd(c, g.cv2, xin, ldin, N, H, W, g.cat + c_, 2 * c_, nullptr, 0);            // C3.cv2 -> second half of the cat
  convbn_fwd(c, g.g1, g.a, c_, N, H, W, g.gc1, 2 * gg, nullptr, 0);                    // GhostConv#1.cv1
  convbn_fwd(c, g.g2, g.gc1, 2 * gg, N, H, W, g.gc1 + gg, 2 * gg, nullptr, 0);         // GhostConv#1.cv2 (dw5)
  convbn_fwd(c, g.h1, g.gc1, 2 * gg, N, H, W, g.hb, c_, nullptr, 0);                   // GhostConv#2.cv1 (linear)
  convbn_fwd(c, g.h2, g.hb, c_, N, H, W, g.hb + hh, c_, nullptr, 0);                   // GhostConv#2.cv2 (dw5, linear)
  double* pool = c.sums((size_t)N * 2 * g.Cout);
  if (!c.dry) launch_add_copy(g.hb, c_, g.a, c_, g.cat, 2 * c_, c_, M, c.s);           // GhostBottleneck: conv(x) + x
  convbn_fwd(c, g.cv3, g.cat, 2 * c_, N, H, W, g.c, g.Cout, nullptr, 0);               // C3.cv3
  if (c.dry) return;
  BnRef none = {};
  launch_col_reduce(2, g.c, g.Cout, nullptr, 0, none, 0, pool, g.Cout, N, HW, c.s);    // ECA: global average pool
  launch_eca_gate(pool, c.P + g.eca_w, g.mean, g.gate, N, g.Cout, HW, c.s);
  launch_scale_rows(g.c, g.Cout, g.gate, nullptr, out, ldo, g.Cout, M, HW, c.s);
  c.launches += 4;
  c.acct((double)M * (3 * c_ + 3 * g.Cout));
}

void ghost_bwd(Ctx& c, Ghost& g, const float* dout, int ldd, float* dxin, int lddx, int dx_ch) {
  const int N = g.cv1.N, H = g.cv1.H, W = g.cv1.W;
  const long long M = (long long)N * H * W, HW = (long long)H * W;
  const int c_ = g.Cout / 2, gg = c_ / 4, hh = c_ / 2;
  float* dC = c.alloc((size_t)M * g.Cout);
  float* dCat = c.alloc((size_t)M * 2 * c_);
  float* dA = c.alloc((size_t)M * c_);
  float* dG = c.alloc((size_t)M * 2 * gg);
  float* dmean = c.alloc((size_t)N * g.Cout);
  double* dsum = c.sums((size_t)N * 2 * g.Cout);
  if (!c.dry) {
    BnRef none = {};
    launch_col_reduce(3, dout, ldd, g.c, g.Cout, none, 0, dsum, g.Cout, N, HW, c.s);   // d gate = sum_px dy * x
    launch_eca_gate_bwd(dsum, c.P + g.eca_w, g.mean, g.gate, dmean, c.G + g.eca_w, N, g.Cout, HW, c.s);
    launch_scale_rows(dout, ldd, g.gate, dmean, dC, g.Cout, g.Cout, M, HW, c.s);       // dx = dy*gate + dmean/HW
    c.launches += 3;
    c.acct((double)M * (4 * g.Cout + 2 * c_));
  }
  convbn_bwd(c, g.cv3, dC, g.Cout, dCat, 2 * c_, 0, 2 * c_);
  convbn_bwd(c, g.cv2, dCat + c_, 2 * c_, dxin, lddx, 0, dx_ch);
  if (!c.dry) { launch_add_copy(dCat, 2 * c_, nullptr, 0, dA, c_, c_, M, c.s); c.launches += 1; }   // shortcut branch
  convbn_bwd(c, g.h2, dCat + hh, 2 * c_, dCat, 2 * c_, 1, hh);
  convbn_bwd(c, g.h1, dCat, 2 * c_, dG, 2 * gg, 0, 2 * gg);
  convbn_bwd(c, g.g2, dG + gg, 2 * gg, dG, 2 * gg, 1, gg);
  convbn_bwd(c, g.g1, dG, 2 * gg, dA, c_, 1, c_);
  convbn_bwd(c, g.cv1, dA, c_, dxin, lddx, 1, dx_ch);
}

// ---- Upsample + DoubleLightConv ---------------------------------------------------------------------------------------------
// Both 1x1 convs that read the upsampled tensor (conv.0.conv1 and residual_conv) run at LOW resolution: bilinear x2 is linear
// with unit weight sums, so  conv1x1(up2(x)) + b == up2(conv1x1(x) + b)  (the identity the inference decoder uses,
// DESIGN §4).  The upsampled Cin-channel tensor is never formed: one GEMM over a quarter of the rows writes
// plow = [conv1 | residual_conv + bias], two up2 passes write z_p and r at full resolution, and the backward takes the
// adjoint route -- up2^T of the two C-channel gradients, then weight / bias / input gradients from GEMMs over the low-res
// rows.  Per full-resolution row this moves 2C + C (+ C for the statistics) floats forward and 2C + 2C backward where the
// direct form moved 4 Cin + 2C and 6 Cin + 5C (Cin = 2C).
void dlc_fwd(Ctx& c, Dlc& d, const float* xl, int ldx, int N, int h, int w, float* out, int ldo) {
  const int H = 2 * h, W = 2 * w;
  const long long M = (long long)N * H * W, Ml = (long long)N * h * w;
  d.xl = xl; d.ldxl = ldx; d.h = h; d.w = w;
  float* plow = c.alloc((size_t)Ml * 2 * d.C);
  float* r = c.alloc((size_t)M * d.C);
  const bool lazy = dw_tiled_shape(H, W, 3);
  static const bool no_dual = getenv("YSP_TRAIN_NO_DUAL") != nullptr;      // A/B switch: the two-output GEMM off
  ConvBN& u = d.p;
  u.x = xl; u.ldx = ldx; u.N = N; u.H = H; u.W = W; u.src = nullptr;
  u.z = c.alloc((size_t)M * d.C);
  u.mean = c.alloc(d.C);
  u.invstd = c.alloc(d.C);
  double* sums = c.sums(2 * (size_t)d.C);
  c.need_dz((size_t)M * d.C);
  float* pb = lazy ? nullptr : c.alloc((size_t)M * d.C);
  if (!c.dry) {
    if (!no_dual) {
      PwDual du;
      du.Jsplit = d.C; du.Wb = c.P + d.r.w; du.ldwb = d.Cin; du.biasb = c.P + d.r.b; du.Cb = plow + d.C; du.ldcb = 2 * d.C;
      launch_pw_gemm(xl, ldx, c.P + u.w, d.Cin, 0, nullptr, plow, 2 * d.C, Ml, d.Cin, 2 * d.C, 0, c.s, nullptr, InTf(), du);
    } else {
      launch_pw_gemm(xl, ldx, c.P + u.w, d.Cin, 0, nullptr, plow, 2 * d.C, Ml, d.Cin, d.C, 0, c.s);
      launch_pw_gemm(xl, ldx, c.P + d.r.w, d.Cin, 0, c.P + d.r.b, plow + d.C, 2 * d.C, Ml, d.Cin, d.C, 0, c.s);
      c.launches += 1;
    }
    BnRef bn = {c.P + u.g, c.P + u.b, u.mean, u.invstd};
    launch_up2_split(plow, 2 * d.C, u.z, d.C, d.C, r, d.C, d.C, N, h, w, sums, c.s);      // z_p (+ its BN statistics) and r
    launch_bn_finalize(sums, d.C, M, kBnEps, c.momentum, u.mean, u.invstd, c.S ? c.S + u.rm : nullptr,
                       c.S ? c.S + u.rv : nullptr, c.s);
    c.launches += 3;
    c.acct((double)Ml * (d.Cin + 2 * d.C) + (double)Ml * 2 * d.C + (double)M * 2 * d.C);
    if (!lazy) {
      launch_bn_apply(u.z, d.C, bn, 0, nullptr, 0, pb, d.C, d.C, M, c.s);
      c.launches += 1;
      c.acct((double)M * d.C * 2);
    }
  }
  // the inner normalised tensors are never written: each consumer applies its producer's BN (+SiLU) on load
  if (lazy) {
    convbn_fwd(c, d.q, nullptr, 0, N, H, W, nullptr, 0, nullptr, 0, &d.p);
    convbn_fwd(c, d.p2, nullptr, 0, N, H, W, nullptr, 0, nullptr, 0, &d.q);
    convbn_fwd(c, d.q2, nullptr, 0, N, H, W, out, ldo, r, d.C, &d.p2);                               // out = conv(x) + residual
  } else {
    float* qb = c.alloc((size_t)M * d.C);
    float* p2b = c.alloc((size_t)M * d.C);
    convbn_fwd(c, d.q, pb, d.C, N, H, W, qb, d.C, nullptr, 0);
    convbn_fwd(c, d.p2, qb, d.C, N, H, W, p2b, d.C, nullptr, 0);
    convbn_fwd(c, d.q2, p2b, d.C, N, H, W, out, ldo, r, d.C);                                        // out = conv(x) + residual
  }
}

void dlc_bwd(Ctx& c, Dlc& d, const float* dout, int ldd, float* dxl, int lddx) {
  const int N = d.p.N, H = d.p.H, W = d.p.W, h = d.h, w = d.w;
  const long long M = (long long)N * H * W, Ml = (long long)N * h * w;
  float* d1 = c.alloc((size_t)M * d.C);
  float* d2 = c.alloc((size_t)M * d.C);
  float* dpl = c.alloc((size_t)Ml * 2 * d.C);        // gradient of plow: [d conv1 | d residual_conv]
  double* bs = c.sums(2 * (size_t)d.C);
  double* ps = c.sums(2 * (size_t)d.C);
  // the depthwise units run fused (dw_bwd_fused_kernel) and leave the BatchNorm-backward sums of their producers behind
  double *s_p2 = nullptr, *s_p = nullptr;
  convbn_bwd(c, d.q2, dout, ldd, d1, d.C, 0, d.C, nullptr, nullptr, 0, nullptr, &s_p2);
  convbn_bwd(c, d.p2, d1, d.C, d2, d.C, 0, d.C, nullptr, nullptr, 0, s_p2);
  convbn_bwd(c, d.q, d2, d.C, d1, d.C, 0, d.C, nullptr, nullptr, 0, nullptr, &s_p);
  if (c.dry) return;
  // conv.0.conv1: BatchNorm backward at full resolution, everything after it on the low-resolution rows
  ConvBN& u = d.p;
  BnRef bn = {c.P + u.g, c.P + u.b, u.mean, u.invstd}, none = {};
  if (s_p) ps = s_p;
  else launch_col_reduce(1, d1, d.C, u.z, d.C, bn, 0, ps, d.C, 1, M, c.s);
  // up2^T of [BN-backward(d1, z_p) | dout] from shared-memory tiles: dz of conv.0.conv1 is formed as it is staged, never written.
  // (The same fusion inside the GATHER form of up2^T measured 0.8 ms slower per step: it reads every full-resolution element
  // four times, on two tensors instead of one.)
  if (!launch_up2_bwd_tiled(d1, d.C, u.z, d.C, bn, ps, M, c.G + u.g, c.G + u.b, d.C, dout, ldd, d.C, dpl, 2 * d.C, N, h, w, c.s)) {
    launch_bn_bwd_apply(d1, d.C, u.z, d.C, bn, 0, ps, c.dz, d.C, c.G + u.g, c.G + u.b, d.C, M, c.s);
    launch_up2_bwd(c.dz, d.C, dpl, 2 * d.C, N, h, w, d.C, c.s);
    launch_up2_bwd(dout, ldd, dpl + d.C, 2 * d.C, N, h, w, d.C, c.s);
    c.launches += 2;
    c.acct((double)M * d.C * 2);
  }
  launch_col_reduce(2, dpl + d.C, 2 * d.C, nullptr, 0, none, 0, bs, d.C, 1, Ml, c.s);     // d bias: up2^T preserves column sums
  launch_add_sums(bs, c.G + d.r.b, d.C, 1, c.s);
  launch_pw_wgrad(dpl, 2 * d.C, d.xl, d.ldxl, c.G + u.w, d.Cin, Ml, d.Cin, d.C, c.s);
  launch_pw_wgrad(dpl + d.C, 2 * d.C, d.xl, d.ldxl, c.G + d.r.w, d.Cin, Ml, d.Cin, d.C, c.s);
  // dxl = d conv1 * W_p + d residual * W_r: one two-operand GEMM
  PwDual du;
  du.A2 = dpl + d.C; du.lda2 = 2 * d.C; du.W2 = c.P + d.r.w; du.ldw2 = d.Cin; du.I2 = d.C;
  launch_pw_gemm(dpl, 2 * d.C, c.P + u.w, d.Cin, 1, nullptr, dxl, lddx, Ml, d.C, d.Cin, 0, c.s, nullptr, InTf(), du);
  c.launches += 6 + (s_p ? 0 : 1);
  c.acct((double)M * d.C * ((s_p ? 0 : 2) + 3) + (double)Ml * (2 * d.C + d.C + 2 * (d.C + d.Cin) + 2 * d.C + d.Cin));
}

// ---- whole step -----------------------------------------------------------------------------------------------------------
struct StepIO {
  const float *skipA, *skipB, *logits, *target;
  float* loss3; float* mask_logits;
  int loss_kind; float grad_scale;
};

void run_step(ysp_trainer* t, Ctx& c, const StepIO& io) {
  const int B = t->B, H = t->H, W = t->W, h8 = H / 8, w8 = W / 8, h4 = H / 4, w4 = W / 4;
  const long long M0 = (long long)B * h8 * w8, M1 = (long long)B * h4 * w4, M3 = (long long)B * H * W;
  // arenas: [double accumulators][dz scratch][bump allocations]
  c.sums_top = 0;
  c.top = c.sums_bytes;
  c.dz = c.alloc(t->dz_floats);
  if (!c.dry) {
    cudaMemsetAsync(c.ws, 0, c.sums_bytes, c.s);
    cudaMemsetAsync(c.G, 0, (size_t)t->n_params * 4, c.s);
  }
  // ---- forward (YOLOSegPlusPlus.py:261-272) ----
  float* in0 = c.alloc((size_t)M0 * 132);        // cat([skipB 128, logits 1]) (:266), row stride padded to 132
  float* d0 = c.alloc((size_t)M0 * 96);
  float* in2 = c.alloc((size_t)M1 * 128);        // cat([decoder.1 out 64, skipA 64]) (:269)
  float* d2 = c.alloc((size_t)M1 * 64);
  float* d3 = c.alloc((size_t)M1 * 4 * 32);
  float* d4 = c.alloc((size_t)M3 * 16);
  float* lg = io.mask_logits ? io.mask_logits : c.alloc((size_t)M3);
  float* dlg = c.alloc((size_t)M3);
  if (!c.dry) {
    launch_add_copy(io.skipB, 128, nullptr, 0, in0, 132, 128, M0, c.s);
    launch_add_copy(io.logits, 1, nullptr, 0, in0 + 128, 132, 1, M0, c.s);
    launch_add_copy(io.skipA, 64, nullptr, 0, in2 + 64, 128, 64, M1, c.s);
    c.launches += 3;
    c.acct((double)M0 * 258 + (double)M1 * 128);
  }
  ghost_fwd(c, t->gh0, in0, 132, B, h8, w8, d0, 96);
  dlc_fwd(c, t->dl1, d0, 96, B, h8, w8, in2, 128);
  ghost_fwd(c, t->gh2, in2, 128, B, h4, w4, d2, 64);
  dlc_fwd(c, t->dl3, d2, 64, B, h4, w4, d3, 32);
  dlc_fwd(c, t->dl4, d3, 32, B, H / 2, W / 2, d4, 16);
  double* lacc = c.sums(4);
  double* obs = c.sums(4);
  // ---- backward ----
  float* dD4 = c.alloc((size_t)M3 * 16);
  float* dD3 = c.alloc((size_t)M1 * 4 * 32);
  float* dD2 = c.alloc((size_t)M1 * 64);
  float* dIn2 = c.alloc((size_t)M1 * 64);
  float* dD0 = c.alloc((size_t)M0 * 96);
  if (!c.dry) {
    const Lin& o = t->out;
    if (!launch_lin1_fwd(d4, 16, c.P + o.w, c.P + o.b, lg, 16, M3, c.s))                           // self.output (:271)
      launch_pw_gemm(d4, 16, c.P + o.w, 16, 0, c.P + o.b, lg, 1, M3, 16, 1, 0, c.s);
    launch_loss(lg, io.target, M3, lacc, io.loss_kind, io.grad_scale, dlg, io.loss3, c.s);
    BnRef none = {};
    launch_col_reduce(2, dlg, 4, nullptr, 0, none, 0, obs, 4, 1, M3 / 4, c.s);                     // d bias = sum dlogits
    launch_add_sums(obs, c.G + o.b, 1, 4, c.s);
    launch_pw_wgrad(dlg, 1, d4, 16, c.G + o.w, 16, M3, 16, 1, c.s);
    if (!launch_lin1_dgrad(dlg, c.P + o.w, dD4, 16, 16, M3, c.s))
      launch_pw_gemm(dlg, 1, c.P + o.w, 16, 1, nullptr, dD4, 16, M3, 1, 16, 0, c.s);
    c.launches += 7;
    c.acct((double)M3 * (17 + 2 + 3 + 1 + 17 + 17));
  }
  dlc_bwd(c, t->dl4, dD4, 16, dD3, 32);
  dlc_bwd(c, t->dl3, dD3, 32, dD2, 64);
  ghost_bwd(c, t->gh2, dD2, 64, dIn2, 64, 64);          // only the decoder.1 half of the concat needs a gradient
  dlc_bwd(c, t->dl1, dIn2, 64, dD0, 96);
  ghost_bwd(c, t->gh0, dD0, 96, nullptr, 0, 0);         // encoder is frozen (train.py:256-260): no input gradient
}

}  // namespace

extern "C" {

int ysp_train_create(ysp_trainer** out, int device, int B, int H, int W) {
  if (!out) return tfail(YSP_EINVAL, "ysp_train_create: null out");
  if (B <= 0 || H <= 0 || W <= 0 || H % 8 || W % 8) return tfail(YSP_EINVAL, "ysp_train_create: B>0 and H, W positive multiples of 8 required (got %d, %dx%d)", B, H, W);
  ysp_trainer* t = new ysp_trainer();
  t->device = device; t->B = B; t->H = H; t->W = W;
  t->gh0 = make_ghost(t, "decoder.0", 129, 96);
  t->dl1 = make_dlc(t, "decoder.1", 96, 64);
  t->gh2 = make_ghost(t, "decoder.2", 128, 64);
  t->dl3 = make_dlc(t, "decoder.3", 64, 32);
  t->dl4 = make_dlc(t, "decoder.4", 32, 16);
  t->out.Cin = 16; t->out.Cout = 1;
  t->out.w = add_param(t, "output.weight", 16);
  t->out.b = add_param(t, "output.bias", 1);
  // dry runs: first sizes the accumulator arena and the dz scratch, second the bump arena behind them
  Ctx c; c.dry = true;
  StepIO io = {};
  run_step(t, c, io);
  t->sums_bytes = (c.sums_top + 255) & ~(size_t)255;
  t->dz_floats = c.dz_max;
  Ctx c2; c2.dry = true; c2.sums_bytes = t->sums_bytes;
  run_step(t, c2, io);
  t->ws_bytes = c2.top + 256;
  *out = t;
  return 0;
}

void ysp_train_destroy(ysp_trainer* t) { delete t; }

int ysp_train_num_tensors(const ysp_trainer* t) { return t ? (int)t->tensors.size() : 0; }

int ysp_train_tensor_info(const ysp_trainer* t, int i, char* name, int name_cap, int* kind, int64_t* offset, int64_t* numel) {
  if (!t || i < 0 || i >= (int)t->tensors.size() || !name || name_cap <= 0) return tfail(YSP_EINVAL, "ysp_train_tensor_info: bad arguments");
  const TInfo& ti = t->tensors[i];
  if ((int)ti.name.size() + 1 > name_cap) return tfail(YSP_EINVAL, "ysp_train_tensor_info: name buffer too small");
  memcpy(name, ti.name.c_str(), ti.name.size() + 1);
  if (kind) *kind = ti.kind;
  if (offset) *offset = ti.off;
  if (numel) *numel = ti.numel;
  return 0;
}

int64_t ysp_train_param_count(const ysp_trainer* t) { return t ? t->n_params : 0; }
int64_t ysp_train_stat_count(const ysp_trainer* t) { return t ? t->n_stats : 0; }
size_t ysp_train_workspace_bytes(const ysp_trainer* t) { return t ? t->ws_bytes : 0; }
int ysp_train_last_launch_count(const ysp_trainer* t) { return t ? t->last_launches : 0; }
double ysp_train_last_step_bytes(const ysp_trainer* t) { return t ? t->last_bytes : 0; }

int ysp_train_step(ysp_trainer* t, const float* d_skipA, const float* d_skipB, const float* d_logits,
                   const float* d_target, const float* d_params, float* d_grads, float* d_stats, float momentum,
                   int loss_kind, float grad_scale, float* d_loss3, float* d_mask_logits, void* d_ws, size_t ws_bytes,
                   void* stream) {
  if (!t) return tfail(YSP_EINVAL, "ysp_train_step: null trainer");
  if (!d_skipA || !d_skipB || !d_logits || !d_target || !d_params || !d_grads || !d_loss3 || !d_ws)
    return tfail(YSP_EINVAL, "ysp_train_step: null pointer");
  if (loss_kind < 0 || loss_kind > 1) return tfail(YSP_EINVAL, "ysp_train_step: loss_kind must be 0 (Dice) or 1 (Dice+BCE)");
  if (ws_bytes < t->ws_bytes) return tfail(YSP_ESTATE, "workspace too small: need %zu bytes, got %zu", t->ws_bytes, ws_bytes);
  int dev = -1;
  cudaGetDevice(&dev);
  if (dev != t->device) return tfail(YSP_ESTATE, "trainer was created for device %d but device %d is current", t->device, dev);
  Ctx c; c.dry = false;
  c.ws = (char*)d_ws; c.s = (cudaStream_t)stream; c.sums_bytes = t->sums_bytes;
  c.P = d_params; c.G = d_grads; c.S = d_stats; c.momentum = momentum;
  StepIO io = {d_skipA, d_skipB, d_logits, d_target, d_loss3, d_mask_logits, loss_kind, grad_scale};
  run_step(t, c, io);
  t->last_launches = c.launches;
  t->last_bytes = c.bytes;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return tfail(YSP_ECUDA, "ysp_train_step: %s", cudaGetErrorString(e));
  return 0;
}

int ysp_seg_loss(const float* d_logits, const float* d_target, int64_t n, int loss_kind, float* d_loss3, void* d_ws32,
                 void* stream) {
  if (!d_logits || !d_target || !d_loss3 || !d_ws32 || n <= 0) return tfail(YSP_EINVAL, "ysp_seg_loss: bad arguments");
  if (loss_kind < 0 || loss_kind > 1) return tfail(YSP_EINVAL, "ysp_seg_loss: loss_kind must be 0 (Dice) or 1 (Dice+BCE)");
  launch_loss_value(d_logits, d_target, (long long)n, (double*)d_ws32, loss_kind, d_loss3, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return tfail(YSP_ECUDA, "ysp_seg_loss: %s", cudaGetErrorString(e));
  return 0;
}

int ysp_grad_sqnorm(const float* d_grads, int64_t n, void* d_out8, void* stream) {
  if (!d_grads || !d_out8 || n <= 0) return tfail(YSP_EINVAL, "ysp_grad_sqnorm: bad arguments");
  launch_sqnorm(d_grads, (long long)n, (double*)d_out8, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return tfail(YSP_ECUDA, "ysp_grad_sqnorm: %s", cudaGetErrorString(e));
  return 0;
}

int ysp_adamw(float* d_params, const float* d_grads, float* d_m, float* d_v, int64_t n, float lr, float beta1, float beta2,
              float eps, float weight_decay, int step, float grad_scale, float max_norm, void* d_ws8, void* stream) {
  if (!d_params || !d_grads || !d_m || !d_v || n < 0 || step < 1) return tfail(YSP_EINVAL, "ysp_adamw: bad arguments");
  if (max_norm > 0.f && !d_ws8) return tfail(YSP_EINVAL, "ysp_adamw: gradient clipping needs an 8-byte device scratch");
  if (n == 0) return 0;
  launch_adamw(d_params, d_grads, d_m, d_v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, max_norm,
               (double*)d_ws8, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return tfail(YSP_ECUDA, "ysp_adamw: %s", cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
