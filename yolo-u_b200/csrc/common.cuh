// common.cuh -- shared device helpers and kernel parameter blocks for libysp (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace ysp {

typedef __nv_bfloat16 bf16;

enum { ACT_NONE = 0, ACT_SILU = 1 };

// ---- scalar conversion ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---- 4-wide vectors (16 B fp32 / 8 B bf16) -------------------------------------------------------------------------
struct F4 { float v[4]; };
template <typename T> __device__ __forceinline__ F4 load4(const T* p);
template <> __device__ __forceinline__ F4 load4<float>(const float* p) {
  float4 t = *reinterpret_cast<const float4*>(p);
  F4 r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
template <> __device__ __forceinline__ F4 load4<bf16>(const bf16* p) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  F4 r; r.v[0] = __low2float(a); r.v[1] = __high2float(a); r.v[2] = __low2float(b); r.v[3] = __high2float(b); return r;
}
template <typename T> __device__ __forceinline__ void store4(T* p, const F4& r);
template <> __device__ __forceinline__ void store4<float>(float* p, const F4& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
}
template <> __device__ __forceinline__ void store4<bf16>(bf16* p, const F4& r) {
  __nv_bfloat162 a = __floats2bfloat162_rn(r.v[0], r.v[1]), b = __floats2bfloat162_rn(r.v[2], r.v[3]);
  uint2 t; t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

// SiLU(x) = x * rcp(1 + 2^(-x log2 e)) on the MUFU unit: ex2.approx (2 ulp) + rcp.approx (1 ulp) -> relative error ~3e-7, the
// size of an fp32 rounding; 5 instructions instead of the ~25 of `x / (1 + expf(-x))` (IEEE divide + range-checked expf),
// which ncu showed to be a quarter of all instructions of the parity-mode conv epilogues.  x << 0: 2^(..) = inf, rcp = 0.
__device__ __forceinline__ float silu_f(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
// MUFU sigmoid (ex2.approx + rcp.approx, ~3e-7): activation DERIVATIVES of the training step; the mask threshold, the loss and the
// detector's class scores keep the exact sigmoid_f
__device__ __forceinline__ float sigmoid_mufu(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float apply_act(float x, int act) { return act == ACT_SILU ? silu_f(x) : x; }
// SiLU(x) = h + h*tanh(h), h = x/2: one MUFU op (tanh.approx, rel. error ~2^-11 -- below bf16 resolution).  Used wherever
// the result is stored as bf16 (throughput mode); the fp32 parity mode keeps the exact expf form above.
__device__ __forceinline__ float silu_approx(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
// fp32 pair -> packed fp16 hi and lo with x = hi + lo (hi = rn16(x), lo = rn16(x - hi): 22 mantissa bits), the operand split of
// the parity-mode tensor-core kernels.  The conversion saturates (F2FP.SATFINITE) instead of clamping with two FMNMX per value.
__device__ __forceinline__ uint32_t f2h2_sat(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
__device__ __forceinline__ void split2_f16(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = f2h2_sat(a, b);
  float2 f;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(f.x), "=f"(f.y) : "r"(hi));
  const float2 d = __ffma2_rn(f, make_float2(-1.f, -1.f), make_float2(a, b));      // (a - fa, b - fb), exact, one FFMA2
  lo = f2h2_sat(d.x, d.y);
}
// fp32 pair -> packed bf16 hi and mid with x = hi + mid + O(2^-18 x): the operand split of the TRAINING GEMMs, whose operands
// (gradients of 1e-5 and below) would underflow the fp16 range; bf16 keeps the fp32 exponent, three MMAs give 2^-17 products.
__device__ __forceinline__ void split2_bf16(float a, float b, uint32_t& hi, uint32_t& mid) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float fa = __uint_as_float(hi << 16), fb = __uint_as_float(hi & 0xffff0000u);
  const float2 d = __ffma2_rn(make_float2(fa, fb), make_float2(-1.f, -1.f), make_float2(a, b));      // exact
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mid) : "f"(d.y), "f"(d.x));
}
// Packed SiLU of two values (FMUL2 / FADD2 around the four MUFU ops): bit-identical to silu_f per lane, 8 instead of 10 issue slots
__device__ __forceinline__ float2 silu2_f(float2 x) {
  const float2 t = __fmul2_rn(x, make_float2(-1.4426950408889634f, -1.4426950408889634f));
  float2 e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(t.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(t.y));
  const float2 d = __fadd2_rn(e, make_float2(1.f, 1.f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(d.y));
  return __fmul2_rn(x, r);
}

template <typename T> __device__ __forceinline__ float silu_for(float x);
template <> __device__ __forceinline__ float silu_for<float>(float x) { return silu_f(x); }
template <> __device__ __forceinline__ float silu_for<bf16>(float x) { return silu_approx(x); }
template <typename T> __device__ __forceinline__ float apply_act_for(float x, int act) { return act == ACT_SILU ? silu_for<T>(x) : x; }

// ---- parameter blocks ----------------------------------------------------------------------------------------------
// Activations are NHWC "views": base pointer already advanced to the first channel of the view, `cs` = pixel stride
// in elements (>= C), so concatenations are formed by construction (producers write into channel slices).
struct ConvP {
  const void* in; void* out; const float* w; const float* bias; const void* res;
  int N, H, W, Cin, in_cs;
  int OH, OW, Cout, out_cs, res_cs;
  int kh, kw, stride, pad, act;
  int M, K, wld;            // M = N*OH*OW, K = kh*kw*Cin, weights [K][wld] fp32
  int cout_store;           // >= Cout: channels [Cout, cout_store) are written as act(0) = 0 (zero channel padding)
  int in_zpad;              // input view has zero-filled channels up to a multiple of 16 (tensor-core K padding)
  int in_pw, in_ph;         // memory pitch of the input view in pixels (0 = dense H x W); tensor-core path only
  const float* in_scale;    // optional per-(slice, input channel) factor [N][Cin] applied to the input on load (the ECA gate);
                            // conv_tc32 flat 1x1 path only
};

struct DwP {
  const void* in; void* out; const float* w; const float* bias; const void* res;
  int N, H, W, C, in_cs, out_cs, res_cs;
  int k, pad, act;
  int grp, grp_stride;      // input channel c lives at (c / grp) * grp_stride + c % grp of the input view
};

struct EwP {                // generic elementwise / resampling ops over NHWC views
  const void* a; const void* b; void* out;
  int N, H, W, C, a_cs, b_cs, out_cs;
  int OH, OW;
};

// ---- programmatic dependent launch --------------------------------------------------------------------------------------
// Every kernel of the layer chain is launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts with
// pdl_sync(): it lets ITS successor begin launching right away and then waits until its predecessor has completed and
// flushed.  Launch latency and CTA ramp-up of kernel N+1 overlap the tail of kernel N; ordering is unchanged.
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE property of a kernel: remember which devices have it (a
// process normally drives one GPU, but a handle per device in one process must work too).  `done` = one static per site.
template <typename Kern>
inline void ensure_dyn_smem(Kern kernel, size_t bytes, unsigned long long& done, const char* name) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (done & bit) return;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) fprintf(stderr, "libysp: cudaFuncSetAttribute(%s, %zu): %s\n", name, bytes, cudaGetErrorString(e));
  done |= bit;
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace ysp
