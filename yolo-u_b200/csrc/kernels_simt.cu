// kernels_simt.cu -- CUDA-core kernels of the YOLO-Seg++ hot path (sm_100a).
//
// These are (a) the whole fp32 "parity mode" (YSP_MODE_FP32) and (b) everything in bf16 "throughput mode" that is
// not a tensor-core GEMM: depthwise convs, resampling, ECA, area-attention core, head decode, layout conversion,
// mask/Dice counters.  All activations are NHWC views (common.cuh).  Reference semantics cited per kernel.
#include <algorithm>
#include <cstdlib>

#include "kernels.h"

namespace ysp {

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// =====================================================================================================================
// Dense convolution as an implicit GEMM on CUDA cores:  out[m, co] = act(bias[co] + sum_k A[m,k] W[k,co]) (+ res[m,co])
//   m = (n, oy, ox), k = (r, s, ci).  Reads beyond H/W return 0, which implements both conv padding and the
//   bottom/right zero-padding of decision D1 without a padded copy.
// Replaces ultralytics Conv.forward_fuse / nn.Conv2d on the detector + seg head.
// =====================================================================================================================
template <typename TI, typename TO, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) conv_dense_kernel(ConvP p) {
  constexpr int BK = 16;
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int AROWS = NT / BK;
  constexpr int APASS = BM / AROWS;
  static_assert(NT % BK == 0 && BM % AROWS == 0, "tile");
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const TI* __restrict__ in = reinterpret_cast<const TI*>(p.in);
  const float* __restrict__ w = p.w;
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int a_k = tid % BK, a_r = tid / BK;

  int pix_n[APASS], pix_y[APASS], pix_x[APASS];
#pragma unroll
  for (int i = 0; i < APASS; ++i) {
    int m = m0 + a_r + i * AROWS;
    if (m < p.M) {
      int ox = m % p.OW, t = m / p.OW;
      int oy = t % p.OH;
      pix_n[i] = t / p.OH;
      pix_y[i] = oy * p.stride - p.pad;
      pix_x[i] = ox * p.stride - p.pad;
    } else {
      pix_n[i] = -1; pix_y[i] = 0; pix_x[i] = 0;
    }
  }
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    const int k = k0 + a_k;
    const bool kv = k < p.K;
    const int tap = kv ? k / p.Cin : 0;
    const int ci = k - tap * p.Cin;
    const int r = tap / p.kw, s = tap - r * p.kw;
#pragma unroll
    for (int i = 0; i < APASS; ++i) {
      float v = 0.f;
      if (kv && pix_n[i] >= 0) {
        int iy = pix_y[i] + r, ix = pix_x[i] + s;
        if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
          v = to_f<TI>(in[((size_t)(pix_n[i] * p.H + iy) * p.W + ix) * p.in_cs + ci]);
      }
      As[a_k][a_r + i * AROWS] = v;
    }
    for (int e = tid; e < BK * BN; e += NT) {
      int kk = e / BN, nn = e - kk * BN;
      int kg = k0 + kk, co = n0 + nn;
      Bs[kk][nn] = (kg < p.K && co < p.Cout) ? w[(size_t)kg * p.wld + co] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  TO* __restrict__ out = reinterpret_cast<TO*>(p.out);
  const TI* __restrict__ res = reinterpret_cast<const TI*>(p.res);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int co = n0 + tx * TN + j;
      if (co >= p.Cout) {
        if (co < p.cout_store) out[(size_t)m * p.out_cs + co] = from_f<TO>(0.f);   // zero channel padding
        continue;
      }
      float v = acc[i][j] + (p.bias ? p.bias[co] : 0.f);
      v = apply_act(v, p.act);
      if (res) v += to_f<TI>(res[(size_t)m * p.res_cs + co]);
      out[(size_t)m * p.out_cs + co] = from_f<TO>(v);
    }
  }
}

template <typename TI, typename TO>
static void conv_dense_dispatch(const ConvP& p, cudaStream_t s) {
  const int cs_ = p.cout_store > p.Cout ? p.cout_store : p.Cout;
  if (cs_ <= 16) {
    dim3 g(cdiv(p.M, 128), cdiv(cs_, 16));
    conv_dense_kernel<TI, TO, 128, 16, 4, 2><<<g, 256, 0, s>>>(p);
  } else if (cs_ <= 32) {
    dim3 g(cdiv(p.M, 128), cdiv(cs_, 32));
    conv_dense_kernel<TI, TO, 128, 32, 4, 4><<<g, 256, 0, s>>>(p);
  } else {
    dim3 g(cdiv(p.M, 64), cdiv(cs_, 64));
    conv_dense_kernel<TI, TO, 64, 64, 4, 4><<<g, 256, 0, s>>>(p);
  }
}

void launch_conv_dense(const ConvP& p, int in_dt, int out_dt, cudaStream_t s) {
  if (in_dt == DT_F32) conv_dense_dispatch<float, float>(p, s);
  else if (out_dt == DT_F32) conv_dense_dispatch<bf16, float>(p, s);
  else conv_dense_dispatch<bf16, bf16>(p, s);
}

// =====================================================================================================================
// Depthwise k x k, stride 1, 'same' padding (ultralytics DWConv / LightConv.conv2 / GhostConv.cv2 / AAttn.pe).
// One thread = one pixel x 4 channels; weights [k*k][C] fp32.  HBM-bound: the k*k re-reads hit L1/L2.
// =====================================================================================================================
template <typename T, int K>   // K = 0: runtime kernel size
__global__ void __launch_bounds__(256) conv_dw_kernel(DwP p, long long total) {
  pdl_sync();
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int C4 = p.C >> 2;
  int c = (int)(e % C4) << 2;
  long long pix = e / C4;
  int x = (int)(pix % p.W);
  long long t = pix / p.W;
  int y = (int)(t % p.H);
  int n = (int)(t / p.H);
  const T* __restrict__ in = reinterpret_cast<const T*>(p.in);
  const int cin = (c / p.grp) * p.grp_stride + (c % p.grp);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (K > 0) {
    // compile-time kernel size: fully unrolled, all K*K loads independent and in flight together (the runtime-k loop
    // serialised one L1 round trip per tap)
    constexpr int KK = K > 0 ? K : 1;
    const T* base = in + ((size_t)(n * p.H + y) * p.W + x) * p.in_cs + cin;
    const float* wb = p.w + c;
#pragma unroll
    for (int r = 0; r < KK; ++r) {
      const int iy = y + r - KK / 2;
      const bool oky = iy >= 0 && iy < p.H;
      F4 v[KK];
#pragma unroll
      for (int s = 0; s < KK; ++s) {                 // issue the whole row of loads before any use
        const int ix = x + s - KK / 2;
        if (oky && ix >= 0 && ix < p.W) v[s] = load4<T>(base + ((long long)(r - KK / 2) * p.W + (s - KK / 2)) * p.in_cs);
        else v[s].v[0] = v[s].v[1] = v[s].v[2] = v[s].v[3] = 0.f;
      }
#pragma unroll
      for (int s = 0; s < KK; ++s) {
        float4 wv = *reinterpret_cast<const float4*>(wb + (size_t)(r * KK + s) * p.C);
        a0 = fmaf(v[s].v[0], wv.x, a0); a1 = fmaf(v[s].v[1], wv.y, a1);
        a2 = fmaf(v[s].v[2], wv.z, a2); a3 = fmaf(v[s].v[3], wv.w, a3);
      }
    }
  } else {
    for (int r = 0; r < p.k; ++r) {
      int iy = y + r - p.pad;
      if (iy < 0 || iy >= p.H) continue;
      for (int s = 0; s < p.k; ++s) {
        int ix = x + s - p.pad;
        if (ix < 0 || ix >= p.W) continue;
        F4 v = load4<T>(in + ((size_t)(n * p.H + iy) * p.W + ix) * p.in_cs + cin);
        float4 wv = *reinterpret_cast<const float4*>(p.w + (size_t)(r * p.k + s) * p.C + c);
        a0 = fmaf(v.v[0], wv.x, a0); a1 = fmaf(v.v[1], wv.y, a1);
        a2 = fmaf(v.v[2], wv.z, a2); a3 = fmaf(v.v[3], wv.w, a3);
      }
    }
  }
  float4 b = *reinterpret_cast<const float4*>(p.bias + c);
  F4 o;
  o.v[0] = apply_act_for<T>(a0 + b.x, p.act); o.v[1] = apply_act_for<T>(a1 + b.y, p.act);
  o.v[2] = apply_act_for<T>(a2 + b.z, p.act); o.v[3] = apply_act_for<T>(a3 + b.w, p.act);
  if (p.res) {
    F4 rv = load4<T>(reinterpret_cast<const T*>(p.res) + (size_t)pix * p.res_cs + c);
#pragma unroll
    for (int i = 0; i < 4; ++i) o.v[i] += rv.v[i];
  }
  store4<T>(reinterpret_cast<T*>(p.out) + (size_t)pix * p.out_cs + c, o);
}

// Tiled variant for the larger kernels (5x5 GhostConv, 7x7 AAttn.pe) and big maps: a TY x TX output tile of 16 channels
// per CTA, input tile + halo staged once in shared memory (fp32, pixel pitch 20 floats: conflict-free float4 reads),
// every thread slides a K-wide window over 4 consecutive output pixels so each staged value is reused from registers.
template <typename T, int K, int TY, int TX>
__global__ void __launch_bounds__(TY * TX) conv_dw_tiled_kernel(DwP p) {
  constexpr int NT = TY * TX;                        // (TX/4) x TY x 4 channel quads
  constexpr int IH = TY + K - 1, IW = TX + K - 1, PS = 20;
  pdl_sync();
  extern __shared__ __align__(16) float dsm[];
  float* sIn = dsm;                                  // [IH][IW][PS]
  float* sW = sIn + IH * IW * PS;                    // [K*K][16]
  const int tid = threadIdx.x;
  const int tiles_x = (p.W + TX - 1) / TX, tiles_y = (p.H + TY - 1) / TY;
  int t = blockIdx.x;
  const int tx0 = (t % tiles_x) * TX; t /= tiles_x;
  const int ty0 = (t % tiles_y) * TY; t /= tiles_y;
  const int ngrp = (p.C + 15) >> 4;
  const int cg = (t % ngrp) << 4;                    // 16-channel group (the last one may be partial: C % 4 == 0)
  const int n = t / ngrp;
  const int nq = min(4, (p.C - cg) >> 2);            // valid channel quads in this group
  const T* __restrict__ in = reinterpret_cast<const T*>(p.in);
  const int cin = (cg / p.grp) * p.grp_stride + (cg % p.grp);
  for (int i = tid; i < K * K * 4; i += NT)
    *reinterpret_cast<float4*>(sW + (i >> 2) * 16 + (i & 3) * 4) =
        (i & 3) < nq ? *reinterpret_cast<const float4*>(p.w + (size_t)(i >> 2) * p.C + cg + (i & 3) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  {
    // batches of 4 loads are issued before their shared-memory stores so the L2 round trips overlap
    constexpr int TOT = IH * IW * 4, UB = 4;
    for (int i0 = tid; i0 < TOT; i0 += NT * UB) {
      F4 f[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int i = i0 + u * NT;
        const int q = i & 3, pp = i >> 2;
        const int iy = ty0 + pp / IW - K / 2, ix = tx0 + pp % IW - K / 2;
        f[u].v[0] = f[u].v[1] = f[u].v[2] = f[u].v[3] = 0.f;
        if (i < TOT && q < nq && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
          f[u] = load4<T>(in + ((size_t)(n * p.H + iy) * p.W + ix) * p.in_cs + cin + q * 4);
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int i = i0 + u * NT;
        if (i < TOT) *reinterpret_cast<float4*>(sIn + (i >> 2) * PS + (i & 3) * 4) = make_float4(f[u].v[0], f[u].v[1], f[u].v[2], f[u].v[3]);
      }
    }
  }
  __syncthreads();
  const int q = tid & 3, txi = (tid >> 2) % (TX / 4), ty = tid / TX;
  float4 acc[4];
  const float4 bias = q < nq ? *reinterpret_cast<const float4*>(p.bias + cg + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int o = 0; o < 4; ++o) acc[o] = bias;
#pragma unroll
  for (int r = 0; r < K; ++r) {
    const float* rp = sIn + ((ty + r) * IW + txi * 4) * PS + q * 4;
    float4 v[K + 3], w[K];
#pragma unroll
    for (int j = 0; j < K + 3; ++j) v[j] = *reinterpret_cast<const float4*>(rp + j * PS);
#pragma unroll
    for (int s2 = 0; s2 < K; ++s2) w[s2] = *reinterpret_cast<const float4*>(sW + (r * K + s2) * 16 + q * 4);
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int s2 = 0; s2 < K; ++s2) {
        // packed fp32 FMAs (FFMA2, sm_100): two channels per instruction, bit-identical to scalar fmaf
        const float2 lo = __ffma2_rn(make_float2(v[o + s2].x, v[o + s2].y), make_float2(w[s2].x, w[s2].y), make_float2(acc[o].x, acc[o].y));
        const float2 hi = __ffma2_rn(make_float2(v[o + s2].z, v[o + s2].w), make_float2(w[s2].z, w[s2].w), make_float2(acc[o].z, acc[o].w));
        acc[o] = make_float4(lo.x, lo.y, hi.x, hi.y);
      }
  }
  const int y = ty0 + ty;
  if (y >= p.H || q >= nq) return;
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const int x = tx0 + txi * 4 + o;
    if (x >= p.W) break;
    const size_t pix = (size_t)(n * p.H + y) * p.W + x;
    F4 ov;
    ov.v[0] = apply_act_for<T>(acc[o].x, p.act); ov.v[1] = apply_act_for<T>(acc[o].y, p.act);
    ov.v[2] = apply_act_for<T>(acc[o].z, p.act); ov.v[3] = apply_act_for<T>(acc[o].w, p.act);
    if (p.res) {
      F4 rv = load4<T>(reinterpret_cast<const T*>(p.res) + pix * p.res_cs + cg + q * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) ov.v[i] += rv.v[i];
    }
    store4<T>(reinterpret_cast<T*>(p.out) + pix * p.out_cs + cg + q * 4, ov);
  }
}

template <typename T, int K, int TY, int TX>
static void conv_dw_tiled_launch(const DwP& p, cudaStream_t s) {
  constexpr size_t smem = sizeof(float) * ((TY + K - 1) * (TX + K - 1) * 20 + K * K * 16);
  static unsigned long long attr_done = 0;
  ensure_dyn_smem(conv_dw_tiled_kernel<T, K, TY, TX>, smem, attr_done, "conv_dw_tiled_kernel");
  const int tiles = ((p.W + TX - 1) / TX) * ((p.H + TY - 1) / TY) * ((p.C + 15) >> 4) * p.N;
  launch_pdl(conv_dw_tiled_kernel<T, K, TY, TX>, dim3(tiles), dim3(TY * TX), smem, s, p);
}

// bf16 variant for maps >= 30 wide (Ghost 5x5, Detect 3x3): the tiled kernel above is bound by shared-memory wavefronts
// (13 LDS.128 per 80 FMAs).  Here the 8 x 32-pixel tile is staged as packed bf16 (pixel pitch 9 words: a warp = one row =
// 8 channel pairs x 4 strips reads 32 distinct banks per LDS.32), a thread owns ONE channel pair of an 8-pixel strip,
// and slides a K-wide window over it: K+7 one-wavefront loads + K weight loads per 16K FMAs.
template <int K>
__global__ void __launch_bounds__(256) conv_dw_row_kernel(DwP p) {
  constexpr int TYR = 8, TXR = 32, IH = TYR + K - 1, IW = TXR + K - 1, PSW = 9;
  pdl_sync();
  __shared__ uint32_t sIn[IH * IW * PSW];
  __shared__ __align__(8) float sW[K * K * 16];
  const int tid = threadIdx.x;
  const int tiles_x = (p.W + TXR - 1) / TXR, tiles_y = (p.H + TYR - 1) / TYR;
  int t = blockIdx.x;
  const int tx0 = (t % tiles_x) * TXR; t /= tiles_x;
  const int ty0 = (t % tiles_y) * TYR; t /= tiles_y;
  const int ngrp = (p.C + 15) >> 4;
  const int cg = (t % ngrp) << 4;
  const int n = t / ngrp;
  const int nq = min(4, (p.C - cg) >> 2);            // valid channel quads in this group
  const bf16* __restrict__ in = reinterpret_cast<const bf16*>(p.in);
  const int cin = (cg / p.grp) * p.grp_stride + (cg % p.grp);
  for (int i = tid; i < K * K * 16; i += 256) sW[i] = ((i & 15) >> 2) < nq ? p.w[(size_t)(i >> 4) * p.C + cg + (i & 15)] : 0.f;
  {
    // all of a thread's 8-byte loads are issued before its first shared-memory store: the L2 round trips overlap
    constexpr int NIT = (IH * IW + 63) / 64;                            // 64 pixels (x 4 quads) per pass
    const int q = tid & 3;
    const int pp0 = tid >> 2;
    const bf16* base = in + (size_t)n * p.H * p.W * p.in_cs + cin + q * 4;
    uint2 d[NIT];
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      const int pp = pp0 + k * 64;
      const int ry = pp / IW, rx = pp - ry * IW;
      const int iy = ty0 + ry - K / 2, ix = tx0 + rx - K / 2;
      d[k] = make_uint2(0u, 0u);
      if (pp < IH * IW && q < nq && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
        d[k] = *reinterpret_cast<const uint2*>(base + (size_t)(iy * p.W + ix) * p.in_cs);
    }
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      const int pp = pp0 + k * 64;
      if (pp < IH * IW) { sIn[pp * PSW + q * 2] = d[k].x; sIn[pp * PSW + q * 2 + 1] = d[k].y; }
    }
  }
  __syncthreads();
  const int cp = tid & 7, strip = (tid >> 3) & 3, ty = tid >> 5;
  float2 acc[8];
  const float2 bias = (cp >> 1) < nq ? *reinterpret_cast<const float2*>(p.bias + cg + cp * 2) : make_float2(0.f, 0.f);
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = bias;
#pragma unroll
  for (int r = 0; r < K; ++r) {
    const uint32_t* rp = sIn + ((ty + r) * IW + strip * 8) * PSW + cp;
    float2 v[K + 7], w[K];
#pragma unroll
    for (int j = 0; j < K + 7; ++j) {
      const uint32_t u = rp[j * PSW];
      v[j] = make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
    }
#pragma unroll
    for (int s2 = 0; s2 < K; ++s2) w[s2] = *reinterpret_cast<const float2*>(sW + (r * K + s2) * 16 + cp * 2);
#pragma unroll
    for (int o = 0; o < 8; ++o)
#pragma unroll
      for (int s2 = 0; s2 < K; ++s2) {
        acc[o] = __ffma2_rn(v[o + s2], w[s2], acc[o]);          // FFMA2: the channel pair in one instruction
      }
  }
  const int y = ty0 + ty;
  if (y >= p.H || (cp >> 1) >= nq) return;
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    const int x = tx0 + strip * 8 + o;
    if (x >= p.W) break;
    const size_t pix = (size_t)(n * p.H + y) * p.W + x;
    float a = apply_act_for<bf16>(acc[o].x, p.act), b = apply_act_for<bf16>(acc[o].y, p.act);
    if (p.res) {
      const __nv_bfloat162 rv = *reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const bf16*>(p.res) + pix * p.res_cs + cg + cp * 2);
      a += __low2float(rv); b += __high2float(rv);
    }
    *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<bf16*>(p.out) + pix * p.out_cs + cg + cp * 2) = __floats2bfloat162_rn(a, b);
  }
}

template <int K>
static void conv_dw_row_launch(const DwP& p, cudaStream_t s) {
  const int tiles = ((p.W + 31) / 32) * ((p.H + 7) / 8) * ((p.C + 15) >> 4) * p.N;
  launch_pdl(conv_dw_row_kernel<K>, dim3(tiles), dim3(256), 0, s, p);
}

template <typename T>
static void conv_dw_dispatch(const DwP& p, long long total, int g, cudaStream_t s) {
  if (sizeof(T) == 2 && p.W >= 30 && (p.C % 4 == 0) && (p.grp % 16 == 0 || p.grp == p.C) && (p.k == 3 || p.k == 5) &&
      getenv("YSP_NO_DWROW") == nullptr) {
    if (p.k == 3) conv_dw_row_launch<3>(p, s); else conv_dw_row_launch<5>(p, s);
    return;
  }
  const bool tileable = (p.C % 4 == 0) && (p.grp % 16 == 0 || p.grp == p.C);
  if (tileable && (p.k == 3 || p.k == 5 || p.k == 7)) {
    const bool small = p.H <= 8 && p.W <= 8;
    if (p.k == 7)      { if (small) conv_dw_tiled_launch<T, 7, 8, 8>(p, s); else conv_dw_tiled_launch<T, 7, 16, 16>(p, s); }
    else if (p.k == 5) { if (small) conv_dw_tiled_launch<T, 5, 8, 8>(p, s); else conv_dw_tiled_launch<T, 5, 16, 16>(p, s); }
    else               { if (small) conv_dw_tiled_launch<T, 3, 8, 8>(p, s); else conv_dw_tiled_launch<T, 3, 16, 16>(p, s); }
    return;
  }
  if (p.k == 3) launch_pdl(conv_dw_kernel<T, 3>, dim3(g), dim3(256), 0, s, p, total);
  else if (p.k == 5) launch_pdl(conv_dw_kernel<T, 5>, dim3(g), dim3(256), 0, s, p, total);
  else if (p.k == 7) launch_pdl(conv_dw_kernel<T, 7>, dim3(g), dim3(256), 0, s, p, total);
  else launch_pdl(conv_dw_kernel<T, 0>, dim3(g), dim3(256), 0, s, p, total);
}

void launch_conv_dw(const DwP& p, int dt, cudaStream_t s) {
  long long total = (long long)p.N * p.H * p.W * (p.C >> 2);
  int g = cdiv(total, 256);
  if (dt == DT_F32) conv_dw_dispatch<float>(p, total, g, s);
  else conv_dw_dispatch<bf16>(p, total, g, s);
}

// =====================================================================================================================
// Elementwise / resampling
// =====================================================================================================================
template <typename T, int MODE>  // 0 add, 1 nearest x2, 2 bilinear x2
__global__ void __launch_bounds__(256) ew_kernel(EwP p, long long total) {
  pdl_sync();
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int C4 = p.C >> 2;
  int c = (int)(e % C4) << 2;
  long long pix = e / C4;
  int ox = (int)(pix % p.OW);
  long long t = pix / p.OW;
  int oy = (int)(t % p.OH);
  int n = (int)(t / p.OH);
  const T* __restrict__ a = reinterpret_cast<const T*>(p.a);
  F4 o;
  if (MODE == 0) {
    F4 x = load4<T>(a + (size_t)pix * p.a_cs + c);
    F4 y = load4<T>(reinterpret_cast<const T*>(p.b) + (size_t)pix * p.b_cs + c);
#pragma unroll
    for (int i = 0; i < 4; ++i) o.v[i] = x.v[i] + y.v[i];
  } else if (MODE == 1) {
    // nn.Upsample(scale_factor=2, mode="nearest") -- detector layers 9, 12
    o = load4<T>(a + ((size_t)(n * p.H + (oy >> 1)) * p.W + (ox >> 1)) * p.a_cs + c);
  } else {
    // nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False) -- YOLOSegPlusPlus.py:155
    // src = (dst + 0.5) * 0.5 - 0.5 clamped at 0; i1 = min(i0 + 1, in - 1); out = l0h*(l0w*a+l1w*b) + l1h*(l0w*c+l1w*d)
    float sy = fmaxf((oy + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf((ox + 0.5f) * 0.5f - 0.5f, 0.f);
    int y0 = (int)sy, x0 = (int)sx;
    int y1 = min(y0 + 1, p.H - 1), x1 = min(x0 + 1, p.W - 1);
    float ly = sy - y0, lx = sx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const size_t rb0 = (size_t)(n * p.H + y0) * p.W, rb1 = (size_t)(n * p.H + y1) * p.W;
    F4 v00 = load4<T>(a + (rb0 + x0) * p.a_cs + c), v01 = load4<T>(a + (rb0 + x1) * p.a_cs + c);
    F4 v10 = load4<T>(a + (rb1 + x0) * p.a_cs + c), v11 = load4<T>(a + (rb1 + x1) * p.a_cs + c);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o.v[i] = hy * (hx * v00.v[i] + lx * v01.v[i]) + ly * (hx * v10.v[i] + lx * v11.v[i]);
  }
  store4<T>(reinterpret_cast<T*>(p.out) + (size_t)pix * p.out_cs + c, o);
}

// bf16 fast path for add / nearest x2: 16-byte vectors (8 channels), 32-bit index arithmetic
template <int MODE>
__global__ void __launch_bounds__(256) ew8_bf16_kernel(EwP p, unsigned total) {
  pdl_sync();
  const unsigned e = blockIdx.x * 256u + threadIdx.x;
  if (e >= total) return;
  const unsigned G = (unsigned)p.C >> 3;
  const unsigned pix = e / G, c = (e - pix * G) << 3;
  const bf16* a = reinterpret_cast<const bf16*>(p.a);
  uint4 o;
  if (MODE == 0) {
    const uint4 x = *reinterpret_cast<const uint4*>(a + (size_t)pix * p.a_cs + c);
    const uint4 y = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.b) + (size_t)pix * p.b_cs + c);
    const __nv_bfloat162* hx = reinterpret_cast<const __nv_bfloat162*>(&x);
    const __nv_bfloat162* hy = reinterpret_cast<const __nv_bfloat162*>(&y);
    __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i)      // fp32 add, one rounding: same result as the generic kernel
      ho[i] = __floats2bfloat162_rn(__low2float(hx[i]) + __low2float(hy[i]), __high2float(hx[i]) + __high2float(hy[i]));
  } else {
    const unsigned ox = pix % (unsigned)p.OW, t = pix / (unsigned)p.OW;
    const unsigned oy = t % (unsigned)p.OH, n = t / (unsigned)p.OH;
    o = *reinterpret_cast<const uint4*>(a + ((size_t)(n * p.H + (oy >> 1)) * p.W + (ox >> 1)) * p.a_cs + c);
  }
  *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + (size_t)pix * p.out_cs + c) = o;
}

template <int MODE>
static void ew_launch(const EwP& p, int dt, cudaStream_t s) {
  long long total = (long long)p.N * p.OH * p.OW * (p.C >> 2);
  if (dt == DT_BF16 && MODE < 2 && p.C % 8 == 0 && p.a_cs % 8 == 0 && p.out_cs % 8 == 0 && (MODE == 1 || p.b_cs % 8 == 0) &&
      total / 2 < (1ll << 32)) {
    const unsigned t8 = (unsigned)(total / 2);
    launch_pdl(ew8_bf16_kernel<MODE>, dim3(cdiv(t8, 256)), dim3(256), 0, s, p, t8);
    return;
  }
  int g = cdiv(total, 256);
  if (dt == DT_F32) launch_pdl(ew_kernel<float, MODE>, dim3(g), dim3(256), 0, s, p, total);
  else launch_pdl(ew_kernel<bf16, MODE>, dim3(g), dim3(256), 0, s, p, total);
}
void launch_add(const EwP& p, int dt, cudaStream_t s) { ew_launch<0>(p, dt, s); }
void launch_up_nearest2(const EwP& p, int dt, cudaStream_t s) { ew_launch<1>(p, dt, s); }
void launch_up_bilinear2(const EwP& p, int dt, cudaStream_t s) { ew_launch<2>(p, dt, s); }

// =====================================================================================================================
// ECA (YOLOSegPlusPlus.py:60-88): global average pool -> conv1d(k=3, zero pad, no bias) over channels -> sigmoid ->
// scale in place.  Pass 1: per (slice, 32-channel group) mean; pass 2: vectorised scale.
// =====================================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) eca_mean_kernel(const T* __restrict__ x, int HW, int C, int cs, float* mean) {
  pdl_sync();
  __shared__ float red[8][33];
  int n = blockIdx.x, c = blockIdx.y * 32 + (threadIdx.x & 31), r = threadIdx.x >> 5;
  float s = 0.f;
  if (c < C)
    for (int i = r; i < HW; i += 8) s += to_f<T>(x[((size_t)n * HW + i) * cs + c]);
  red[r][threadIdx.x & 31] = s;
  __syncthreads();
  if (r == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
    mean[n * C + c] = t / (float)HW;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) eca_scale_kernel(T* x, int HW, int C, int cs, const float* __restrict__ mean,
                                                        const float* __restrict__ w3, long long total) {
  pdl_sync();
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int C4 = C >> 2;
  int c = (int)(e % C4) << 2;
  long long pix = e / C4;
  int n = (int)(pix / HW);
  const float* m = mean + n * C;
  F4 v = load4<T>(x + (size_t)pix * cs + c);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int cc = c + i;
    float z = w3[1] * m[cc];
    if (cc > 0) z += w3[0] * m[cc - 1];
    if (cc + 1 < C) z += w3[2] * m[cc + 1];
    v.v[i] *= sigmoid_f(z);
  }
  store4<T>(x + (size_t)pix * cs + c, v);
}
// bf16 fast path: one CTA per slice pools with 16-byte loads (8 channels per thread) and finishes the whole gate
// (mean -> conv1d k3 -> sigmoid) itself, so the scale pass is a pure 16-byte read-multiply-write stream.
__global__ void __launch_bounds__(512) eca_gate_bf16_kernel(const bf16* __restrict__ x, int HW, int C, int cs,
                                                            const float* __restrict__ w3, float* gate) {
  pdl_sync();
  extern __shared__ float esm[];               // [lanes][C] partial sums, then [C] means
  const int G = C >> 3, lanes = blockDim.x / G;
  const int g = threadIdx.x % G, lane = threadIdx.x / G;
  const bf16* xb = x + (size_t)blockIdx.x * HW * cs + g * 8;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (lane < lanes) {
    int i = lane;
    for (; i + 3 * lanes < HW; i += 4 * lanes) {           // four independent 16-byte loads in flight per thread
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(xb + (size_t)(i + u * lanes) * cs);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) { a[2 * j] += __low2float(h[j]); a[2 * j + 1] += __high2float(h[j]); }
      }
    }
    for (; i < HW; i += lanes) {
      const uint4 v = *reinterpret_cast<const uint4*>(xb + (size_t)i * cs);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) { a[2 * j] += __low2float(h[j]); a[2 * j + 1] += __high2float(h[j]); }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) esm[lane * C + g * 8 + j] = a[j];
  }
  __syncthreads();
  float* mean = esm + lanes * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += esm[l * C + c];
    mean[c] = t / (float)HW;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float z = w3[1] * mean[c];
    if (c > 0) z += w3[0] * mean[c - 1];
    if (c + 1 < C) z += w3[2] * mean[c + 1];
    gate[(size_t)blockIdx.x * C + c] = sigmoid_f(z);
  }
}
// fp32 twin of eca_gate_bf16_kernel (parity mode): one CTA per slice pools with 16-byte loads (4 channels per thread) and
// finishes mean -> conv1d k3 -> sigmoid itself.  The consumer conv multiplies by the gate while it converts its operands
// (conv_tc32.cu, ConvP::in_scale), so the scaled tensor is never written.
__global__ void __launch_bounds__(512) eca_gate_f32_kernel(const float* __restrict__ x, int HW, int C, int cs,
                                                           const float* __restrict__ w3, float* gate) {
  pdl_sync();
  extern __shared__ float esm32[];             // [lanes][C] partial sums, then [C] means
  const int G = C >> 2, lanes = blockDim.x / G;
  const int g = threadIdx.x % G, lane = threadIdx.x / G;
  const float* xb = x + (size_t)blockIdx.x * HW * cs + g * 4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < lanes) {
    int i = lane;
    for (; i + 3 * lanes < HW; i += 4 * lanes) {           // four independent 16-byte loads in flight per thread
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(xb + (size_t)(i + u * lanes) * cs);
#pragma unroll
      for (int u = 0; u < 4; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
    }
    for (; i < HW; i += lanes) {
      const float4 v = *reinterpret_cast<const float4*>(xb + (size_t)i * cs);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    *reinterpret_cast<float4*>(esm32 + lane * C + g * 4) = a;
  }
  __syncthreads();
  float* mean = esm32 + lanes * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += esm32[l * C + c];
    mean[c] = t / (float)HW;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float z = w3[1] * mean[c];
    if (c > 0) z += w3[0] * mean[c - 1];
    if (c + 1 < C) z += w3[2] * mean[c + 1];
    gate[(size_t)blockIdx.x * C + c] = sigmoid_f(z);
  }
}
__global__ void __launch_bounds__(256) eca_scale_bf16_kernel(bf16* x, int HW, int C, int cs, const float* __restrict__ gate,
                                                             unsigned total) {
  pdl_sync();
  const unsigned e = blockIdx.x * 256u + threadIdx.x;
  if (e >= total) return;
  const unsigned G = (unsigned)C >> 3;
  const unsigned pix = e / G, g = e - pix * G;
  const unsigned n = pix / (unsigned)HW;
  const float* gt = gate + (size_t)n * C + g * 8;
  const float4 g0 = *reinterpret_cast<const float4*>(gt), g1 = *reinterpret_cast<const float4*>(gt + 4);
  uint4* px = reinterpret_cast<uint4*>(x + (size_t)pix * cs + g * 8);
  uint4 v = *px;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
  h[0] = __floats2bfloat162_rn(__low2float(h[0]) * g0.x, __high2float(h[0]) * g0.y);
  h[1] = __floats2bfloat162_rn(__low2float(h[1]) * g0.z, __high2float(h[1]) * g0.w);
  h[2] = __floats2bfloat162_rn(__low2float(h[2]) * g1.x, __high2float(h[2]) * g1.y);
  h[3] = __floats2bfloat162_rn(__low2float(h[3]) * g1.z, __high2float(h[3]) * g1.w);
  *px = v;
}

void launch_eca(void* x, int N, int HW, int C, int cs, const float* w3, float* mean_ws, int dt, cudaStream_t s) {
  dim3 g(N, cdiv(C, 32));
  long long total = (long long)N * HW * (C >> 2);
  if (dt == DT_F32) {
    launch_pdl(eca_mean_kernel<float>, g, dim3(256), 0, s, (const float*)x, HW, C, cs, mean_ws);
    launch_pdl(eca_scale_kernel<float>, dim3(cdiv(total, 256)), dim3(256), 0, s, (float*)x, HW, C, cs, (const float*)mean_ws, w3, total);
  } else if (C % 8 == 0 && cs % 8 == 0 && C <= 512 && (long long)N * HW * (C >> 3) < (1ll << 32)) {
    const int G = C >> 3, lanes = 512 / G;
    launch_pdl(eca_gate_bf16_kernel, dim3(N), dim3(512), (size_t)(lanes + 1) * C * 4, s, (const bf16*)x, HW, C, cs, w3, mean_ws);
    const unsigned tot8 = (unsigned)((long long)N * HW * G);
    launch_pdl(eca_scale_bf16_kernel, dim3(cdiv(tot8, 256)), dim3(256), 0, s, (bf16*)x, HW, C, cs, (const float*)mean_ws, tot8);
  } else {
    launch_pdl(eca_mean_kernel<bf16>, g, dim3(256), 0, s, (const bf16*)x, HW, C, cs, mean_ws);
    launch_pdl(eca_scale_kernel<bf16>, dim3(cdiv(total, 256)), dim3(256), 0, s, (bf16*)x, HW, C, cs, (const float*)mean_ws, w3, total);
  }
}

bool eca_gate_f32_supported(int C, int cs) { return C % 4 == 0 && cs % 4 == 0 && C <= 512; }
void launch_eca_gate_f32(const void* x, int N, int HW, int C, int cs, const float* w3, float* gate, cudaStream_t s) {
  const int G = C >> 2, lanes = 512 / G;
  launch_pdl(eca_gate_f32_kernel, dim3(N), dim3(512), (size_t)(lanes + 1) * C * 4, s, (const float*)x, HW, C, cs, w3, gate);
}

// =====================================================================================================================
// Area-attention core (ultralytics AAttn.forward): per (slice, area, head)
//   out[i, h*32+d] = sum_j softmax_j(q_i . k_j * 32^-0.5) v_j[d],  head_dim = 32,
//   qkv channel layout per head = [q32 | k32 | v32]; areas = contiguous chunks of the flattened H*W index.
// One thread per query, K/V chunk staged in shared memory, online softmax in fp32.
// =====================================================================================================================
template <typename T>
__global__ void __launch_bounds__(128) attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, int Ntok,
                                                        int area, int qkv_cs, int out_cs) {
  constexpr int HD = 32, KC = 64;
  __shared__ float Ks[KC][HD];
  __shared__ float Vs[KC][HD];
  const int nt = Ntok / area;
  const int b = blockIdx.x / area, ar = blockIdx.x % area, h = blockIdx.y;
  const size_t tok0 = (size_t)b * Ntok + (size_t)ar * nt;
  const T* base = qkv + tok0 * qkv_cs + h * 3 * HD;
  const float scale = 0.17677669529663687f;  // 32 ** -0.5
  for (int q0 = 0; q0 < nt; q0 += blockDim.x) {
    const int qi = q0 + threadIdx.x;
    const bool qv = qi < nt;
    float q[HD], acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) { q[d] = qv ? to_f<T>(base[(size_t)qi * qkv_cs + d]) : 0.f; acc[d] = 0.f; }
    float mx = -INFINITY, l = 0.f;
    for (int j0 = 0; j0 < nt; j0 += KC) {
      __syncthreads();
      for (int e = threadIdx.x; e < KC * HD; e += blockDim.x) {
        int j = e / HD, d = e % HD;
        bool ok = j0 + j < nt;
        Ks[j][d] = ok ? to_f<T>(base[(size_t)(j0 + j) * qkv_cs + HD + d]) : 0.f;
        Vs[j][d] = ok ? to_f<T>(base[(size_t)(j0 + j) * qkv_cs + 2 * HD + d]) : 0.f;
      }
      __syncthreads();
      const int jn = min(KC, nt - j0);
      for (int j = 0; j < jn; ++j) {
        float sdot = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) sdot = fmaf(q[d], Ks[j][d], sdot);
        sdot *= scale;
        float mn = fmaxf(mx, sdot);
        float corr = expf(mx - mn), pj = expf(sdot - mn);
        l = l * corr + pj;
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] = acc[d] * corr + pj * Vs[j][d];
        mx = mn;
      }
    }
    if (qv) {
      float inv = 1.f / l;
      T* o = out + (tok0 + qi) * out_cs + h * HD;
#pragma unroll
      for (int d = 0; d < HD; ++d) o[d] = from_f<T>(acc[d] * inv);
    }
  }
}
void launch_attention(const void* qkv, void* out, int B, int Ntok, int C, int heads, int area, int qkv_cs, int out_cs,
                      int dt, cudaStream_t s) {
  (void)C;
  dim3 g(B * area, heads);
  if (dt == DT_F32) attention_kernel<float><<<g, 128, 0, s>>>((const float*)qkv, (float*)out, Ntok, area, qkv_cs, out_cs);
  else attention_kernel<bf16><<<g, 128, 0, s>>>((const bf16*)qkv, (bf16*)out, Ntok, area, qkv_cs, out_cs);
}

// =====================================================================================================================
// Layout / input kernels
// =====================================================================================================================
// fp32 NCHW -> NHWC view (C <= 4 handled per pixel; general C falls back to per-element)
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, T* __restrict__ out, int C,
                                                           int HW, int out_cs, long long total) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over N*HW
  if (e >= total) return;
  long long n = e / HW;
  int i = (int)(e % HW);
  for (int c = 0; c < C; ++c) out[(size_t)e * out_cs + c] = from_f<T>(in[((size_t)n * C + c) * HW + i]);
}
void launch_nchw_to_nhwc(const float* in, void* out, int N, int C, int H, int W, int out_cs, int dt, cudaStream_t s) {
  long long total = (long long)N * H * W;
  if (dt == DT_F32) nchw_to_nhwc_kernel<float><<<cdiv(total, 256), 256, 0, s>>>(in, (float*)out, C, H * W, out_cs, total);
  else nchw_to_nhwc_kernel<bf16><<<cdiv(total, 256), 256, 0, s>>>(in, (bf16*)out, C, H * W, out_cs, total);
}

// a1 (dataset.py:53-68 ToTensor): u8 HWC4 -> x/255.  True division so the value is bit-identical to torch's.
template <typename T>
__global__ void __launch_bounds__(256) u8_to_nhwc_kernel(const uchar4* __restrict__ in, T* __restrict__ out, int out_cs,
                                                         long long total) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  uchar4 v = in[e];
  F4 o; o.v[0] = v.x / 255.f; o.v[1] = v.y / 255.f; o.v[2] = v.z / 255.f; o.v[3] = v.w / 255.f;
  store4<T>(out + (size_t)e * out_cs, o);
}
void launch_u8_to_nhwc(const uint8_t* in, void* out, int N, int H, int W, int out_cs, int dt, cudaStream_t s) {
  long long total = (long long)N * H * W;
  if (dt == DT_F32) u8_to_nhwc_kernel<float><<<cdiv(total, 256), 256, 0, s>>>((const uchar4*)in, (float*)out, out_cs, total);
  else u8_to_nhwc_kernel<bf16><<<cdiv(total, 256), 256, 0, s>>>((const uchar4*)in, (bf16*)out, out_cs, total);
}
// u8 [N,H,W,4] -> fp32 NCHW [N,4,H,W] (the ToTensor output the reference hands to the model)
__global__ void __launch_bounds__(256) normalize_u8_kernel(const uchar4* __restrict__ in, float* __restrict__ out,
                                                           int HW, long long total) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  long long n = e / HW;
  int i = (int)(e % HW);
  uchar4 v = in[e];
  float* o = out + (size_t)n * 4 * HW + i;
  o[0] = v.x / 255.f; o[(size_t)HW] = v.y / 255.f; o[(size_t)2 * HW] = v.z / 255.f; o[(size_t)3 * HW] = v.w / 255.f;
}
void launch_normalize_u8(const uint8_t* in, float* out_nchw, int N, int H, int W, cudaStream_t s) {
  long long total = (long long)N * H * W;
  normalize_u8_kernel<<<cdiv(total, 256), 256, 0, s>>>((const uchar4*)in, out_nchw, H * W, total);
}

template <typename T>
__global__ void __launch_bounds__(256) logits_to_nhwc_kernel(const float* __restrict__ lg, T* __restrict__ out,
                                                             int out_cs, int zero_pad, long long total) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  T* o = out + (size_t)e * out_cs;
  o[0] = from_f<T>(lg[e]);
  for (int i = 1; i <= zero_pad; ++i) o[i] = from_f<T>(0.f);
}
// bf16 fast path: logit + 15 zero channels = one 32-byte record per pixel, written as two 16-byte stores
__global__ void __launch_bounds__(256) logits_to_nhwc16_kernel(const float* __restrict__ lg, bf16* __restrict__ out, int out_cs,
                                                               int total) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= total) return;
  const __nv_bfloat162 h = __floats2bfloat162_rn(lg[e], 0.f);
  uint4* o = reinterpret_cast<uint4*>(out + (size_t)e * out_cs);
  o[0] = make_uint4(*reinterpret_cast<const uint32_t*>(&h), 0u, 0u, 0u);
  o[1] = make_uint4(0u, 0u, 0u, 0u);
}
void launch_logits_to_nhwc(const float* logits, void* out, int N, int h, int w, int out_cs, int zero_pad, int dt,
                           cudaStream_t s) {
  long long total = (long long)N * h * w;
  if (dt == DT_BF16 && zero_pad == 15 && out_cs % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && total < (1ll << 31)) {
    logits_to_nhwc16_kernel<<<cdiv(total, 256), 256, 0, s>>>(logits, (bf16*)out, out_cs, (int)total);
    return;
  }
  if (dt == DT_F32) logits_to_nhwc_kernel<float><<<cdiv(total, 256), 256, 0, s>>>(logits, (float*)out, out_cs, zero_pad, total);
  else logits_to_nhwc_kernel<bf16><<<cdiv(total, 256), 256, 0, s>>>(logits, (bf16*)out, out_cs, zero_pad, total);
}

template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int HW,
                                                           int C, int in_cs, long long total) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over N*C*HW (NCHW order: coalesced writes)
  if (e >= total) return;
  int i = (int)(e % HW);
  long long t = e / HW;
  int c = (int)(t % C);
  long long n = t / C;
  out[e] = to_f<T>(in[((size_t)n * HW + i) * in_cs + c]);
}
void launch_nhwc_to_nchw_f32(const void* in, float* out, int N, int H, int W, int C, int in_cs, int dt, cudaStream_t s) {
  long long total = (long long)N * C * H * W;
  if (dt == DT_F32) nhwc_to_nchw_kernel<float><<<cdiv(total, 256), 256, 0, s>>>((const float*)in, out, H * W, C, in_cs, total);
  else nhwc_to_nchw_kernel<bf16><<<cdiv(total, 256), 256, 0, s>>>((const bf16*)in, out, H * W, C, in_cs, total);
}

// =====================================================================================================================
// Detect head decode (ultralytics Detect._inference + DFL + dist2bbox) fused with the NCHW copies of
// the raw maps the reference returns, and with the bottleneck extraction of evaluate_model.py:142-144.
// One thread per (slice, anchor).
// =====================================================================================================================
__global__ void __launch_bounds__(128) detect_decode_kernel(DecodeP p) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)p.B * p.A) return;
  int b = (int)(e / p.A), a = (int)(e % p.A);
  int lvl = 0, a0 = 0;
  int n0 = p.h[0] * p.w[0], n1 = p.h[1] * p.w[1];
  if (a >= n0 + n1) { lvl = 2; a0 = n0 + n1; }
  else if (a >= n0) { lvl = 1; a0 = n0; }
  const int al = a - a0;
  const int hw = p.h[lvl] * p.w[lvl];
  const int no = 64 + p.nc;
  const float* r = p.raw[lvl] + ((size_t)b * hw + al) * p.cs;
  float d[4];
#pragma unroll
  for (int sd = 0; sd < 4; ++sd) {
    float v[16], mx = -INFINITY;
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {                    // rows are 16-byte aligned (cs % 4 == 0): 16-byte loads
      const float4 t = *reinterpret_cast<const float4*>(r + sd * 16 + i4 * 4);
      v[i4 * 4] = t.x; v[i4 * 4 + 1] = t.y; v[i4 * 4 + 2] = t.z; v[i4 * 4 + 3] = t.w;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) mx = fmaxf(mx, v[i]);
    float se = 0.f, sw = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { float ex = expf(v[i] - mx); se += ex; sw = fmaf((float)i, ex, sw); }
    d[sd] = sw / se;
  }
  const int ax = al % p.w[lvl], ay = al / p.w[lvl];
  const float cx = ax + 0.5f, cy = ay + 0.5f, st = p.stride[lvl];
  const float x1 = cx - d[0], y1 = cy - d[1], x2 = cx + d[2], y2 = cy + d[3];
  if (p.y) {
    float* y = p.y + (size_t)b * (4 + p.nc) * p.A + a;
    y[0] = (x1 + x2) / 2.f * st;
    y[(size_t)p.A] = (y1 + y2) / 2.f * st;
    y[(size_t)2 * p.A] = (x2 - x1) * st;
    y[(size_t)3 * p.A] = (y2 - y1) * st;
    for (int c = 0; c < p.nc; ++c) y[(size_t)(4 + c) * p.A] = sigmoid_f(r[64 + c]);
  }
  if (p.p[lvl]) {
    float* o = p.p[lvl] + (size_t)b * no * hw + al;
    for (int c = 0; c < no; ++c) o[(size_t)c * hw] = r[c];
  }
  if (p.bott && lvl == 0 && ay < p.bh && ax < p.bw)
    p.bott[((size_t)b * p.bh + ay) * p.bw + ax] = sigmoid_f(r[no - 1]);
}
void launch_detect_decode(const DecodeP& p, cudaStream_t s) {
  long long total = (long long)p.B * p.A;
  detect_decode_kernel<<<cdiv(total, 128), 128, 0, s>>>(p);
}

// evaluate_model.py:142-144 as a standalone op on the NCHW raw map
__global__ void __launch_bounds__(256) bottleneck_kernel(const float* __restrict__ p3, int C, int Hs, int Ws,
                                                         float* __restrict__ out, int h, int w, long long total) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  int x = (int)(e % w);
  long long t = e / w;
  int y = (int)(t % h);
  long long b = t / h;
  out[e] = sigmoid_f(p3[(((size_t)b * C + (C - 1)) * Hs + y) * Ws + x]);
}
// same op on the engine's NHWC raw P3 map (fp32, pixel stride cs, class logit at channel ch): lets the pipeline publish
// the bottleneck as soon as the P3 class branch is done, before the rest of the Detect head
__global__ void __launch_bounds__(256) bottleneck_nhwc_kernel(const float* __restrict__ raw, int cs, int ch, int Hs, int Ws,
                                                              float* __restrict__ out, int h, int w, int total) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int x = e % w, t = e / w;
  const int y = t % h, b = t / h;
  out[e] = sigmoid_f(raw[((size_t)(b * Hs + y) * Ws + x) * cs + ch]);
}
void launch_bottleneck_nhwc(const float* raw, int cs, int ch, int B, int Hs, int Ws, float* logits, int h, int w, cudaStream_t s) {
  const int total = B * h * w;
  bottleneck_nhwc_kernel<<<cdiv(total, 256), 256, 0, s>>>(raw, cs, ch, Hs, Ws, logits, h, w, total);
}
void launch_bottleneck(const float* p3, int B, int C, int Hs, int Ws, float* logits, int h, int w, cudaStream_t s) {
  long long total = (long long)B * h * w;
  bottleneck_kernel<<<cdiv(total, 256), 256, 0, s>>>(p3, C, Hs, Ws, logits, h, w, total);
}

// =====================================================================================================================
// a9 mask + Dice counters (evaluate_model.py:157-158,166-174): P = sigmoid(x) > 0.5 evaluated in fp32 (NOT x > 0:
// logits in (0, ~9e-8] give sigmoid == 0.5), T = target > 0.5.  counts[b] = (|P&T|, |P|, |T|).  HBM-bound:
// 8 B read per pixel (+1 B optional mask write).  grid = (chunks, B); warp-shuffle reduce then 3 atomics per block.
// =====================================================================================================================
// TT = float (reference target tensor, T = t > 0.5) or uint8_t (the mask PNG as stored, before ToTensor: T = v/255 > 0.5
// <=> v >= 128; a quarter of the H2D bytes).  `bits` (optional): the mask bit-packed, pixel i of slice b = bit i%32 of word
// b*HW/32 + i/32 (needs HW % 128 == 0) -- 1/32 of the fp32 logits for the device->host read of a predict() caller.
template <typename TT>
__global__ void __launch_bounds__(256) mask_dice_kernel(const float* __restrict__ lg, const TT* __restrict__ tg,
                                                        int HW, int32_t* counts, uint8_t* mask, uint32_t* bits) {
  const int b = blockIdx.y;
  const float* x = lg + (size_t)b * HW;
  const TT* t = tg ? tg + (size_t)b * HW : nullptr;
  int ci = 0, cp = 0, ct = 0;
  auto one = [&](float xv, int ti, size_t idx) -> int {
    float s = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-xv)));
    int pi = s > 0.5f;
    ci += pi & ti; cp += pi; ct += ti;
    if (mask) mask[idx] = (uint8_t)pi;
    return pi;
  };
  if ((HW & 3) == 0) {                     // 16-byte loads (HW = 57600 on the hot path)
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < HW; i += gridDim.x * blockDim.x * 4) {
      const float4 xv = *reinterpret_cast<const float4*>(x + i);
      int t0 = 0, t1 = 0, t2 = 0, t3 = 0;
      if (t) {
        if (sizeof(TT) == 4) {
          const float4 tv = *reinterpret_cast<const float4*>(t + i);
          t0 = tv.x > 0.5f; t1 = tv.y > 0.5f; t2 = tv.z > 0.5f; t3 = tv.w > 0.5f;
        } else {
          const uchar4 tv = *reinterpret_cast<const uchar4*>(t + i);
          t0 = tv.x >= 128; t1 = tv.y >= 128; t2 = tv.z >= 128; t3 = tv.w >= 128;
        }
      }
      const size_t o = (size_t)b * HW + i;
      const int p0 = one(xv.x, t0, o), p1 = one(xv.y, t1, o + 1), p2 = one(xv.z, t2, o + 2), p3 = one(xv.w, t3, o + 3);
      if (bits) {                          // HW % 128 == 0: a warp covers 128 consecutive pixels = 4 words, all lanes active
        const int lane = threadIdx.x & 31;
        uint32_t v = (uint32_t)(p0 | (p1 << 1) | (p2 << 2) | (p3 << 3)) << (4 * (lane & 7));
        v |= __shfl_xor_sync(0xffffffffu, v, 1);
        v |= __shfl_xor_sync(0xffffffffu, v, 2);
        v |= __shfl_xor_sync(0xffffffffu, v, 4);
        if ((lane & 7) == 0) bits[o >> 5] = v;
      }
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
      int ti = 0;
      if (t) ti = sizeof(TT) == 4 ? (float)t[i] > 0.5f : (int)t[i] >= 128;
      one(x[i], ti, (size_t)b * HW + i);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ci += __shfl_xor_sync(0xffffffffu, ci, o);
    cp += __shfl_xor_sync(0xffffffffu, cp, o);
    ct += __shfl_xor_sync(0xffffffffu, ct, o);
  }
  __shared__ int red[3][8];
  int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][wid] = ci; red[1][wid] = cp; red[2][wid] = ct; }
  __syncthreads();
  if (threadIdx.x < 3) {
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
    if (s) atomicAdd(&counts[b * 3 + threadIdx.x], s);
  }
}
// The detection-confidence gate sketched (commented out) at evaluate_model.py:149-155: a slice whose best detection is
// missing or not above `thres` gets an all-zero predicted mask.  Applied after NMS + mask/Dice: zeroes |P&T| and |P| of
// the gated slices (|T| stays) and, if given, their binary mask.  det rows are [x1,y1,x2,y2,conf,cls], best first.
__global__ void __launch_bounds__(256) conf_gate_kernel(const float* __restrict__ det_boxes, const int32_t* __restrict__ det_count,
                                                        int max_det, int row, float thres, int32_t* counts, uint8_t* mask,
                                                        int HW, uint8_t* gated) {
  const int b = blockIdx.y;
  const bool keep = det_count[b] > 0 && det_boxes[(size_t)b * max_det * row + 4] > thres;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (gated) gated[b] = keep ? 0 : 1;
    if (!keep) { counts[b * 3] = 0; counts[b * 3 + 1] = 0; }
  }
  if (keep || !mask) return;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < HW; i += gridDim.x * 256) mask[(size_t)b * HW + i] = 0;
}
void launch_conf_gate(const float* det_boxes, const int32_t* det_count, int B, int max_det, int row, float thres, int32_t* counts,
                      uint8_t* mask, int HW, uint8_t* gated, cudaStream_t s) {
  conf_gate_kernel<<<dim3(mask ? 8 : 1, B), 256, 0, s>>>(det_boxes, det_count, max_det, row, thres, counts, mask, HW, gated);
}

void launch_mask_dice(const float* logits, const float* target, const uint8_t* target_u8, int B, int HW, int32_t* counts,
                      uint8_t* mask, uint32_t* bits, cudaStream_t s) {
  cudaMemsetAsync(counts, 0, sizeof(int32_t) * 3 * (size_t)B, s);
  int chunks = cdiv(HW, 256 * 8);
  if (chunks < 1) chunks = 1;
  dim3 g(chunks, B);
  if (target_u8 && !target) mask_dice_kernel<uint8_t><<<g, 256, 0, s>>>(logits, target_u8, HW, counts, mask, bits);
  else mask_dice_kernel<float><<<g, 256, 0, s>>>(logits, target, HW, counts, mask, bits);
}


// =====================================================================================================================
// bottleneck producer/consumer format.  generate_objectmaps.py:91-106 stores the raw P3 class-logit map
// per image; dataset.py:86-97 turns it into the training-time bottleneck  sigmoid((x - mean) / std)  per map (torch.std:
// unbiased; std == 0 -> x - mean).  One CTA per map, two-pass mean / variance in fp32 with fp64 block totals.
// =====================================================================================================================
__global__ void __launch_bounds__(256) objectmap_transform_kernel(const float* __restrict__ in, float* __restrict__ out, int n) {
  __shared__ double red[8];
  __shared__ float s_mean, s_inv;
  const float* x = in + (size_t)blockIdx.x * n;
  float* y = out + (size_t)blockIdx.x * n;
  auto block_sum = [&](double v) -> double {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    return t;
  };
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += x[i];
  const double mean = block_sum(s) / n;
  double q = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) { double d = (double)x[i] - mean; q += d * d; }
  const double var = n > 1 ? block_sum(q) / (n - 1) : 0.0;
  if (threadIdx.x == 0) { s_mean = (float)mean; float sd = (float)sqrt(var); s_inv = sd > 0.f ? 1.f / sd : 1.f; }
  __syncthreads();
  const float m = s_mean, inv = s_inv;
  for (int i = threadIdx.x; i < n; i += 256) y[i] = sigmoid_f((x[i] - m) * inv);
}
void launch_objectmap_transform(const float* in, float* out, int B, int n, cudaStream_t s) {
  if (B > 0) objectmap_transform_kernel<<<B, 256, 0, s>>>(in, out, n);
}

// ultralytics ops.scale_boxes as used by custom_detseg_predictor.py:177 -- xyxy boxes from the network
// canvas back to the original image: subtract the letterbox pad, divide by the gain, clip.  In place on [n, row] rows.
__global__ void __launch_bounds__(256) scale_boxes_kernel(float* boxes, long long n, int row, float gain, float pad_x,
                                                          float pad_y, float w0, float h0) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float* b = boxes + e * row;
  float x1 = (b[0] - pad_x) / gain, y1 = (b[1] - pad_y) / gain, x2 = (b[2] - pad_x) / gain, y2 = (b[3] - pad_y) / gain;
  b[0] = fminf(fmaxf(x1, 0.f), w0); b[1] = fminf(fmaxf(y1, 0.f), h0);
  b[2] = fminf(fmaxf(x2, 0.f), w0); b[3] = fminf(fmaxf(y2, 0.f), h0);
}
void launch_scale_boxes(float* boxes, long long n, int row, float gain, float pad_x, float pad_y, float w0, float h0, cudaStream_t s) {
  if (n > 0) scale_boxes_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(boxes, n, row, gain, pad_x, pad_y, w0, h0);
}

}  // namespace ysp

namespace ysp {
// =====================================================================================================================
// Slice ingest: cv2.resize on uint8 + transforms.ToTensor (dataset.py:59-70), bit-exact with OpenCV's
// generic uint8 path.  INTER_LINEAR: per-axis source index / weight from float((d + .5) * scale - .5) (double product,
// no FMA contraction), weights rounded to 11-bit fixed point, horizontal pass in int32, vertical pass
// (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2.  An exact 2x downscale is the rounded 2x2 mean (cv2 reroutes it to
// INTER_AREA).  INTER_NEAREST: min(floor(d * (1/(dn/sn))), sn-1).  One thread per output pixel, C in {1, 4}.
// =====================================================================================================================
struct AxisTap { int i0, i1, w0, w1; };

__device__ __forceinline__ AxisTap linear_tap(int d, double scale, int sn, bool clamp_weight) {
  float f = __double2float_rn(__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5));
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  AxisTap t;
  if (clamp_weight) {                       // x axis: cv2 resets the weight at the borders
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= sn - 1) { f = 0.f; s = sn - 1; }
    t.i0 = s; t.i1 = min(s + 1, sn - 1);
  } else {                                  // y axis: rows are clamped, weights kept
    t.i0 = min(max(s, 0), sn - 1); t.i1 = min(max(s + 1, 0), sn - 1);
  }
  t.w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  t.w1 = __float2int_rn(__fmul_rn(f, 2048.f));
  return t;
}

template <int C>
__global__ void __launch_bounds__(256) resize_u8_kernel(const uint8_t* __restrict__ src, int B, int sh, int sw, int dh,
                                                        int dw, int interp, double sx, double sy, uint8_t* dst_u8,
                                                        float* dst_f32) {
  const long long total = (long long)B * dh * dw;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const int x = (int)(e % dw);
    const long long t = e / dw;
    const int y = (int)(t % dh);
    const long long n = t / dh;
    const uint8_t* img = src + (size_t)n * sh * sw * C;
    int out[C];
    if (interp == 0) {
      int ix = min((int)floor(__dmul_rn((double)x, sx)), sw - 1), iy = min((int)floor(__dmul_rn((double)y, sy)), sh - 1);
#pragma unroll
      for (int c = 0; c < C; ++c) out[c] = img[((size_t)iy * sw + ix) * C + c];
    } else if (sh == 2 * dh && sw == 2 * dw) {
      const uint8_t* p = img + ((size_t)(2 * y) * sw + 2 * x) * C;
#pragma unroll
      for (int c = 0; c < C; ++c) out[c] = (p[c] + p[C + c] + p[(size_t)sw * C + c] + p[(size_t)sw * C + C + c] + 2) >> 2;
    } else {
      const AxisTap tx = linear_tap(x, sx, sw, true), ty = linear_tap(y, sy, sh, false);
      const uint8_t* r0 = img + (size_t)ty.i0 * sw * C;
      const uint8_t* r1 = img + (size_t)ty.i1 * sw * C;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        int h0 = r0[tx.i0 * C + c] * tx.w0 + r0[tx.i1 * C + c] * tx.w1;
        int h1 = r1[tx.i0 * C + c] * tx.w0 + r1[tx.i1 * C + c] * tx.w1;
        int v = (((ty.w0 * (h0 >> 4)) >> 16) + ((ty.w1 * (h1 >> 4)) >> 16) + 2) >> 2;
        out[c] = min(max(v, 0), 255);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (dst_u8) dst_u8[(size_t)e * C + c] = (uint8_t)out[c];
      if (dst_f32) dst_f32[(((size_t)n * C + c) * dh + y) * dw + x] = __fdiv_rn((float)out[c], 255.f);
    }
  }
}

void launch_resize_u8(const uint8_t* src, int B, int sh, int sw, int C, int dh, int dw, int interp, uint8_t* dst_u8,
                      float* dst_f32, cudaStream_t s) {
  double sx, sy;
  if (interp == 0) { sx = 1.0 / ((double)dw / sw); sy = 1.0 / ((double)dh / sh); }
  else { sx = (double)sw / dw; sy = (double)sh / dh; }
  long long total = (long long)B * dh * dw;
  int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  if (C == 4) resize_u8_kernel<4><<<grid, 256, 0, s>>>(src, B, sh, sw, dh, dw, interp, sx, sy, dst_u8, dst_f32);
  else resize_u8_kernel<1><<<grid, 256, 0, s>>>(src, B, sh, sw, dh, dw, interp, sx, sy, dst_u8, dst_f32);
}
}  // namespace ysp
