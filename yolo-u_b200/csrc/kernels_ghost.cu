// kernels_ghost.cu -- the GhostBottleneck of the seg decoder's C3Ghost blocks as ONE kernel (bf16 mode).
//
// ultralytics GhostBottleneck(c, c, s=1) (YOLOSegPlusPlus.py:158,166 via C3Ghost):
//     g1 = SiLU(W1 a + b1)            1x1   c   -> c/4
//     g2 = SiLU(DW5(g1) + bd1)        5x5 depthwise
//     h1 = W3 [g1|g2] + b3            1x1   c/2 -> c/2   (no activation)
//     h2 = DW5(h1) + bd2              5x5 depthwise      (no activation)
//     out = [h1|h2] + a
// Unfused these are four convolutions and an add on 8..24-channel tensors: every one of them launch- or latency-bound
// (0.41 ms of the 5.5 ms step for 0.15 GFLOP).  Here a CTA owns a 16x16 output tile: it stages the 24x24 input patch
// (two 5x5 halos) once and keeps g1, g2, h1 in shared memory as bf16 -- the same rounding points as the unfused chain,
// which stores those tensors as bf16 in HBM.  The depthwise convs zero-pad their INPUT feature map, so g1 and h1 are
// forced to zero outside the image rather than computed from padded pixels.
#include <algorithm>

#include "kernels.h"

namespace ysp {

namespace {
constexpr int GT = 16, GRA = GT + 8, GRB = GT + 4;          // output tile, a/g1 region, g2/h1 region

__device__ __forceinline__ float2 bf2(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }
// Packed fp32 FMAs (FFMA2, sm_100): two of a thread's four channel accumulators per instruction -- half the FMA issue
// slots, bit-identical to scalar fmaf.
__device__ __forceinline__ void fma_s4(float (&a)[4], float x, const float4& w) {          // a += x * w
  const float2 lo = __ffma2_rn(make_float2(x, x), make_float2(w.x, w.y), make_float2(a[0], a[1]));
  const float2 hi = __ffma2_rn(make_float2(x, x), make_float2(w.z, w.w), make_float2(a[2], a[3]));
  a[0] = lo.x; a[1] = lo.y; a[2] = hi.x; a[3] = hi.y;
}
__device__ __forceinline__ void fma_v4(float (&a)[4], const float2& v0, const float2& v1, const float4& w) {   // a += [v0|v1] * w
  const float2 lo = __ffma2_rn(v0, make_float2(w.x, w.y), make_float2(a[0], a[1]));
  const float2 hi = __ffma2_rn(v1, make_float2(w.z, w.w), make_float2(a[2], a[3]));
  a[0] = lo.x; a[1] = lo.y; a[2] = hi.x; a[3] = hi.y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
}  // namespace

// Every stage gives a thread 4 horizontally adjacent pixels x 4 channels, so a weight vector loaded from shared memory
// feeds 16 FMAs (1x1 stages) and a 5-tap row feeds 80 (depthwise stages: 8 loaded pixels slide under 4 outputs).
template <int G>   // G = c/4: 8 (decoder.2) or 12 (decoder.0)
__global__ void __launch_bounds__(256) ghost_fused_kernel(GhostP p) {
  constexpr int CA = 4 * G, HC = 2 * G;
  constexpr int PA = CA / 2 + 1;          // words per pixel of the a tile (bf16 pairs; odd pitch spreads the 4-pixel strips over banks)
  constexpr int PG = G + 2;               // words per pixel of [g1|g2] (2G bf16 = G words); even: a thread's 4 channels are
  constexpr int PH = G + 2;               // one aligned 8-byte access.  Same for h1 (HC bf16 = G words)
  extern __shared__ __align__(16) uint32_t gsm[];
  float* sW1 = reinterpret_cast<float*>(gsm);          // [CA][G]
  float* sW3 = sW1 + CA * G;                           // [2G][HC]
  float* sD1 = sW3 + 2 * G * HC;                       // [25][G]
  float* sD2 = sD1 + 25 * G;                           // [25][HC]
  float* sB = sD2 + 25 * HC;                           // b1[G] bd1[G] b3[HC] bd2[HC]
  uint32_t* sA = reinterpret_cast<uint32_t*>(sB + 2 * G + 2 * HC);   // [GRA*GRA][PA]
  uint32_t* sG = sA + GRA * GRA * PA;                  // [GRA*GRA][PG]   g1 everywhere, g2 inside the GRB region
  uint32_t* sH = sA;                                   // [GRB*GRB][PH]   aliases the a tile: `a` is dead after stage A (the
                                                       // residual is re-read from global/L2 in stage D) -> 3 CTAs per SM for G = 8
  pdl_sync();
  const int tid = threadIdx.x;
  const int tiles_x = (p.W + GT - 1) / GT, tiles_y = (p.H + GT - 1) / GT;
  int t = blockIdx.x;
  const int X0 = (t % tiles_x) * GT; t /= tiles_x;
  const int Y0 = (t % tiles_y) * GT;
  const int n = t / tiles_y;
  // ---- weights (fp32, engine layouts: dense [K][ld], depthwise [tap][C]) ----
  for (int i = tid; i < CA * G; i += 256) sW1[i] = p.w1[(size_t)(i / G) * p.w1ld + i % G];
  for (int i = tid; i < 2 * G * HC; i += 256) sW3[i] = p.w3[(size_t)(i / HC) * p.w3ld + i % HC];
  for (int i = tid; i < 25 * G; i += 256) sD1[i] = p.dw1[i];
  for (int i = tid; i < 25 * HC; i += 256) sD2[i] = p.dw2[i];
  for (int i = tid; i < G; i += 256) { sB[i] = p.b1[i]; sB[G + i] = p.bd1[i]; }
  for (int i = tid; i < HC; i += 256) { sB[2 * G + i] = p.b3[i]; sB[2 * G + HC + i] = p.bd2[i]; }
  // ---- stage 0: a patch (24x24, zero outside the image).  All of a thread's 16-byte loads are issued before the first
  //      store, so the ~10 L2 round trips overlap instead of queueing behind each other ----
  {
    constexpr int V = CA / 8;                                           // 16-byte vectors per pixel
    constexpr int NIT = (GRA * GRA * V + 255) / 256;
    const bf16* base = p.a + (size_t)n * p.H * p.W * p.a_cs;
    uint4 d[NIT];
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      const int i = tid + k * 256;
      const int v = i % V, pp = i / V;
      const int y = Y0 - 4 + pp / GRA, x = X0 - 4 + pp % GRA;
      d[k] = make_uint4(0u, 0u, 0u, 0u);
      if (i < GRA * GRA * V && y >= 0 && y < p.H && x >= 0 && x < p.W)
        d[k] = *reinterpret_cast<const uint4*>(base + ((size_t)y * p.W + x) * p.a_cs + v * 8);
    }
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      const int i = tid + k * 256;
      if (i < GRA * GRA * V) {
        uint32_t* o = sA + (i / V) * PA + (i % V) * 4;
        o[0] = d[k].x; o[1] = d[k].y; o[2] = d[k].z; o[3] = d[k].w;
      }
    }
  }
  __syncthreads();
  // ---- stage A: g1 = SiLU(W1 a + b1) on the 24x24 region.  Strip width chosen so that all items fit ONE pass of the
  //      256 threads (G=8: 6-pixel strips -> 192 items; G=12: 8-pixel strips -> 216 items) ----
  {
    constexpr int Q = G / 4, SW = G == 8 ? 6 : 8, SPR = GRA / SW;
    for (int i = tid; i < GRA * SPR * Q; i += 256) {
      const int q = i % Q, st = i / Q;
      const int ry = st / SPR, rx = (st % SPR) * SW;
      const int pp0 = ry * GRA + rx;
      float acc[SW][4];
#pragma unroll
      for (int j = 0; j < SW; ++j) { acc[j][0] = sB[q * 4]; acc[j][1] = sB[q * 4 + 1]; acc[j][2] = sB[q * 4 + 2]; acc[j][3] = sB[q * 4 + 3]; }
      const uint32_t* ap = sA + pp0 * PA;
#pragma unroll 2
      for (int k2 = 0; k2 < CA / 2; ++k2) {
        const float4 w0 = *reinterpret_cast<const float4*>(sW1 + (2 * k2) * G + q * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(sW1 + (2 * k2 + 1) * G + q * 4);
#pragma unroll
        for (int j = 0; j < SW; ++j) {
          const float2 av = bf2(ap[j * PA + k2]);
          fma_s4(acc[j], av.x, w0);
          fma_s4(acc[j], av.y, w1);
        }
      }
      const int y = Y0 - 4 + ry;
#pragma unroll
      for (int j = 0; j < SW; ++j) {
        const int x = X0 - 4 + rx + j;
        const bool in = y >= 0 && y < p.H && x >= 0 && x < p.W;
        uint32_t* o = sG + (pp0 + j) * PG + q * 2;
        *reinterpret_cast<uint2*>(o) = in ? make_uint2(pack2(silu_approx(acc[j][0]), silu_approx(acc[j][1])), pack2(silu_approx(acc[j][2]), silu_approx(acc[j][3])))
                                          : make_uint2(0u, 0u);
      }
    }
  }
  __syncthreads();
  // ---- stage B: g2 = SiLU(DW5(g1) + bd1) on the 20x20 region ----
  {
    constexpr int Q = G / 4, SPR = GRB / 4;
    for (int i = tid; i < GRB * SPR * Q; i += 256) {
      const int q = i % Q, st = i / Q;
      const int by = st / SPR, bx = (st % SPR) * 4;                     // region-B coordinates; region-A = +2
      float acc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[j][0] = sB[G + q * 4]; acc[j][1] = sB[G + q * 4 + 1]; acc[j][2] = sB[G + q * 4 + 2]; acc[j][3] = sB[G + q * 4 + 3]; }
#pragma unroll
      for (int r = 0; r < 5; ++r) {
        const uint32_t* gp = sG + ((by + r) * GRA + bx) * PG + q * 2;
        float2 v0[8], v1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const uint2 u = *reinterpret_cast<const uint2*>(gp + j * PG); v0[j] = bf2(u.x); v1[j] = bf2(u.y); }
#pragma unroll
        for (int s = 0; s < 5; ++s) {
          const float4 w = *reinterpret_cast<const float4*>(sD1 + (r * 5 + s) * G + q * 4);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            fma_v4(acc[j], v0[j + s], v1[j + s], w);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t* o = sG + ((by + 2) * GRA + bx + 2 + j) * PG + Q * 2 + q * 2;   // g2 sits behind g1 in the same pixel record
        *reinterpret_cast<uint2*>(o) = make_uint2(pack2(silu_approx(acc[j][0]), silu_approx(acc[j][1])), pack2(silu_approx(acc[j][2]), silu_approx(acc[j][3])));
      }
    }
  }
  __syncthreads();
  // ---- stage C: h1 = W3 [g1|g2] + b3 on the 20x20 region (zero outside the image) ----
  {
    constexpr int Q = HC / 4, SPR = GRB / 4;
    for (int i = tid; i < GRB * SPR * Q; i += 256) {
      const int q = i % Q, st = i / Q;
      const int by = st / SPR, bx = (st % SPR) * 4;
      float acc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[j][0] = sB[2 * G + q * 4]; acc[j][1] = sB[2 * G + q * 4 + 1]; acc[j][2] = sB[2 * G + q * 4 + 2]; acc[j][3] = sB[2 * G + q * 4 + 3]; }
      const uint32_t* gp = sG + ((by + 2) * GRA + bx + 2) * PG;
#pragma unroll
      for (int k2 = 0; k2 < G; ++k2) {                                   // 2G inputs = G bf16 pairs
        const float4 w0 = *reinterpret_cast<const float4*>(sW3 + (2 * k2) * HC + q * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(sW3 + (2 * k2 + 1) * HC + q * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 gv = bf2(gp[j * PG + k2]);
          fma_s4(acc[j], gv.x, w0);
          fma_s4(acc[j], gv.y, w1);
        }
      }
      const int y = Y0 - 2 + by;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = X0 - 2 + bx + j;
        const bool in = y >= 0 && y < p.H && x >= 0 && x < p.W;
        uint32_t* o = sH + (by * GRB + bx + j) * PH + q * 2;
        *reinterpret_cast<uint2*>(o) = in ? make_uint2(pack2(acc[j][0], acc[j][1]), pack2(acc[j][2], acc[j][3])) : make_uint2(0u, 0u);
      }
    }
  }
  __syncthreads();
  // ---- stage D: h2 = DW5(h1) + bd2; out = [h1 | h2] + a ----
  {
    constexpr int Q = HC / 4, SPR = GT / 4;
    bf16* obase = p.out + (size_t)n * p.H * p.W * p.out_cs;
    for (int i = tid; i < GT * SPR * Q; i += 256) {
      const int q = i % Q, st = i / Q;
      const int ty = st / SPR, tx = (st % SPR) * 4;
      float acc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j][0] = sB[2 * G + HC + q * 4]; acc[j][1] = sB[2 * G + HC + q * 4 + 1];
        acc[j][2] = sB[2 * G + HC + q * 4 + 2]; acc[j][3] = sB[2 * G + HC + q * 4 + 3];
      }
#pragma unroll
      for (int r = 0; r < 5; ++r) {
        const uint32_t* hp = sH + ((ty + r) * GRB + tx) * PH + q * 2;
        float2 v0[8], v1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const uint2 u = *reinterpret_cast<const uint2*>(hp + j * PH); v0[j] = bf2(u.x); v1[j] = bf2(u.y); }
#pragma unroll
        for (int s = 0; s < 5; ++s) {
          const float4 w = *reinterpret_cast<const float4*>(sD2 + (r * 5 + s) * HC + q * 4);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            fma_v4(acc[j], v0[j + s], v1[j + s], w);
          }
        }
      }
      const int y = Y0 + ty;
      if (y >= p.H) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = X0 + tx + j;
        if (x >= p.W) break;
        const uint32_t* hc = sH + ((ty + 2) * GRB + tx + j + 2) * PH + q * 2;           // h1 at the output pixel
        const bf16* ag = p.a + ((size_t)(n * p.H + y) * p.W + x) * p.a_cs;              // a at the output pixel (L2-resident)
        const uint2 alo = *reinterpret_cast<const uint2*>(ag + q * 4), ahi = *reinterpret_cast<const uint2*>(ag + HC + q * 4);
        const uint2 hu = *reinterpret_cast<const uint2*>(hc);
        const float2 h0 = bf2(hu.x), h1v = bf2(hu.y);
        const float2 a0 = bf2(alo.x), a1 = bf2(alo.y);                                  // a[4q .. 4q+3]
        const float2 a2 = bf2(ahi.x), a3 = bf2(ahi.y);                                  // a[HC + 4q ..]
        bf16* op = obase + ((size_t)y * p.W + x) * p.out_cs;
        // rounding points of the unfused chain: h1 is a bf16 tensor, h2 + a is rounded once
        *reinterpret_cast<uint2*>(op + q * 4) = make_uint2(pack2(h0.x + a0.x, h0.y + a0.y), pack2(h1v.x + a1.x, h1v.y + a1.y));
        *reinterpret_cast<uint2*>(op + HC + q * 4) = make_uint2(pack2(acc[j][0] + a2.x, acc[j][1] + a2.y), pack2(acc[j][2] + a3.x, acc[j][3] + a3.y));
      }
    }
  }
}

static size_t ghost_smem(int G) {
  const int CA = 4 * G, HC = 2 * G, PA = CA / 2 + 1, PG = G + 2, PH = G + 2;
  (void)PH;                                            // the h1 tile aliases the (larger) a tile
  return 4 * ((size_t)GRA * GRA * PA + (size_t)GRA * GRA * PG) +
         4 * ((size_t)CA * G + 2 * G * HC + 25 * G + 25 * HC + 2 * G + 2 * HC);
}

bool ghost_fused_supported(int c, int a_cs, int out_cs) { return (c == 32 || c == 48) && a_cs % 8 == 0 && out_cs % 4 == 0; }

void launch_ghost_fused(const GhostP& p, int c, cudaStream_t s) {
  const int tiles = ((p.W + GT - 1) / GT) * ((p.H + GT - 1) / GT) * p.N;
  const int G = c / 4;
  const size_t smem = ghost_smem(G);
  if (G == 8) {
    static unsigned long long attr_done = 0;
    ensure_dyn_smem(ghost_fused_kernel<8>, smem, attr_done, "ghost_fused_kernel<8>");
    launch_pdl(ghost_fused_kernel<8>, dim3(tiles), dim3(256), smem, s, p);
  } else {
    static unsigned long long attr_done = 0;
    ensure_dyn_smem(ghost_fused_kernel<12>, smem, attr_done, "ghost_fused_kernel<12>");
    launch_pdl(ghost_fused_kernel<12>, dim3(tiles), dim3(256), smem, s, p);
  }
}

}  // namespace ysp
