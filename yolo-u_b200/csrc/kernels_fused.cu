// kernels_fused.cu -- the decoder's DoubleLightConv stages (YOLOSegPlusPlus.py:33-58 behind nn.Upsample(bilinear x2),
// :155-175) as ONE kernel per stage that keeps every full-resolution intermediate in shared memory.
//
// Algebra used (exact in real arithmetic): both 1x1 convs applied directly to the upsampled tensor -- LightConv.conv1
// (Conv-BN, no activation) and residual_conv (+bias) -- are linear and bilinear weights sum to 1, so
//      conv1x1(up(x)) == up(conv1x1(x)).
// They are therefore evaluated at LOW resolution (4x fewer pixels) by one tensor-core GEMM with concatenated weights
// (P = [conv1 | residual], 2C channels), and this kernel does the rest per hi-res output tile:
//      a = up2(P[:, :C])                (bilinear, align_corners=False; zero outside the image = conv padding)
//      b = SiLU(DW3x3(a) + b1)          LightConv 0 depthwise
//      c = W2 b + b2                    LightConv 1 pointwise (BN folded, linear)
//      d = SiLU(DW3x3(c) + b3)          LightConv 1 depthwise
//      out = d + up2(P[:, C:])          residual add (YOLOSegPlusPlus.py:55-57)
//      [head] logit = wo . out + bo     self.output (1x1, 16 -> 1), fused for the last stage
// HBM traffic per stage: read P (low-res, 2C ch) + write out -- the unfused chain moved ~8 full-resolution tensors.
#include "kernels.h"

namespace ysp {

// SiLU(x) = h + h * tanh(h), h = x/2: one MUFU op (tanh.approx, rel. error ~2^-11 -- below bf16 resolution)
__device__ __forceinline__ float silu_fast(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
template <bool FAST> __device__ __forceinline__ float silu_t(float x) { return FAST ? silu_fast(x) : silu_f(x); }

// Packed fp32 arithmetic (FFMA2 / FMUL2, sm_100): two channels per instruction, half the issue slots of the depthwise
// windows and the bilinear blends.
__device__ __forceinline__ void fma4(float4& acc, const float4& v, const float4& w) {
  const float2 lo = __ffma2_rn(make_float2(v.x, v.y), make_float2(w.x, w.y), make_float2(acc.x, acc.y));
  const float2 hi = __ffma2_rn(make_float2(v.z, v.w), make_float2(w.z, w.w), make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 lerp4(const float4& a, float wa, const float4& b, float wb) {
  const float2 wa2 = make_float2(wa, wa), wb2 = make_float2(wb, wb);
  const float2 lo = __ffma2_rn(wa2, make_float2(a.x, a.y), __fmul2_rn(wb2, make_float2(b.x, b.y)));
  const float2 hi = __ffma2_rn(wa2, make_float2(a.z, a.w), __fmul2_rn(wb2, make_float2(b.z, b.w)));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Depthwise 3x3 over a column strip of S outputs: each input row is loaded once (3 float4) and feeds up to 3 outputs.
// src points at the top-left input pixel of the strip (pixel stride SC floats, row stride RW pixels).
template <int S>
__device__ __forceinline__ void dw_strip(const float* src, int RW, int SC, const float4 (&wk)[9], const float4& bias, float4 (&acc)[S]) {
#pragma unroll
  for (int o = 0; o < S; ++o) acc[o] = bias;
#pragma unroll
  for (int ir = 0; ir < S + 2; ++ir) {
    const float* rp = src + ir * RW * SC;
    float4 v0 = *reinterpret_cast<const float4*>(rp), v1 = *reinterpret_cast<const float4*>(rp + SC), v2 = *reinterpret_cast<const float4*>(rp + 2 * SC);
#pragma unroll
    for (int o = 0; o < S; ++o) {
      const int r = ir - o;
      if (r >= 0 && r < 3) { fma4(acc[o], v0, wk[r * 3]); fma4(acc[o], v1, wk[r * 3 + 1]); fma4(acc[o], v2, wk[r * 3 + 2]); }
    }
  }
}

// FAST = bf16 "throughput mode": fast SiLU, pointwise conv on mma.sync bf16 tensor-core tiles (b and W2 rounded to
// bf16, exactly what the unfused bf16 path stores); !FAST = fp32 parity mode: everything in fp32 FFMA.
template <typename T, int C, int TH, int TW, int S2, int S4, bool FAST>
__global__ void __launch_bounds__(256, (FAST && C < 64) ? 3 : 2) dlc_fused_kernel(DlcP p) {   // C = 64: shared memory allows two CTAs per SM anyway
  constexpr int C4 = C / 4;
  constexpr int PH = TH / 2 + 4, PW = TW / 2 + 4;      // low-res tile incl. halo
  constexpr int AH = TH + 4, AW = TW + 4;              // a: tile + 2
  constexpr int BH = TH + 2, BW = TW + 2;              // b, c: tile + 1
  constexpr int NB = BH * BW;
  constexpr int CS = C + 4;                            // fp32 b: padded pixel stride (conflict-free float4 per-pixel reads)
  constexpr int CSH = C + 8;                           // bf16 b: padded pixel stride (conflict-free mma A-fragment loads)
  constexpr int CC = (FAST && C == 64) ? C + 8 : C;    // c: pixel stride (padded: the float2 accumulator stores of stage 3 hit
                                                       // 16 distinct 8-byte slots per half-warp instead of 4; fits inside the a tile)
  static_assert(BH * BW * CC <= AH * AW * C, "c tile must fit in the a tile");
  static_assert(BH % S2 == 0 && TH % S4 == 0 && S4 % 2 == 0, "strip heights");
  pdl_sync();
  extern __shared__ __align__(16) float sm[];
  float* sP = sm;                                      // [2][PH*PW][C]: plane 0 = conv1 half, plane 1 = residual half
                                                       // (two planes: consecutive pixels are contiguous -> no bank conflicts)
  float* sA = sP + PH * PW * 2 * C;                    // [AH*AW][C]      (re-used for c: [BH*BW][C])
  float* sB = sA + AH * AW * C;                        // fp32 [NB][CS]  or  bf16 [NB16][CSH]
  constexpr int NB16 = (NB + 15) / 16 * 16;
  constexpr int SB_FLOATS = FAST ? (NB16 * CSH + 1) / 2 : NB * CS;
  float* sW2 = sB + SB_FLOATS;                         // fp32 [C][C] k-major (sW2[k*C+co])  (!FAST only)
  float* sB2 = sW2 + (FAST ? 0 : C * C);               // [C]
  bf16* sW2h = reinterpret_cast<bf16*>(sB2 + C);       // bf16 [co][CSH]: W2^T for the mma B fragments (FAST only)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = 2 * p.h, W = 2 * p.w;
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int n = t / tiles_y;
  const int X0 = tx * TW, Y0 = ty * TH;                // hi-res tile origin (even)
  const int px0 = X0 / 2 - 2, py0 = Y0 / 2 - 2;        // low-res tile origin
  const T* __restrict__ P = reinterpret_cast<const T*>(p.P);

  // ---- stage 0: pointwise weights + the low-res P tile (edge-clamped: out-of-range neighbours replicate the border,
  //      which is exactly torch's index clamping for align_corners=False) -> smem ----
  if (!FAST) for (int i = tid; i < C * C; i += 256) sW2[i] = p.w2[(i / C) * p.w2ld + (i % C)];
  if (FAST) {
    // W2^T as bf16 pairs (k, k+1): a quarter of a warp covers 8 output channels x 4 k-pairs, so the global loads use
    // whole 32-byte sectors and the 4-byte shared stores fall into 32 different banks (word = co * (CSH/2) + k/2).
    // (Before: every thread built its B fragments from 128 scalar global loads -- 18 % of the kernel's instructions.)
    for (int i = tid; i < C * C / 2; i += 256) {
      const int rest = i >> 5;
      const int co = (rest % (C / 8)) * 8 + (i & 7), k = ((rest / (C / 8)) * 4 + ((i >> 3) & 3)) * 2;
      const __nv_bfloat162 w = __floats2bfloat162_rn(p.w2[(size_t)k * p.w2ld + co], p.w2[(size_t)(k + 1) * p.w2ld + co]);
      *reinterpret_cast<__nv_bfloat162*>(sW2h + co * CSH + k) = w;
    }
  }
  for (int i = tid; i < C; i += 256) sB2[i] = p.b2[i];
  {
    // batches of 4 global loads are issued before their shared-memory stores so the L2 round trips overlap
    constexpr int TOT = PH * PW * (2 * C4), UB = 4;
    for (int i0 = tid; i0 < TOT; i0 += 256 * UB) {
      F4 v[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int i = i0 + u * 256;
        if (i < TOT) {
          const int c4 = i % (2 * C4), pp = i / (2 * C4);
          const int gx = min(max(px0 + pp % PW, 0), p.w - 1), gy = min(max(py0 + pp / PW, 0), p.h - 1);
          v[u] = load4<T>(P + ((size_t)(n * p.h + gy) * p.w + gx) * p.p_cs + c4 * 4);
        }
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int i = i0 + u * 256;
        if (i < TOT) {
          const int c4 = i % (2 * C4), pp = i / (2 * C4);
          *reinterpret_cast<float4*>(sP + (c4 / C4) * (PH * PW * C) + pp * C + (c4 % C4) * 4) = make_float4(v[u].v[0], v[u].v[1], v[u].v[2], v[u].v[3]);
        }
      }
    }
  }
  __syncthreads();

  // ---- stage 1: a = up2(P[:, :C]) on tile + 2, one thread per (low-res pixel, 4 channels) -> a 2x2 hi-res block.
  //      x2 bilinear (align_corners=False) has fixed weights: even output = .25*prev + .75*cur, odd = .75*cur + .25*next.
  //      Blocks outside the image are zero (= padding of the first depthwise conv). ----
  for (int i = tid; i < (AH / 2) * (AW / 2) * C4; i += 256) {
    const int c4 = i % C4, bb = i / C4;
    const int bi = bb / (AW / 2), bj = bb % (AW / 2);
    const int li = bi + 1, lj = bj + 1;                // position in the P tile (low-res coordinate py0 + li)
    const int gi = py0 + li, gj = px0 + lj;
    float4 o00, o01, o10, o11;
    if (gi >= 0 && gi < p.h && gj >= 0 && gj < p.w) {
      const float* c = sP + (li * PW + lj) * C + c4 * 4;
      constexpr int RS = PW * C, PS = C;
      float4 m0 = *reinterpret_cast<const float4*>(c - RS - PS), m1 = *reinterpret_cast<const float4*>(c - RS), m2 = *reinterpret_cast<const float4*>(c - RS + PS);
      float4 z0 = *reinterpret_cast<const float4*>(c - PS), z1 = *reinterpret_cast<const float4*>(c), z2 = *reinterpret_cast<const float4*>(c + PS);
      float4 q0 = *reinterpret_cast<const float4*>(c + RS - PS), q1 = *reinterpret_cast<const float4*>(c + RS), q2 = *reinterpret_cast<const float4*>(c + RS + PS);
      float4 t0 = lerp4(m0, 0.25f, z0, 0.75f), t1 = lerp4(m1, 0.25f, z1, 0.75f), t2 = lerp4(m2, 0.25f, z2, 0.75f);   // even row
      float4 u0 = lerp4(z0, 0.75f, q0, 0.25f), u1 = lerp4(z1, 0.75f, q1, 0.25f), u2 = lerp4(z2, 0.75f, q2, 0.25f);   // odd row
      o00 = lerp4(t0, 0.25f, t1, 0.75f); o01 = lerp4(t1, 0.75f, t2, 0.25f);
      o10 = lerp4(u0, 0.25f, u1, 0.75f); o11 = lerp4(u1, 0.75f, u2, 0.25f);
    } else {
      o00 = o01 = o10 = o11 = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float* d = sA + ((2 * bi) * AW + 2 * bj) * C + c4 * 4;
    *reinterpret_cast<float4*>(d) = o00; *reinterpret_cast<float4*>(d + C) = o01;
    *reinterpret_cast<float4*>(d + AW * C) = o10; *reinterpret_cast<float4*>(d + AW * C + C) = o11;
  }
  __syncthreads();

  // ---- stage 2: b = SiLU(DW3x3(a) + b1) on tile + 1, column strips of S2 outputs ----
  {
    const int c4 = tid % C4;                           // fixed per thread (C4 divides 256)
    float4 wk[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wk[k] = *reinterpret_cast<const float4*>(p.dw1 + k * C + c4 * 4);
    const float4 bias = *reinterpret_cast<const float4*>(p.b1 + c4 * 4);
    for (int i = tid; i < (BH / S2) * BW * C4; i += 256) {
      const int it = i / C4;
      const int bx = it % BW, by = (it / BW) * S2;
      float4 acc[S2];
      dw_strip<S2>(sA + (by * AW + bx) * C + c4 * 4, AW, C, wk, bias, acc);
#pragma unroll
      for (int o = 0; o < S2; ++o) {
        const int pp = (by + o) * BW + bx;
        float4 r = make_float4(silu_t<FAST>(acc[o].x), silu_t<FAST>(acc[o].y), silu_t<FAST>(acc[o].z), silu_t<FAST>(acc[o].w));
        if (FAST) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(r.x, r.y), h1 = __floats2bfloat162_rn(r.z, r.w);
          uint2 u; u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(sB) + pp * CSH + c4 * 4) = u;
        } else {
          *reinterpret_cast<float4*>(sB + pp * CS + c4 * 4) = r;
        }
      }
    }
  }
  __syncthreads();

  // ---- stage 3: c = W2 b + b2 on tile + 1 (zero outside the image: padding of the second depthwise conv) -> sA ----
  if (FAST) {
    // mma.sync m16n8k16 bf16: A = b (16 pixels x 16 k, row-major in smem), B = W2^T fragments held in registers.
    constexpr int KS = C / 16, NT = C / 8;
    constexpr int NTH = NT > 4 ? NT / 2 : NT;          // n-tiles per pass (<= 32 B-fragment registers live at once)
    const int g = lane >> 2, tig = lane & 3;
    const bf16* sBh = reinterpret_cast<const bf16*>(sB);
#pragma unroll 1
    for (int nh = 0; nh < NT / NTH; ++nh) {
      uint32_t bf[KS][NTH][2];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int nt = 0; nt < NTH; ++nt) {
          const int co = (nh * NTH + nt) * 8 + g;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)
            bf[ks][nt][hh] = *reinterpret_cast<const uint32_t*>(sW2h + co * CSH + ks * 16 + tig * 2 + hh * 8);
        }
      for (int mt = warp; mt < NB16 / 16; mt += 8) {
        float d[NTH][4];
#pragma unroll
        for (int nt = 0; nt < NTH; ++nt) {
          const int co = (nh * NTH + nt) * 8 + tig * 2;
          d[nt][0] = d[nt][2] = sB2[co]; d[nt][1] = d[nt][3] = sB2[co + 1];
        }
        const int r0 = mt * 16 + g, r1 = r0 + 8;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t a[4];
          const bf16* ap = sBh + ks * 16 + tig * 2;
          a[0] = *reinterpret_cast<const uint32_t*>(ap + r0 * CSH); a[1] = *reinterpret_cast<const uint32_t*>(ap + r1 * CSH);
          a[2] = *reinterpret_cast<const uint32_t*>(ap + r0 * CSH + 8); a[3] = *reinterpret_cast<const uint32_t*>(ap + r1 * CSH + 8);
#pragma unroll
          for (int nt = 0; nt < NTH; ++nt) mma_bf16_16816(d[nt], a, bf[ks][nt][0], bf[ks][nt][1]);
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int pp = hh ? r1 : r0;
          if (pp < NB) {
            const int Y = Y0 - 1 + pp / BW, X = X0 - 1 + pp % BW;
            const bool inside = Y >= 0 && Y < H && X >= 0 && X < W;
#pragma unroll
            for (int nt = 0; nt < NTH; ++nt) {
              const int co = (nh * NTH + nt) * 8 + tig * 2;
              float2 v = inside ? make_float2(d[nt][hh * 2], d[nt][hh * 2 + 1]) : make_float2(0.f, 0.f);
              *reinterpret_cast<float2*>(sA + pp * CC + co) = v;
            }
          }
        }
      }
    }
  } else {
    // fp32: lane = pixel (b rows read at a padded stride: conflict free), 16 output channels per item (W2 broadcast)
    constexpr int G = C / 16;
    for (int i = tid; i < NB * G; i += 256) {
      const int pp = i % NB, g = i / NB;
      const int Y = Y0 - 1 + pp / BW, X = X0 - 1 + pp % BW;
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = sB2[g * 16 + j];
      const float* brow = sB + pp * CS;
#pragma unroll 4
      for (int k = 0; k < C; k += 4) {
        float4 bv = *reinterpret_cast<const float4*>(brow + k);
        const float bk[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float* wr = sW2 + (k + kk) * C + g * 16;
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            float4 w = *reinterpret_cast<const float4*>(wr + j4 * 4);
            acc[j4 * 4 + 0] = fmaf(bk[kk], w.x, acc[j4 * 4 + 0]); acc[j4 * 4 + 1] = fmaf(bk[kk], w.y, acc[j4 * 4 + 1]);
            acc[j4 * 4 + 2] = fmaf(bk[kk], w.z, acc[j4 * 4 + 2]); acc[j4 * 4 + 3] = fmaf(bk[kk], w.w, acc[j4 * 4 + 3]);
          }
        }
      }
      const bool inside = Y >= 0 && Y < H && X >= 0 && X < W;
      float* crow = sA + pp * CC + g * 16;
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4)
        *reinterpret_cast<float4*>(crow + j4 * 4) = inside ? make_float4(acc[j4 * 4], acc[j4 * 4 + 1], acc[j4 * 4 + 2], acc[j4 * 4 + 3])
                                                           : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();

  // ---- stage 4: d = SiLU(DW3x3(c) + b3); out = d + up2(P[:, C:]); optional 1x1 head.  Column strips of S4. ----
  {
    const int c4 = tid % C4;
    float4 wk[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wk[k] = *reinterpret_cast<const float4*>(p.dw2 + k * C + c4 * 4);
    const float4 bias = *reinterpret_cast<const float4*>(p.b3 + c4 * 4);
    float4 wo = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.wo) wo = make_float4(p.wo[(c4 * 4 + 0) * p.wo_ld], p.wo[(c4 * 4 + 1) * p.wo_ld], p.wo[(c4 * 4 + 2) * p.wo_ld],
                               p.wo[(c4 * 4 + 3) * p.wo_ld]);   // dense-conv layout [K=C][ld], Cout = 1
    const float bo = p.wo ? p.bo[0] : 0.f;
    // (TH/S4)*TW*C4 is a multiple of 256 (static_assert at launch): whole warps stay converged for the head shuffle
    for (int i = tid; i < (TH / S4) * TW * C4; i += 256) {
      const int it = i / C4;
      const int ox = it % TW, oy0 = (it / TW) * S4;
      float4 acc[S4];
      dw_strip<S4>(sA + (oy0 * BW + ox) * CC + c4 * 4, BW, CC, wk, bias, acc);
      const int X = X0 + ox;
      // residual: bilinear x2 of P[:, C:].  Column pair + weights are fixed by the parity of X; the strip starts on an
      // even row, so its S4 rows need S4/2 + 2 low-res rows: lerp each once horizontally, then blend vertically.
      const int lj = (X >> 1) - px0;                   // low-res column of X in the P tile
      const int ja = (X & 1) ? lj : lj - 1, jb = (X & 1) ? lj + 1 : lj;
      const float wa = (X & 1) ? 0.75f : 0.25f, wb = 1.f - wa;
      const int li0 = ((Y0 + oy0) >> 1) - py0;         // low-res row of the strip's first output
      const float* rp0 = sP + PH * PW * C + ((li0 - 1) * PW) * C + c4 * 4;
      float4 hrow[S4 / 2 + 2];
#pragma unroll
      for (int k = 0; k < S4 / 2 + 2; ++k)
        hrow[k] = lerp4(*reinterpret_cast<const float4*>(rp0 + (k * PW + ja) * C), wa, *reinterpret_cast<const float4*>(rp0 + (k * PW + jb) * C), wb);
#pragma unroll
      for (int o = 0; o < S4; ++o) {
        const int Y = Y0 + oy0 + o;
        const bool inside = Y < H && X < W;
        const float4 rs = (o & 1) ? lerp4(hrow[o / 2 + 1], 0.75f, hrow[o / 2 + 2], 0.25f) : lerp4(hrow[o / 2], 0.25f, hrow[o / 2 + 1], 0.75f);
        F4 ov;
        ov.v[0] = silu_t<FAST>(acc[o].x) + rs.x; ov.v[1] = silu_t<FAST>(acc[o].y) + rs.y;
        ov.v[2] = silu_t<FAST>(acc[o].z) + rs.z; ov.v[3] = silu_t<FAST>(acc[o].w) + rs.w;
        if (p.wo) {
          float part = ov.v[0] * wo.x + ov.v[1] * wo.y + ov.v[2] * wo.z + ov.v[3] * wo.w;
#pragma unroll
          for (int off = 1; off < C4; off <<= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
          if (c4 == 0 && inside) reinterpret_cast<float*>(p.out)[((size_t)n * H + Y) * W + X] = part + bo;
        } else if (inside) {
          store4<T>(reinterpret_cast<T*>(p.out) + (((size_t)n * H + Y) * W + X) * p.out_cs + c4 * 4, ov);
        }
      }
    }
  }
}

template <typename T, int C, int TH, int TW, int S2, int S4, bool FAST>
static void dlc_launch(const DlcP& p, cudaStream_t s) {
  constexpr int PH = TH / 2 + 4, PW = TW / 2 + 4, AH = TH + 4, AW = TW + 4, BH = TH + 2, BW = TW + 2, CS = C + 4, CSH = C + 8;
  constexpr int NB = BH * BW, NB16 = (NB + 15) / 16 * 16;
  constexpr int SB_FLOATS = FAST ? (NB16 * CSH + 1) / 2 : NB * CS;
  constexpr size_t smem = sizeof(float) * (PH * PW * 2 * C + AH * AW * C + SB_FLOATS + (FAST ? 0 : C * C) + C) + (FAST ? (size_t)C * CSH * 2 : 0);
  static_assert(((TH / S4) * TW * (C / 4)) % 256 == 0, "stage 4 must keep warps converged for the head shuffle");
  static unsigned long long attr_done = 0;
  ensure_dyn_smem(dlc_fused_kernel<T, C, TH, TW, S2, S4, FAST>, smem, attr_done, "dlc_fused_kernel");
  const int H = 2 * p.h, W = 2 * p.w;
  const int tiles = ((W + TW - 1) / TW) * ((H + TH - 1) / TH) * p.N;
  launch_pdl(dlc_fused_kernel<T, C, TH, TW, S2, S4, FAST>, dim3(tiles), dim3(256), smem, s, p);
}

void launch_dlc_fused(const DlcP& p, int dt, cudaStream_t s) {
  if (dt == DT_F32) {
    if (p.C == 16) dlc_launch<float, 16, 16, 16, 3, 4, false>(p, s);
    else if (p.C == 32) dlc_launch<float, 32, 8, 16, 2, 4, false>(p, s);
    else dlc_launch<float, 64, 8, 8, 2, 2, false>(p, s);
  } else {
    if (p.C == 16) dlc_launch<bf16, 16, 16, 16, 3, 4, true>(p, s);
    else if (p.C == 32) dlc_launch<bf16, 32, 8, 16, 2, 4, true>(p, s);
    else dlc_launch<bf16, 64, 8, 8, 2, 2, true>(p, s);
  }
}

}  // namespace ysp
