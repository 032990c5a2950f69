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

template <typename T, int C, int TH, int TW>
__global__ void __launch_bounds__(256) dlc_fused_kernel(DlcP p) {
  constexpr int C4 = C / 4;
  constexpr int PH = TH / 2 + 4, PW = TW / 2 + 4;      // low-res tile incl. halo
  constexpr int AH = TH + 4, AW = TW + 4;              // a: tile + 2
  constexpr int BH = TH + 2, BW = TW + 2;              // b, c: tile + 1
  constexpr int CS = C + 4;                            // padded pixel stride of b (conflict-free float4 per-pixel reads)
  extern __shared__ __align__(16) float sm[];
  float* sP = sm;                                      // [PH*PW][2C]
  float* sA = sP + PH * PW * 2 * C;                    // [AH*AW][C]      (re-used for c: [BH*BW][C])
  float* sB = sA + AH * AW * C;                        // [BH*BW][CS]
  float* sW2 = sB + BH * BW * CS;                      // [C][C]  (k-major: sW2[k*C + co])
  float* sB2 = sW2 + C * C;                            // [C]
  const int tid = threadIdx.x;
  const int H = 2 * p.h, W = 2 * p.w;
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int n = t / tiles_y;
  const int X0 = tx * TW, Y0 = ty * TH;                // hi-res tile origin (even)
  const int px0 = X0 / 2 - 2, py0 = Y0 / 2 - 2;        // low-res tile origin
  const T* __restrict__ P = reinterpret_cast<const T*>(p.P);

  // ---- stage 0: weights of the pointwise conv + the low-res P tile (edge-clamped) -> smem ----
  for (int i = tid; i < C * C; i += 256) sW2[i] = p.w2[(i / C) * p.w2ld + (i % C)];
  for (int i = tid; i < C; i += 256) sB2[i] = p.b2[i];
  for (int i = tid; i < PH * PW * (2 * C4); i += 256) {
    int c4 = i % (2 * C4), pp = i / (2 * C4);
    int gx = min(max(px0 + pp % PW, 0), p.w - 1), gy = min(max(py0 + pp / PW, 0), p.h - 1);
    F4 v = load4<T>(P + ((size_t)(n * p.h + gy) * p.w + gx) * p.p_cs + c4 * 4);
    *reinterpret_cast<float4*>(sP + pp * 2 * C + c4 * 4) = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
  }
  __syncthreads();

  // bilinear sample of channel group [cb, cb+4) of sP at hi-res pixel (Y, X) inside the image
  auto up4 = [&](int Y, int X, int cb) -> float4 {
    float sy = fmaxf(Y * 0.5f - 0.25f, 0.f), sx = fmaxf(X * 0.5f - 0.25f, 0.f);
    int y0 = (int)sy, x0 = (int)sx;
    int y1 = min(y0 + 1, p.h - 1), x1 = min(x0 + 1, p.w - 1);
    float ly = sy - y0, lx = sx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const float* r0 = sP + ((y0 - py0) * PW) * 2 * C + cb;
    const float* r1 = sP + ((y1 - py0) * PW) * 2 * C + cb;
    float4 v00 = *reinterpret_cast<const float4*>(r0 + (x0 - px0) * 2 * C), v01 = *reinterpret_cast<const float4*>(r0 + (x1 - px0) * 2 * C);
    float4 v10 = *reinterpret_cast<const float4*>(r1 + (x0 - px0) * 2 * C), v11 = *reinterpret_cast<const float4*>(r1 + (x1 - px0) * 2 * C);
    float4 o;
    o.x = hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
    o.y = hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
    o.z = hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z);
    o.w = hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w);
    return o;
  };

  // ---- stage 1: a = up2(P[:, :C]) on tile + 2 (zero outside the image) ----
  for (int i = tid; i < AH * AW * C4; i += 256) {
    int c4 = i % C4, pp = i / C4;
    int Y = Y0 - 2 + pp / AW, X = X0 - 2 + pp % AW;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (Y >= 0 && Y < H && X >= 0 && X < W) v = up4(Y, X, c4 * 4);
    *reinterpret_cast<float4*>(sA + pp * C + c4 * 4) = v;
  }
  __syncthreads();

  // ---- stage 2: b = SiLU(DW3x3(a) + b1) on tile + 1 ----
  {
    const int c4 = tid % C4;                           // fixed per thread (C4 divides 256)
    float4 wk[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wk[k] = *reinterpret_cast<const float4*>(p.dw1 + k * C + c4 * 4);
    const float4 bb = *reinterpret_cast<const float4*>(p.b1 + c4 * 4);
    for (int i = tid; i < BH * BW * C4; i += 256) {
      int pp = i / C4;
      int by = pp / BW, bx = pp % BW;                  // b(by,bx) <-> a(by..by+2, bx..bx+2)
      float4 acc = bb;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          float4 v = *reinterpret_cast<const float4*>(sA + ((by + r) * AW + bx + s) * C + c4 * 4);
          float4 w = wk[r * 3 + s];
          acc.x = fmaf(v.x, w.x, acc.x); acc.y = fmaf(v.y, w.y, acc.y); acc.z = fmaf(v.z, w.z, acc.z); acc.w = fmaf(v.w, w.w, acc.w);
        }
      acc.x = silu_f(acc.x); acc.y = silu_f(acc.y); acc.z = silu_f(acc.z); acc.w = silu_f(acc.w);
      *reinterpret_cast<float4*>(sB + pp * CS + c4 * 4) = acc;
    }
  }
  __syncthreads();

  // ---- stage 3: c = W2 b + b2 on tile + 1 (zero outside the image: padding of the second depthwise conv) -> sA ----
  // lane = pixel (b rows read at a padded stride: conflict free), 16 output channels per item (W2 reads broadcast)
  {
    constexpr int G = C / 16;
    for (int i = tid; i < BH * BW * G; i += 256) {
      const int pp = i % (BH * BW), g = i / (BH * BW);
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
      float* crow = sA + pp * C + g * 16;
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4)
        *reinterpret_cast<float4*>(crow + j4 * 4) = inside ? make_float4(acc[j4 * 4], acc[j4 * 4 + 1], acc[j4 * 4 + 2], acc[j4 * 4 + 3])
                                                           : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();

  // ---- stage 4: d = SiLU(DW3x3(c) + b3); out = d + up2(P[:, C:]); optional 1x1 head ----
  {
    const int c4 = tid % C4;
    float4 wk[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wk[k] = *reinterpret_cast<const float4*>(p.dw2 + k * C + c4 * 4);
    const float4 bb = *reinterpret_cast<const float4*>(p.b3 + c4 * 4);
    float4 wo = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.wo) wo = make_float4(p.wo[(c4 * 4 + 0) * p.wo_ld], p.wo[(c4 * 4 + 1) * p.wo_ld], p.wo[(c4 * 4 + 2) * p.wo_ld],
                               p.wo[(c4 * 4 + 3) * p.wo_ld]);   // dense-conv layout [K=C][ld], Cout = 1
    for (int i = tid; i < TH * TW * C4; i += 256) {     // TH*TW*C4 is a multiple of 256: whole warps stay converged
      int pp = i / C4;
      int oy = pp / TW, ox = pp % TW;
      int Y = Y0 + oy, X = X0 + ox;
      const bool inside = Y < H && X < W;
      float4 acc = bb;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          float4 v = *reinterpret_cast<const float4*>(sA + ((oy + r) * BW + ox + s) * C + c4 * 4);
          float4 w = wk[r * 3 + s];
          acc.x = fmaf(v.x, w.x, acc.x); acc.y = fmaf(v.y, w.y, acc.y); acc.z = fmaf(v.z, w.z, acc.z); acc.w = fmaf(v.w, w.w, acc.w);
        }
      float4 rs = inside ? up4(Y, X, C + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      F4 o;
      o.v[0] = silu_f(acc.x) + rs.x; o.v[1] = silu_f(acc.y) + rs.y; o.v[2] = silu_f(acc.z) + rs.z; o.v[3] = silu_f(acc.w) + rs.w;
      if (p.wo) {
        float part = o.v[0] * wo.x + o.v[1] * wo.y + o.v[2] * wo.z + o.v[3] * wo.w;
#pragma unroll
        for (int off = 1; off < C4; off <<= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        if (c4 == 0 && inside) reinterpret_cast<float*>(p.out)[((size_t)n * H + Y) * W + X] = part + p.bo[0];
      } else if (inside) {
        store4<T>(reinterpret_cast<T*>(p.out) + (((size_t)n * H + Y) * W + X) * p.out_cs + c4 * 4, o);
      }
    }
  }
}

template <typename T, int C, int TH, int TW>
static void dlc_launch(const DlcP& p, cudaStream_t s) {
  constexpr int PH = TH / 2 + 4, PW = TW / 2 + 4, AH = TH + 4, AW = TW + 4, BH = TH + 2, BW = TW + 2, CS = C + 4;
  constexpr size_t smem = sizeof(float) * (PH * PW * 2 * C + AH * AW * C + BH * BW * CS + C * C + C);
  static_assert((TH * TW * (C / 4)) % 256 == 0, "stage 4 must keep warps converged for the head shuffle");
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(dlc_fused_kernel<T, C, TH, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  const int H = 2 * p.h, W = 2 * p.w;
  const int tiles = ((W + TW - 1) / TW) * ((H + TH - 1) / TH) * p.N;
  dlc_fused_kernel<T, C, TH, TW><<<tiles, 256, smem, s>>>(p);
}

void launch_dlc_fused(const DlcP& p, int dt, cudaStream_t s) {
  if (dt == DT_F32) {
    if (p.C == 16) dlc_launch<float, 16, 16, 16>(p, s);
    else if (p.C == 32) dlc_launch<float, 32, 8, 16>(p, s);
    else dlc_launch<float, 64, 8, 8>(p, s);
  } else {
    if (p.C == 16) dlc_launch<bf16, 16, 16, 16>(p, s);
    else if (p.C == 32) dlc_launch<bf16, 32, 8, 16>(p, s);
    else dlc_launch<bf16, 64, 8, 8>(p, s);
  }
}

}  // namespace ysp
