// kernels_dlc32.cu -- decoder stage (bilinear x2 + DoubleLightConv [+ 1x1 mask head]) for the PARITY mode (YSP_MODE_TC32):
// fp32 activations, fp32-accurate arithmetic, the pointwise GEMM on tcgen05 with fp16 hi/lo operand splits.
//
// Reference stage (YOLOSegPlusPlus.py:33-58 behind nn.Upsample(bilinear x2), :155-175):
//     u = up2(x);  a = W1 u + c1;  b = SiLU(DW1(a) + b1);  c = W2 b + c2;  d = SiLU(DW2(c) + b3);  out = d + (Wr u + cr)
// Algebra (exact in real arithmetic):
//   * both 1x1 convs on u commute with up2 (bilinear weights sum to 1): P = [W1 x + c1 | Wr x + cr] is ONE low-resolution
//     GEMM (conv_tc32.cu, 4x fewer pixels), a = up2(P[:, :C]), residual = up2(P[:, C:]);
//   * DW1(up2(P)) is a linear map from the 3x3 low-res neighbourhood of a pixel to the pixel, with weights that depend only
//     on the pixel's (row, column) PARITY:  w[di][dj] = sum_{r,s} dw1[r][s] * V[pY][r][di] * V[pX][s][dj]  (V = the x2
//     bilinear coefficients of the three rows Y-1, Y, Y+1 on the low-res rows i-1, i, i+1).  The hi-res tensor `a` is never
//     formed: 9 FMAs per pixel and channel straight from the low-res tile instead of 4-tap up2 + 9-tap depthwise.  Only the
//     outermost image rows / columns differ (the reference zero-pads `a`, the composite would extrapolate it): those
//     pixels are recomputed by a generic path after the main pass.
// Per 14 x (BW-2) output tile (b/c region 16 x BW = 4 or 2 MMA row blocks of 128 pixels):
//   A. low-res P tile -> shared memory by TMA (one 4-D box per channel half, zero-filled outside the map; tiles on the image
//      border then replicate the edge rows / columns into the halo = torch's index clamping).  The conv1 half of tile i+1 and
//      the (double-buffered) residual half are fetched while tile i is computed, so no load latency is exposed.
//   B. b = SiLU(composite 3x3 on P + b1) -> fp16 hi / lo tiles in the canonical K-major NO-SWIZZLE UMMA layout
//      [8-channel plane][pixel] x 16 B                                                                   (CUDA cores, FFMA2)
//   C. c = W2 b + c2: 3 x C/16 tcgen05.mma (M = 128, N = C) per row block (hi.hi + lo.hi + hi.lo, fp32 accumulation in
//      TMEM); epilogue tcgen05.ld -> * 2^-e + c2, zero outside the image -> fp32 channel-plane tile      (tensor cores)
//   D. d = SiLU(DW2(c) + b3) + up2(P[:, C:]) -> NHWC fp32 store, or the fused 1x1 head (16 -> 1)          (CUDA cores)
// Column strips with a register sliding window: every low-res / c row is loaded once per strip (3 x 16 B) and feeds three
// output rows.  SiLU = x * rcp(1 + ex2(-x log2 e)) on the MUFU unit (2 ulp; the tolerance is 1e-3 on the mask logits).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include "kernels.h"

namespace ysp {

namespace {

constexpr int TH = 14, BH = 16, PH = 11;     // output rows, b/c rows, low-res rows per tile

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWL32:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DN32;\n\tbra WL32;\n\tDN32:\n\t}"
               ::"r"(s32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ float4 silu4(const float4& v) {
  const float2 a = silu2_f(make_float2(v.x, v.y)), b = silu2_f(make_float2(v.z, v.w));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(s32(dst)), "l"(tm), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
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
__device__ __forceinline__ void split4(const float4& v, uint2& hi, uint2& lo) {
  split2_f16(v.x, v.y, hi.x, lo.x);
  split2_f16(v.z, v.w, hi.y, lo.y);
}
// x2 bilinear (align_corners=False) coefficient of hi-res row (Y - 1 + r), r = 0..2, on low-res row (i - 1 + di), i = Y >> 1:
//   Y even: r0 = (.75,.25,0)  r1 = (.25,.75,0)  r2 = (0,.75,.25);   Y odd: r0 = (.25,.75,0)  r1 = (0,.75,.25)  r2 = (0,.25,.75)
__device__ __forceinline__ float vcoef(int parity, int r, int di) {
  const int yr = parity - 1 + r;                       // row relative to 2i: -1 .. 2
  const int k = yr >> 1;                               // arithmetic shift: floor
  const int ra = (yr & 1) ? k : k - 1;                 // rows (ra, ra + 1) relative to i
  const float wa = (yr & 1) ? 0.75f : 0.25f;
  const int d = di - 1;
  return d == ra ? wa : (d == ra + 1 ? 1.0f - wa : 0.0f);
}

template <int C, int BW, int RESB, bool ALIAS>
struct Geo {
  static constexpr int TW = BW - 2, PW = BW / 2 + 3, NPIX = BH * BW, NBLK = NPIX / 128;
  static constexpr int LBO = NPIX * 16 + 16;                        // plane pitch inside a 16-channel k-step (== 16 mod 128: the
  static constexpr int KSTEP = 2 * LBO + (C == 32 ? 32 : 0);        //  8-byte stores of a half-warp fall into 16 distinct slots)
  static constexpr int B_BYTES = (C / 16) * KSTEP;                  // one of the hi / lo tiles
  static constexpr int CPL = NPIX * 16 + 32;                        // c tile: [C/4 planes][pixel] x 16 B, plane pitch == 32 mod 128
  static constexpr int C_BYTES = (C / 4) * CPL;
  static constexpr int PHALF = PH * PW * C * 4;                     // one channel half of the low-res tile: [li][lj][C] fp32
  static constexpr int PPITCH = (PHALF + 127) / 128 * 128;          // 128-byte aligned TMA destinations
  static constexpr int P_BYTES = (1 + RESB) * PPITCH;               // conv1 half + RESB buffers of the residual half
  static constexpr int KC = C % 32 == 0 ? 32 : 16;                  // Cin chunk of the tc32 weight pack (tc32_tiling)
  static constexpr int W_BYTES = 2 * C * C * 2;                     // hi + lo
  static constexpr int T_FLOATS = 4 * 9 * C + 9 * C + 4 * C + 4;    // composite weights [cls][tap][C], dw2 [tap][C], b1, b2, b3, head w + b
  // ALIAS: the c tile lies over the b tiles (dead once every MMA of the tile has completed): less shared memory, one more barrier
  static constexpr int BC_BYTES = ALIAS ? (2 * B_BYTES > C_BYTES ? 2 * B_BYTES : C_BYTES) : 2 * B_BYTES + C_BYTES;
  static constexpr size_t SMEM = (size_t)P_BYTES + BC_BYTES + W_BYTES + T_FLOATS * 4 + 128;
  static constexpr int TCOLS = NBLK * C < 32 ? 32 : NBLK * C;       // TMEM columns (power of two: 64 / 128 / 256)
};

}  // namespace

// NT threads per CTA (512 / NT CTAs per SM: the register file holds 512 threads of this kernel); RESB = 2: the residual half
// of tile i+1 is prefetched with the conv1 half during tile i; RESB = 1 (less shared memory): fetched during phase C of its tile
template <int C, int BW, int NT, int RESB, bool ALIAS, bool HEAD>
__global__ void __launch_bounds__(NT, 512 / NT) dlc32_kernel(const __grid_constant__ CUtensorMap tmP, const Dlc32P p) {
  using G = Geo<C, BW, RESB, ALIAS>;
  constexpr int kD32Threads = NT;
  constexpr int C4 = C / 4, TW = G::TW, PW = G::PW, NPIX = G::NPIX, NBLK = G::NBLK;
  extern __shared__ __align__(128) uint8_t dsm32[];
  float* sP = reinterpret_cast<float*>(dsm32);                         // conv1 half [PH][PW][C], then residual half x 2 buffers
  uint8_t* sBh = dsm32 + G::P_BYTES;                                   // b hi: k-step ks, plane j at ks*KSTEP + j*LBO, pixel m at m*16
  uint8_t* sBl = sBh + G::B_BYTES;
  uint8_t* sC = ALIAS ? sBh : sBl + G::B_BYTES;                        // [C4][NPIX] x 16 B (pitch CPL)
  uint8_t* sW = sBh + G::BC_BYTES;                                     // tc32 weight pack of conv.1.conv1
  float* sWc = reinterpret_cast<float*>(sW + G::W_BYTES);              // composite up2 o DW1 weights [pY*2+pX][di*3+dj][C]
  float* sWk = sWc + 4 * 9 * C;                                        // dw2 [9][C]
  float* sB1 = sWk + 9 * C;                                            // b1, b2 (= c2), b3
  float* sB2 = sB1 + C;
  float* sB3 = sB2 + C;
  float* sWo = sB3 + C;                                               // head weights [C] + bias (HEAD only)
  __shared__ __align__(8) uint64_t bar[4], pbar[3];                    // MMA row blocks done; TMA: conv1 half, residual buffers 0/1
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = 2 * p.h, W = 2 * p.w;
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  const int total_tiles = tiles_x * tiles_y * p.N;

  // ---- prologue: constants only (overlaps the previous kernel under PDL) ----
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmP) : "memory");
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
    for (int i = 0; i < 3; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&pbar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_s)), "r"((uint32_t)G::TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (int i = tid; i < G::W_BYTES / 16; i += kD32Threads) reinterpret_cast<uint4*>(sW)[i] = reinterpret_cast<const uint4*>(p.wpack)[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  for (int i = tid; i < 4 * 9 * C; i += kD32Threads) {
    const int c = i % C, tap = (i / C) % 9, cls = i / (9 * C);
    const int di = tap / 3, dj = tap % 3, pY = cls >> 1, pX = cls & 1;
    float a = 0.f;
    for (int r = 0; r < 3; ++r)
      for (int q = 0; q < 3; ++q) a = fmaf(p.dw1[(r * 3 + q) * C + c], vcoef(pY, r, di) * vcoef(pX, q, dj), a);
    sWc[i] = a;
  }
  for (int i = tid; i < 9 * C; i += kD32Threads) sWk[i] = p.dw2[i];
  for (int i = tid; i < C; i += kD32Threads) { sB1[i] = p.b1[i]; sB2[i] = p.b2[i]; sB3[i] = p.b3[i]; if (HEAD) sWo[i] = p.wo[(size_t)i * p.wo_ld]; }
  if (HEAD && tid == 0) sWo[C] = p.bo[0];

  // phase-B item of this thread: (c4, cx, pX, pY); composite weights in registers
  constexpr int ITEMS_B = 4 * (BW / 2) * C4, PASS_B = (ITEMS_B + kD32Threads - 1) / kD32Threads;
  static_assert(ITEMS_B % kD32Threads == 0, "phase B items must fill whole passes");
  // phase-D slot: (c4, ox) + chunk (7 output rows each), chunk stride padded to whole warps
  constexpr int CH_STRIDE = (C4 * TW + 31) / 32 * 32, PASS_D = (2 * CH_STRIDE + kD32Threads - 1) / kD32Threads;

  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // D = F32, A = B = F16, K-major
  uint32_t tpar = 0;
  auto decode = [&](int tile, int& tx, int& ty, int& n) {
    const int r = p.magic_x ? (int)__umulhi((unsigned)tile, p.magic_x) : tile;
    tx = tile - r * tiles_x;
    n = p.magic_y ? (int)__umulhi((unsigned)r, p.magic_y) : r;
    ty = r - n * tiles_y;
  };
  // TMA fetches of a tile's low-res P halves (issued by thread 0): conv1 half -> sP, residual half -> buffer `rb`
  auto fetch_a = [&](int tile) {
    int tx, ty, n;
    decode(tile, tx, ty, n);
    mbar_expect_tx(&pbar[0], G::PHALF);
    tma_load_4d(sP, &tmP, &pbar[0], 0, tx * TW / 2 - 2, ty * TH / 2 - 2, n);
  };
  auto fetch_r = [&](int tile, int rb) {
    int tx, ty, n;
    decode(tile, tx, ty, n);
    mbar_expect_tx(&pbar[1 + rb], G::PHALF);
    tma_load_4d(sP + (1 + rb) * (G::PPITCH / 4), &tmP, &pbar[1 + rb], C, tx * TW / 2 - 2, ty * TH / 2 - 2, n);
  };
  if (tid == 0 && (int)blockIdx.x < total_tiles) { fetch_a(blockIdx.x); if (RESB == 2) fetch_r(blockIdx.x, 0); }

  int it = 0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
    int tx, ty, n;
    decode(tile, tx, ty, n);
    const int X0 = tx * TW, Y0 = ty * TH;                  // even
    const int px0 = X0 / 2 - 2, py0 = Y0 / 2 - 2;
    const int rb = RESB == 2 ? (it & 1) : 0;
    float* sPr = sP + (1 + rb) * (G::PPITCH / 4);          // this tile's residual half
    const bool border_tile = Y0 == 0 || Y0 + TH + 1 >= H || X0 == 0 || X0 + TW + 1 >= W;
    const bool clamp_tile = px0 < 0 || py0 < 0 || px0 + PW > p.w || py0 + PH > p.h;   // the low-res window leaves the map (zero-filled)
    // replicate the map's edge rows / columns into the zero-filled halo of a P half (= torch's index clamping)
    auto replicate = [&](float* half) {
      const int lo_x = max(0, -px0), hi_x = min(PW - 1, p.w - 1 - px0), lo_y = max(0, -py0), hi_y = min(PH - 1, p.h - 1 - py0);
#pragma unroll 1
      for (int i = tid; i < PH * PW * C4; i += kD32Threads) {
        const int c4 = i % C4, pp = i / C4;
        const int lj = pp % PW, li = pp / PW;
        const int cj = min(max(lj, lo_x), hi_x), ci = min(max(li, lo_y), hi_y);
        if (cj != lj || ci != li)
          *reinterpret_cast<float4*>(half + pp * C + c4 * 4) = *reinterpret_cast<const float4*>(half + (ci * PW + cj) * C + c4 * 4);
      }
    };

    // ---- A: the tile's conv1 half (and, RESB = 2, its residual half) were fetched during the previous tile ----
    mbar_wait_parity(&pbar[0], tpar);
    if (RESB == 2) mbar_wait_parity(&pbar[1 + rb], (uint32_t)(it >> 1) & 1u);
    if (clamp_tile) {
      replicate(sP);
      if (RESB == 2) replicate(sPr);
      __syncthreads();
    }

    // ---- B: b = SiLU(composite(P) + b1) -> fp16 hi / lo UMMA tiles ----
#pragma unroll 1
    for (int pass = 0; pass < PASS_B; ++pass) {
      const int item = tid + pass * kD32Threads;
      const int c4 = item % C4, cx = (item / C4) % (BW / 2), cls = item / (C4 * (BW / 2));
      const int pX = cls & 1, pY = cls >> 1;               // absolute parities of X and Y (cls = pY * 2 + pX)
      float4 wc[3][3];
#pragma unroll
      for (int k = 0; k < 9; ++k) wc[k / 3][k % 3] = *reinterpret_cast<const float4*>(sWc + (cls * 9 + k) * C + c4 * 4);
      const float4 bias = *reinterpret_cast<const float4*>(sB1 + c4 * 4);
      float4 acc[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[o] = bias;
      // output cy (b row by = 2 cy + 1 - pY) is centred on low-res tile row cy + 2 - pY; column cx on lj = cx + 2 - pX
      const float* base = sP + ((1 - pY) * PW + (cx + 1 - pX)) * C + c4 * 4;
#pragma unroll
      for (int rr = 0; rr < 10; ++rr) {
        const float* rp = base + rr * PW * C;
        const float4 v0 = *reinterpret_cast<const float4*>(rp), v1 = *reinterpret_cast<const float4*>(rp + C), v2 = *reinterpret_cast<const float4*>(rp + 2 * C);
#pragma unroll
        for (int di = 0; di < 3; ++di) {
          const int o = rr - di;
          if (o >= 0 && o < 8) { fma4(acc[o], v0, wc[di][0]); fma4(acc[o], v1, wc[di][1]); fma4(acc[o], v2, wc[di][2]); }
        }
      }
      const int bx = 2 * cx + 1 - pX;
      uint8_t* dh = sBh + (c4 >> 2) * G::KSTEP + ((c4 >> 1) & 1) * G::LBO + (c4 & 1) * 8;
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const int m = (2 * o + 1 - pY) * BW + bx;
        uint2 hi, lo;
        split4(silu4(acc[o]), hi, lo);
        *reinterpret_cast<uint2*>(dh + m * 16) = hi;
        *reinterpret_cast<uint2*>(dh + G::B_BYTES + m * 16) = lo;
      }
    }
    // image-border pixels: the reference zero-pads `a` (the composite extrapolates it) -> generic recomputation
    if (border_tile) {
      __syncthreads();
#pragma unroll 1
      for (int i = tid; i < NPIX * C4; i += kD32Threads) {
        const int c4 = i % C4, m = i / C4;
        const int by = m / BW, bx = m % BW;
        const int Y = Y0 - 1 + by, X = X0 - 1 + bx;
        if (Y < 0 || Y >= H || X < 0 || X >= W) continue;
        if (Y != 0 && Y != H - 1 && X != 0 && X != W - 1) continue;
        float4 acc = *reinterpret_cast<const float4*>(sB1 + c4 * 4);
#pragma unroll 1
        for (int r = 0; r < 3; ++r) {
          const int y = Y - 1 + r;
          if (y < 0 || y >= H) continue;
          const int ia = ((y + 1) >> 1) - 1 - py0;          // low-res tile rows (ia, ia + 1) with weights (wya, 1 - wya)
          const float wya = (y & 1) ? 0.75f : 0.25f;
#pragma unroll 1
          for (int s = 0; s < 3; ++s) {
            const int x = X - 1 + s;
            if (x < 0 || x >= W) continue;
            const int ja = ((x + 1) >> 1) - 1 - px0;
            const float wxa = (x & 1) ? 0.75f : 0.25f;
            const float* q = sP + (ia * PW + ja) * C + c4 * 4;
            const float4 t0 = lerp4(*reinterpret_cast<const float4*>(q), wxa, *reinterpret_cast<const float4*>(q + C), 1.f - wxa);
            const float4 t1 = lerp4(*reinterpret_cast<const float4*>(q + PW * C), wxa, *reinterpret_cast<const float4*>(q + PW * C + C), 1.f - wxa);
            const float4 a = lerp4(t0, wya, t1, 1.f - wya);
            fma4(acc, a, *reinterpret_cast<const float4*>(p.dw1 + (r * 3 + s) * C + c4 * 4));
          }
        }
        uint2 hi, lo;
        split4(silu4(acc), hi, lo);
        uint8_t* dh = sBh + (c4 >> 2) * G::KSTEP + ((c4 >> 1) & 1) * G::LBO + (c4 & 1) * 8 + m * 16;
        *reinterpret_cast<uint2*>(dh) = hi;
        *reinterpret_cast<uint2*>(dh + G::B_BYTES) = lo;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();

    // ---- C: c = W2 b (+ c2 in the epilogue): 3 MMAs per 16-channel k-step and 128-pixel row block ----
    if (tid == 0) {
      // every warp is past phase B of this tile and phase D of the previous one: the conv1 half and the OTHER residual
      // buffer are free -> fetch the next tile now, it lands while this tile finishes
      if (tile + (int)gridDim.x < total_tiles) { fetch_a(tile + gridDim.x); if (RESB == 2) fetch_r(tile + gridDim.x, rb ^ 1); }
      if (RESB == 1) fetch_r(tile, 0);                     // single residual buffer: this tile's half lands during phase C
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ah0 = s32(sBh), al0 = s32(sBl), w0 = s32(sW);
      // issue order: the same product term of ALL row blocks back to back, so consecutive MMAs never accumulate into the same
      // TMEM columns (a chain of dependent small MMAs costs ~190 cycles per link, independent ones ~65)
#pragma unroll
      for (int ks = 0; ks < C / 16; ++ks) {
        // weight pack: per Cin chunk of KC channels [hi: KC/8 groups][lo: KC/8 groups], group = [C][8] fp16
        const int kc = (ks * 16) / G::KC, kl = (ks * 16 % G::KC) / 8;
        const uint32_t wb = w0 + kc * (2 * G::KC * C * 2);
        const uint64_t bh = desc_nosw(wb + kl * (C * 16), C * 16, 128u);
        const uint64_t bl = desc_nosw(wb + (G::KC / 8 + kl) * (C * 16), C * 16, 128u);
#pragma unroll
        for (int term = 0; term < 3; ++term)
#pragma unroll
          for (int hb = 0; hb < NBLK; ++hb) {
            const uint64_t a = desc_nosw((term == 1 ? al0 : ah0) + ks * G::KSTEP + hb * 2048, G::LBO, 128u);
            umma_f16(tmem + hb * C, a, term == 2 ? bl : bh, idesc, (ks | term) ? 1u : 0u);
          }
      }
#pragma unroll
      for (int hb = 0; hb < NBLK; ++hb)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar[hb])) : "memory");
    }
    {
      const int q = warp & 3;                               // TMEM lane quarter of this warp
#pragma unroll 1
      for (int hb = warp >> 2; hb < NBLK; hb += kD32Threads / 128) {
        mbar_wait_parity(&bar[ALIAS ? NBLK - 1 : hb], tpar);      // ALIAS: the c tile overwrites b -> every MMA must be done
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int m = hb * 128 + q * 32 + lane;
        const int by = m / BW, bx = m % BW;
        const int Y = Y0 - 1 + by, X = X0 - 1 + bx;
        const bool inside = Y >= 0 && Y < H && X >= 0 && X < W;
#pragma unroll
        for (int c0 = 0; c0 < C; c0 += 16) {
          uint32_t v[16];
          ld16(tmem + ((uint32_t)(q * 32) << 16) + hb * C + c0, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(sB2 + c0 + 4 * j);
            float4 o;
            o.x = inside ? fmaf(__uint_as_float(v[4 * j]), p.w_unscale, b4.x) : 0.f;
            o.y = inside ? fmaf(__uint_as_float(v[4 * j + 1]), p.w_unscale, b4.y) : 0.f;
            o.z = inside ? fmaf(__uint_as_float(v[4 * j + 2]), p.w_unscale, b4.z) : 0.f;
            o.w = inside ? fmaf(__uint_as_float(v[4 * j + 3]), p.w_unscale, b4.w) : 0.f;
            *reinterpret_cast<float4*>(sC + (c0 / 4 + j) * G::CPL + m * 16) = o;
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (RESB == 1) {
      mbar_wait_parity(&pbar[1], tpar);
      if (clamp_tile) { replicate(sPr); __syncthreads(); }
    }
    // ---- D: d = SiLU(DW2(c) + b3) + up2(P[:, C:]) -> store / head ----
#pragma unroll 1
    for (int pass = 0; pass < PASS_D; ++pass) {
      const int slot = tid + pass * kD32Threads;
      const int chunk = slot / CH_STRIDE, it = slot % CH_STRIDE;
      const int c4 = it % C4, ox = it / C4;
      const bool active = chunk < 2 && ox < TW;
      float4 acc[7];
      if (active) {
        {
          float4 wk[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) wk[k] = *reinterpret_cast<const float4*>(sWk + k * C + c4 * 4);
          const float4 bias = *reinterpret_cast<const float4*>(sB3 + c4 * 4);
#pragma unroll
          for (int o = 0; o < 7; ++o) acc[o] = bias;
          // output (oy, ox) <-> c rows oy .. oy + 2, columns ox .. ox + 2 of the 16 x BW tile
          const uint8_t* cp = sC + c4 * G::CPL + (chunk * 7 * BW + ox) * 16;
#pragma unroll
          for (int ir = 0; ir < 9; ++ir) {
            const float4 v0 = *reinterpret_cast<const float4*>(cp + ir * BW * 16), v1 = *reinterpret_cast<const float4*>(cp + ir * BW * 16 + 16),
                         v2 = *reinterpret_cast<const float4*>(cp + ir * BW * 16 + 32);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              const int o = ir - r;
              if (o >= 0 && o < 7) { fma4(acc[o], v0, wk[r * 3]); fma4(acc[o], v1, wk[r * 3 + 1]); fma4(acc[o], v2, wk[r * 3 + 2]); }
            }
          }
        }
#pragma unroll
        for (int o = 0; o < 7; ++o) acc[o] = silu4(acc[o]);
        // residual: x2 bilinear of P[:, C:].  Column pair + weights from the parity of X (= parity of ox, X0 is even); the 7
        // rows of a chunk need 5 low-res rows: chunk 0 -> tile rows 1..5, chunk 1 -> 5..9; lerp each once horizontally.
        const int lj = (ox >> 1) + 2;
        const int ja = (ox & 1) ? lj : lj - 1;
        const float wa = (ox & 1) ? 0.75f : 0.25f;
        const float* rp0 = sPr + ((chunk * 4 + 1) * PW + ja) * C + c4 * 4;
        float4 hrow[5];
#pragma unroll
        for (int k = 0; k < 5; ++k)
          hrow[k] = lerp4(*reinterpret_cast<const float4*>(rp0 + k * PW * C), wa, *reinterpret_cast<const float4*>(rp0 + k * PW * C + C), 1.f - wa);
        auto add4 = [](float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; };
        if (chunk == 0) {
          // oy = o.  o even: low-res tile rows (o/2 + 1, o/2 + 2) = hrow (o/2, o/2 + 1), weights (.25, .75); o odd: hrow ((o+1)/2, +1), (.75, .25)
#pragma unroll
          for (int o = 0; o < 7; ++o)
            add4(acc[o], (o & 1) ? lerp4(hrow[(o + 1) / 2], 0.75f, hrow[(o + 1) / 2 + 1], 0.25f) : lerp4(hrow[o / 2], 0.25f, hrow[o / 2 + 1], 0.75f));
        } else {
          // oy = 7 + o.  oy odd (o even): tile rows (o/2 + 5, o/2 + 6) = hrow (o/2, o/2 + 1), weights (.75, .25)
          //              oy even (o odd): tile rows ((o+7)/2 + 1, + 2) = hrow ((o+7)/2 - 4, - 3), weights (.25, .75)
#pragma unroll
          for (int o = 0; o < 7; ++o)
            add4(acc[o], (o & 1) ? lerp4(hrow[(o + 7) / 2 - 4], 0.25f, hrow[(o + 7) / 2 - 3], 0.75f) : lerp4(hrow[o / 2], 0.75f, hrow[o / 2 + 1], 0.25f));
        }
      }
      const int X = X0 + ox;
      if (HEAD) {
        // 1x1 head: each of the C4 = 4 lanes of a pixel column holds the partial dot products of its 4 channels for 7 rows.
        // Butterfly exchange over the lane quad (6 shuffles instead of 14): after the xor-1 step a lane keeps 4 rows, after
        // the xor-2 step 2 rows, fully reduced -- and the stores are spread over all four lanes.
        static_assert(!HEAD || C4 == 4, "the head reduction is written for 16 channels");
        const float4 wo = *reinterpret_cast<const float4*>(sWo + c4 * 4);
        float part[8];
#pragma unroll
        for (int o = 0; o < 7; ++o) part[o] = active ? acc[o].x * wo.x + acc[o].y * wo.y + acc[o].z * wo.z + acc[o].w * wo.w : 0.f;
        part[7] = 0.f;
        const bool b0 = (lane & 1) != 0, b1 = (lane & 2) != 0;
        float q4[4], q2[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) q4[i] = (b0 ? part[i + 4] : part[i]) + __shfl_xor_sync(0xffffffffu, b0 ? part[i] : part[i + 4], 1);
#pragma unroll
        for (int i = 0; i < 2; ++i) q2[i] = (b1 ? q4[i + 2] : q4[i]) + __shfl_xor_sync(0xffffffffu, b1 ? q4[i] : q4[i + 2], 2);
        const int r0 = (b0 ? 4 : 0) + (b1 ? 2 : 0);          // this lane's rows r0, r0 + 1 of the chunk
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int Y = Y0 + chunk * 7 + r0 + i;
          if (active && r0 + i < 7 && Y < H && X < W) reinterpret_cast<float*>(p.out)[((size_t)n * H + Y) * W + X] = q2[i] + sWo[C];
        }
      } else if (active) {
#pragma unroll
        for (int o = 0; o < 7; ++o) {
          const int Y = Y0 + chunk * 7 + o;
          if (Y < H && X < W)
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (((size_t)n * H + Y) * W + X) * p.out_cs + c4 * 4) = acc[o];
        }
      }
    }
    // no barrier here: the next tile's phase B touches sP (conv1 half, refetched), sBh / sBl (their MMAs are complete) and
    // the other residual buffer only; sC and the accumulators are rewritten after the barrier that ends that phase B
    if (ALIAS) __syncthreads();                             // ... unless the c tile lies over the b tiles
    tpar ^= 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)G::TCOLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFnD)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnD get_encode_d() {
  static EncodeTiledFnD fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qres) == cudaSuccess && q) fn = (EncodeTiledFnD)q;
  }
  return fn;
}

template <int C, int BW, int NT, int RESB, bool ALIAS, bool HEAD>
static void dlc32_launch(Dlc32P p, cudaStream_t s) {
  using G = Geo<C, BW, RESB, ALIAS>;
  static unsigned long long attr_done = 0;
  ensure_dyn_smem(dlc32_kernel<C, BW, NT, RESB, ALIAS, HEAD>, G::SMEM, attr_done, "dlc32_kernel");
  const int H = 2 * p.h, W = 2 * p.w;
  const unsigned txs = (W + G::TW - 1) / G::TW, tys = (H + TH - 1) / TH;
  p.magic_x = txs > 1 ? (unsigned)((0x100000000ull + txs - 1) / txs) : 0u;
  p.magic_y = tys > 1 ? (unsigned)((0x100000000ull + tys - 1) / tys) : 0u;
  // low-res P as a 4-D tensor {2C channels, w, h, N}; box = one channel half of a tile's low-res window (zero fill outside)
  CUtensorMap tm;
  {
    const cuuint64_t px = (cuuint64_t)p.p_cs * 4;
    cuuint64_t dims[4] = {(cuuint64_t)(2 * C), (cuuint64_t)p.w, (cuuint64_t)p.h, (cuuint64_t)p.N};
    cuuint64_t strides[3] = {px, px * p.w, px * p.w * p.h};
    cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)G::PW, (cuuint32_t)PH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = get_encode_d()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(p.P), dims, strides, box, es,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "libysp: cuTensorMapEncodeTiled(dlc32 P) failed: %d\n", (int)r); return; }
  }
  const int tiles = (int)(txs * tys) * p.N;
  static int sms = 0;
  if (!sms) { int d = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d); if (sms <= 0) sms = 148; }
  int per_sm = (int)((227 * 1024) / (G::SMEM + 1024));
  if (per_sm > 512 / NT) per_sm = 512 / NT;
  if (per_sm < 1) per_sm = 1;
  const int ctas = per_sm * sms;
  launch_pdl(dlc32_kernel<C, BW, NT, RESB, ALIAS, HEAD>, dim3(tiles < ctas ? tiles : ctas), dim3(NT), G::SMEM, s, tm, p);
}

bool dlc32_supported(int C, bool head) { return get_encode_d() != nullptr && ((C == 16 && head) || ((C == 32 || C == 64) && !head)); }

void launch_dlc32(const Dlc32P& p, cudaStream_t s) {
  // Measured at B = 256 (tools/profile_layers.py tc32, profiles/r7_*): with CTA-wide barriers between the phases, small CTAs win
  // where they fit -- stage 4 (C = 16): four 128-thread CTAs per SM on 14 x 14 tiles 1.09 ms, two 256-thread CTAs on 14 x 30
  // tiles 2.12 ms (ncu: 28 % of the stall samples were barrier waits); stage 3 (C = 32): two 256-thread CTAs 0.67 ms, three
  // 128-thread CTAs with the c tile over the b tiles 0.75 ms, four 128-thread CTAs on 14 x 6 tiles 0.665 ms (YSP_DLC32_CFG=2);
  // stage 1 (C = 64): 14 x 14 tiles need 157 KB = ONE CTA per SM, whose phases then run strictly one after the other (512
  // threads 0.57 ms, 256 threads 0.67 ms); 14 x 6 tiles fit twice (102 KB) and the two CTAs cover each other's barriers:
  // 0.43 ms although the halo overhead grows from 1.31x to 1.52x.  YSP_DLC32_CFG=1 / 3 select the alternatives (A/B timing only).
  static const int cfg = getenv("YSP_DLC32_CFG") ? atoi(getenv("YSP_DLC32_CFG")) : 0;
  if (p.C == 16) {
    if (cfg == 1) dlc32_launch<16, 32, 256, 2, false, true>(p, s);
    else dlc32_launch<16, 16, 128, 1, false, true>(p, s);
  } else if (p.C == 32) {
    if (cfg == 1) dlc32_launch<32, 16, 128, 1, true, false>(p, s);
    else if (cfg == 2) dlc32_launch<32, 8, 128, 1, true, false>(p, s);      // 14 x 6 tiles: four CTAs per SM
    else dlc32_launch<32, 16, 256, 1, false, false>(p, s);
  } else {
    if (cfg == 1) dlc32_launch<64, 16, 256, 1, true, false>(p, s);
    else if (cfg == 3) dlc32_launch<64, 16, 512, 1, true, false>(p, s);
    else dlc32_launch<64, 8, 256, 1, true, false>(p, s);                    // 14 x 6 tiles: 102 KB -> TWO CTAs per SM
  }
}

}  // namespace ysp
