// kernels_dlc_tc.cu -- decoder stage (bilinear x2 + DoubleLightConv [+ 1x1 mask head]) on tcgen05 tensor cores.
//
// The reference stage (YOLOSegPlusPlus.py:33-58 behind nn.Upsample, :155-175) is
//     u = up2(x);  b = SiLU(DW1(W1 u + c1) + b1);  d = SiLU(DW2(W2 b + c2) + b3);  out = d + (Wr u + cr)   [; logit = wo.out + bo]
// Both "pointwise then depthwise" pairs are compositions of linear maps, i.e. ONE dense 3x3 convolution with rank-1
// weights  Weff[tap][co][k] = dw[tap][co] * W[co][k],  and up2 commutes with the pointwise biases (bilinear weights sum
// to 1).  So per 14 x 30 output tile this kernel
//   1. builds u = up2(x) (bf16, zero outside the image) in shared memory from a low-res x tile    (CUDA cores, bf16x2 math)
//   2. conv1: 9 taps x Cin/16 tcgen05.mma (M=128, N=C) per 8-pixel-wide column block, + the residual 1x1 as one more tap
//   3. epilogue 1: TMEM -> + position-dependent bias -> SiLU -> bf16 b tile in shared memory
//   4. conv2: 9 taps x C/16 MMAs on the b tile;  5. epilogue 2: SiLU, + residual, head dot product / NHWC store.
// The MMA A operands are SHIFTED VIEWS of one shared-memory tile: activations are stored channel-group-major
// ([8-channel plane][row][col] x 16 B), which is the canonical K-major no-swizzle UMMA layout when a core matrix
// (8 M-rows x 16 B) is 8 horizontally adjacent pixels; a tap (r,s) only moves the descriptor start address by
// (r*pitch + s) * 16 B, SBO = row pitch, LBO = plane pitch.  No im2col, no TMA needed: the tile never leaves the SM.
// The pointwise bias counts only for taps inside the image (the reference zero-pads the pointwise OUTPUT), hence the
// 9 position classes of effective bias (3 row cases x 3 column cases).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "kernels.h"

namespace ysp {

namespace {

// Tile geometry for NBLK column blocks of 8 pixels (NBLK = 4: 14 x 30 outputs, 320 threads; NBLK = 2: 14 x 14 outputs, 160 threads
// and twice as many co-resident CTAs -- the CTA-wide barriers between the phases then synchronise half as many warps)
constexpr int TH = 14;                     // output rows per tile
constexpr int AR = 18;                     // u / b tile rows
constexpr int XR = 11;                     // low-res x tile rows incl. halo
__host__ __device__ constexpr int tw_of(int nblk) { return 8 * nblk - 2; }       // output columns
__host__ __device__ constexpr int ap_of(int nblk) { return 8 * nblk + 2; }       // u / b tile pitch (pixels)
__host__ __device__ constexpr int xc_of(int nblk) { return tw_of(nblk) / 2 + 4; } // low-res x tile columns
// Bytes of one 8-channel plane of the u / b tile (= the descriptors' LBO).  With 8 planes (CIN = 64) the pitch is padded by
// 16 B so that the 8 lanes of a quarter-warp, which write the 8 planes of one pixel in phase 1, hit 8 different 16-byte
// bank groups (unpadded: 9792 = 64 mod 128 -> 4-way conflicts; ncu r6: 22.6 M store wavefronts for 5.6 M ideal).
__host__ __device__ constexpr int plane_bytes(int cin, int nblk) { return AR * ap_of(nblk) * 16 + (cin == 64 ? 16 : 0); }

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Descriptor = {lo: start>>4 | (LBO>>4)<<16, hi: SBO>>4 | version<<14}.  Only the start address changes between the MMAs
// of a tile, by compile-time byte offsets: the issuing thread adds (offset >> 4) to the low word -- one IADD per operand.
__device__ __forceinline__ void umma2(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, bool acc) {
  if (acc)
    asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, 1, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                 ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, 0, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                 ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DN;\n\tbra WL;\n\tDN:\n\t}"
               ::"r"(s32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ float silu_th(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
// Packed fp32 (FADD2 / FMUL2 / FFMA2, sm_100): the epilogues handle two adjacent channels per instruction; the operation
// order per channel is the scalar one, so results are bit-identical.
__device__ __forceinline__ float2 f2(const uint32_t (&r)[16], int i) { return make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])); }
__device__ __forceinline__ float2 silu_th2(float2 x) {
  const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
  return __ffma2_rn(h, t, h);
}
__device__ __forceinline__ __nv_bfloat162 lerp2(__nv_bfloat162 a, __nv_bfloat162 wa, __nv_bfloat162 b, __nv_bfloat162 wb) {
  return __hfma2(a, wa, __hmul2(b, wb));
}
__device__ __forceinline__ uint4 lerp8(const uint4& a, __nv_bfloat162 wa, const uint4& b, __nv_bfloat162 wb) {
  uint4 o;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) po[i] = lerp2(pa[i], wa, pb[i], wb);
  return o;
}

}  // namespace

// Timing probes (drop MMAs / arithmetic, results WRONG) exist only in -DYSP_PROBES builds: the shipped library cannot skip work.
#ifdef YSP_PROBES
#define PROBE(bit) ((p.probe & (bit)) != 0)
#else
#define PROBE(bit) false
#endif

// warps 0 .. 2 NBLK - 1: CUDA-core phases + epilogues; then ISSUERS MMA-issue warps (a whole number of warps: 320 / 160 threads)
__host__ __device__ constexpr int dlc_threads(int nblk) { return 32 * (2 * nblk + (nblk == 4 ? 2 : 1)); }

template <int CIN, int C, bool HEAD, int NBLK>
__global__ void __launch_bounds__(dlc_threads(NBLK), (C <= 16 ? 2 : 1) * (4 / NBLK)) dlc_tc_kernel(DlcTcP p) {
  constexpr int kDlcThreads = dlc_threads(NBLK), EW = 2 * NBLK;      // EW = epilogue / CUDA-core warps
  constexpr int TW = tw_of(NBLK), AP = ap_of(NBLK), XC = xc_of(NBLK);
  constexpr int KP1 = CIN / 8, KP2 = C / 8;          // 8-channel planes of u and of b
  constexpr int PLANE = plane_bytes(CIN, NBLK);
  constexpr int W1B = 9 * KP1 * C * 16, W2B = 9 * KP2 * C * 16, WRB = KP1 * C * 16;
  extern __shared__ __align__(128) uint8_t dsm[];
  uint8_t* sU = dsm;                                 // [KP1][AR][AP] x 16 B
  uint8_t* sB = sU + KP1 * PLANE;                    // [KP2][AR][AP] x 16 B
  uint8_t* sX = sB + KP2 * PLANE;                    // [XR*XC][CIN] bf16 (edge-clamped low-res tile)
  uint8_t* sW1 = sX + XR * XC * CIN * 2;             // [9][KP1][C][8] bf16
  uint8_t* sW2 = sW1 + W1B;                          // [9][KP2][C][8]
  uint8_t* sWr = sW2 + W2B;                          // [KP1][C][8]
  float* sBe1 = reinterpret_cast<float*>(sWr + WRB); // [9][C]
  float* sBe2 = sBe1 + 9 * C;                        // [9][C]
  float* sCr = sBe2 + 9 * C;                         // [C] residual bias
  float* sWo = sCr + C;                              // [C] head weights
  __shared__ __align__(8) uint64_t bar1[4], bar2[4];   // per column block: conv1(+residual) done / conv2 done
  __shared__ uint32_t tmem_s;
  // TMEM columns: D1[h][r] (h = column block, r = tap row: three independent accumulation chains, summed in the epilogue --
  // a chain of tiny dependent MMAs costs ~190 cycles per link) at (3h + r) * C, re-used for D2[h][r]; Dr[h] at (3 NBLK + h) * C
  constexpr uint32_t TCOLS = 4 * NBLK * C <= 32 ? 32 : (4 * NBLK * C <= 64 ? 64 : (4 * NBLK * C <= 128 ? 128 : (4 * NBLK * C <= 256 ? 256 : 512)));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = 2 * p.h, W = 2 * p.w;
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  const int total_tiles = tiles_x * tiles_y * p.N;
  // tile -> (tx, ty, n) without runtime division (every thread decodes three tiles per iteration)
  auto decode = [&](int t, int& tx, int& ty, int& n) {
    const int r = p.magic_x ? (int)__umulhi((unsigned)t, p.magic_x) : t;        // t / tiles_x
    tx = t - r * tiles_x;
    n = p.magic_y ? (int)__umulhi((unsigned)r, p.magic_y) : r;                  // r / tiles_y
    ty = r - n * tiles_y;
  };

  // ---- prologue: constants only (overlaps the previous kernel under PDL) ----
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar1[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar2[i])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_s)), "r"(TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  {
    const uint4* g1 = reinterpret_cast<const uint4*>(p.wpack);
    uint4* d1 = reinterpret_cast<uint4*>(sW1);
    for (int i = tid; i < (W1B + W2B + WRB) / 16; i += kDlcThreads) d1[i] = g1[i];
    const float* gb = reinterpret_cast<const float*>(p.wpack + W1B + W2B + WRB);
    for (int i = tid; i < 18 * C + 2 * C; i += kDlcThreads) sBe1[i] = gb[i];
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t a_hi = ((uint32_t)(AP * 16) >> 4) | (1u << 14);          // SBO = row pitch, version 1
  const uint32_t b_hi = (128u >> 4) | (1u << 14);                          // SBO = 8 rows x 16 B
  auto prefetch_x = [&](int tl) {
    if (tl >= total_tiles) return;
    int tx, ty, n;
    decode(tl, tx, ty, n);
    const int px0 = tx * TW / 2 - 2, py0 = ty * TH / 2 - 2;
    const bf16* xg = reinterpret_cast<const bf16*>(p.x);
    for (int i = tid; i < XR * XC * KP1; i += kDlcThreads) {
      const int kc = i % KP1, pp = i / KP1;
      const int gx = min(max(px0 + pp % XC, 0), p.w - 1), gy = min(max(py0 + pp / XC, 0), p.h - 1);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(sX + (pp * CIN + kc * 8) * 2)),
                   "l"(xg + ((size_t)(n * p.h + gy) * p.w + gx) * p.x_cs + kc * 8) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // phase 1 of a tile: u = up2(x) on the 18 x 34 tile, one thread per (2x2 hi-res block, 8 channels); zero outside the image
  auto build_u = [&](int tl) {
    int tx, ty, n_unused;
    decode(tl, tx, ty, n_unused);
    const int px0 = tx * TW / 2 - 2, py0 = ty * TH / 2 - 2;
    const __nv_bfloat162 q25 = __floats2bfloat162_rn(0.25f, 0.25f), q75 = __floats2bfloat162_rn(0.75f, 0.75f);
#pragma unroll 2
    for (int i = tid; i < (AR / 2) * (AP / 2) * KP1; i += kDlcThreads) {
      const int kc = i % KP1, bb = i / KP1;
      const int bi = bb / (AP / 2), bj = bb % (AP / 2);
      const int li = bi + 1, lj = bj + 1;
      const int gi = py0 + li, gj = px0 + lj;
      uint4 o00, o01, o10, o11;
      if (gi >= 0 && gi < p.h && gj >= 0 && gj < p.w && !PROBE(4)) {
        const uint8_t* c = sX + ((li * XC + lj) * CIN + kc * 8) * 2;
        constexpr int RS = XC * CIN * 2, PS = CIN * 2;
        const uint4 m0 = *reinterpret_cast<const uint4*>(c - RS - PS), m1 = *reinterpret_cast<const uint4*>(c - RS), m2 = *reinterpret_cast<const uint4*>(c - RS + PS);
        const uint4 z0 = *reinterpret_cast<const uint4*>(c - PS), z1 = *reinterpret_cast<const uint4*>(c), z2 = *reinterpret_cast<const uint4*>(c + PS);
        const uint4 w0 = *reinterpret_cast<const uint4*>(c + RS - PS), w1 = *reinterpret_cast<const uint4*>(c + RS), w2 = *reinterpret_cast<const uint4*>(c + RS + PS);
        const uint4 t0 = lerp8(m0, q25, z0, q75), t1 = lerp8(m1, q25, z1, q75), t2 = lerp8(m2, q25, z2, q75);   // even row
        const uint4 u0 = lerp8(z0, q75, w0, q25), u1 = lerp8(z1, q75, w1, q25), u2 = lerp8(z2, q75, w2, q25);   // odd row
        o00 = lerp8(t0, q25, t1, q75); o01 = lerp8(t1, q75, t2, q25);
        o10 = lerp8(u0, q25, u1, q75); o11 = lerp8(u1, q75, u2, q25);
      } else {
        o00 = o01 = o10 = o11 = make_uint4(0u, 0u, 0u, 0u);
      }
      uint8_t* d = sU + kc * PLANE + ((2 * bi) * AP + 2 * bj) * 16;
      if (KP1 == 4) {
        // 4 planes: a quarter-warp is (4 planes) x (2 blocks) = bank groups {0,4} + {0,2}.  Planes 2,3 store the right
        // pixel first, which adds the odd groups: 8 lanes -> 8 distinct groups (was a 2-way conflict).
        const bool sw = (kc & 2) != 0;
        const int o = sw ? 16 : 0;
        *reinterpret_cast<uint4*>(d + o) = sw ? o01 : o00; *reinterpret_cast<uint4*>(d + 16 - o) = sw ? o00 : o01;
        *reinterpret_cast<uint4*>(d + AP * 16 + o) = sw ? o11 : o10; *reinterpret_cast<uint4*>(d + AP * 16 + 16 - o) = sw ? o10 : o11;
      } else {
        *reinterpret_cast<uint4*>(d) = o00; *reinterpret_cast<uint4*>(d + 16) = o01;
        *reinterpret_cast<uint4*>(d + AP * 16) = o10; *reinterpret_cast<uint4*>(d + AP * 16 + 16) = o11;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the tensor core
  };
  // Software pipeline over the CTA's tiles: the u tile of tile i+1 is built while the conv2 MMAs of tile i run (sU is
  // only read by conv1 + residual, all complete by then; sX holds x(i+1), prefetched during tile i's first half).
  prefetch_x(blockIdx.x);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if ((int)blockIdx.x < total_tiles) build_u(blockIdx.x);
  __syncthreads();
  prefetch_x(blockIdx.x + gridDim.x);
  // persistent CTA: weights, TMEM and the barrier are set up once; tiles are strided over the grid
  uint32_t tpar = 0;                                                   // mbarrier phase parity: every barrier completes once per tile
#pragma unroll 1
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
  int tx, ty, n;
  decode(tile, tx, ty, n);
  const int X0 = tx * TW, Y0 = ty * TH;

  // ---- phase 2: conv1 (D1[h][r], h = 8-pixel column block of the 16 x 32 b region) and the residual 1x1 (Dr[h]),
  //      issued block by block by warp 8 with one commit per block, so the epilogue of block h (warps 0-7 below)
  //      overlaps the MMAs of blocks h+1..  All descriptors are base + compile-time offset. ----
  constexpr int ISSUERS = (C <= 16 || NBLK < 4) ? 1 : 2;   // wide layers need two issuing threads to keep the tensor pipe fed
  if (warp >= EW && warp < EW + ISSUERS && lane == 0) {
#pragma unroll
    for (int h = warp - EW; h < NBLK; h += ISSUERS) {
      const uint32_t a_lo = ((s32(sU) & 0x3FFFF) >> 4) + (uint32_t)(8 * h) + (((uint32_t)PLANE >> 4) << 16);
      const uint32_t w_lo = ((s32(sW1) & 0x3FFFF) >> 4) + (((uint32_t)(C * 16) >> 4) << 16);
      const uint32_t r_lo = ((s32(sWr) & 0x3FFFF) >> 4) + (((uint32_t)(C * 16) >> 4) << 16);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        if (PROBE(1) && tap % 3) continue;             // timing probe (wrong results; -DYSP_PROBES builds only)
#pragma unroll
        for (int ks = 0; ks < CIN / 16; ++ks)
          umma2(tmem + (3 * h + tap / 3) * C, a_lo + (((2 * ks) * PLANE) >> 4) + (tap / 3) * AP + (tap % 3), a_hi,
                w_lo + (((tap * KP1 + 2 * ks) * C * 16) >> 4), b_hi, idesc, ((tap % 3) | ks) != 0);
      }
#pragma unroll
      for (int ks = 0; ks < CIN / 16; ++ks)            // residual: output pixel (oy, ox) <-> u(oy + 2, ox + 2)
        umma2(tmem + (3 * NBLK + h) * C, a_lo + (((2 * ks) * PLANE) >> 4) + 2 * AP + 2, a_hi, r_lo + (((2 * ks) * C * 16) >> 4), b_hi, idesc, ks != 0);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar1[h])) : "memory");
    }
  }

  // ---- phase 3: epilogue 1: b = SiLU(D1 + bias_eff1), zero outside the image, bf16 into the b tile ----
  const int q = warp & 3;                             // TMEM lane quarter this warp may read
  const int row = q * 32 + lane;                      // M row: (by = row / 8, bxl = row % 8)
  const int by = row >> 3, bxl = row & 7;
#pragma unroll 1
  for (int hh = 0; hh < 2 && warp < EW; ++hh) {
    const int h = (warp >> 2) + (EW / 4) * hh;        // blocks complete in order: NBLK = 4: warps 0-3 take 0 and 2, warps 4-7 take 1 and 3
    mbar_wait_parity(&bar1[h], tpar);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int bx = 8 * h + bxl;
    const int Y = Y0 - 1 + by, X = X0 - 1 + bx;
    const bool inside = Y >= 0 && Y < H && X >= 0 && X < W;
    const int cls = (Y <= 0 ? 0 : (Y >= H - 1 ? 2 : 1)) * 3 + (X <= 0 ? 0 : (X >= W - 1 ? 2 : 1));
    const float* be = sBe1 + cls * C;
#pragma unroll
    for (int c0 = 0; c0 < C; c0 += 16) {
      uint32_t v[16], v1[16], v2[16];
      ld16(tmem + ((uint32_t)(q * 32) << 16) + (3 * h) * C + c0, v);
      ld16(tmem + ((uint32_t)(q * 32) << 16) + (3 * h + 1) * C + c0, v1);
      ld16(tmem + ((uint32_t)(q * 32) << 16) + (3 * h + 2) * C + c0, v2);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      uint32_t w[8];
      float bb[16];                                   // 16-byte shared loads (the scalar form was 16 LDS per thread)
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(&bb[4 * j]) = *reinterpret_cast<const float4*>(be + c0 + 4 * j);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 sm = __fadd2_rn(__fadd2_rn(f2(v, 2 * j), f2(v1, 2 * j)), f2(v2, 2 * j));
        const float2 ff = PROBE(8) ? sm : silu_th2(__fadd2_rn(sm, make_float2(bb[2 * j], bb[2 * j + 1])));
        __nv_bfloat162 hv = __floats2bfloat162_rn(inside ? ff.x : 0.f, inside ? ff.y : 0.f);
        w[j] = *reinterpret_cast<uint32_t*>(&hv);
      }
      uint8_t* d = sB + (c0 / 8) * PLANE + (by * AP + bx) * 16;
      *reinterpret_cast<uint4*>(d) = make_uint4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<uint4*>(d + PLANE) = make_uint4(w[4], w[5], w[6], w[7]);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");             // x(i+1) has landed (this thread's copies; the barrier publishes them)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();

  // ---- phase 4: conv2 on the b tile: D2[h][r], output pixel (oy, 8h + oxl) <-> b(oy + r, 8h + oxl + s) ----
  if (warp >= EW && warp < EW + ISSUERS && lane == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int h = warp - EW; h < NBLK; h += ISSUERS) {
      const uint32_t a_lo = ((s32(sB) & 0x3FFFF) >> 4) + (uint32_t)(8 * h) + (((uint32_t)PLANE >> 4) << 16);
      const uint32_t w_lo = ((s32(sW2) & 0x3FFFF) >> 4) + (((uint32_t)(C * 16) >> 4) << 16);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        if (PROBE(2) && tap % 3) continue;
#pragma unroll
        for (int ks = 0; ks < C / 16; ++ks)
          umma2(tmem + (3 * h + tap / 3) * C, a_lo + (((2 * ks) * PLANE) >> 4) + (tap / 3) * AP + (tap % 3), a_hi,
                w_lo + (((tap * KP2 + 2 * ks) * C * 16) >> 4), b_hi, idesc, ((tap % 3) | ks) != 0);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar2[h])) : "memory");
    }
  }

  // ---- phase 1 of the NEXT tile, in the shadow of the conv2 MMAs ----
  if (tile + (int)gridDim.x < total_tiles) build_u(tile + gridDim.x);

  // ---- phase 5: epilogue 2: out = SiLU(D2 + bias_eff2) + Dr + cr;  head: logit = wo . out + bo ----
  const int oy = by, oxl = bxl;
#pragma unroll 1
  for (int hh = 0; hh < 2 && warp < EW; ++hh) {
    const int h = (warp >> 2) + (EW / 4) * hh;
    mbar_wait_parity(&bar2[h], tpar);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int ox = 8 * h + oxl;
    const int Y = Y0 + oy, X = X0 + ox;
    const bool valid = oy < TH && ox < TW && Y < H && X < W;
    const int cls = (Y <= 0 ? 0 : (Y >= H - 1 ? 2 : 1)) * 3 + (X <= 0 ? 0 : (X >= W - 1 ? 2 : 1));
    const float* be = sBe2 + cls * C;
    float part = 0.f;
#pragma unroll
    for (int c0 = 0; c0 < C; c0 += 16) {
      uint32_t v[16], v1[16], v2[16], rsd[16];
      ld16(tmem + ((uint32_t)(q * 32) << 16) + (3 * h) * C + c0, v);
      ld16(tmem + ((uint32_t)(q * 32) << 16) + (3 * h + 1) * C + c0, v1);
      ld16(tmem + ((uint32_t)(q * 32) << 16) + (3 * h + 2) * C + c0, v2);
      ld16(tmem + ((uint32_t)(q * 32) << 16) + (3 * NBLK + h) * C + c0, rsd);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float f[16];
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 b4 = *reinterpret_cast<const float4*>(be + c0 + 4 * j4), c4 = *reinterpret_cast<const float4*>(sCr + c0 + 4 * j4);
        const float bq[4] = {b4.x, b4.y, b4.z, b4.w}, cq[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int jj = 0; jj < 4; jj += 2) {
          const int j = 4 * j4 + jj;
          const float2 sm = __fadd2_rn(__fadd2_rn(__fadd2_rn(f2(v, j), f2(v1, j)), f2(v2, j)), make_float2(bq[jj], bq[jj + 1]));
          const float2 o2 = __fadd2_rn(__fadd2_rn(silu_th2(sm), f2(rsd, j)), make_float2(cq[jj], cq[jj + 1]));
          f[j] = o2.x; f[j + 1] = o2.y;
        }
      }
      if (HEAD) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 w4 = *reinterpret_cast<const float4*>(sWo + c0 + 4 * j4);
          part = fmaf(f[4 * j4], w4.x, part); part = fmaf(f[4 * j4 + 1], w4.y, part);
          part = fmaf(f[4 * j4 + 2], w4.z, part); part = fmaf(f[4 * j4 + 3], w4.w, part);
        }
      } else if (valid) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __nv_bfloat162 hv = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
          w[j] = *reinterpret_cast<uint32_t*>(&hv);
        }
        uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + (((size_t)n * H + Y) * W + X) * p.out_cs + c0);
        op[0] = make_uint4(w[0], w[1], w[2], w[3]);
        op[1] = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
    if (HEAD && valid) reinterpret_cast<float*>(p.out)[((size_t)n * H + Y) * W + X] = part + p.bo[0];
  }
  // the next tile overwrites sX / sU / sB and the TMEM accumulators: every warp must be done reading them
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  prefetch_x(tile + 2 * gridDim.x);                                  // sX is free: every thread is past build_u(i+1)
  tpar ^= 1;
  }  // tile loop
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS) : "memory");
  }
}

// ---- one-time weight preparation (device): the packed buffer is the exact shared-memory image of the kernel ----------
//   [W1eff 9*KP1*C*8 bf16][W2eff 9*KP2*C*8][Wr KP1*C*8][be1 9*C f32][be2 9*C][cr C][wo C]
__global__ void dlc_tc_prepare_kernel(DlcTcPrep q, uint8_t* out) {
  const int CIN = q.Cin, C = q.C, KP1 = CIN / 8, KP2 = C / 8;
  bf16* w1 = reinterpret_cast<bf16*>(out);
  bf16* w2 = w1 + 9 * KP1 * C * 8;
  bf16* wr = w2 + 9 * KP2 * C * 8;
  float* fb = reinterpret_cast<float*>(wr + KP1 * C * 8);
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = t0; i < 9 * KP1 * C * 8; i += stride) {
    const int j = i % 8, co = (i / 8) % C, kc = (i / (8 * C)) % KP1, tap = i / (8 * C * KP1);
    w1[i] = __float2bfloat16_rn(q.dw1[tap * C + co] * q.w1[(size_t)(kc * 8 + j) * q.w1ld + co]);
  }
  for (int i = t0; i < 9 * KP2 * C * 8; i += stride) {
    const int j = i % 8, co = (i / 8) % C, kc = (i / (8 * C)) % KP2, tap = i / (8 * C * KP2);
    w2[i] = __float2bfloat16_rn(q.dw2[tap * C + co] * q.w2[(size_t)(kc * 8 + j) * q.w2ld + co]);
  }
  for (int i = t0; i < KP1 * C * 8; i += stride) {
    const int j = i % 8, co = (i / 8) % C, kc = i / (8 * C);
    wr[i] = __float2bfloat16_rn(q.wr[(size_t)(kc * 8 + j) * q.wrld + co]);
  }
  for (int i = t0; i < 9 * C; i += stride) {
    const int co = i % C, cls = i / C, cy = cls / 3, cx = cls % 3;
    float s1 = 0.f, s2 = 0.f;
    for (int r = 0; r < 3; ++r)
      for (int s = 0; s < 3; ++s) {
        const bool ok = !(cy == 0 && r == 0) && !(cy == 2 && r == 2) && !(cx == 0 && s == 0) && !(cx == 2 && s == 2);
        if (ok) { s1 += q.dw1[(r * 3 + s) * C + co]; s2 += q.dw2[(r * 3 + s) * C + co]; }
      }
    fb[i] = q.b1[co] + q.c1[co] * s1;
    fb[9 * C + i] = q.b3[co] + q.c2[co] * s2;
  }
  for (int i = t0; i < C; i += stride) {
    fb[18 * C + i] = q.cr[i];
    fb[19 * C + i] = q.wo ? q.wo[(size_t)i * q.wold] : 0.f;
  }
}

size_t dlc_tc_pack_bytes(int Cin, int C) {
  return (size_t)(9 * (Cin / 8) * C * 8 + 9 * (C / 8) * C * 8 + (Cin / 8) * C * 8) * 2 + (size_t)20 * C * 4;
}
void launch_dlc_tc_prepare(const DlcTcPrep& q, void* out, cudaStream_t s) {
  dlc_tc_prepare_kernel<<<64, 256, 0, s>>>(q, reinterpret_cast<uint8_t*>(out));
}
bool dlc_tc_supported(int Cin, int C, bool head) {
  return (Cin == 32 && C == 16 && head) || (Cin == 64 && C == 32 && !head);
}

template <int CIN, int C, bool HEAD, int NBLK>
static void dlc_tc_launch(DlcTcP p, cudaStream_t s) {
  constexpr int TW = tw_of(NBLK), XC = xc_of(NBLK);
  constexpr size_t smem = (size_t)(CIN / 8 + C / 8) * plane_bytes(CIN, NBLK) + (size_t)XR * XC * CIN * 2 +
                          (size_t)(9 * (CIN / 8) * C * 8 + 9 * (C / 8) * C * 8 + (CIN / 8) * C * 8) * 2 + (size_t)20 * C * 4 + 128;
  static unsigned long long attr_done = 0;
  ensure_dyn_smem(dlc_tc_kernel<CIN, C, HEAD, NBLK>, smem, attr_done, "dlc_tc_kernel");
  const int H = 2 * p.h, W = 2 * p.w;
  const unsigned txs = (W + TW - 1) / TW, tys = (H + TH - 1) / TH;
  p.magic_x = txs > 1 ? (unsigned)((0x100000000ull + txs - 1) / txs) : 0u;
  p.magic_y = tys > 1 ? (unsigned)((0x100000000ull + tys - 1) / tys) : 0u;
  const int tiles = (int)(txs * tys) * p.N;
  static int sms = 0;
  if (!sms) { int d = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d); if (sms <= 0) sms = 148; }
  int per_sm = (C <= 16 ? 2 : 1) * (4 / NBLK);
  const int by_smem = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > by_smem) per_sm = by_smem;
  if (per_sm < 1) per_sm = 1;
  const int ctas = per_sm * sms;
  launch_pdl(dlc_tc_kernel<CIN, C, HEAD, NBLK>, dim3(tiles < ctas ? tiles : ctas), dim3(dlc_threads(NBLK)), smem, s, p);
}

void launch_dlc_tc(const DlcTcP& p0, cudaStream_t s) {
  DlcTcP p = p0;
  p.probe = 0;
#ifdef YSP_PROBES
  // YSP_DLC_PROBE (timing experiments only, results are wrong; compiled in with -DYSP_PROBES, never in build.py's flags):
  // 1 = conv1 issues 3 of its 9 taps, 2 = same for conv2, 4 = no up2 arithmetic in phase 1, 8 = no SiLU in epilogue 1
  static const int probe = getenv("YSP_DLC_PROBE") ? atoi(getenv("YSP_DLC_PROBE")) : 0;
  p.probe = probe;
#endif
  // Both stages run the 14 x 30-tile / 320-thread shape.  The 14 x 14-tile / 160-thread shape (four CTAs per SM for stage 4 --
  // the change that halved the parity-mode dlc32 kernel) measured 0.934 vs 0.914 ms here (stage 4) and 0.906 vs 0.568 ms (stage
  // 3, one CTA per SM either way: 59 KB of weights): this kernel is bound by its chains of small MMAs, not by the barriers.
  // YSP_DLC_TC_CFG (A/B timing): 1 = small shape for stage 4, 2 = for stage 3.
  static const int cfg = getenv("YSP_DLC_TC_CFG") ? atoi(getenv("YSP_DLC_TC_CFG")) : 0;
  if (p.Cin == 32 && p.C == 16) {
    if (cfg == 1) dlc_tc_launch<32, 16, true, 2>(p, s);
    else dlc_tc_launch<32, 16, true, 4>(p, s);
  } else {
    if (cfg == 2) dlc_tc_launch<64, 32, false, 2>(p, s);
    else dlc_tc_launch<64, 32, false, 4>(p, s);
  }
}

}  // namespace ysp
