// conv_halo32.cu -- the PARITY-mode (YSP_MODE_TC32) twin of conv_halo.cu: 3x3 stride-1 convolutions with 16 / 32 input channels
// on maps 8..64 pixels wide, fp32 activations, fp32-accurate products on tcgen05 (fp16 hi/lo operand splits, 3 MMAs per product).
//
// conv_tc32.cu converts every fp32 input element to its hi/lo pair once PER TAP (nine TMA boxes and nine conversions per
// 128-pixel tile); for the small-channel 3x3 layers (C3k2 / C3k bottlenecks of detector layers 2, 4, 11, 14, 17 and encoder
// layers 2, 4) that conversion, not HBM or the tensor pipe, sets the time.  Here a 16-row x 32-column input tile is loaded ONCE
// with its 1-pixel halo, split ONCE, and stored as two channel-group-major tiles (hi, lo) [8-channel plane][row][col] x 16 B --
// the canonical K-major no-swizzle UMMA layout when a core matrix (8 M-rows x 16 B) is 8 horizontally adjacent pixels.  The nine
// taps are SHIFTED DESCRIPTOR VIEWS of those tiles (start address + (r * pitch + s) * 16 B, SBO = row pitch, LBO = plane pitch).
// One MMA block = 8 pixels x 16 rows (M = 128); a tile = up to 4 such blocks with independent TMEM accumulators; the issue order
// (tap, k-step, product term, block) never puts two MMAs on the same accumulator back to back.  Weights: the engine's tc32 pack
// ([tap][hi | lo][8-channel group][Cout][8] fp16, scaled by a power of two), resident in shared memory for the CTA's lifetime.
// Epilogue: acc * 2^-e + bias (BN folded), SiLU, + residual, fp32 NHWC store.  Replaces ultralytics Conv.forward_fuse for those
// layers; out-of-image taps read zeros = padding.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "kernels.h"

namespace ysp {

namespace {

constexpr int HROWS = 18;                    // staged rows for 16 output rows (+ halo)

__device__ __forceinline__ uint32_t hs32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void h_umma(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, bool acc) {
  if (acc)
    asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, 1, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                 ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, 0, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                 ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void h_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void h_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tHW3:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra HD3;\n\tbra HW3;\n\tHD3:\n\t}"
               ::"r"(hs32(bar)), "r"(parity) : "memory");
}

}  // namespace

constexpr int kHalo32Threads = 288;          // warps 0-7: staging + epilogue; warp 8: MMA issue

struct Halo32P {
  const float* in; float* out; const float* res; const uint8_t* wpack; const float* bias;
  int N, H, W, in_cs, out_cs, res_cs, Cout, act;
  float w_unscale;
  int nblk, AP, plane, tcols, nbuf;          // column blocks per tile, tile pitch (pixels), bytes per plane, TMEM columns, tile buffers
  int tiles_x;                               // column tiles of 8 * nblk output pixels
};

template <int CIN, int NOUT>
__global__ void __launch_bounds__(kHalo32Threads) conv_halo32_kernel(Halo32P p) {
  constexpr int KP = CIN / 8;
  constexpr int TAPB = 2 * CIN * NOUT * 2;                   // one tap of the tc32 pack: [hi | lo][KP][NOUT] x 16 B
  constexpr int WB = 9 * TAPB;
  extern __shared__ __align__(128) uint8_t hsm32[];
  uint8_t* sW = hsm32;
  float* sBias = reinterpret_cast<float*>(sW + WB);          // [NOUT]
  uint8_t* sT = sW + WB + NOUT * 4;                          // nbuf x { hi [KP][HROWS][AP] x 16 B, lo [KP][HROWS][AP] x 16 B }
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_y = (p.H + 15) >> 4;
  const int total = tiles_y * p.tiles_x * p.N;
  const int half_bytes = KP * p.plane;                       // the hi (or lo) tile
  const int tile_bytes = 2 * half_bytes;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(hs32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(hs32(&tmem_s)), "r"((uint32_t)p.tcols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (int i = tid; i < WB / 16; i += kHalo32Threads) reinterpret_cast<uint4*>(sW)[i] = reinterpret_cast<const uint4*>(p.wpack)[i];
  for (int i = tid; i < NOUT; i += kHalo32Threads) sBias[i] = i < p.Cout ? p.bias[i] : 0.f;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // load + split + store the input tile of `tl` (zero outside the image): item = (pixel slot, 8-channel plane)
  auto stage = [&](int tl, int buf) {
    if (tl >= total) return;
    const int owb = 8 * p.nblk;
    const int X0 = (tl % p.tiles_x) * owb;
    const int tr = tl / p.tiles_x;
    const int n = tr / tiles_y, Y0 = (tr % tiles_y) << 4;
    uint8_t* dst = sT + (size_t)buf * tile_bytes;
    const int items = HROWS * p.AP * KP;
    constexpr int UB = 4;                                    // loads of UB items are in flight before the first conversion
#pragma unroll 1
    for (int i0 = tid; i0 < items; i0 += 256 * UB) {
      float4 v[UB][2];
      int off[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int i = i0 + u * 256;
        off[u] = -1;
        v[u][0] = v[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < items) {
          const int kc = i % KP, pp = i / KP;
          const int rr = pp / p.AP, cc = pp - rr * p.AP;
          const int y = Y0 - 1 + rr, x = X0 + cc - 1;
          off[u] = kc * p.plane + pp * 16;
          if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
            const float4* src = reinterpret_cast<const float4*>(p.in + ((size_t)(n * p.H + y) * p.W + x) * p.in_cs + kc * 8);
            v[u][0] = src[0]; v[u][1] = src[1];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        if (off[u] < 0) continue;
        uint4 h, l;
        split2_f16(v[u][0].x, v[u][0].y, h.x, l.x); split2_f16(v[u][0].z, v[u][0].w, h.y, l.y);
        split2_f16(v[u][1].x, v[u][1].y, h.z, l.z); split2_f16(v[u][1].z, v[u][1].w, h.w, l.w);
        *reinterpret_cast<uint4*>(dst + off[u]) = h;
        *reinterpret_cast<uint4*>(dst + half_bytes + off[u]) = l;
      }
    }
  };
  if (warp < 8) stage(blockIdx.x, 0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(NOUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // D = F32, A = B = F16
  const uint32_t a_hi = (uint32_t)p.AP | (1u << 14);                     // SBO = row pitch (16-byte units), version 1
  const uint32_t b_hi = (128u >> 4) | (1u << 14);                        // SBO = 8 rows x 16 B
  const float2 us2 = make_float2(p.w_unscale, p.w_unscale);
  uint32_t par = 0;
  int buf = 0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int X0 = (tile % p.tiles_x) * 8 * p.nblk;
    const int n = (tile / p.tiles_x) / tiles_y, Y0 = ((tile / p.tiles_x) % tiles_y) << 4;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // generic-proxy tile writes -> visible to the tensor core
    __syncthreads();
    if (warp == 8 && lane == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t t_lo = ((hs32(sT + (size_t)buf * tile_bytes) & 0x3FFFF) >> 4) + (((uint32_t)p.plane >> 4) << 16);
      const uint32_t l_off = (uint32_t)half_bytes >> 4;                   // lo tile behind the hi tile
      const uint32_t w_lo = ((hs32(sW) & 0x3FFFF) >> 4) + (((uint32_t)(NOUT * 16) >> 4) << 16);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
        for (int ks = 0; ks < CIN / 16; ++ks) {
          const uint32_t a_t = t_lo + (uint32_t)((2 * ks * p.plane) >> 4) + (uint32_t)((tap / 3) * p.AP + (tap % 3));
          const uint32_t bh = w_lo + ((tap * TAPB + 2 * ks * NOUT * 16) >> 4);
          const uint32_t bl = bh + ((KP * NOUT * 16) >> 4);
#pragma unroll
          for (int term = 0; term < 3; ++term)
            for (int h = 0; h < p.nblk; ++h)
              h_umma(tmem + h * NOUT, a_t + 8 * h + (term == 1 ? l_off : 0u), a_hi, term == 2 ? bl : bh, b_hi, idesc, (tap | ks | term) != 0);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(hs32(&bar)) : "memory");
    }
    if (warp < 8) {
      if (p.nbuf > 1) stage(tile + gridDim.x, buf ^ 1);                   // the next tile is converted behind the MMAs
      h_wait(&bar, par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int q = warp & 3, row = q * 32 + lane;
      const int by = row >> 3, bxl = row & 7;
      const int Y = Y0 + by;
      for (int h = warp >> 2; h < p.nblk; h += 2) {
        const int X = X0 + 8 * h + bxl;
        const bool valid = Y < p.H && X < p.W;
        const size_t pix = ((size_t)n * p.H + Y) * p.W + X;
#pragma unroll
        for (int c0 = 0; c0 < NOUT; c0 += 16) {
          uint32_t v[16];
          h_ld16(tmem + ((uint32_t)(q * 32) << 16) + h * NOUT + c0, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (!valid) continue;
          float f[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sBias + c0 + 4 * j4);
            float2 x0 = __ffma2_rn(make_float2(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1])), us2, make_float2(b4.x, b4.y));
            float2 x1 = __ffma2_rn(make_float2(__uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])), us2, make_float2(b4.z, b4.w));
            if (p.act) { x0 = silu2_f(x0); x1 = silu2_f(x1); }
            f[4 * j4] = x0.x; f[4 * j4 + 1] = x0.y; f[4 * j4 + 2] = x1.x; f[4 * j4 + 3] = x1.y;
          }
          if (p.res) {
            const float4* rp = reinterpret_cast<const float4*>(p.res + pix * p.res_cs + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 r4 = rp[j];
              f[4 * j] += r4.x; f[4 * j + 1] += r4.y; f[4 * j + 2] += r4.z; f[4 * j + 3] += r4.w;
            }
          }
          float4* op = reinterpret_cast<float4*>(p.out + pix * p.out_cs + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
      }
    }
    par ^= 1;
    // the next tile's MMAs overwrite the accumulators (and, single-buffered, the tile): everyone must be done reading
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (p.nbuf > 1) buf ^= 1;
    else if (warp < 8) stage(tile + gridDim.x, 0);
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tcols) : "memory");
}

bool conv_halo32_supported(const ConvP& p) {
  if (p.kh != 3 || p.kw != 3 || p.pad != 1 || p.stride != 1 || p.in_pw || p.in_ph) return false;
  if (p.H != p.OH || p.W != p.OW || p.OW > 64 || p.OW < 8) return false;
  const int cin = (p.Cin + 15) / 16 * 16;
  if (cin != 16 && cin != 32) return false;
  if (p.Cin != cin && !p.in_zpad) return false;                     // padded input channels must really be zeros
  const int nout = (p.Cout + 15) / 16 * 16;                         // = N_tile of the tc32 weight pack (tc32_tiling)
  if (nout != 16 && nout != 32 && nout != 64) return false;
  if (p.cout_store != p.Cout && p.cout_store != nout) return false;
  if (p.in_cs % 4 || p.out_cs % 4 || (p.res_cs % 4) || p.in_cs < cin || p.out_cs < (p.cout_store > p.Cout ? nout : p.Cout)) return false;
  if (p.Cout % 16 != 0 && p.cout_store != nout) return false;       // the epilogue stores whole 16-channel groups
  return true;
}

template <int CIN, int NOUT>
static void conv_halo32_launch(const ConvP& p, const void* wpack, float w_unscale, cudaStream_t s) {
  Halo32P q = {};
  q.in = (const float*)p.in; q.out = (float*)p.out; q.res = (const float*)p.res; q.wpack = (const uint8_t*)wpack; q.bias = p.bias;
  q.N = p.N; q.H = p.H; q.W = p.W; q.in_cs = p.in_cs; q.out_cs = p.out_cs; q.res_cs = p.res_cs; q.Cout = p.Cout; q.act = p.act;
  q.w_unscale = w_unscale;
  q.nblk = (p.OW + 7) / 8;
  q.tiles_x = 1;
  if (q.nblk > 4) { q.tiles_x = (q.nblk + 3) / 4; q.nblk = (q.nblk + q.tiles_x - 1) / q.tiles_x; }   // <= 32 columns per tile
  q.AP = 8 * q.nblk + 2;
  q.plane = HROWS * q.AP * 16;
  int cols = q.nblk * NOUT;
  q.tcols = 32;
  while (q.tcols < cols) q.tcols <<= 1;
  const size_t fixed = (size_t)9 * 2 * CIN * NOUT * 2 + NOUT * 4;
  const size_t tile = (size_t)2 * (CIN / 8) * q.plane;
  q.nbuf = (fixed + 2 * tile <= 110 * 1024) ? 2 : 1;               // double-buffer when two CTAs still fit per SM
  const size_t smem = fixed + q.nbuf * tile;
  static unsigned long long attr_done = 0;
  ensure_dyn_smem(conv_halo32_kernel<CIN, NOUT>, 200 * 1024, attr_done, "conv_halo32_kernel");
  const int total = ((p.H + 15) / 16) * q.tiles_x * p.N;
  int per_sm = (int)std::min<size_t>(std::min<size_t>(512 / q.tcols, (220 * 1024) / (smem + 1024)), 4);
  if (per_sm < 1) per_sm = 1;
  static int sms = 0;
  if (!sms) { int d = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d); if (sms <= 0) sms = 148; }
  const int grid = std::min(total, sms * per_sm);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kHalo32Threads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = getenv("YSP_NO_PDL") ? 0 : 1;
  cudaLaunchKernelEx(&cfg, conv_halo32_kernel<CIN, NOUT>, q);
}

void launch_conv_halo32(const ConvP& p, const void* wpack, float w_unscale, cudaStream_t s) {
  const int cin = (p.Cin + 15) / 16 * 16, nout = (p.Cout + 15) / 16 * 16;
#define YSP_HALO32(a, b) if (cin == a && nout == b) return conv_halo32_launch<a, b>(p, wpack, w_unscale, s)
  YSP_HALO32(16, 16); YSP_HALO32(16, 32); YSP_HALO32(16, 64); YSP_HALO32(32, 16); YSP_HALO32(32, 32); YSP_HALO32(32, 64);
#undef YSP_HALO32
}

}  // namespace ysp
