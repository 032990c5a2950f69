// conv_halo.cu -- 3x3 stride-1 convolutions with few channels (Cin, Cout <= 64) on maps up to 64 pixels wide, on tcgen05.
//
// The TMA implicit-GEMM kernel (conv_tc.cu) fetches one box per tap and 128-pixel tile; for the small-channel 3x3 layers
// (C3k2 / C3k bottlenecks at 64x64..16x16, Detect's 64->64 box branch) that is nine 2-4 KB boxes per tile and the TMA
// box rate, not HBM or the tensor pipe, sets the time (profiles/r2_layers.md: 100 us for 33 MB of traffic).  Here the
// input tile is staged ONCE per 16 output rows, with its 1-pixel halo, in the channel-group-major layout
// [8-channel plane][row][col] x 16 B -- the canonical K-major no-swizzle UMMA layout when a core matrix (8 M-rows x
// 16 B) is 8 horizontally adjacent pixels -- and the nine taps are SHIFTED DESCRIPTOR VIEWS of that one tile (start
// address + (r*pitch + s) * 16 B, SBO = row pitch, LBO = plane pitch), the trick kernels_dlc_tc.cu introduced.
// One MMA block = 8 pixels x 16 rows (M = 128); a tile = ceil(W/8) such blocks with independent TMEM accumulators, issued
// tap-major so consecutive MMAs never depend on each other.  Weights (bf16, resident in smem for the CTA's lifetime) are
// read straight from the engine's K-major tensor-core pack.  Epilogue: + bias, SiLU, + residual, bf16 NHWC store.
// Replaces ultralytics Conv.forward_fuse for those layers; out-of-image taps read zeros = padding.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "kernels.h"

namespace ysp {

namespace {

constexpr int HR1 = 18, HR2 = 33;            // staged rows for 16 output rows: stride 1 (+ halo) / stride 2

__device__ __forceinline__ uint32_t hs32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void h_umma(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                       bool acc) {
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
  asm volatile("{\n\t.reg .pred p;\n\tHW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra HD;\n\tbra HW;\n\tHD:\n\t}"
               ::"r"(hs32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ float h_tanh(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return t;
}
__device__ __forceinline__ float h_silu(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

}  // namespace

constexpr int kHaloThreads = 288;            // warps 0-7: staging + epilogue; warp 8: MMA issue

struct HaloP {
  const bf16* in; bf16* out; const bf16* res; const bf16* w; const float* bias;
  int N, H, W, in_cs, out_cs, res_cs, Cout, Ktc, act;
  int nblk, AP, plane, tcols, nbuf;          // column blocks, tile pitch (pixels), bytes per plane, TMEM columns, tile buffers
  // stride 2 (3x3, pad 1): a staged row holds the ODD input columns (x = 2j-1, j = 0..8*nblk) followed by the EVEN ones
  // (x = 2j), so the 8 output pixels of a core matrix are again 8 adjacent 16-byte slots for every tap; consecutive output
  // rows are two staged rows apart (SBO = 2 * row pitch).
  int stride, OH, OW, rows, sbo16;           // rows = staged rows per tile; sbo16 = SBO in 16-byte units
  int tiles_x;                               // column tiles of 8 * nblk output pixels
  int tap_off[9];                            // descriptor start offset of tap (r, s) in 16-byte units
};

template <int CIN, int NOUT>
__global__ void __launch_bounds__(kHaloThreads) conv_halo_kernel(HaloP p) {
  constexpr int KP = CIN / 8;
  constexpr int WB = 9 * KP * NOUT * 16;                     // resident weights [9][KP][NOUT] x 16 B
  extern __shared__ __align__(128) uint8_t hsm[];
  uint8_t* sW = hsm;
  float* sBias = reinterpret_cast<float*>(sW + WB);          // [NOUT]
  uint8_t* sT = sW + WB + NOUT * 4;                          // nbuf x [KP][HR][AP] x 16 B
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_y = (p.OH + 15) >> 4;
  const int total = tiles_y * p.tiles_x * p.N;
  const int tile_bytes = KP * p.plane;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(hs32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(hs32(&tmem_s)), "r"((uint32_t)p.tcols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // weights: 16-byte chunk (n, tap, kp) of the K-major pack [n][tap*CIN + ci] -> [tap][kp][n]
  for (int i = tid; i < 9 * KP * NOUT; i += kHaloThreads) {
    const int n = i % NOUT, tk = i / NOUT;                   // tk = tap*KP + kp
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (n < p.Cout) v = *reinterpret_cast<const uint4*>(p.w + (size_t)n * p.Ktc + tk * 8);
    *reinterpret_cast<uint4*>(sW + (size_t)i * 16) = v;
  }
  for (int i = tid; i < NOUT; i += kHaloThreads) sBias[i] = i < p.Cout ? p.bias[i] : 0.f;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  auto stage = [&](int tl, int buf) {                        // cp.async the input tile (zero-filled outside the image)
    if (tl >= total) return;
    const int owb = 8 * p.nblk;
    const int X0 = (tl % p.tiles_x) * owb;
    const int tr = tl / p.tiles_x;
    const int n = tr / tiles_y, Y0 = (tr % tiles_y) << 4;
    uint8_t* dst = sT + (size_t)buf * tile_bytes;
    const int per_plane = p.rows * p.AP;
    // 256 % KP == 0: a thread keeps its channel plane; its pixel slot advances by 256/KP per pass, tracked as (rr, cc)
    // incrementally instead of dividing by the runtime pitch for every 16-byte copy
    const int kc = tid % KP;
    constexpr int STEP = 256 / KP;
    const int drr = STEP / p.AP, dcc = STEP - drr * p.AP;
    int pp = tid / KP, rr = pp / p.AP, cc = pp - rr * p.AP;
    for (; pp < per_plane; pp += STEP, rr += drr, cc += dcc) {
      if (cc >= p.AP) { cc -= p.AP; ++rr; }
      int y, x;
      if (p.stride == 1) { y = Y0 - 1 + rr; x = X0 + cc - 1; }
      else { y = 2 * Y0 - 1 + rr; x = 2 * X0 + (cc <= owb ? 2 * cc - 1 : 2 * (cc - owb - 1)); }
      const bool ok = y >= 0 && y < p.H && x >= 0 && x < p.W;
      const bf16* src = p.in + ((size_t)(n * p.H + (ok ? y : 0)) * p.W + (ok ? x : 0)) * p.in_cs + kc * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(hs32(dst + kc * p.plane + pp * 16)), "l"(src),
                   "r"(ok ? 16 : 0) : "memory");
    }
  };
  if (warp < 8) stage(blockIdx.x, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NOUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t a_hi = (uint32_t)p.sbo16 | (1u << 14);                  // SBO = (stride x) row pitch, version 1
  const uint32_t b_hi = (128u >> 4) | (1u << 14);                        // SBO = 8 rows x 16 B
  uint32_t par = 0;
  int buf = 0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int X0 = (tile % p.tiles_x) * 8 * p.nblk;
    const int n = (tile / p.tiles_x) / tiles_y, Y0 = ((tile / p.tiles_x) % tiles_y) << 4;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // cp.async writes -> visible to the tensor core
    __syncthreads();
    if (p.nbuf > 1) {                                                      // next tile streams in behind the MMAs
      if (warp < 8) stage(tile + gridDim.x, buf ^ 1);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (warp == 8 && lane == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t t_lo = ((hs32(sT + (size_t)buf * tile_bytes) & 0x3FFFF) >> 4) + (((uint32_t)p.plane >> 4) << 16);
      const uint32_t w_lo = ((hs32(sW) & 0x3FFFF) >> 4) + (((uint32_t)(NOUT * 16) >> 4) << 16);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
        for (int ks = 0; ks < CIN / 16; ++ks) {
          const uint32_t a_t = t_lo + (uint32_t)((2 * ks * p.plane) >> 4) + (uint32_t)p.tap_off[tap];
          const uint32_t b_t = w_lo + (((tap * KP + 2 * ks) * NOUT * 16) >> 4);
          for (int h = 0; h < p.nblk; ++h) h_umma(tmem + h * NOUT, a_t + 8 * h, a_hi, b_t, b_hi, idesc, (tap | ks) != 0);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(hs32(&bar)) : "memory");
    }
    if (warp < 8) {
      h_wait(&bar, par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int q = warp & 3, row = q * 32 + lane;
      const int by = row >> 3, bxl = row & 7;
      const int Y = Y0 + by;
      for (int h = warp >> 2; h < p.nblk; h += 2) {
        const int X = X0 + 8 * h + bxl;
        const bool valid = Y < p.OH && X < p.OW;
        const size_t pix = ((size_t)n * p.OH + Y) * p.OW + X;
#pragma unroll
        for (int c0 = 0; c0 < NOUT; c0 += 16) {
          uint32_t v[16];
          h_ld16(tmem + ((uint32_t)(q * 32) << 16) + h * NOUT + c0, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (!valid) continue;
          float f[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {           // 16-byte shared loads (sBias sits at a multiple of 16 B, c0 % 16 == 0)
            const float4 b4 = *reinterpret_cast<const float4*>(sBias + c0 + 4 * j4);
            const float bq[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int jj = 0; jj < 4; jj += 2) {         // packed fp32 (FADD2 / FMUL2 / FFMA2): two channels per instruction
              const int j = 4 * j4 + jj;
              float2 a = __fadd2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), make_float2(bq[jj], bq[jj + 1]));
              if (p.act) {
                const float2 h = __fmul2_rn(a, make_float2(0.5f, 0.5f));
                a = __ffma2_rn(h, make_float2(h_tanh(h.x), h_tanh(h.y)), h);
              }
              f[j] = a.x; f[j + 1] = a.y;
            }
          }
          if (p.res) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.res + pix * p.res_cs + c0);
            const uint4 r0 = rp[0], r1 = rp[1];
            const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
            const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              f[2 * j] += __low2float(h0[j]); f[2 * j + 1] += __high2float(h0[j]);
              f[8 + 2 * j] += __low2float(h1[j]); f[8 + 2 * j + 1] += __high2float(h1[j]);
            }
          }
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 hv = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            w[j] = *reinterpret_cast<uint32_t*>(&hv);
          }
          uint4* op = reinterpret_cast<uint4*>(p.out + pix * p.out_cs + c0);
          op[0] = make_uint4(w[0], w[1], w[2], w[3]);
          op[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
    }
    par ^= 1;
    // the next tile's MMAs overwrite the accumulators (and, single-buffered, the tile): everyone must be done reading
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (p.nbuf > 1) buf ^= 1;
    else {
      if (warp < 8) stage(tile + gridDim.x, 0);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tcols) : "memory");
}

bool conv_halo_supported(const ConvP& p, int in_dt, int out_dt) {
  if (in_dt != DT_BF16 || out_dt != DT_BF16) return false;
  if (p.kh != 3 || p.kw != 3 || p.pad != 1 || p.in_pw || p.in_ph) return false;
  if (p.stride == 1) { if (p.H != p.OH || p.W != p.OW) return false; }
  else if (p.stride == 2) { if (p.H != 2 * p.OH || p.W != 2 * p.OW || p.res_cs) return false; }
  else return false;
  if (p.OW > 64 || p.OW < 8) return false;
  const int cin = (p.Cin + 15) / 16 * 16;
  // measured (profiles/r3_layers.md): with 64 input channels the staged tile + resident weights leave one CTA per SM
  // and the TMA kernel is as fast or faster; the halo path wins for 16 / 32 channels (2.0-3.0x on the 60x60 / 64x64 maps)
  if (cin != 16 && cin != 32) return false;
  if (p.stride == 2 && cin != 16) return false;                     // the stride-2 tile is 4x larger: 16 channels only
  if (p.Cin != cin && !p.in_zpad) return false;                     // padded input channels must really be zeros
  const int nout = p.cout_store;
  if (nout != 16 && nout != 32 && nout != 64) return false;
  if (p.Cout > nout || p.in_cs % 8 || p.out_cs % 8 || (p.res_cs % 8)) return false;
  return true;
}

template <int CIN, int NOUT>
static void conv_halo_launch(const ConvP& p, const void* w_tc, int Ktc, cudaStream_t s) {
  HaloP q = {};
  q.in = (const bf16*)p.in; q.out = (bf16*)p.out; q.res = (const bf16*)p.res; q.w = (const bf16*)w_tc; q.bias = p.bias;
  q.N = p.N; q.H = p.H; q.W = p.W; q.in_cs = p.in_cs; q.out_cs = p.out_cs; q.res_cs = p.res_cs; q.Cout = p.Cout; q.Ktc = Ktc;
  q.act = p.act;
  q.stride = p.stride; q.OH = p.OH; q.OW = p.OW;
  q.nblk = (p.OW + 7) / 8;
  q.tiles_x = 1;
  if (p.stride == 2 && q.nblk > 4) { q.tiles_x = (q.nblk + 3) / 4; q.nblk = 4; }   // stride-2 tiles are 4x larger: 32 columns each
  const int owb = 8 * q.nblk;
  if (p.stride == 1) {
    q.AP = owb + 2; q.rows = HR1; q.sbo16 = q.AP;
    for (int t = 0; t < 9; ++t) q.tap_off[t] = (t / 3) * q.AP + (t % 3);
  } else {
    q.AP = 2 * owb + 1; q.rows = HR2; q.sbo16 = 2 * q.AP;
    const int col[3] = {0, owb + 1, 1};                              // s = 0: odd[ox], s = 1: even[ox], s = 2: odd[ox + 1]
    for (int t = 0; t < 9; ++t) q.tap_off[t] = (t / 3) * q.AP + col[t % 3];
  }
  q.plane = q.rows * q.AP * 16;
  int cols = q.nblk * NOUT;
  q.tcols = 32;
  while (q.tcols < cols) q.tcols <<= 1;
  const size_t fixed = (size_t)9 * (CIN / 8) * NOUT * 16 + NOUT * 4;
  const size_t tile = (size_t)(CIN / 8) * q.plane;
  q.nbuf = (fixed + 2 * tile <= 100 * 1024) ? 2 : 1;               // double-buffer when two CTAs still fit per SM
  const size_t smem = fixed + q.nbuf * tile;
  static unsigned long long attr_done = 0;
  ensure_dyn_smem(conv_halo_kernel<CIN, NOUT>, 200 * 1024, attr_done, "conv_halo_kernel");
  const int total = ((p.OH + 15) / 16) * q.tiles_x * p.N;
  int per_sm = (int)std::min<size_t>(std::min<size_t>(512 / q.tcols, (220 * 1024) / (smem + 1024)), 4);
  if (per_sm < 1) per_sm = 1;
  const int grid = std::min(total, 148 * per_sm);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kHaloThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = getenv("YSP_NO_PDL") ? 0 : 1;
  cudaLaunchKernelEx(&cfg, conv_halo_kernel<CIN, NOUT>, q);
}

void launch_conv_halo(const ConvP& p, const void* w_tc, int Ktc, cudaStream_t s) {
  const int cin = (p.Cin + 15) / 16 * 16, nout = p.cout_store;
#define YSP_HALO(a, b) if (cin == a && nout == b) return conv_halo_launch<a, b>(p, w_tc, Ktc, s)
  YSP_HALO(16, 16); YSP_HALO(16, 32); YSP_HALO(16, 64); YSP_HALO(32, 16); YSP_HALO(32, 32); YSP_HALO(32, 64);
#undef YSP_HALO
}

}  // namespace ysp
