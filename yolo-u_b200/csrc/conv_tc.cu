// conv_tc.cu -- implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05) fed by TMA, for sm_100a.
//
//   D[128 pixels x N couts] (fp32, TMEM) = sum over taps (r,s) and Cin chunks of  A_tap[128 x Kc] * W_tap[N x Kc]^T
//
// * A tiles come straight from the NHWC bf16 activation tensor with ONE 4-D tiled TMA load per (tap, chunk): box
//   {Kc, TW, TH, TN} at coordinates {c, ox0*s + s_tap - pad, oy0*s + r_tap - pad, n0}; TMA's out-of-bounds zero fill
//   implements the conv padding (and the bottom/right zero-padding of decision D1), its element strides implement
//   stride-2 convs.  No im2col buffer ever exists.  1x1/stride-1 convs use the flat [pixels, C] view (box {Kc,128}).
// * W tiles: 2-D TMA from the packed [Cout_pad][taps*Cin_pad] bf16 matrix (K-major), L2-resident.
// * Both land in 128B/64B/32B-swizzled shared memory (swizzle = Kc*2 bytes) and are consumed by
//   tcgen05.mma.cta_group::1.kind::f16 (M=128, N=N_tile, K=16) issued by one thread; accumulators are double-buffered
//   in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
// * Epilogue (8 warps): tcgen05.ld 32x32b -> +bias (BN folded) -> SiLU -> +residual -> bf16/fp32 NHWC store, written
//   into a channel slice of the consumer's buffer (concat by construction).
// * Persistent grid (<= #SMs CTAs), warp-specialised: warp0 TMA producer, warp1 MMA issuer + TMEM owner, warps2-9 epilogue.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "kernels.h"

namespace ysp {

struct TcParams {
  int kw, ntaps, stride, pad, kchunks, Kc, cin_pad;
  int TW, TH, TN, tiles_w, tiles_h, tiles_n, n_tiles_m, n_tiles_n, N_tile, flat;
  int tw_shift, th_shift; unsigned magic_w, magic_h;   // row decode shifts; ceil(2^32 / tiles_{w,h}) for exact small divisions
  int OH, OW, NB, Cout, Cout_st;
  long long M;
  const float* bias; const bf16* res; void* out;
  int res_cs, out_cs, out_f32, act;
  uint32_t idesc, sbo, layout_type, a_bytes, b_bytes, stage_bytes;
  int stages, tmem_cols;
  int w_resident; uint32_t w_bytes;          // weights kept in smem for the whole kernel (loaded once per CTA)
};

struct TcConvPlan {
  TcParams p;
  CUtensorMap tmB;
  mutable CUtensorMap tmA;
  mutable const void* last_in = nullptr;
  int in_cs, H, W, NB, swizzle, pw, ph;
  size_t smem;
  int grid;
};

// ---- PTX wrappers -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major swizzled operand descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46
// | layout_type <<61.  LBO is unused for K-major swizzled tiles whose K extent equals the swizzle span.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo, uint32_t layout_type) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)layout_type << 61);
}

// SiLU(x) = x * sigmoid(x) = h + h * tanh(h), h = x/2: one MUFU (tanh.approx, rel. error ~2^-11, below bf16 resolution)
__device__ __forceinline__ float silu_tanh(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// two adjacent channels at once on the packed fp32 pipe (FMUL2 / FFMA2, sm_100); per channel the same operations as silu_tanh
__device__ __forceinline__ float2 silu_tanh2(float2 x) {
  const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
  return __ffma2_rn(h, t, h);
}

constexpr int kMaxStages = 8;

constexpr int kTcThreads = 320;   // warp0 TMA, warp1 MMA, warps 2-9 epilogue

__global__ void __launch_bounds__(kTcThreads, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tfull_bar[2], tempty_bar[2], wfull_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[528];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
    mbar_init(&wfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 528; i += kTcThreads) s_bias[i] = i < p.Cout ? p.bias[i] : 0.f;
  // Programmatic dependent launch: everything up to griddepcontrol.wait touches constants only (barrier init, TMEM
  // allocation, descriptor prefetch, bias staging, and the resident-weight TMA loads) and overlaps the tail of the
  // previous kernel in the stream; activations (A tiles, residuals) are read only after the wait.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int total_tiles = p.n_tiles_m * p.n_tiles_n;
  const int kiters = p.ntaps * p.kchunks;
  uint8_t* const wsm = smem;                          // resident weights: [n_tiles_n][tap][kchunk] tiles of b_bytes
  uint8_t* const ring = smem + p.w_bytes;             // A (+B when streaming) stage ring
  if (warp == 0 && lane == 0 && p.w_resident) {
    // every CTA used to re-fetch the same few-KB weight tile from the same L2 lines on every k-iteration (an L2
    // hot spot that cost up to half the kernel); now the whole packed weight matrix is loaded once per CTA, and
    // (weights being constants) before the dependency wait.
    mbar_expect_tx(&wfull_bar, (uint32_t)(p.n_tiles_n * kiters) * (uint32_t)p.N_tile * p.Kc * 2u);
    for (int t = 0; t < p.n_tiles_n; ++t)
      for (int it = 0; it < kiters; ++it)
        tma_load_2d(wsm + (size_t)(t * kiters + it) * p.b_bytes, &tmB, &wfull_bar, (it / p.kchunks) * p.cin_pad + (it % p.kchunks) * p.Kc,
                    t * p.N_tile);
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===== TMA producer (one lane) =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tn_i = tile % p.n_tiles_n, tm = tile / p.n_tiles_n;
        int cw, ch, cn;
        if (p.flat) { cw = tm * 128; ch = 0; cn = 0; }
        else {
          int ti = tm % p.tiles_w, r = tm / p.tiles_w;
          int tj = r % p.tiles_h, tk = r / p.tiles_h;
          cw = ti * p.TW * p.stride - p.pad; ch = tj * p.TH * p.stride - p.pad; cn = tk * p.TN;
        }
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int tr = tap / p.kw, ts = tap - tr * p.kw;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = ring + (size_t)stage * p.stage_bytes;
            if (p.w_resident) {
              mbar_expect_tx(&full_bar[stage], 128u * p.Kc * 2u);
              tma_load_4d(sa, &tmA, &full_bar[stage], kc * p.Kc, cw + ts, ch + tr, cn);
            } else {
              mbar_expect_tx(&full_bar[stage], 128u * p.Kc * 2u + (uint32_t)p.N_tile * p.Kc * 2u);
              tma_load_4d(sa, &tmA, &full_bar[stage], kc * p.Kc, cw + ts, ch + tr, cn);
              tma_load_2d(sa + p.a_bytes, &tmB, &full_bar[stage], tap * p.cin_pad + kc * p.Kc, tn_i * p.N_tile);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one lane) =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      if (p.w_resident) mbar_wait(&wfull_bar, 0);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tn_i = tile % p.n_tiles_n;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.N_tile);
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + (size_t)stage * p.stage_bytes);
          const uint32_t sb = p.w_resident ? smem_u32(wsm + (size_t)(tn_i * kiters + it) * p.b_bytes) : sa + p.a_bytes;
          const int nk = p.Kc >> 4;
          for (int k = 0; k < nk; ++k) {
            uint64_t ad = make_desc(sa + k * 32, p.sbo, p.layout_type);
            uint64_t bd = make_desc(sb + k * 32, p.sbo, p.layout_type);
            umma_bf16(d_tmem, ad, bd, p.idesc, (it | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);            // frees the smem slot when these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);                // accumulator ready for the epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue: 8 warps.  Warp w owns TMEM lanes [32q, 32q+32), q = w & 3 (hardware lane-quarter rule); the two
    // warps sharing a quarter split the 16-column chunks of the tile between them (even / odd chunk index). =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    // tile-invariant part of the row -> pixel map (TW, TH are powers of two)
    const int tw = row & (p.TW - 1), r2 = row >> p.tw_shift;
    const int th = r2 & (p.TH - 1), tn = r2 >> p.th_shift;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int tn_i = 0, tm = tile;
      if (p.n_tiles_n > 1) { tn_i = tile % p.n_tiles_n; tm = tile / p.n_tiles_n; }
      long long pix; bool valid;
      if (p.flat) { pix = (long long)tm * 128 + row; valid = pix < p.M; }
      else {
        const int r = p.magic_w ? (int)__umulhi((unsigned)tm, p.magic_w) : tm;     // tm / tiles_w
        const int ti = tm - r * p.tiles_w;
        const int tk = p.magic_h ? (int)__umulhi((unsigned)r, p.magic_h) : r;       // r / tiles_h
        const int tj = r - tk * p.tiles_h;
        const int ox = ti * p.TW + tw, oy = tj * p.TH + th, n = tk * p.TN + tn;
        valid = ox < p.OW && oy < p.OH && n < p.NB;
        pix = ((long long)n * p.OH + oy) * p.OW + ox;
      }
      const int n_base = tn_i * p.N_tile;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.N_tile);
      for (int c0 = half * 16; c0 < p.N_tile; c0 += 32) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        const int cg = n_base + c0;               // first global output channel of this chunk
        if (valid && cg < p.Cout_st) {
          const int nvalid = p.Cout_st - cg;      // >= 16 means the whole chunk is stored (zero-padded channels included)
          float f[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {           // cg % 16 == 0: four 16-byte shared loads instead of 16 scalar ones
            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + cg + 4 * j4);
            const float bq[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int jj = 0; jj < 4; jj += 2) {
              const int j = 4 * j4 + jj;
              float2 x = __fadd2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), make_float2(bq[jj], bq[jj + 1]));
              if (p.act == ACT_SILU) x = silu_tanh2(x);
              f[j] = x.x; f[j + 1] = x.y;
            }
          }
          if (p.res) {
            const bf16* rp = p.res + (size_t)pix * p.res_cs + cg;
            if (nvalid >= 16 && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
              uint4 r0 = *reinterpret_cast<const uint4*>(rp), r1 = *reinterpret_cast<const uint4*>(rp + 8);
              const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
              const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                f[2 * j] += __low2float(h0[j]); f[2 * j + 1] += __high2float(h0[j]);
                f[8 + 2 * j] += __low2float(h1[j]); f[8 + 2 * j + 1] += __high2float(h1[j]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) if (j < nvalid) f[j] += __bfloat162float(rp[j]);
            }
          }
          if (p.out_f32) {
            float* op = reinterpret_cast<float*>(p.out) + (size_t)pix * p.out_cs + cg;
            if (nvalid >= 16 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                reinterpret_cast<float4*>(op)[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) if (j < nvalid) op[j] = f[j];
            }
          } else {
            bf16* op = reinterpret_cast<bf16*>(p.out) + (size_t)pix * p.out_cs + cg;
            if (nvalid >= 16 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
              uint32_t w[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                w[j] = *reinterpret_cast<uint32_t*>(&h);
              }
              reinterpret_cast<uint4*>(op)[0] = make_uint4(w[0], w[1], w[2], w[3]);
              reinterpret_cast<uint4*>(op)[1] = make_uint4(w[4], w[5], w[6], w[7]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) if (j < nvalid) op[j] = __float2bfloat16_rn(f[j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && p) fn = (EncodeTiledFn)p;
  }
  return fn;
}
static CUtensorMapSwizzle swz(int bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}
static int num_sms() {
  static int n = 0;
  if (!n) { int d = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); if (n <= 0) n = 148; }
  return n;
}

bool tc_conv_supported(const ConvP& p) {
  if (p.kh != p.kw || (p.kh != 1 && p.kh != 3) || (p.stride != 1 && p.stride != 2)) return false;
  const int cin_pad = (p.Cin + 15) / 16 * 16;
  if (p.Cin % 16 != 0 && !(p.in_zpad && p.in_cs >= cin_pad)) return false;   // padded K only over zero-filled channels
  if (p.in_cs % 8 != 0) return false;
  if (p.M < 128) return false;
  return get_encode() != nullptr;
}

static bool encode_A(const TcConvPlan* pl, const void* in) {
  const TcParams& p = pl->p;
  cuuint64_t dims[4]; cuuint64_t strides[3]; cuuint32_t box[4]; cuuint32_t es[4];
  const cuuint64_t px = (cuuint64_t)pl->in_cs * 2;
  if (p.flat) {
    dims[0] = p.cin_pad; dims[1] = (cuuint64_t)p.M; dims[2] = 1; dims[3] = 1;
    strides[0] = px; strides[1] = px * (cuuint64_t)p.M; strides[2] = strides[1];
    box[0] = p.Kc; box[1] = 128; box[2] = 1; box[3] = 1;
    es[0] = es[1] = es[2] = es[3] = 1;
  } else {
    dims[0] = p.cin_pad; dims[1] = pl->W; dims[2] = pl->H; dims[3] = pl->NB;
    strides[0] = px; strides[1] = px * pl->pw; strides[2] = px * pl->pw * pl->ph;
    box[0] = p.Kc; box[1] = p.TW * p.stride; box[2] = p.TH * p.stride; box[3] = p.TN;
    es[0] = 1; es[1] = p.stride; es[2] = p.stride; es[3] = 1;
  }
  CUresult r = get_encode()(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz(pl->swizzle), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "libysp: cuTensorMapEncodeTiled(A) failed: %d\n", (int)r); return false; }
  pl->last_in = in;
  return true;
}

TcConvPlan* tc_conv_plan_create(const ConvP& c, const void* w_bf16, int out_dt) {
  TcConvPlan* pl = new TcConvPlan();
  TcParams& p = pl->p;
  p = TcParams();
  const int cin_pad = (c.Cin + 15) / 16 * 16;
  const int cout_pad = (c.Cout + 15) / 16 * 16;
  p.kw = c.kw; p.ntaps = c.kh * c.kw; p.stride = c.stride; p.pad = c.pad; p.cin_pad = cin_pad;
  p.Kc = cin_pad % 64 == 0 ? 64 : (cin_pad % 32 == 0 ? 32 : 16);
  p.kchunks = cin_pad / p.Kc;
  p.OH = c.OH; p.OW = c.OW; p.NB = c.N; p.Cout = c.Cout; p.Cout_st = c.cout_store > c.Cout ? c.cout_store : c.Cout; p.M = c.M;
  p.n_tiles_n = (cout_pad + 255) / 256;
  p.N_tile = ((cout_pad + p.n_tiles_n - 1) / p.n_tiles_n + 15) / 16 * 16;
  const bool pitched = (c.in_pw && c.in_pw != c.W) || (c.in_ph && c.in_ph != c.H);
  p.flat = (c.kh == 1 && c.stride == 1 && c.OH == c.H && c.OW == c.W && !pitched) ? 1 : 0;
  if (p.flat) {
    p.TW = 128; p.TH = 1; p.TN = 1; p.tiles_w = p.tiles_h = p.tiles_n = 1;
    p.n_tiles_m = (int)((c.M + 127) / 128);
  } else {
    // choose the 128-pixel output tile (TW x TH x TN, powers of two) that wastes the fewest pixels; prefer wide rows
    double best = 1e30; int bw = 8, bh = 16, bn = 1;
    for (int tw = 128; tw >= 1; tw >>= 1) {
      if (tw < 8 && c.OW >= 8) continue;
      if (tw * c.stride > 256) continue;
      for (int th = 128 / tw; th >= 1; th >>= 1) {
        int tn = 128 / (tw * th);
        if (th * c.stride > 256 || tn > 256) continue;
        double cover = (double)((c.OW + tw - 1) / tw * tw) * ((c.OH + th - 1) / th * th) * ((c.N + tn - 1) / tn * tn);
        double cost = cover * (1.0 + 0.01 * tn);
        if (cost < best) { best = cost; bw = tw; bh = th; bn = tn; }
      }
    }
    p.TW = bw; p.TH = bh; p.TN = bn;
    p.tiles_w = (c.OW + bw - 1) / bw; p.tiles_h = (c.OH + bh - 1) / bh; p.tiles_n = (c.N + bn - 1) / bn;
    p.n_tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  }
  p.tw_shift = 0; while ((1 << p.tw_shift) < p.TW) ++p.tw_shift;
  p.th_shift = 0; while ((1 << p.th_shift) < p.TH) ++p.th_shift;
  p.magic_w = p.tiles_w > 1 ? (unsigned)((0x100000000ull + p.tiles_w - 1) / p.tiles_w) : 0u;
  p.magic_h = p.tiles_h > 1 ? (unsigned)((0x100000000ull + p.tiles_h - 1) / p.tiles_h) : 0u;
  pl->swizzle = p.Kc * 2;
  p.sbo = 8u * pl->swizzle;
  p.layout_type = pl->swizzle == 128 ? 2u : (pl->swizzle == 64 ? 4u : 6u);
  p.a_bytes = (128u * p.Kc * 2u + 1023u) & ~1023u;
  p.b_bytes = ((uint32_t)p.N_tile * p.Kc * 2u + 1023u) & ~1023u;
  int cols = 32;
  while (cols < 2 * p.N_tile) cols <<= 1;
  p.tmem_cols = cols;
  // weights stay resident in smem when the whole packed matrix is small (true for all but the widest 3x3 layers)
  const uint32_t w_all = (uint32_t)(p.n_tiles_n * p.ntaps * p.kchunks) * p.b_bytes;
  p.w_resident = w_all <= 80u * 1024u ? 1 : 0;
  p.w_bytes = p.w_resident ? w_all : 0u;
  p.stage_bytes = p.w_resident ? p.a_bytes : p.a_bytes + p.b_bytes;
  // two CTAs share an SM (TMEM <= 256 columns and <= ~105 KB smem each) when at least 3 stages fit next to the weights
  uint32_t budget = 104u * 1024u;
  if (cols > 256 || p.w_bytes + 3u * p.stage_bytes > budget) budget = 200u * 1024u;
  int st = (int)((budget - p.w_bytes) / p.stage_bytes);
  p.stages = st > kMaxStages ? kMaxStages : (st < 2 ? 2 : st);
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=BF16, K-major both, N>>3 @17, M>>4 @24
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N_tile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.bias = c.bias; p.res_cs = c.res_cs; p.out_cs = c.out_cs; p.out_f32 = out_dt == DT_F32; p.act = c.act;
  pl->in_cs = c.in_cs; pl->H = c.H; pl->W = c.W; pl->NB = c.N;
  pl->pw = c.in_pw ? c.in_pw : c.W; pl->ph = c.in_ph ? c.in_ph : c.H;
  pl->smem = (size_t)p.w_bytes + (size_t)p.stages * p.stage_bytes + 1024;
  const int total = p.n_tiles_m * p.n_tiles_n;
  const int ctas_per_sm = (p.tmem_cols <= 256 && pl->smem <= 110 * 1024) ? 2 : 1;
  pl->grid = total < ctas_per_sm * num_sms() ? total : ctas_per_sm * num_sms();
  // weights: [cout_pad][Ktc] bf16
  {
    cuuint64_t dims[2] = {(cuuint64_t)p.ntaps * cin_pad, (cuuint64_t)cout_pad};
    cuuint64_t strides[1] = {(cuuint64_t)p.ntaps * cin_pad * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.Kc, (cuuint32_t)p.N_tile};
    cuuint32_t es[2] = {1, 1};
    CUresult r = get_encode()(&pl->tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_bf16), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swz(pl->swizzle), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "libysp: cuTensorMapEncodeTiled(W) failed: %d\n", (int)r); delete pl; return nullptr; }
  }
  static unsigned long long attr_done = 0;
  ensure_dyn_smem(conv_tc_kernel, 208 * 1024, attr_done, "conv_tc_kernel");
  return pl;
}

void tc_conv_plan_destroy(TcConvPlan* p) { delete p; }

void launch_conv_tc(const TcConvPlan* pl, const ConvP& c, cudaStream_t s) {
  if (pl->last_in != c.in && !encode_A(pl, c.in)) return;
  TcParams p = pl->p;
  p.res = reinterpret_cast<const bf16*>(c.res); p.out = c.out;
  static const bool no_pdl = getenv("YSP_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl->grid); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = pl->smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
  cudaLaunchKernelEx(&cfg, conv_tc_kernel, pl->tmA, pl->tmB, p);
}

}  // namespace ysp
