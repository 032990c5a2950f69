// kernels_train_tc.cu -- the 1x1-conv GEMMs of the seg-head TRAINING step (forward and input gradient) on tcgen05 tensor cores.
//
//   C[m][j] = beta * C[m][j] + bias[j] + sum_i A[m][i] * Wop(i, j)          (same contract as pw_gemm_kernel, kernels_train.cu)
//
// ncu of the FFMA version (profiles/r7_train_gemm.md): the 32 pointwise GEMMs of a step take 8.8 ms of its 23 ms at ~1 TB/s of
// DRAM traffic -- compute-bound on the fp32 pipe (K, N = 16..144), not bandwidth-bound.  Here every operand is split into bf16
// hi + mid (x = hi + mid to 2^-18; bf16 because the gradient operands, 1e-5 and below, underflow the fp16 range that the inference
// path's hi/lo split uses) and each product is three bf16 MMAs (hi.hi + mid.hi + hi.mid) with fp32 accumulation in TMEM: products
// accurate to 2^-17, far inside the 2e-3 gradient tolerance checked against torch autograd (the reference itself trains under fp16
// autocast, train.py:302-341), and the GEMM runs at the rate of its loads:
//   * prologue (once per persistent CTA): the weight matrix -- either orientation, the optional second reduction range (A2/W2)
//     and the optional rider columns (Jsplit/Wb) -- is split and stored as the B operand
//     [8-k plane][n][8] (K-major, no swizzle);
//   * per 128-row tile: 256 threads load the fp32 rows (two threads per row, alternate 8-channel planes), apply the optional
//     input transform x = act(z * sc + sh) (the producer's BatchNorm + SiLU, never materialised), split, store the A operand
//     [8-k plane][row] x 16 B; one thread issues 3 * K/16 MMAs (M = 128, N = padded J); 8 warps read the accumulator back
//     (tcgen05.ld), un-scale, add bias / the previous C (beta), store; BatchNorm statistics (column sum and sum of squares of the
//     bias-free product) go through a column-major shared-memory tile and are accumulated per thread across the CTA's tiles.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "kernels.h"
#include "kernels_train.h"

namespace ysp {

namespace {

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
  asm volatile("{\n\t.reg .pred p;\n\tWLT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DNT;\n\tbra WLT;\n\tDNT:\n\t}"
               ::"r"(s32(bar)), "r"(parity) : "memory");
}

struct GemmTcP {
  const float* A0; int lda0; const float* W0; int ldw0; int trans; const float* bias; float* C; int ldc;
  long long M; int I0, J, beta; double* sums; InTf tf; PwDual du;
  int K0, K2, Kt, N, tcols, NB;      // padded reduction ranges (K0 + K2 = Kt), padded columns, TMEM columns, 128-row blocks per iteration
  int a_half, b_half, out_pitch;     // bytes of one A / B half (hi or lo); floats per column of the statistics tile
};


// weight of (reduction index k in the padded range, output column j); 0 in the padding
__device__ __forceinline__ float weight_at(const GemmTcP& p, int k, int j) {
  if (j >= p.J) return 0.f;
  if (k < p.K0) {
    if (k >= p.I0) return 0.f;
    if (p.du.Jsplit && j >= p.du.Jsplit) return p.du.Wb[(size_t)(j - p.du.Jsplit) * p.du.ldwb + k];   // rider: forward layout [Cout][Cin]
    return p.trans ? p.W0[(size_t)k * p.ldw0 + j] : p.W0[(size_t)j * p.ldw0 + k];
  }
  const int k2 = k - p.K0;
  if (k2 >= p.du.I2) return 0.f;
  return p.trans ? p.du.W2[(size_t)k2 * p.du.ldw2 + j] : p.du.W2[(size_t)j * p.du.ldw2 + k2];
}

}  // namespace

// F16 = true: fp16 hi/lo operands (22 mantissa bits; weights scaled by a power of two so their lo parts stay normal) -- the FORWARD
// GEMMs, whose A operand is an O(1) activation and whose result must reproduce the fp32 loss to 1e-5.  F16 = false: bf16 hi/mid
// operands (17-bit products, fp32 exponent range) -- the INPUT-GRADIENT GEMMs, whose A operand holds gradients of 1e-5 and below
// that underflow fp16 (measured: 0.2 % gradient errors and a training curve that drifts from the fp32 one with fp16 operands).
// kGT threads per CTA (512 / kGT CTAs per SM: 113 registers per thread): the phases of an iteration (load + convert, MMA, epilogue,
// statistics) run one after the other inside a CTA, so several small CTAs overlap them better than two large ones.
template <bool F16, int kGT>
__global__ void __launch_bounds__(kGT, 512 / kGT) pw_gemm_tc_kernel(const GemmTcP p) {
  static_assert(kGT == 256, "the weight-range reduction and the lane-quarter mapping assume eight warps");
  extern __shared__ __align__(128) uint8_t gsm[];
  uint8_t* sA = gsm;                                         // hi [Kt/8][128] x 16 B, then lo; re-used as the statistics tile
  uint8_t* sB = gsm + 2 * p.a_half;                          // hi [Kt/8][N] x 16 B, then lo
  float* sSc = reinterpret_cast<float*>(sB + 2 * p.b_half);  // input transform scale / shift [K0]
  float* sSh = sSc + p.K0;
  float* sBias = sSh + p.K0;                                 // [N]
  double* sAcc = reinterpret_cast<double*>(sBias + p.N + ((2 * p.K0 + p.N) & 1));   // [2][N] column sum / sum of squares (8-byte aligned)
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  __shared__ float red[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool tfon = p.tf.gamma != nullptr;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_s)), "r"((uint32_t)p.tcols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- weights -> B operand (fp16: scaled by 2^e so that the lo parts stay normal) ----
  float wmax = 0.f;
  if (F16) {
    for (int e = tid; e < p.Kt * p.N; e += kGT) wmax = fmaxf(wmax, fabsf(weight_at(p, e % p.Kt, e / p.Kt)));
#pragma unroll
    for (int o = 16; o; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) red[warp] = wmax;                          // kGT = 256: eight warps
  }
  for (int i = tid; i < p.K0; i += kGT) {
    float sc = 1.f, sh = 0.f;
    if (tfon && i < p.I0) { sc = p.tf.gamma[i] * p.tf.invstd[i]; sh = p.tf.beta[i] - p.tf.mean[i] * sc; }
    sSc[i] = sc; sSh[i] = sh;
  }
  for (int j = tid; j < p.N; j += kGT) {
    float b = 0.f;
    if (j < p.J) {
      if (p.du.Jsplit && j >= p.du.Jsplit) b = p.du.biasb ? p.du.biasb[j - p.du.Jsplit] : 0.f;
      else b = p.bias ? p.bias[j] : 0.f;
    }
    sBias[j] = b;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  int e2 = 0;
  if (F16) {
    wmax = fmaxf(fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])), fmaxf(fmaxf(red[4], red[5]), fmaxf(red[6], red[7])));
    if (wmax > 0.f) { int ex; frexpf(wmax, &ex); e2 = 14 - ex; }         // max |w| * 2^e2 in [2^13, 2^14)
  }
  const float wsc = ldexpf(1.f, e2), unscale = ldexpf(1.f, -e2);
  for (int e = tid; e < (p.Kt >> 1) * p.N; e += kGT) {                    // two consecutive k per thread: one 32-bit store per half
    const int kp = e % (p.Kt >> 1), j = e / (p.Kt >> 1), k = 2 * kp;
    uint32_t hi, lo;
    if (F16) split2_f16(weight_at(p, k, j) * wsc, weight_at(p, k + 1, j) * wsc, hi, lo);
    else split2_bf16(weight_at(p, k, j), weight_at(p, k + 1, j), hi, lo);
    const uint32_t off = (uint32_t)(k >> 3) * (uint32_t)(p.N * 16) + (uint32_t)j * 16u + (uint32_t)(k & 7) * 2u;
    *reinterpret_cast<uint32_t*>(sB + off) = hi;
    *reinterpret_cast<uint32_t*>(sB + p.b_half + off) = lo;
  }
  const uint32_t tmem = tmem_s;
  const uint32_t idesc = (1u << 4) | (F16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // D = F32, A = B = F16 / BF16
  const int row = tid & 127, hsel = tid >> 7;                // A load: kGT / 128 threads per row, interleaved 8-channel planes
  constexpr int PSTEP = kGT / 128;
  const int q = warp & 3, chalf = warp >> 2;                 // epilogue: TMEM lane quarter, interleaved 16-column chunks
  const bool vecA0 = ((p.lda0 | p.I0) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.A0) & 15) == 0;
  const bool vecA2 = p.du.A2 && ((p.du.lda2 | p.du.I2) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.du.A2) & 15) == 0;
  const bool vecC = ((p.ldc | p.J | p.du.Jsplit | p.du.ldcb) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(p.du.Cb) & 15) == 0;
  const int Jstat = p.du.Jsplit ? p.du.Jsplit : p.J;
  for (int i = tid; i < 2 * p.N; i += kGT) sAcc[i] = 0.0;     // BatchNorm statistics of this CTA's rows (double)
  uint32_t par = 0;
  const int planes = p.Kt >> 3;
  const int nbs = p.NB == 4 ? 2 : (p.NB == 2 ? 1 : 0);
  const int blk_bytes = planes * 2048;                       // one 128-row block of the hi (or lo) A tile
  const long long mtiles = (p.M + 128 * p.NB - 1) / (128 * p.NB);
#pragma unroll 1
  for (long long mt = blockIdx.x; mt < mtiles; mt += gridDim.x) {
    const long long mbase = mt * 128 * p.NB;
    // ---- A operand: NB row blocks x this thread's planes; the loads of UB items (32 B each) are in flight before the first
    //      conversion: 256 threads x 4 x 32 B x 2 CTAs = 64 KB per SM, what the HBM latency-bandwidth product needs ----
    constexpr int UB = 4;
    const int items = p.NB * ((planes - hsel + PSTEP - 1) / PSTEP);   // (block, plane) pairs of this thread
#pragma unroll 1
    for (int it0 = 0; it0 < items; it0 += UB) {
      float v[UB][8];
      int dst[UB];
      bool tfm[UB];
      int k0s[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int it = it0 + u;
        dst[u] = -1; tfm[u] = false; k0s[u] = 0;
        if (it < items) {
          const int blk = it & (p.NB - 1), g = hsel + PSTEP * (it >> nbs);   // NB is 1, 2 or 4
          const long long m = mbase + blk * 128 + row;
          const int k0 = g * 8;
          const bool part2 = k0 >= p.K0;
          const float* src = part2 ? p.du.A2 + m * p.du.lda2 + (k0 - p.K0) : p.A0 + m * p.lda0 + k0;
          const int valid = part2 ? p.du.I2 - (k0 - p.K0) : p.I0 - k0;       // channels of this plane that exist
          if (m < p.M && valid >= 8 && (part2 ? vecA2 : vecA0)) {
            const float4 a = *reinterpret_cast<const float4*>(src), bq = *reinterpret_cast<const float4*>(src + 4);
            v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w; v[u][4] = bq.x; v[u][5] = bq.y; v[u][6] = bq.z; v[u][7] = bq.w;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[u][j] = (m < p.M && j < valid) ? src[j] : 0.f;
          }
          dst[u] = blk * blk_bytes + g * 2048 + row * 16;
          tfm[u] = tfon && !part2 && m < p.M;
          k0s[u] = k0;
        }
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        if (dst[u] < 0) continue;
        if (tfm[u]) {                                        // x = act(z * sc + sh); channels past I0 have sc = 1, sh = 0 and z = 0
          const float4 c0 = *reinterpret_cast<const float4*>(sSc + k0s[u]), c1 = *reinterpret_cast<const float4*>(sSc + k0s[u] + 4);
          const float4 h0 = *reinterpret_cast<const float4*>(sSh + k0s[u]), h1 = *reinterpret_cast<const float4*>(sSh + k0s[u] + 4);
          const float sc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w}, sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            float2 x = __ffma2_rn(make_float2(v[u][j], v[u][j + 1]), make_float2(sc[j], sc[j + 1]), make_float2(sh[j], sh[j + 1]));
            if (p.tf.act) x = silu2_f(x);                      // MUFU SiLU (3e-7), as every fp32-storage inference kernel
            const bool e0 = k0s[u] + j < p.I0, e1 = k0s[u] + j + 1 < p.I0;
            v[u][j] = e0 ? x.x : 0.f; v[u][j + 1] = e1 ? x.y : 0.f;
          }
        }
        uint4 h, l;
        if (F16) {
          split2_f16(v[u][0], v[u][1], h.x, l.x); split2_f16(v[u][2], v[u][3], h.y, l.y);
          split2_f16(v[u][4], v[u][5], h.z, l.z); split2_f16(v[u][6], v[u][7], h.w, l.w);
        } else {
          split2_bf16(v[u][0], v[u][1], h.x, l.x); split2_bf16(v[u][2], v[u][3], h.y, l.y);
          split2_bf16(v[u][4], v[u][5], h.z, l.z); split2_bf16(v[u][6], v[u][7], h.w, l.w);
        }
        *reinterpret_cast<uint4*>(sA + dst[u]) = h;
        *reinterpret_cast<uint4*>(sA + p.a_half + dst[u]) = l;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = s32(sA), b0 = s32(sB);
      // per row block two accumulators: the hi.hi products and the two cross terms (lo.hi + hi.lo) -- the tensor core truncates
      // its fp32 accumulation, so the long chain carries only K/16 steps and the small terms cannot disturb it; block-major
      // inner loop so consecutive MMAs never hit the same accumulator
      for (int ks = 0; ks < (p.Kt >> 4); ++ks) {
        const uint64_t bh = desc_nosw(b0 + ks * 2 * p.N * 16, (uint32_t)p.N * 16u, 128u);
        const uint64_t bl = desc_nosw(b0 + p.b_half + ks * 2 * p.N * 16, (uint32_t)p.N * 16u, 128u);
        for (int term = 0; term < 3; ++term)
          for (int blk = 0; blk < p.NB; ++blk) {
            const uint64_t ad = desc_nosw(a0 + (term == 1 ? p.a_half : 0) + blk * blk_bytes + ks * 4096, 2048u, 128u);
            umma_f16(tmem + (2 * blk + (term ? 1 : 0)) * p.N, ad, term == 2 ? bl : bh, idesc, term == 2 ? 1u : (ks ? 1u : 0u));
          }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    }
    mbar_wait_parity(&bar, par);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    par ^= 1;
    // ---- epilogue: row = TMEM lane, alternate 16-column chunks per warp half; the A tiles are dead -> statistics tile ----
    float* sOut = reinterpret_cast<float*>(sA);              // [column][out_pitch] (column-major: conflict-free both ways)
    for (int blk = 0; blk < p.NB; ++blk) {
      const int erow = q * 32 + lane;
      const long long em = mbase + blk * 128 + erow;
      for (int c0 = chalf * 16; c0 < p.N; c0 += 16 * PSTEP) {
        uint32_t v[16], w[16];
        ld16(tmem + ((uint32_t)(q * 32) << 16) + (2 * blk) * p.N + c0, v);
        ld16(tmem + ((uint32_t)(q * 32) << 16) + (2 * blk + 1) * p.N + c0, w);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = (__uint_as_float(v[j]) + __uint_as_float(w[j])) * unscale;
        if (p.sums) {
#pragma unroll
          for (int j = 0; j < 16; ++j) sOut[(c0 + j) * p.out_pitch + blk * 128 + erow] = f[j];
        }
        if (em < p.M && c0 < p.J) {
          const bool second = p.du.Jsplit && c0 >= p.du.Jsplit;            // Jsplit is a multiple of 16 when the tensor-core path is taken
          float* crow = second ? p.du.Cb + em * p.du.ldcb + (c0 - p.du.Jsplit) : p.C + em * p.ldc + c0;
          const int nvalid = (second ? p.J : (p.du.Jsplit ? p.du.Jsplit : p.J)) - c0;
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sBias + c0 + 4 * j4);
            f[4 * j4] += b4.x; f[4 * j4 + 1] += b4.y; f[4 * j4 + 2] += b4.z; f[4 * j4 + 3] += b4.w;
          }
          if (vecC && nvalid >= 16) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              float4 o = make_float4(f[4 * j4], f[4 * j4 + 1], f[4 * j4 + 2], f[4 * j4 + 3]);
              float4* op = reinterpret_cast<float4*>(crow) + j4;
              if (p.beta) { const float4 pv = *op; o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w; }
              *op = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < nvalid) crow[j] = p.beta ? crow[j] + f[j] : f[j];
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                         // accumulators read by everyone; statistics tile complete
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (p.sums) {
      // column sums of the bias-free product: a warp per column, lanes stride over the rows (consecutive addresses), shuffle
      // reduction, lane 0 adds into the CTA's double accumulators (each column is owned by one warp: no atomics)
      const int rows = 128 * p.NB;
      for (int col = warp; col < Jstat; col += kGT / 32) {
        const float* cp = sOut + col * p.out_pitch;
        float a1 = 0.f, a2 = 0.f;
        for (int r = lane; r < rows; r += 32) { const float x = cp[r]; a1 += x; a2 = fmaf(x, x, a2); }   // rows >= M hold exact zeros
#pragma unroll
        for (int o = 16; o; o >>= 1) { a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); }
        if (lane == 0) { sAcc[col] += (double)a1; sAcc[p.N + col] += (double)a2; }
      }
      __syncthreads();                                       // the next tile's A operand overwrites the statistics tile
    }
  }
  if (p.sums) {
    for (int col = warp; col < Jstat; col += kGT / 32)
      if (lane == 0) { atomicAdd(&p.sums[col], sAcc[col]); atomicAdd(&p.sums[Jstat + col], sAcc[p.N + col]); }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tcols) : "memory");
  }
}

static bool launch_pw_gemm_tc2(GemmTcP p, int trans, cudaStream_t s);       // the warp-specialised version below

// true when the launch was taken (tensor-core path); false -> the caller runs the FFMA kernel
bool launch_pw_gemm_tc(const float* A, int lda, const float* W, int ldw, int trans, const float* bias, float* C, int ldc, long long M,
                       int I, int J, int beta, cudaStream_t s, double* sums, InTf tf, PwDual du) {
  static const bool off = getenv("YSP_TRAIN_NO_TC") != nullptr;
  if (off || M < 128) return false;
  GemmTcP p = {};
  p.A0 = A; p.lda0 = lda; p.W0 = W; p.ldw0 = ldw; p.trans = trans; p.bias = bias; p.C = C; p.ldc = ldc; p.M = M; p.I0 = I; p.J = J;
  p.beta = beta; p.sums = sums; p.tf = tf; p.du = du;
  p.K0 = (I + 15) / 16 * 16; p.K2 = du.A2 ? (du.I2 + 15) / 16 * 16 : 0; p.Kt = p.K0 + p.K2;
  p.N = (J + 15) / 16 * 16;
  if (p.N > 256 || p.Kt > 288 || J < 8) return false;                        // narrow outputs (the 16 -> 1 head) stay on CUDA cores
  if (du.Jsplit && (du.Jsplit & 15)) return false;                            // rider columns must start on a 16-column chunk
  // large launches with N <= 64 go to the warp-specialised ring version below (15.14 -> 14.89 ms per step; YSP_TRAIN_GEMM_RING=0: A/B)
  static const int ring = getenv("YSP_TRAIN_GEMM_RING") ? atoi(getenv("YSP_TRAIN_GEMM_RING")) : 1;
  if (ring && M >= 128 * 296 && launch_pw_gemm_tc2(p, trans, s)) return true;
  // Rows per iteration: as many 128-row blocks (1, 2 or 4) as keep the A tiles within 64 KB and the 2 NB accumulators within 256
  // TMEM columns (two 256-thread CTAs per SM) -- the narrow GEMMs of the full-resolution stages move 16 KB per block and are
  // latency-bound one block at a time.  (Four 128-thread CTAs per SM, the shape that halved the inference decoder kernel,
  // measured the same here: 19.5 ms per step either way.)
  p.NB = 1;
  for (int c = 2; c <= 4; c *= 2)
    if (2 * c * p.N <= 256 && (size_t)(2 * c) * (p.Kt / 8) * 2048 <= 64 * 1024 && (long long)128 * c * 592 <= M) p.NB = c;
  p.tcols = 32;
  while (p.tcols < 2 * p.NB * p.N) p.tcols <<= 1;
  if (p.tcols > 512) return false;
  p.a_half = p.NB * (p.Kt / 8) * 2048;
  p.b_half = (p.Kt / 8) * p.N * 16;
  p.out_pitch = 128 * p.NB + 1;
  const size_t stats_bytes = sums ? (size_t)p.N * p.out_pitch * 4 : 0;
  const size_t a_bytes = std::max<size_t>(2 * (size_t)p.a_half, stats_bytes);
  if (a_bytes > 2 * (size_t)p.a_half) p.a_half = (int)((a_bytes / 2 + 127) / 128 * 128);     // the statistics tile needs more room than the A tiles
  const size_t smem = 2 * (size_t)p.a_half + 2 * (size_t)p.b_half + (size_t)(2 * p.K0 + p.N + 1) * 4 + (size_t)2 * p.N * 8 + 128;
  if (smem > 200 * 1024) return false;
  static int sms = 0;
  if (!sms) { int d = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d); if (sms <= 0) sms = 148; }
  const long long mtiles = (M + 128 * p.NB - 1) / (128 * p.NB);
  int per_sm = (int)std::min<size_t>(std::min<size_t>(512 / p.tcols, (220 * 1024) / (smem + 1024)), 2);
  if (per_sm < 1) per_sm = 1;
  const int grid = (int)std::min<long long>(mtiles, (long long)sms * per_sm);
  static unsigned long long attr_f = 0, attr_b = 0;
  ensure_dyn_smem(pw_gemm_tc_kernel<true, 256>, 200 * 1024, attr_f, "pw_gemm_tc_kernel<f16>");
  ensure_dyn_smem(pw_gemm_tc_kernel<false, 256>, 200 * 1024, attr_b, "pw_gemm_tc_kernel<bf16>");
  if (trans) pw_gemm_tc_kernel<false, 256><<<grid, 256, smem, s>>>(p);      // input gradient: bf16 hi/mid operands
  else pw_gemm_tc_kernel<true, 256><<<grid, 256, smem, s>>>(p);            // forward: fp16 hi/lo operands
  return true;
}

// =====================================================================================================================
// The same GEMM, warp-specialised (the kernel above runs load -> MMA -> epilogue -> statistics one after the other inside a CTA;
// ncu: 25-35 % of its stall samples wait for the loads, 10-14 % at the CTA barriers, and every instruction added to its epilogue is
// on the critical path).  Here a 128-row block travels through a ring:
//   loader groups (NLG x 4 warps, thread = row): fp32 row -> [transform] -> split -> operand slot; group g takes the blocks with
//                 (block index % NLG) == g, so NLG blocks are being loaded at once;                          full[s] / empty[s]
//   MMA warp:     3 x Kt/16 MMAs per block into one of two accumulator pairs in TMEM;                          tfull[a] / tempty[a]
//   epilogue (4 warps): tcgen05.ld -> un-scale, bias, beta -> store; BatchNorm statistics of the output by a reduce-scatter
//                 butterfly over the warp's 32 rows (31 shuffles for 16 columns x {sum, sum of squares}) into per-warp double
//                 accumulators.
// Same arithmetic, operand layouts and weight prologue as pw_gemm_tc_kernel.
// =====================================================================================================================
namespace {
constexpr int kG2LG = 2;                                  // loader groups (3 groups = 544 threads leave 56 registers: spills, 15.4 ms)
constexpr int kG2EG = 1;                                  // epilogue groups of four warps (group g drains accumulator pair g of two);
                                                          // one loader group + two epilogue groups measured 15.3 ms against 14.9
constexpr int kG2Threads = 32 * (4 * kG2LG + 1 + 4 * kG2EG);
constexpr int kG2MaxS = 6;
__device__ __forceinline__ void mbar_arrive2(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
}  // namespace

template <bool F16>
__global__ void __launch_bounds__(kG2Threads, 2) pw_gemm_tc2_kernel(const GemmTcP p, const int S) {
  extern __shared__ __align__(128) uint8_t gsm2[];
  const int planes = p.Kt >> 3;
  const int slot_bytes = 2 * planes * 2048;                  // hi then lo: [plane][128 rows] x 16 B
  uint8_t* sB = gsm2;                                        // hi [Kt/8][N] x 16 B, then lo
  uint8_t* sA = gsm2 + ((2 * p.b_half + 127) & ~127);        // S slots
  float* sSc = reinterpret_cast<float*>(sA + (size_t)S * slot_bytes);
  float* sSh = sSc + p.K0;
  float* sBias = sSh + p.K0;
  double* sAcc = reinterpret_cast<double*>(sBias + p.N + ((2 * p.K0 + p.N) & 1));   // [4 * kG2EG epilogue warps][2][N]
  __shared__ __align__(8) uint64_t full_bar[kG2MaxS], empty_bar[kG2MaxS], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_s;
  __shared__ float red[16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool tfon = p.tf.gamma != nullptr;
  if (tid == 0) {
    for (int i = 0; i < S; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(s32(&full_bar[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty_bar[i])));
    }
    for (int i = 0; i < 2; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&tfull_bar[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(s32(&tempty_bar[i])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4 * kG2LG) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_s)), "r"((uint32_t)p.tcols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- weights -> B operand (as in pw_gemm_tc_kernel) ----
  float wmax = 0.f;
  if (F16) {
    for (int e = tid; e < p.Kt * p.N; e += kG2Threads) wmax = fmaxf(wmax, fabsf(weight_at(p, e % p.Kt, e / p.Kt)));
#pragma unroll
    for (int o = 16; o; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) red[warp] = wmax;
  }
  for (int i = tid; i < p.K0; i += kG2Threads) {
    float sc = 1.f, sh = 0.f;
    if (tfon && i < p.I0) { sc = p.tf.gamma[i] * p.tf.invstd[i]; sh = p.tf.beta[i] - p.tf.mean[i] * sc; }
    sSc[i] = sc; sSh[i] = sh;
  }
  for (int j = tid; j < p.N; j += kG2Threads) {
    float b = 0.f;
    if (j < p.J) {
      if (p.du.Jsplit && j >= p.du.Jsplit) b = p.du.biasb ? p.du.biasb[j - p.du.Jsplit] : 0.f;
      else b = p.bias ? p.bias[j] : 0.f;
    }
    sBias[j] = b;
  }
  for (int i = tid; i < 8 * kG2EG * p.N; i += kG2Threads) sAcc[i] = 0.0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  int e2 = 0;
  if (F16) {
    wmax = 0.f;
    for (int w = 0; w < kG2Threads / 32; ++w) wmax = fmaxf(wmax, red[w]);
    if (wmax > 0.f) { int ex; frexpf(wmax, &ex); e2 = 14 - ex; }
  }
  const float wsc = ldexpf(1.f, e2), unscale = ldexpf(1.f, -e2);
  for (int e = tid; e < (p.Kt >> 1) * p.N; e += kG2Threads) {
    const int kp = e % (p.Kt >> 1), j = e / (p.Kt >> 1), k = 2 * kp;
    uint32_t hi, lo;
    if (F16) split2_f16(weight_at(p, k, j) * wsc, weight_at(p, k + 1, j) * wsc, hi, lo);
    else split2_bf16(weight_at(p, k, j), weight_at(p, k + 1, j), hi, lo);
    const uint32_t off = (uint32_t)(k >> 3) * (uint32_t)(p.N * 16) + (uint32_t)j * 16u + (uint32_t)(k & 7) * 2u;
    *reinterpret_cast<uint32_t*>(sB + off) = hi;
    *reinterpret_cast<uint32_t*>(sB + p.b_half + off) = lo;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t tmem = tmem_s;
  const long long mtiles = (p.M + 127) / 128;
  const int Jstat = p.du.Jsplit ? p.du.Jsplit : p.J;

  if (warp < 4 * kG2LG) {
    // ===== loaders: fp32 row -> [transform] -> split -> operand slot =====
    const int grp = warp >> 2, row = (warp & 3) * 32 + lane;
    const bool vecA0 = ((p.lda0 | p.I0) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.A0) & 15) == 0;
    const bool vecA2 = p.du.A2 && ((p.du.lda2 | p.du.I2) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.du.A2) & 15) == 0;
    int s = 0; uint32_t ph = 0; int gi = 0;
    for (long long mt = blockIdx.x; mt < mtiles; mt += gridDim.x) {
      if (gi == grp) {
        const long long m = mt * 128 + row;
        uint32_t par = ph ^ 1u;
        asm volatile("{\n\t.reg .pred q;\n\tWE2:\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t@q bra DE2;\n\tbra WE2;\n\tDE2:\n\t}"
                     ::"r"(s32(&empty_bar[s])), "r"(par) : "memory");
        uint8_t* slot = sA + (size_t)s * slot_bytes;
        constexpr int UB = 4;
#pragma unroll 1
        for (int g0 = 0; g0 < planes; g0 += UB) {
          float v[UB][8];
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            const int g = g0 + u;
            if (g < planes) {
              const int k0 = g * 8;
              const bool part2 = k0 >= p.K0;
              const float* src = part2 ? p.du.A2 + m * p.du.lda2 + (k0 - p.K0) : p.A0 + m * p.lda0 + k0;
              const int valid = part2 ? p.du.I2 - (k0 - p.K0) : p.I0 - k0;
              if (m < p.M && valid >= 8 && (part2 ? vecA2 : vecA0)) {
                const float4 a = *reinterpret_cast<const float4*>(src), bq = *reinterpret_cast<const float4*>(src + 4);
                v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w; v[u][4] = bq.x; v[u][5] = bq.y; v[u][6] = bq.z; v[u][7] = bq.w;
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[u][j] = (m < p.M && j < valid) ? src[j] : 0.f;
              }
            }
          }
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            const int g = g0 + u;
            if (g >= planes) break;
            const int k0 = g * 8;
            if (tfon && k0 < p.K0 && m < p.M) {
              const float4 c0 = *reinterpret_cast<const float4*>(sSc + k0), c1 = *reinterpret_cast<const float4*>(sSc + k0 + 4);
              const float4 h0 = *reinterpret_cast<const float4*>(sSh + k0), h1 = *reinterpret_cast<const float4*>(sSh + k0 + 4);
              const float sc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w}, sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                float2 x = __ffma2_rn(make_float2(v[u][j], v[u][j + 1]), make_float2(sc[j], sc[j + 1]), make_float2(sh[j], sh[j + 1]));
                if (p.tf.act) x = silu2_f(x);
                v[u][j] = k0 + j < p.I0 ? x.x : 0.f; v[u][j + 1] = k0 + j + 1 < p.I0 ? x.y : 0.f;
              }
            }
            uint4 h, l;
            if (F16) {
              split2_f16(v[u][0], v[u][1], h.x, l.x); split2_f16(v[u][2], v[u][3], h.y, l.y);
              split2_f16(v[u][4], v[u][5], h.z, l.z); split2_f16(v[u][6], v[u][7], h.w, l.w);
            } else {
              split2_bf16(v[u][0], v[u][1], h.x, l.x); split2_bf16(v[u][2], v[u][3], h.y, l.y);
              split2_bf16(v[u][4], v[u][5], h.z, l.z); split2_bf16(v[u][6], v[u][7], h.w, l.w);
            }
            *reinterpret_cast<uint4*>(slot + g * 2048 + row * 16) = h;
            *reinterpret_cast<uint4*>(slot + planes * 2048 + g * 2048 + row * 16) = l;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive2(&full_bar[s]);
      }
      if (++gi == kG2LG) gi = 0;
      if (++s == S) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 4 * kG2LG) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (F16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t b0 = s32(sB);
      int s = 0; uint32_t ph = 0; int acc = 0; uint32_t aph = 0;
      for (long long mt = blockIdx.x; mt < mtiles; mt += gridDim.x) {
        mbar_wait_parity(&tempty_bar[acc], aph ^ 1u);
        mbar_wait_parity(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = s32(sA + (size_t)s * slot_bytes), al0 = a0 + planes * 2048;
        const uint32_t d0 = tmem + (uint32_t)(acc * 2 * p.N);
        for (int ks = 0; ks < (p.Kt >> 4); ++ks) {
          const uint64_t bh = desc_nosw(b0 + ks * 2 * p.N * 16, (uint32_t)p.N * 16u, 128u);
          const uint64_t bl = desc_nosw(b0 + p.b_half + ks * 2 * p.N * 16, (uint32_t)p.N * 16u, 128u);
          const uint64_t ah = desc_nosw(a0 + ks * 4096, 2048u, 128u), am = desc_nosw(al0 + ks * 4096, 2048u, 128u);
          umma_f16(d0, ah, bh, idesc, ks ? 1u : 0u);                  // hi . hi
          umma_f16(d0 + p.N, am, bh, idesc, ks ? 1u : 0u);            // lo . hi
          umma_f16(d0 + p.N, ah, bl, idesc, 1u);                      // hi . lo
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&empty_bar[s])) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&tfull_bar[acc])) : "memory");
        if (++s == S) { s = 0; ph ^= 1u; }
        if (++acc == 2) { acc = 0; aph ^= 1u; }
      }
    }
  } else {
    // ===== epilogue: warp w owns TMEM lanes [32q, 32q + 32), q = w & 3 =====
    const int q = warp & 3, ew = warp - (4 * kG2LG + 1);
    double* myAcc = sAcc + (size_t)ew * 2 * p.N;
    const bool vecC = ((p.ldc | p.J | p.du.Jsplit | p.du.ldcb) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(p.du.Cb) & 15) == 0;
    const int egrp = ew >> 2;
    int acc = 0; uint32_t aph = 0;
    for (long long mt = blockIdx.x; mt < mtiles; mt += gridDim.x) {
      if (kG2EG == 2 && acc != egrp) {                     // the other group's accumulator pair
        if (++acc == 2) { acc = 0; aph ^= 1u; }
        continue;
      }
      const long long em = mt * 128 + q * 32 + lane;
      mbar_wait_parity(&tfull_bar[acc], aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * p.N);
      for (int c0 = 0; c0 < p.N; c0 += 16) {
        uint32_t v[16], w[16];
        ld16(taddr + c0, v);
        ld16(taddr + p.N + c0, w);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = (__uint_as_float(v[j]) + __uint_as_float(w[j])) * unscale;
        if (p.sums && c0 < Jstat) {
          // column sums of the bias-free product over this warp's 32 rows (rows >= M hold exact zeros)
          float r[32];
#pragma unroll
          for (int j = 0; j < 16; ++j) { r[j] = f[j]; r[16 + j] = f[j] * f[j]; }
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int i = 0; i < o; ++i) {
              const float send = up ? r[i] : r[i + o], keep = up ? r[i + o] : r[i];
              r[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          myAcc[(lane >> 4) * p.N + c0 + (lane & 15)] += (double)r[0];
        }
        if (em < p.M && c0 < p.J) {
          const bool second = p.du.Jsplit && c0 >= p.du.Jsplit;
          float* crow = second ? p.du.Cb + em * p.du.ldcb + (c0 - p.du.Jsplit) : p.C + em * p.ldc + c0;
          const int nvalid = (second ? p.J : (p.du.Jsplit ? p.du.Jsplit : p.J)) - c0;
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sBias + c0 + 4 * j4);
            f[4 * j4] += b4.x; f[4 * j4 + 1] += b4.y; f[4 * j4 + 2] += b4.z; f[4 * j4 + 3] += b4.w;
          }
          if (vecC && nvalid >= 16) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              float4 o = make_float4(f[4 * j4], f[4 * j4 + 1], f[4 * j4 + 2], f[4 * j4 + 3]);
              float4* op = reinterpret_cast<float4*>(crow) + j4;
              if (p.beta) { const float4 pv = *op; o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w; }
              *op = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < nvalid) crow[j] = p.beta ? crow[j] + f[j] : f[j];
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive2(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; aph ^= 1u; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (p.sums) {
    for (int e = tid; e < 2 * p.N; e += kG2Threads) {
      const int h = e / p.N, col = e - h * p.N;
      if (col >= Jstat) continue;
      double a = 0.0;
#pragma unroll
      for (int w = 0; w < 4 * kG2EG; ++w) a += sAcc[(size_t)w * 2 * p.N + e];
      atomicAdd(&p.sums[h * Jstat + col], a);
    }
  }
  if (warp == 4 * kG2LG) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tcols) : "memory");
  }
}

// false: shape outside the envelope (two accumulator pairs must fit 256 TMEM columns, at least three ring slots in ~100 KB)
static bool launch_pw_gemm_tc2(GemmTcP p, int trans, cudaStream_t s) {
  if (4 * p.N > 256) return false;
  p.NB = 1;
  p.tcols = 32;
  while (p.tcols < 4 * p.N) p.tcols <<= 1;
  p.b_half = (p.Kt / 8) * p.N * 16;
  const size_t slot = 2 * (size_t)(p.Kt / 8) * 2048;
  const size_t fixed = ((2 * (size_t)p.b_half + 127) & ~(size_t)127) + (size_t)(2 * p.K0 + p.N + 1) * 4 + (size_t)8 * kG2EG * p.N * 8 + 256;
  // two CTAs per SM when three ring slots fit ~100 KB; wide reductions (K >= 96: 48 KB and more per slot) run one CTA per SM
  int per_sm = 2;
  int S = fixed < 100 * 1024 ? (int)((100 * 1024 - fixed) / slot) : 0;
  if (S < 3) { per_sm = 1; S = fixed < 208 * 1024 ? (int)((208 * 1024 - fixed) / slot) : 0; }
  if (S > kG2MaxS) S = kG2MaxS;
  if (S < 2) return false;                                 // K = 144 (the 129-channel concat): two 74-KB slots, one per loader group
  const size_t smem = fixed + (size_t)S * slot;
  static int sms = 0;
  if (!sms) { int d = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d); if (sms <= 0) sms = 148; }
  const long long mtiles = (p.M + 127) / 128;
  const int grid = (int)std::min<long long>(mtiles, (long long)sms * per_sm);
  static unsigned long long attr_f = 0, attr_b = 0;
  ensure_dyn_smem(pw_gemm_tc2_kernel<true>, 212 * 1024, attr_f, "pw_gemm_tc2_kernel<f16>");
  ensure_dyn_smem(pw_gemm_tc2_kernel<false>, 212 * 1024, attr_b, "pw_gemm_tc2_kernel<bf16>");
  if (trans) pw_gemm_tc2_kernel<false><<<grid, kG2Threads, smem, s>>>(p, S);
  else pw_gemm_tc2_kernel<true><<<grid, kG2Threads, smem, s>>>(p, S);
  return true;
}

// =====================================================================================================================
// Weight gradient of a 1x1 conv on tcgen05:   dW[j][i] += sum_m D[m][j] * X'[m][i]        (X' = optional input transform of X)
//
// As a GEMM the reduction runs over PIXELS: A = D^T (M' = output channels, K' = pixels), B = X'^T (N' = input channels).  Both
// tensors are pixel-major in memory, i.e. their 8-channel groups are contiguous for a fixed pixel -- exactly the canonical
// MN-MAJOR no-swizzle UMMA layout ((8 MN elements = 16 B contiguous) x (8 K rows at 16 B) core matrices): the SAME shared-memory
// tiles [8-channel plane][pixel] x 16 B that every K-major kernel of this library builds are read here with the instruction
// descriptor's a_major / b_major bits set, SBO = plane pitch (next 8 channels), LBO = 128 B (next 8 pixels).  No transposition.
//   * per 128-pixel tile: 256 threads load their rows of D and X (two threads per pixel, alternate planes), transform, split into
//     bf16 hi / mid (gradients underflow fp16), store; one thread issues 8 k-steps x 3 MMAs (M = 128, N = padded I) that
//     ACCUMULATE ACROSS TILES in TMEM (hi.hi in one accumulator, the cross terms in a second); tiles are double-buffered so the
//     next tile is loaded while the MMAs of this one run;
//   * every 32 tiles and at the end the accumulators are flushed with float atomics into dW (pre-zeroed by the caller).
// Rows j >= J of the A operand are whatever follows the D tile in shared memory: their accumulator rows are never read.
// =====================================================================================================================
namespace {
struct WgradTcP {
  const float* D; int ldd; const float* X; int ldx; float* dW; int ldw; long long M; int I, J; InTf tf;
  int JP, IP, N, tcols, tile_bytes, nbuf;    // 8-channel planes of D and X (IP even), padded I, TMEM columns, bytes per tile buffer
};
}  // namespace

__global__ void __launch_bounds__(256, 2) pw_wgrad_tc_kernel(const WgradTcP p) {
  extern __shared__ __align__(128) uint8_t wsm[];
  // tile buffer b: [D hi JP planes][X hi IP planes][D mid JP planes][X mid IP planes], plane = 128 pixels x 16 B
  float* sSc = reinterpret_cast<float*>(wsm + (size_t)p.nbuf * p.tile_bytes + 32 * 1024);   // behind the buffers + the A over-read margin
  float* sSh = sSc + p.IP * 8;
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool tfon = p.tf.gamma != nullptr;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_s)), "r"((uint32_t)p.tcols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < p.IP * 8; i += 256) {
    float sc = 1.f, sh = 0.f;
    if (tfon && i < p.I) { sc = p.tf.gamma[i] * p.tf.invstd[i]; sh = p.tf.beta[i] - p.tf.mean[i] * sc; }
    sSc[i] = sc; sSh[i] = sh;
  }
  // the over-read margin of the A operand must hold finite numbers nowhere in particular; zero it once (it also covers the
  // first use of a buffer's unused tail)
  for (int i = tid; i < (p.nbuf * p.tile_bytes + 32 * 1024) / 16; i += 256) reinterpret_cast<uint4*>(wsm)[i] = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  // D = F32, A = B = BF16, both MN-major (bits 15, 16), N = padded I, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const int row = tid & 127, hsel = tid >> 7;
  const bool vecD = ((p.ldd | p.J) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.D) & 15) == 0;
  const bool vecX = ((p.ldx | p.I) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.X) & 15) == 0;
  const int planes = p.JP + p.IP, half_bytes = planes * 2048;
  const long long mtiles = (p.M + 127) / 128;
  uint32_t par[2] = {0u, 0u};
  int used[2] = {0, 0};                                    // MMAs in flight on a buffer
  bool fresh = true;                                       // the accumulators start from zero
  int since_flush = 0, t = 0;
  auto flush = [&]() {
    // consume the outstanding per-buffer commits (a commit covers every MMA issued before it): afterwards nothing is in flight
    for (int bb = 0; bb < 2; ++bb)
      if (used[bb]) { mbar_wait_parity(&bar[bb], par[bb]); par[bb] ^= 1; used[bb] = 0; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3, chalf = warp >> 2, j = q * 32 + lane;
    for (int c0 = chalf * 16; c0 < p.N; c0 += 32) {
      uint32_t v[16], w[16];
      ld16(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
      ld16(tmem + ((uint32_t)(q * 32) << 16) + p.N + c0, w);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (j < p.J) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < p.I) atomicAdd(&p.dW[(size_t)j * p.ldw + c0 + i], __uint_as_float(v[i]) + __uint_as_float(w[i]));
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    fresh = true; since_flush = 0;
  };
#pragma unroll 1
  for (long long mt = blockIdx.x; mt < mtiles; mt += gridDim.x, ++t) {
    const int b = p.nbuf > 1 ? (t & 1) : 0;
    if (used[b]) {                                         // the MMAs that last read this buffer
      mbar_wait_parity(&bar[b], par[b]);
      par[b] ^= 1; used[b] = 0;
    }
    uint8_t* buf = wsm + (size_t)b * p.tile_bytes;
    const long long m = mt * 128 + row;
    constexpr int UB = 4;
    const int items = (planes - hsel + 1) >> 1;
#pragma unroll 1
    for (int it0 = 0; it0 < items; it0 += UB) {
      float v[UB][8];
      int gs[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int it = it0 + u;
        gs[u] = -1;
        if (it < items) {
          const int g = hsel + 2 * it;
          gs[u] = g;
          const bool isx = g >= p.JP;
          const int k0 = (isx ? g - p.JP : g) * 8;
          const float* src = isx ? p.X + m * p.ldx + k0 : p.D + m * p.ldd + k0;
          const int valid = (isx ? p.I : p.J) - k0;
          if (m < p.M && valid >= 8 && (isx ? vecX : vecD)) {
            const float4 a = *reinterpret_cast<const float4*>(src), bq = *reinterpret_cast<const float4*>(src + 4);
            v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w; v[u][4] = bq.x; v[u][5] = bq.y; v[u][6] = bq.z; v[u][7] = bq.w;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[u][j] = (m < p.M && j < valid) ? src[j] : 0.f;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        if (gs[u] < 0) continue;
        const int g = gs[u];
        if (tfon && g >= p.JP && m < p.M) {
          const int k0 = (g - p.JP) * 8;
          const float4 c0 = *reinterpret_cast<const float4*>(sSc + k0), c1 = *reinterpret_cast<const float4*>(sSc + k0 + 4);
          const float4 h0 = *reinterpret_cast<const float4*>(sSh + k0), h1 = *reinterpret_cast<const float4*>(sSh + k0 + 4);
          const float sc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w}, sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            float2 x = __ffma2_rn(make_float2(v[u][j], v[u][j + 1]), make_float2(sc[j], sc[j + 1]), make_float2(sh[j], sh[j + 1]));
            if (p.tf.act) x = silu2_f(x);
            v[u][j] = k0 + j < p.I ? x.x : 0.f; v[u][j + 1] = k0 + j + 1 < p.I ? x.y : 0.f;
          }
        }
        uint4 h, l;
        split2_bf16(v[u][0], v[u][1], h.x, l.x); split2_bf16(v[u][2], v[u][3], h.y, l.y);
        split2_bf16(v[u][4], v[u][5], h.z, l.z); split2_bf16(v[u][6], v[u][7], h.w, l.w);
        *reinterpret_cast<uint4*>(buf + g * 2048 + row * 16) = h;
        *reinterpret_cast<uint4*>(buf + half_bytes + g * 2048 + row * 16) = l;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t dh = s32(buf), xh = dh + p.JP * 2048, dm = dh + half_bytes, xm = dm + p.JP * 2048;
      for (int ks = 0; ks < 8; ++ks) {                       // 16 pixels per step: +256 B inside every plane
        const uint64_t ah = desc_nosw(dh + ks * 256, 128u, 2048u), am = desc_nosw(dm + ks * 256, 128u, 2048u);
        const uint64_t bh = desc_nosw(xh + ks * 256, 128u, 2048u), bm = desc_nosw(xm + ks * 256, 128u, 2048u);
        const uint32_t acc0 = (fresh && ks == 0) ? 0u : 1u;
        umma_f16(tmem, ah, bh, idesc, acc0);                 // hi . hi
        umma_f16(tmem + p.N, am, bh, idesc, acc0);           // mid . hi
        umma_f16(tmem + p.N, ah, bm, idesc, 1u);             // hi . mid
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar[b])) : "memory");
    }
    used[b] = 1; fresh = false;
    if (++since_flush == 32) flush();
  }
  if (since_flush) flush();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tcols) : "memory");
  }
}

bool launch_pw_wgrad_tc(const float* D, int ldd, const float* X, int ldx, float* dW, int ldw, long long M, int I, int J, cudaStream_t s,
                        InTf tf) {
  static const bool off = getenv("YSP_TRAIN_NO_TC") != nullptr || getenv("YSP_TRAIN_NO_WGRAD_TC") != nullptr;
  // widest layers only (I + J >= 128; measured 19.16 -> 18.95 ms per step, any lower threshold was slower): the FFMA kernel already streams the narrow full-resolution ones (J = 16, I = 32) at 3-4 TB/s, and 24 MMAs
  // with 7/8 of the M = 128 rows unused per 24 KB tile would make those tensor-pipe-bound
  static const int minw = getenv("YSP_WGRAD_TC_MINW") ? atoi(getenv("YSP_WGRAD_TC_MINW")) : 128;
  if (off || M < 4096 || J > 128 || J < 8 || I < 8 || I + J < minw) return false;
  WgradTcP p = {};
  p.D = D; p.ldd = ldd; p.X = X; p.ldx = ldx; p.dW = dW; p.ldw = ldw; p.M = M; p.I = I; p.J = J; p.tf = tf;
  p.JP = (J + 7) / 8;
  p.N = (I + 15) / 16 * 16;
  p.IP = p.N / 8;
  if (p.N > 256) return false;
  p.tcols = 32;
  while (p.tcols < 2 * p.N) p.tcols <<= 1;
  p.tile_bytes = 2 * (p.JP + p.IP) * 2048;
  p.nbuf = 2 * (size_t)p.tile_bytes + 32 * 1024 + (size_t)p.IP * 64 + 256 <= 200 * 1024 ? 2 : 1;
  const size_t smem = (size_t)p.nbuf * p.tile_bytes + 32 * 1024 + (size_t)p.IP * 64 + 256;
  if (smem > 200 * 1024) return false;
  static unsigned long long attr_done = 0;
  ensure_dyn_smem(pw_wgrad_tc_kernel, 200 * 1024, attr_done, "pw_wgrad_tc_kernel");
  static int sms = 0;
  if (!sms) { int d = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d); if (sms <= 0) sms = 148; }
  int per_sm = (int)std::min<size_t>(std::min<size_t>(512 / p.tcols, (220 * 1024) / (smem + 1024)), 2);
  if (per_sm < 1) per_sm = 1;
  const long long mtiles = (M + 127) / 128;
  const int grid = (int)std::min<long long>(mtiles, (long long)sms * per_sm);
  pw_wgrad_tc_kernel<<<grid, 256, smem, s>>>(p);
  return true;
}

}  // namespace ysp
