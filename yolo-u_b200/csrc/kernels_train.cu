// kernels_train.cu -- fp32 CUDA-core kernels of the seg-head TRAINING step (BASELINE cfg 4).
//
// The trainable part of YOLO-Seg++ is the 63.8 K-parameter decoder (reference train.py:256-267 excludes `encoder.*`),
// run in train() mode: every ultralytics Conv is conv(no bias) -> BatchNorm2d with BATCH statistics -> SiLU/identity.
// The work is HBM-bound (full-resolution activations, channel counts 8..129), so these are coalesced NHWC kernels:
//   pw_gemm / pw_wgrad          1x1 convs: forward, input-gradient (same kernel, transposed weight view), weight-gradient
//   dw_conv / dw_wgrad          depthwise k x k: forward, input-gradient (flipped taps), weight-gradient
//   col_reduce<MODE>            per-channel double-precision reductions: BN statistics, BN backward sums, bias
//                               gradients, ECA pooling and its backward
//   bn_finalize / bn_apply / bn_bwd_apply, up2 / up2_bwd (bilinear, align_corners=False), ECA gate fwd/bwd,
//   Dice(+BCE) loss fwd/bwd (monai DiceLoss(sigmoid, soft_label, batch=True), train.py:98-104), AdamW.
// Weights stay in PyTorch's own layouts ([Cout][Cin] and [C][k*k]) so the flat parameter buffer IS the state_dict.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "kernels.h"
#include "kernels_train.h"

namespace ysp {

static inline int cdivl(long long a, long long b) { return (int)((a + b - 1) / b); }

template <int K, int MODE>
static void dw_tiled_launch(const float* X, int ldx, const float* W, float* Y, int ldy, const float* D, int ldd, float* dW,
                            double* sums, int N, int H, int Wd, int C, int beta, cudaStream_t s, InTf tf = InTf());

// =====================================================================================================================
// C[m][j] = beta*C[m][j] + bias[j] + sum_i A[m][i] * Wop(i,j)      Wop(i,j) = trans ? W[i*ldw + j] : W[j*ldw + i]
//   forward 1x1 conv:   i = ci, j = co, trans = 0 (W = [Cout][Cin]);   input gradient: i = co, j = ci, trans = 1.
// Thread micro-tile 4x4; BN in {16,32,64} columns per CTA and 4*(256/(BN/4)) rows, so narrow outputs waste nothing.
// =====================================================================================================================
template <int BN>
__global__ void __launch_bounds__(256, 4) pw_gemm_kernel(const float* __restrict__ A0, int lda0, const float* __restrict__ W0,
                                                      int ldw0, int trans, const float* __restrict__ bias, float* C,
                                                      int ldc, long long M, int I0, int J, int beta, double* sums, InTf tf,
                                                      PwDual du) {
  constexpr int BK = 16, TX = BN / 4, TY = 256 / TX, BM = TY * 4;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  __shared__ double sRed[2][BN];
  __shared__ float sSc[144], sSh[144];                 // input transform x = act(z*sc + sh), I <= 144
  const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
  const bool tfon = tf.gamma != nullptr;
  if (tfon) {
    for (int i = tid; i < I0; i += 256) {
      const float sc = tf.gamma[i] * tf.invstd[i];
      sSc[i] = sc; sSh[i] = tf.beta[i] - tf.mean[i] * sc;
    }
    __syncthreads();
  }
  const int j0 = blockIdx.y * BN;
  const int Jstat = du.Jsplit ? du.Jsplit : J;         // columns that carry BN statistics
  const bool vecC = ((ldc | J | du.Jsplit | du.ldcb) & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(du.Cb) & 15) == 0;
  const long long mtiles = (M + BM - 1) / BM;
  const int nparts = du.A2 ? 2 : 1;
  float cs1[4] = {0.f, 0.f, 0.f, 0.f}, cs2[4] = {0.f, 0.f, 0.f, 0.f};   // column sums of this thread's outputs (BN statistics)
  for (long long mt = blockIdx.x; mt < mtiles; mt += gridDim.x) {
    const long long m0 = mt * BM;
    float acc[4][4] = {};
    for (int part = 0; part < nparts; ++part) {
      const float* __restrict__ A = part ? du.A2 : A0;
      const float* __restrict__ W = part ? du.W2 : W0;
      const int lda = part ? du.lda2 : lda0, ldw = part ? du.ldw2 : ldw0, I = part ? du.I2 : I0;
      const bool vecA = ((lda | I) & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
      const bool xf = tfon && part == 0;
      for (int i0 = 0; i0 < I; i0 += BK) {
        if (vecA) {
          // all of this thread's 16-byte loads of the chunk are issued before the first shared-memory store
          constexpr int NL = BM * (BK / 4) / 256;
          float4 va[NL];
#pragma unroll
          for (int u = 0; u < NL; ++u) {
            const int e = tid + u * 256;
            const int r = e / (BK / 4), i = (e % (BK / 4)) * 4;
            const long long m = m0 + r;
            va[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < M && i0 + i < I) va[u] = *reinterpret_cast<const float4*>(A + m * lda + i0 + i);      // I % 4 == 0
          }
#pragma unroll
          for (int u = 0; u < NL; ++u) {
            const int e = tid + u * 256;
            const int r = e / (BK / 4), i = (e % (BK / 4)) * 4;
            float4 v = va[u];
            if (xf && m0 + r < M && i0 + i < I) {
              const int ig = i0 + i;
              v.x = fmaf(v.x, sSc[ig], sSh[ig]); v.y = fmaf(v.y, sSc[ig + 1], sSh[ig + 1]);
              v.z = fmaf(v.z, sSc[ig + 2], sSh[ig + 2]); v.w = fmaf(v.w, sSc[ig + 3], sSh[ig + 3]);
              if (tf.act) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
            }
            As[i][r] = v.x; As[i + 1][r] = v.y; As[i + 2][r] = v.z; As[i + 3][r] = v.w;
          }
        } else {
          for (int e = tid; e < BM * BK; e += 256) {
            int r = e / BK, i = e % BK;
            long long m = m0 + r;
            float v = 0.f;
            if (m < M && i0 + i < I) {
              v = A[m * lda + i0 + i];
              if (xf) { v = fmaf(v, sSc[i0 + i], sSh[i0 + i]); if (tf.act) v = silu_f(v); }
            }
            As[i][r] = v;
          }
        }
        for (int e = tid; e < BK * BN; e += 256) {
          int i = e / BN, jj = e % BN;
          int ig = i0 + i, jg = j0 + jj;
          float w = 0.f;
          if (ig < I && jg < J) {
            if (du.Jsplit && jg >= du.Jsplit) w = du.Wb[(size_t)(jg - du.Jsplit) * du.ldwb + ig];      // forward layout [Cout][Cin]
            else w = trans ? W[(size_t)ig * ldw + jg] : W[(size_t)jg * ldw + ig];
          }
          Bs[i][jj] = w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
          float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
          float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
          const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int r = 0; r < 4; ++r) {      // FFMA2 (sm_100 packed fp32 FMA, scalar a broadcast): 8 instead of 16 issue slots per k
            if (BN == 16) {                    // (the 64-register BN = 16 variant spills with register pairs: scalar FMAs there)
#pragma unroll
              for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
              continue;
            }
            const float2 lo = __ffma2_rn(make_float2(av[r], av[r]), make_float2(bv[0], bv[1]), make_float2(acc[r][0], acc[r][1]));
            const float2 hi = __ffma2_rn(make_float2(av[r], av[r]), make_float2(bv[2], bv[3]), make_float2(acc[r][2], acc[r][3]));
            acc[r][0] = lo.x; acc[r][1] = lo.y; acc[r][2] = hi.x; acc[r][3] = hi.y;
          }
        }
        __syncthreads();
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      long long m = m0 + ty * 4 + r;
      if (m >= M) continue;
      const int jb = j0 + tx * 4;
      if (sums && jb < Jstat) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { cs1[c] += acc[r][c]; cs2[c] = fmaf(acc[r][c], acc[r][c], cs2[c]); }
      }
      const bool second = du.Jsplit && jb >= du.Jsplit;
      float* crow = second ? du.Cb + m * du.ldcb + (jb - du.Jsplit) : C + m * ldc + jb;
      const float* brow = second ? (du.biasb ? du.biasb + (jb - du.Jsplit) : nullptr) : (bias ? bias + jb : nullptr);
      if (vecC) {
        if (jb >= J) continue;
        float4* o = reinterpret_cast<float4*>(crow);
        float4 v = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        if (brow) { v.x += brow[0]; v.y += brow[1]; v.z += brow[2]; v.w += brow[3]; }
        if (beta) { float4 pv = *o; v.x += pv.x; v.y += pv.y; v.z += pv.z; v.w += pv.w; }
        *o = v;
        continue;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int jx = jb + c;
        if (jx >= J) continue;
        float v = acc[r][c] + (brow ? brow[c] : 0.f);
        float* o = crow + c;
        *o = beta ? *o + v : v;
      }
    }
  }
  if (sums) {   // BatchNorm statistics of the (bias-free) output: warp shuffles over equal tx, then shared, then global atomics
    for (int e = tid; e < 2 * BN; e += 256) sRed[e / BN][e % BN] = 0.0;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double a1 = cs1[c], a2 = cs2[c];
#pragma unroll
      for (int m = TX; m < 32; m <<= 1) { a1 += __shfl_xor_sync(0xffffffffu, a1, m); a2 += __shfl_xor_sync(0xffffffffu, a2, m); }
      if ((tid & 31) < TX) { atomicAdd(&sRed[0][tx * 4 + c], a1); atomicAdd(&sRed[1][tx * 4 + c], a2); }
    }
    __syncthreads();
    for (int e = tid; e < 2 * BN; e += 256)
      if (j0 + e % BN < Jstat) atomicAdd(&sums[(e / BN) * Jstat + j0 + e % BN], sRed[e / BN][e % BN]);
  }
}

// resident CTAs per SM of a kernel (queried once): persistent grids are sized to exactly one full wave
template <typename Kern>
static int resident_ctas(Kern kern, int threads, size_t smem) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem) != cudaSuccess || n < 1) n = 1;
  return n;
}

template <int BN>
static void pw_gemm_launch(const float* A, int lda, const float* W, int ldw, int trans, const float* bias, float* C, int ldc,
                           long long M, int I, int J, int beta, cudaStream_t s, double* sums, InTf tf, PwDual du) {
  constexpr int BM = (256 / (BN / 4)) * 4;
  static const int occ = resident_ctas(pw_gemm_kernel<BN>, 256, 0);
  const int jt = cdivl(J, BN);
  const int cap = std::max(1, 148 * occ / jt);
  pw_gemm_kernel<BN><<<dim3(std::min(cdivl(M, BM), cap), jt), 256, 0, s>>>(A, lda, W, ldw, trans, bias, C, ldc, M, I, J, beta, sums, tf, du);
}

void launch_pw_gemm(const float* A, int lda, const float* W, int ldw, int trans, const float* bias, float* C, int ldc,
                    long long M, int I, int J, int beta, cudaStream_t s, double* sums, InTf tf, PwDual du) {
  if (launch_pw_gemm_tc(A, lda, W, ldw, trans, bias, C, ldc, M, I, J, beta, s, sums, tf, du)) return;   // tensor cores (fp16 hi/lo splits)
  if (J <= 16) pw_gemm_launch<16>(A, lda, W, ldw, trans, bias, C, ldc, M, I, J, beta, s, sums, tf, du);
  else if (J <= 32) pw_gemm_launch<32>(A, lda, W, ldw, trans, bias, C, ldc, M, I, J, beta, s, sums, tf, du);
  else pw_gemm_launch<64>(A, lda, W, ldw, trans, bias, C, ldc, M, I, J, beta, s, sums, tf, du);
}

// =====================================================================================================================
// The 1x1 mask head (C -> 1, YOLOSegPlusPlus.py:178) as two streaming kernels: as GEMMs its forward (N = 1) ran on the FFMA
// kernel at 1.7 TB/s and its input gradient (K = 1, an outer product) on the tensor-core kernel at 2.6 TB/s, both bound by
// instruction issue.   y[m] = b + sum_c x[m][c] * w[c];   dx[m][c] = dy[m] * w[c].   One thread per pixel, C % 4 == 0, C <= 64.
// =====================================================================================================================
__global__ void __launch_bounds__(256) lin1_fwd_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ w,
                                                       const float* __restrict__ b, float* __restrict__ Y, int C, long long M) {
  __shared__ __align__(16) float sw[64];
  if (threadIdx.x < C) sw[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  const float bias = b ? b[0] : 0.f;
  for (long long m = (long long)blockIdx.x * 256 + threadIdx.x; m < M; m += (long long)gridDim.x * 256) {
    const float4* xp = reinterpret_cast<const float4*>(X + m * ldx);
    float acc = bias;
    for (int c4 = 0; c4 < C / 4; ++c4) {
      const float4 v = xp[c4], ww = *reinterpret_cast<const float4*>(sw + 4 * c4);
      acc = fmaf(v.x, ww.x, acc); acc = fmaf(v.y, ww.y, acc); acc = fmaf(v.z, ww.z, acc); acc = fmaf(v.w, ww.w, acc);
    }
    Y[m] = acc;
  }
}
__global__ void __launch_bounds__(256) lin1_dgrad_kernel(const float* __restrict__ DY, const float* __restrict__ w,
                                                         float* __restrict__ DX, int ldx, int C, long long M) {
  __shared__ __align__(16) float sw[64];
  if (threadIdx.x < C) sw[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  for (long long m = (long long)blockIdx.x * 256 + threadIdx.x; m < M; m += (long long)gridDim.x * 256) {
    const float d = DY[m];
    float4* op = reinterpret_cast<float4*>(DX + m * ldx);
    for (int c4 = 0; c4 < C / 4; ++c4) {
      const float4 ww = *reinterpret_cast<const float4*>(sw + 4 * c4);
      op[c4] = make_float4(d * ww.x, d * ww.y, d * ww.z, d * ww.w);
    }
  }
}
bool launch_lin1_fwd(const float* X, int ldx, const float* w, const float* b, float* Y, int C, long long M, cudaStream_t s) {
  if ((C & 3) || C > 64 || (ldx & 3) || (reinterpret_cast<uintptr_t>(X) & 15)) return false;
  lin1_fwd_kernel<<<(int)std::min<long long>(cdivl(M, 256), 148 * 16), 256, 0, s>>>(X, ldx, w, b, Y, C, M);
  return true;
}
bool launch_lin1_dgrad(const float* DY, const float* w, float* DX, int ldx, int C, long long M, cudaStream_t s) {
  if ((C & 3) || C > 64 || (ldx & 3) || (reinterpret_cast<uintptr_t>(DX) & 15)) return false;
  lin1_dgrad_kernel<<<(int)std::min<long long>(cdivl(M, 256), 148 * 16), 256, 0, s>>>(DY, w, DX, ldx, C, M);
  return true;
}

// =====================================================================================================================
// dW[j][i] += sum_m D[m][j] * X[m][i]   (weight gradient of a 1x1 conv; dW in PyTorch layout, ld = I)
// CTA = one (TJ x TI) tile of dW and one chunk of rows; the 256 threads form G = 256/((TJ/4)(TI/4)) row groups, each
// with a 4x4 micro-tile; groups combine through shared-memory atomics, CTAs through global atomics (dW pre-zeroed).
// =====================================================================================================================
template <int TJ, int TI>
__global__ void __launch_bounds__(256) pw_wgrad_kernel(const float* __restrict__ D, int ldd, const float* __restrict__ X,
                                                       int ldx, float* dW, int ldw, long long M, int I, int J,
                                                       int rows_per_cta, InTf tf) {
  constexpr int TPG = (TJ / 4) * (TI / 4), G = 256 / TPG;
  __shared__ float sSc[TI], sSh[TI];
  const bool tfon = tf.gamma != nullptr;
  constexpr int RS = (TJ + TI <= 48) ? 128 : (TJ + TI <= 96 ? 64 : 32);                       // rows staged per pass
  __shared__ __align__(16) float sD[RS][TJ + 4];
  __shared__ __align__(16) float sX[RS][TI + 4];
  __shared__ float sAcc[TJ * TI];
  const int tid = threadIdx.x, g = tid / TPG, t = tid % TPG;
  const int tj = t / (TI / 4), ti = t % (TI / 4);
  const int j0 = blockIdx.y * TJ, i0 = blockIdx.z * TI;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = (r0 + rows_per_cta < M) ? r0 + rows_per_cta : M;
  const bool vecD = (ldd & 3) == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0;
  const bool vecX = (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0;
  for (int e = tid; e < TJ * TI; e += 256) sAcc[e] = 0.f;
  if (tfon)
    for (int e = tid; e < TI; e += 256) {
      const int i = i0 + e;
      const float sc = i < I ? tf.gamma[i] * tf.invstd[i] : 0.f;
      sSc[e] = sc; sSh[e] = i < I ? tf.beta[i] - tf.mean[i] * sc : 0.f;
    }
  float acc[4][4] = {};
  // one staged tile: rows [rb, rb+RS) x columns [c0, c0+TC) of a row-major matrix, zero-filled outside
  // one staged tile: rows [rb, rb+RS) x columns [c0, c0+TC) of a row-major matrix, zero-filled outside.  `fetch` issues
  // the global loads into registers, `put` (transform +) stores them: both matrices are fetched before anything is stored,
  // so a thread has 4-6 independent 16-byte loads in flight instead of one.
  auto fetch = [&](const float* __restrict__ src, int ld, int c0, int cols, bool vec, int q4, int e, long long rb) -> float4 {
    const int r = e / q4, c = (e % q4) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rb + r < r1) {
      const float* p = src + (rb + r) * ld + c0 + c;
      if (vec && c0 + c + 3 < cols) v = *reinterpret_cast<const float4*>(p);
      else {
        if (c0 + c < cols) v.x = p[0];
        if (c0 + c + 1 < cols) v.y = p[1];
        if (c0 + c + 2 < cols) v.z = p[2];
        if (c0 + c + 3 < cols) v.w = p[3];
      }
    }
    return v;
  };
  auto put = [&](float4 v, float* dst, int dld, int q4, int e, long long rb, bool xf) {
    const int r = e / q4, c = (e % q4) * 4;
    if (xf && rb + r < r1) {       // columns beyond the matrix have sc = sh = 0 and stay 0 (SiLU(0) = 0)
      v.x = fmaf(v.x, sSc[c], sSh[c]); v.y = fmaf(v.y, sSc[c + 1], sSh[c + 1]);
      v.z = fmaf(v.z, sSc[c + 2], sSh[c + 2]); v.w = fmaf(v.w, sSc[c + 3], sSh[c + 3]);
      if (tf.act) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
    }
    *reinterpret_cast<float4*>(dst + r * dld + c) = v;
  };
  constexpr int NLD = RS * (TJ / 4) / 256, NLX = RS * (TI / 4) / 256;
  static_assert(NLD >= 1 && NLX >= 1 && RS * (TJ / 4) % 256 == 0 && RS * (TI / 4) % 256 == 0, "staging loop shape");
  for (long long rb = r0; rb < r1; rb += RS) {
    __syncthreads();
    float4 vd[NLD], vx[NLX];
#pragma unroll
    for (int u = 0; u < NLD; ++u) vd[u] = fetch(D, ldd, j0, J, vecD, TJ / 4, tid + u * 256, rb);
#pragma unroll
    for (int u = 0; u < NLX; ++u) vx[u] = fetch(X, ldx, i0, I, vecX, TI / 4, tid + u * 256, rb);
#pragma unroll
    for (int u = 0; u < NLD; ++u) put(vd[u], &sD[0][0], TJ + 4, TJ / 4, tid + u * 256, rb, false);
#pragma unroll
    for (int u = 0; u < NLX; ++u) put(vx[u], &sX[0][0], TI + 4, TI / 4, tid + u * 256, rb, tfon);
    __syncthreads();
#pragma unroll 4
    for (int r = g; r < RS; r += G) {
      float4 d = *reinterpret_cast<const float4*>(&sD[r][tj * 4]);
      float4 x = *reinterpret_cast<const float4*>(&sX[r][ti * 4]);
      const float dv[4] = {d.x, d.y, d.z, d.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float2 lo = __ffma2_rn(make_float2(dv[a], dv[a]), make_float2(xv[0], xv[1]), make_float2(acc[a][0], acc[a][1]));
        const float2 hi = __ffma2_rn(make_float2(dv[a], dv[a]), make_float2(xv[2], xv[3]), make_float2(acc[a][2], acc[a][3]));
        acc[a][0] = lo.x; acc[a][1] = lo.y; acc[a][2] = hi.x; acc[a][3] = hi.y;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) atomicAdd(&sAcc[(tj * 4 + a) * TI + ti * 4 + b], acc[a][b]);
  __syncthreads();
  for (int e = tid; e < TJ * TI; e += 256) {
    int j = j0 + e / TI, i = i0 + e % TI;
    if (j < J && i < I) atomicAdd(&dW[(size_t)j * ldw + i], sAcc[e]);
  }
}

template <int TJ, int TI>
static void pw_wgrad_launch(const float* D, int ldd, const float* X, int ldx, float* dW, int ldw, long long M, int I,
                            int J, cudaStream_t s, InTf tf) {
  int tiles = cdivl(J, TJ) * cdivl(I, TI);
  long long want = (148 * 4 + tiles - 1) / tiles;                       // ~4 CTAs per SM in total
  long long rows = (M + want - 1) / want;
  rows = ((rows + 127) / 128) * 128;
  if (rows < 256) rows = 256;
  pw_wgrad_kernel<TJ, TI><<<dim3(cdivl(M, rows), cdivl(J, TJ), cdivl(I, TI)), 256, 0, s>>>(D, ldd, X, ldx, dW, ldw, M, I,
                                                                                          J, (int)rows, tf);
}

void launch_pw_wgrad(const float* D, int ldd, const float* X, int ldx, float* dW, int ldw, long long M, int I, int J,
                     cudaStream_t s, InTf tf) {
  if (launch_pw_wgrad_tc(D, ldd, X, ldx, dW, ldw, M, I, J, s, tf)) return;      // tensor cores (bf16 hi/mid splits, MN-major operands)
  const int cj = J <= 16 ? 16 : J <= 32 ? 32 : 64, ci = I <= 16 ? 16 : I <= 32 ? 32 : 64;
#define YSP_WG(a, b) if (cj == a && ci == b) return pw_wgrad_launch<a, b>(D, ldd, X, ldx, dW, ldw, M, I, J, s, tf)
  YSP_WG(16, 16); YSP_WG(16, 32); YSP_WG(16, 64); YSP_WG(32, 16); YSP_WG(32, 32); YSP_WG(32, 64);
  YSP_WG(64, 16); YSP_WG(64, 32); YSP_WG(64, 64);
#undef YSP_WG
}

// =====================================================================================================================
// Depthwise k x k, stride 1, "same" padding, weights [C][k*k] (PyTorch [C,1,k,k]).  flip = 1 gives the input gradient.
// One thread = one pixel x 4 channels (float4); taps staged transposed in shared memory.
// =====================================================================================================================
__global__ void __launch_bounds__(256) dw_conv_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W,
                                                      float* Y, int ldy, int N, int H, int Wd, int C, int k, int flip,
                                                      int beta) {
  extern __shared__ float sW[];   // [k*k][C]
  const int kk = k * k, nthr = blockDim.x * blockDim.y, tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int e = tid; e < kk * C; e += nthr) {
    int c = e / kk, tp = e % kk;
    sW[(flip ? kk - 1 - tp : tp) * C + c] = W[e];
  }
  __syncthreads();
  const int pad = k >> 1, q = threadIdx.x;
  const unsigned total = (unsigned)N * H * Wd;
  for (unsigned pix = blockIdx.x * blockDim.y + threadIdx.y; pix < total; pix += gridDim.x * blockDim.y) {
    const int x = pix % (unsigned)Wd;
    const unsigned t = pix / (unsigned)Wd;
    const int y = t % (unsigned)H;
    const float* xb = X + (size_t)(pix - (unsigned)(y * Wd + x)) * ldx + q * 4;   // image base
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < k; ++r) {
      int iy = y + r - pad;
      if (iy < 0 || iy >= H) continue;
      for (int s = 0; s < k; ++s) {
        int ix = x + s - pad;
        if (ix < 0 || ix >= Wd) continue;
        float4 v = *reinterpret_cast<const float4*>(xb + (size_t)(iy * Wd + ix) * ldx);
        float4 w = *reinterpret_cast<const float4*>(sW + (r * k + s) * C + q * 4);
        acc.x = fmaf(v.x, w.x, acc.x); acc.y = fmaf(v.y, w.y, acc.y);
        acc.z = fmaf(v.z, w.z, acc.z); acc.w = fmaf(v.w, w.w, acc.w);
      }
    }
    float4* o = reinterpret_cast<float4*>(Y + (size_t)pix * ldy + q * 4);
    if (beta) { float4 p = *o; acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w; }
    *o = acc;
  }
}

// block = (C/4 channel quads, 256/(C/4) rows): consecutive threads walk consecutive quads of consecutive rows, so a warp's
// accesses are contiguous whenever the row stride equals C, and no kernel needs an integer division per element.
static inline dim3 quad_block(int C) { int Q = C / 4; return dim3(Q, 256 / Q); }
static inline int quad_grid(long long rows, int C, int per_sm = 16) {
  int rpb = 256 / (C / 4);
  return (int)std::min<long long>(cdivl(rows, rpb), 148 * per_sm);
}

void launch_dw_conv(const float* X, int ldx, const float* W, float* Y, int ldy, int N, int H, int Wd, int C, int k,
                    int flip, int beta, cudaStream_t s) {
  long long total = (long long)N * H * Wd;
  if (flip && dw_tiled_shape(H, Wd, k)) {
    if (k == 3) dw_tiled_launch<3, 1>(X, ldx, W, Y, ldy, nullptr, 0, nullptr, nullptr, N, H, Wd, C, beta, s);
    else dw_tiled_launch<5, 1>(X, ldx, W, Y, ldy, nullptr, 0, nullptr, nullptr, N, H, Wd, C, beta, s);
    return;
  }
  dw_conv_kernel<<<quad_grid(total, C), quad_block(C), (size_t)k * k * C * 4, s>>>(X, ldx, W, Y, ldy, N, H, Wd, C, k, flip, beta);
}

// dW[c][tap] += sum_{n,y,x} D[n,y,x,c] * X[n, y+r-pad, x+s-pad, c]
template <int K>
__global__ void __launch_bounds__(256) dw_wgrad_kernel(const float* __restrict__ D, int ldd, const float* __restrict__ X,
                                                       int ldx, float* dW, int N, int H, int Wd, int C,
                                                       unsigned pix_per_cta) {
  constexpr int KK = K * K, PAD = K / 2;
  extern __shared__ float sAcc[];   // [C][KK]
  const int nthr = blockDim.x * blockDim.y, tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int e = tid; e < C * KK; e += nthr) sAcc[e] = 0.f;
  __syncthreads();
  const int q = threadIdx.x;
  const unsigned total = (unsigned)N * H * Wd;
  const unsigned p0 = blockIdx.x * pix_per_cta;
  const unsigned p1 = (p0 + pix_per_cta < total) ? p0 + pix_per_cta : total;
  float acc[KK][4];
#pragma unroll
  for (int tp = 0; tp < KK; ++tp) acc[tp][0] = acc[tp][1] = acc[tp][2] = acc[tp][3] = 0.f;
  for (unsigned pix = p0 + threadIdx.y; pix < p1; pix += blockDim.y) {
    const int x = pix % (unsigned)Wd;
    const unsigned t = pix / (unsigned)Wd;
    const int y = t % (unsigned)H;
    const float* xb = X + (size_t)(pix - (unsigned)(y * Wd + x)) * ldx + q * 4;
    float4 d = *reinterpret_cast<const float4*>(D + (size_t)pix * ldd + q * 4);
#pragma unroll
    for (int r = 0; r < K; ++r) {
      int iy = y + r - PAD;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int s = 0; s < K; ++s) {
        int ix = x + s - PAD;
        if (ix < 0 || ix >= Wd) continue;
        float4 v = *reinterpret_cast<const float4*>(xb + (size_t)(iy * Wd + ix) * ldx);
        acc[r * K + s][0] = fmaf(d.x, v.x, acc[r * K + s][0]);
        acc[r * K + s][1] = fmaf(d.y, v.y, acc[r * K + s][1]);
        acc[r * K + s][2] = fmaf(d.z, v.z, acc[r * K + s][2]);
        acc[r * K + s][3] = fmaf(d.w, v.w, acc[r * K + s][3]);
      }
    }
  }
#pragma unroll
  for (int tp = 0; tp < KK; ++tp)
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(&sAcc[(q * 4 + j) * KK + tp], acc[tp][j]);
  __syncthreads();
  for (int e = tid; e < C * KK; e += nthr) atomicAdd(&dW[e], sAcc[e]);
}

void launch_dw_wgrad(const float* D, int ldd, const float* X, int ldx, float* dW, int N, int H, int Wd, int C, int k,
                     cudaStream_t s, InTf tf) {
  long long total = (long long)N * H * Wd;
  if (dw_tiled_shape(H, Wd, k)) {
    if (k == 3) dw_tiled_launch<3, 2>(X, ldx, nullptr, nullptr, 0, D, ldd, dW, nullptr, N, H, Wd, C, 0, s, tf);
    else dw_tiled_launch<5, 2>(X, ldx, nullptr, nullptr, 0, D, ldd, dW, nullptr, N, H, Wd, C, 0, s, tf);
    return;
  }
  long long per = std::max<long long>(cdivl(total, 148 * 4), 256);
  int grid = cdivl(total, per);
  size_t sm = (size_t)C * k * k * 4;
  if (k == 3) dw_wgrad_kernel<3><<<grid, quad_block(C), sm, s>>>(D, ldd, X, ldx, dW, N, H, Wd, C, (unsigned)per);
  else dw_wgrad_kernel<5><<<grid, quad_block(C), sm, s>>>(D, ldd, X, ldx, dW, N, H, Wd, C, (unsigned)per);
}

// =====================================================================================================================
// Tiled depthwise family (the hot depthwise launches: every map >= 32x32).  One CTA = one 16-channel group, looping over
// 16x16-pixel tiles: input tile + halo staged once in shared memory (pixel pitch 20 floats: conflict-free float4 reads),
// each thread owns one channel quad of 4 consecutive pixels and slides a K-wide register window over them.
//   MODE 0  forward, and the BatchNorm batch statistics of its own output (sum, sum of squares -> `sums`, double)
//   MODE 1  input gradient: flipped taps, optional accumulate into Y
//   MODE 2  weight gradient: dW[c][tap] += sum D[pix][c] * X[pix + tap][c], accumulated in registers across the tiles
// Cross-thread reductions: warp shuffles over the lanes that share a quad, then shared, then global atomics per CTA.
// =====================================================================================================================
template <int K, int MODE>
__global__ void __launch_bounds__(256) dw_tiled_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W,
                                                       float* Y, int ldy, const float* __restrict__ D, int ldd, float* dW,
                                                       double* sums, int N, int H, int Wd, int C, int beta, InTf tf) {
  constexpr int TY = 16, TX = 16, IH = TY + K - 1, IW = TX + K - 1, PS = 20, KK = K * K, PAD = K / 2;
  extern __shared__ __align__(16) float dsm[];
  float* sIn = dsm;                      // [IH*IW][PS]
  float* sW = sIn + IH * IW * PS;        // MODE 0/1: [KK][16] taps;  MODE 2: [16][KK] block accumulators
  __shared__ double sRed[2][16];
  __shared__ __align__(16) float sSc[16], sSh[16];     // input transform of this CTA's 16 channels
  const bool tfon = MODE != 1 && tf.gamma != nullptr;
  const int tid = threadIdx.x;
  const int tiles_x = (Wd + TX - 1) / TX, tiles_y = (H + TY - 1) / TY;
  const int ntiles = N * tiles_y * tiles_x;
  const int cg = blockIdx.y * 16;
  const int nq = min(4, (C - cg) >> 2);
  const int q = tid & 3, txi = (tid >> 2) & 3, ty = tid >> 4;
  if (MODE < 2) {
    for (int e = tid; e < KK * 16; e += 256) {
      int tp = e >> 4, c = e & 15;
      sW[(MODE == 1 ? KK - 1 - tp : tp) * 16 + c] = (cg + c < C) ? W[(size_t)(cg + c) * KK + tp] : 0.f;
    }
  } else {
    for (int e = tid; e < 16 * KK; e += 256) sW[e] = 0.f;
  }
  if (tid < 32) sRed[tid >> 4][tid & 15] = 0.0;
  if (tfon && tid < 16) {
    const int ch = cg + tid;
    const float sc = ch < C ? tf.gamma[ch] * tf.invstd[ch] : 0.f;
    sSc[tid] = sc; sSh[tid] = ch < C ? tf.beta[ch] - tf.mean[ch] * sc : 0.f;
  }
  if (tfon) __syncthreads();
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  float wacc[MODE == 2 ? KK : 1][4];
#pragma unroll
  for (int tp = 0; tp < (MODE == 2 ? KK : 1); ++tp) wacc[tp][0] = wacc[tp][1] = wacc[tp][2] = wacc[tp][3] = 0.f;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int t = tile;
    const int tx0 = (t % tiles_x) * TX; t /= tiles_x;
    const int ty0 = (t % tiles_y) * TY;
    const int n = t / tiles_y;
    __syncthreads();
    {
      // the input tile: all of a thread's 16-byte loads are issued before its first shared-memory store
      constexpr int TOT = IH * IW * 4, NL = (TOT + 255) / 256;
      float4 vv[NL];
      bool inimg[NL];
#pragma unroll
      for (int u = 0; u < NL; ++u) {
        const int i = tid + u * 256;
        const int qq = i & 3, pp = i >> 2;
        const int iy = ty0 + pp / IW - PAD, ix = tx0 + pp % IW - PAD;
        vv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        inimg[u] = i < TOT && qq < nq && iy >= 0 && iy < H && ix >= 0 && ix < Wd;
        if (inimg[u]) vv[u] = *reinterpret_cast<const float4*>(X + ((size_t)(n * H + iy) * Wd + ix) * ldx + cg + qq * 4);
      }
#pragma unroll
      for (int u = 0; u < NL; ++u) {
        const int i = tid + u * 256;
        if (i >= TOT) break;
        const int qq = i & 3, pp = i >> 2;
        float4 v = vv[u];
        if (tfon && inimg[u]) {        // the conv zero-pads the TRANSFORMED tensor: only in-image pixels are mapped
          const float4 sc = *reinterpret_cast<const float4*>(sSc + qq * 4), sh = *reinterpret_cast<const float4*>(sSh + qq * 4);
          v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
          if (tf.act) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
        }
        *reinterpret_cast<float4*>(sIn + pp * PS + qq * 4) = v;
      }
    }
    __syncthreads();
    const int y = ty0 + ty;
    const bool rowok = y < H && q < nq;
    if (MODE < 2) {
      float4 acc[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < K; ++r) {
        const float* rp = sIn + ((ty + r) * IW + txi * 4) * PS + q * 4;
        float4 v[K + 3], w[K];
#pragma unroll
        for (int j = 0; j < K + 3; ++j) v[j] = *reinterpret_cast<const float4*>(rp + j * PS);
#pragma unroll
        for (int s2i = 0; s2i < K; ++s2i) w[s2i] = *reinterpret_cast<const float4*>(sW + (r * K + s2i) * 16 + q * 4);
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
          for (int s2i = 0; s2i < K; ++s2i) {
            const float2 lo = __ffma2_rn(make_float2(v[o + s2i].x, v[o + s2i].y), make_float2(w[s2i].x, w[s2i].y), make_float2(acc[o].x, acc[o].y));
            const float2 hi = __ffma2_rn(make_float2(v[o + s2i].z, v[o + s2i].w), make_float2(w[s2i].z, w[s2i].w), make_float2(acc[o].z, acc[o].w));
            acc[o] = make_float4(lo.x, lo.y, hi.x, hi.y);
          }
      }
      if (rowok) {
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const int x = tx0 + txi * 4 + o;
          if (x >= Wd) break;
          float4* op = reinterpret_cast<float4*>(Y + ((size_t)(n * H + y) * Wd + x) * ldy + cg + q * 4);
          float4 a = acc[o];
          if (MODE == 1 && beta) { float4 pv = *op; a.x += pv.x; a.y += pv.y; a.z += pv.z; a.w += pv.w; }
          *op = a;
          if (MODE == 0) {
            s1[0] += a.x; s1[1] += a.y; s1[2] += a.z; s1[3] += a.w;
            s2[0] = fmaf(a.x, a.x, s2[0]); s2[1] = fmaf(a.y, a.y, s2[1]);
            s2[2] = fmaf(a.z, a.z, s2[2]); s2[3] = fmaf(a.w, a.w, s2[3]);
          }
        }
      }
    } else {
      float4 d[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const int x = tx0 + txi * 4 + o;
        d[o] = (rowok && x < Wd) ? *reinterpret_cast<const float4*>(D + ((size_t)(n * H + y) * Wd + x) * ldd + cg + q * 4)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int r = 0; r < K; ++r) {
        const float* rp = sIn + ((ty + r) * IW + txi * 4) * PS + q * 4;
        float4 v[K + 3];
#pragma unroll
        for (int j = 0; j < K + 3; ++j) v[j] = *reinterpret_cast<const float4*>(rp + j * PS);
#pragma unroll
        for (int s2i = 0; s2i < K; ++s2i)
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            wacc[MODE == 2 ? r * K + s2i : 0][0] = fmaf(d[o].x, v[o + s2i].x, wacc[MODE == 2 ? r * K + s2i : 0][0]);
            wacc[MODE == 2 ? r * K + s2i : 0][1] = fmaf(d[o].y, v[o + s2i].y, wacc[MODE == 2 ? r * K + s2i : 0][1]);
            wacc[MODE == 2 ? r * K + s2i : 0][2] = fmaf(d[o].z, v[o + s2i].z, wacc[MODE == 2 ? r * K + s2i : 0][2]);
            wacc[MODE == 2 ? r * K + s2i : 0][3] = fmaf(d[o].w, v[o + s2i].w, wacc[MODE == 2 ? r * K + s2i : 0][3]);
          }
      }
    }
  }
  // ---- block reductions (lanes with equal q: xor 4, 8, 16) ----
  if (MODE == 0) {
    __syncthreads();
    double a1[4], a2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a1[j] = s1[j]; a2[j] = s2[j];
#pragma unroll
      for (int m = 4; m < 32; m <<= 1) {
        a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], m);
        a2[j] += __shfl_xor_sync(0xffffffffu, a2[j], m);
      }
    }
    if ((tid & 31) < 4 && q < nq)
#pragma unroll
      for (int j = 0; j < 4; ++j) { atomicAdd(&sRed[0][q * 4 + j], a1[j]); atomicAdd(&sRed[1][q * 4 + j], a2[j]); }
    __syncthreads();
    if (tid < 32 && cg + (tid & 15) < C) atomicAdd(&sums[(tid >> 4) * C + cg + (tid & 15)], sRed[tid >> 4][tid & 15]);
  }
  if (MODE == 2) {
    __syncthreads();
#pragma unroll
    for (int tp = 0; tp < (MODE == 2 ? KK : 1); ++tp)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = wacc[tp][j];
#pragma unroll
        for (int m = 4; m < 32; m <<= 1) a += __shfl_xor_sync(0xffffffffu, a, m);
        if ((tid & 31) < 4) atomicAdd(&sW[(q * 4 + j) * KK + tp], a);
      }
    __syncthreads();
    for (int e = tid; e < 16 * KK; e += 256)
      if (cg + e / KK < C) atomicAdd(&dW[(size_t)cg * KK + e], sW[e]);
  }
}

template <int K, int MODE>
static void dw_tiled_launch(const float* X, int ldx, const float* W, float* Y, int ldy, const float* D, int ldd, float* dW,
                            double* sums, int N, int H, int Wd, int C, int beta, cudaStream_t s, InTf tf) {
  constexpr size_t smem = sizeof(float) * ((16 + K - 1) * (16 + K - 1) * 20 + K * K * 16);
  const int ngrp = (C + 15) / 16;
  const int ntiles = N * ((H + 15) / 16) * ((Wd + 15) / 16);
  static const int occ = resident_ctas(dw_tiled_kernel<K, MODE>, 256, smem);
  int gx = std::min(ntiles, std::max(1, 148 * occ / ngrp));
  dw_tiled_kernel<K, MODE><<<dim3(gx, ngrp), 256, smem, s>>>(X, ldx, W, Y, ldy, D, ldd, dW, sums, N, H, Wd, C, beta, tf);
}

// forward with fused BatchNorm statistics (sums = [2][C] doubles, pre-zeroed) -- returns false if the shape is not tiled
// maps from 24 x 24 up: the 30 x 30 stage of the decoder (3 of its 4 tiles are partial) still runs 3-7x faster here than on the
// per-pixel kernels (dw_wgrad_kernel<5> streamed it at 0.18 TB/s)
bool dw_tiled_shape(int H, int Wd, int k) { return H >= 24 && Wd >= 24 && (k == 3 || k == 5); }

bool launch_dw_fwd_stats(const float* X, int ldx, const float* W, float* Y, int ldy, double* sums, int N, int H, int Wd,
                         int C, int k, cudaStream_t s, InTf tf) {
  if (!dw_tiled_shape(H, Wd, k)) return false;
  if (k == 3) dw_tiled_launch<3, 0>(X, ldx, W, Y, ldy, nullptr, 0, nullptr, sums, N, H, Wd, C, 0, s, tf);
  else dw_tiled_launch<5, 0>(X, ldx, W, Y, ldy, nullptr, 0, nullptr, sums, N, H, Wd, C, 0, s, tf);
  return true;
}

// =====================================================================================================================
__device__ __forceinline__ float silu_grad(float t) {
  const float sg = sigmoid_mufu(t);
  return sg * (1.f + t * (1.f - sg));
}

// Backward of a depthwise ConvBN unit in ONE pass (the DoubleLightConv units, K = 3): BatchNorm backward, input gradient,
// weight gradient and -- for the unit that PRODUCED the conv input -- that unit's BatchNorm-backward sums.
//   dz = gamma * invstd * (dt - S1/M - zhat * S2/M),  dt = dy * act'(gamma * zhat + beta)      (S1, S2: col_reduce MODE 1)
//   dx[p]      = sum_tap W[tap] * dz[p - tap]                                                  (zero padding)
//   dW[c][tap] = sum_p dz[p] * x'[p + tap],  x' = producer's BN (+act) applied to its raw output (InTf), zero outside the image
//   producer sums: sum dx * act_p'(.), sum dx * act_p'(.) * zhat_p
// The unfused chain (bn_bwd_apply -> dz, dw_tiled<2>(dz, x), dw_tiled<1>(dz) and col_reduce<1>(dx, z_p) for the producer) moved
// 3 + 2.3 + 2.3 + 2 floats per element; here dy, z, x come in once with a one-pixel halo (3 x 1.27) and dx goes out: 4.8.
// dz is computed at the halo pixels too (each CTA recomputes its border), never written.  Same tiling as dw_tiled_kernel:
// CTA = one 16-channel group, persistent over 16x16-pixel tiles, thread = one channel quad x 4 consecutive pixels.
// =====================================================================================================================
struct DwBwdP {
  const float* DY; int ldd; const float* Z; int ldz; BnRef bn; int act; const double* sums;
  const float* X; int ldx; InTf tf; const float* W; float* DX; int lddx; float* dW; float* dgamma; float* dbeta; double* psums;
  int N, H, Wd, C; double invM;
};

template <int K>
__global__ void __launch_bounds__(256, 2) dw_bwd_fused_kernel(const DwBwdP p) {
  constexpr int TY = 16, TX = 16, IH = TY + K - 1, IW = TX + K - 1, PS = 20, KK = K * K, PAD = K / 2;
  extern __shared__ __align__(16) float fsm[];
  float* sDz = fsm;                      // [IH*IW][PS]  dz with halo
  float* sIn = sDz + IH * IW * PS;       // [IH*IW][PS]  x' with halo
  float* sW = sIn + IH * IW * PS;        // [KK][16] flipped taps
  float* sWa = sW + KK * 16;             // [16][KK] block accumulators of dW
  float* sRd = sWa + 16 * KK;            // [IH*IW*4] x 16 B: raw dy of the NEXT tile (cp.async; item i of a thread = its own slot)
  float* sRz = sRd + IH * IW * 16;       // raw z, same layout
  __shared__ double sRed[2][16];
  __shared__ __align__(16) float sBn[6][16];      // this unit: mean, invstd, gamma, beta, S1/M, S2/M
  __shared__ __align__(16) float sPr[4][16];      // producer: scale, shift (x' = z * sc + sh), mean, invstd
  const int tid = threadIdx.x;
  const int H = p.H, Wd = p.Wd, C = p.C;
  const int tiles_x = (Wd + TX - 1) / TX, tiles_y = (H + TY - 1) / TY;
  const int ntiles = p.N * tiles_y * tiles_x;
  const int cg = blockIdx.y * 16;
  const int nq = min(4, (C - cg) >> 2);
  const int q = tid & 3, txi = (tid >> 2) & 3, ty = tid >> 4;
  const bool tfon = p.tf.gamma != nullptr;
  for (int e = tid; e < KK * 16; e += 256) {
    const int tp = e >> 4, c = e & 15;
    sW[(KK - 1 - tp) * 16 + c] = (cg + c < C) ? p.W[(size_t)(cg + c) * KK + tp] : 0.f;
  }
  for (int e = tid; e < 16 * KK; e += 256) sWa[e] = 0.f;
  if (tid < 32) sRed[tid >> 4][tid & 15] = 0.0;
  if (tid < 16) {
    const int ch = cg + tid;
    const bool ok = ch < C;
    sBn[0][tid] = ok ? p.bn.mean[ch] : 0.f; sBn[1][tid] = ok ? p.bn.invstd[ch] : 0.f;
    sBn[2][tid] = ok ? p.bn.gamma[ch] : 0.f; sBn[3][tid] = ok ? p.bn.beta[ch] : 0.f;
    sBn[4][tid] = ok ? (float)(p.sums[ch] * p.invM) : 0.f; sBn[5][tid] = ok ? (float)(p.sums[C + ch] * p.invM) : 0.f;
    float sc = 1.f, sh = 0.f, mu = 0.f, is = 0.f;
    if (tfon && ok) { is = p.tf.invstd[ch]; mu = p.tf.mean[ch]; sc = p.tf.gamma[ch] * is; sh = p.tf.beta[ch] - mu * sc; }
    sPr[0][tid] = sc; sPr[1][tid] = sh; sPr[2][tid] = mu; sPr[3][tid] = is;
    if (blockIdx.x == 0 && ok) {          // d gamma = S2, d beta = S1 (accumulated: the flat gradient buffer is pre-zeroed)
      p.dbeta[ch] += (float)p.sums[ch]; p.dgamma[ch] += (float)p.sums[C + ch];
    }
  }
  __syncthreads();
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  float wacc[KK][4];
#pragma unroll
  for (int tp = 0; tp < KK; ++tp) wacc[tp][0] = wacc[tp][1] = wacc[tp][2] = wacc[tp][3] = 0.f;

  constexpr int TOT = IH * IW * 4, NL = (TOT + 255) / 256;       // every item of a thread has channel quad q (256 % 4 == 0)
  // dy and z of a tile arrive by cp.async one tile ahead (zero-filled outside the image); a thread only ever touches its own
  // 16-byte slots of the raw buffers, so the prefetch needs no barrier of its own
  auto prefetch = [&](int tile) {
    int t = tile;
    const int tx0 = (t % tiles_x) * TX; t /= tiles_x;
    const int ty0 = (t % tiles_y) * TY;
    const int n = t / tiles_y;
#pragma unroll
    for (int u = 0; u < NL; ++u) {
      const int i = tid + u * 256, pp = i >> 2;
      if (i >= TOT) break;
      const int iy = ty0 + pp / IW - PAD, ix = tx0 + pp % IW - PAD;
      const bool in = q < nq && iy >= 0 && iy < H && ix >= 0 && ix < Wd;
      const size_t pix = in ? (size_t)(n * H + iy) * Wd + ix : 0;
      const uint32_t nb = in ? 16u : 0u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(sRd + i * 4)),
                   "l"(p.DY + pix * p.ldd + cg + (in ? q * 4 : 0)), "r"(nb) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(sRz + i * 4)),
                   "l"(p.Z + pix * p.ldz + cg + (in ? q * 4 : 0)), "r"(nb) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if ((int)blockIdx.x < ntiles) prefetch(blockIdx.x);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int t = tile;
    const int tx0 = (t % tiles_x) * TX; t /= tiles_x;
    const int ty0 = (t % tiles_y) * TY;
    const int n = t / tiles_y;
    __syncthreads();
    // the x loads are in flight while dz is computed
    float4 xv[NL];
    bool xin[NL];
#pragma unroll
    for (int u = 0; u < NL; ++u) {
      const int i = tid + u * 256, pp = i >> 2;
      const int iy = ty0 + pp / IW - PAD, ix = tx0 + pp % IW - PAD;
      xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      xin[u] = i < TOT && q < nq && iy >= 0 && iy < H && ix >= 0 && ix < Wd;
      if (xin[u]) xv[u] = *reinterpret_cast<const float4*>(p.X + ((size_t)(n * H + iy) * Wd + ix) * p.ldx + cg + q * 4);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    {
      // ---- dz tile from the raw buffers ----
      const float4 mu = *reinterpret_cast<const float4*>(&sBn[0][q * 4]), is = *reinterpret_cast<const float4*>(&sBn[1][q * 4]);
      const float4 ga = *reinterpret_cast<const float4*>(&sBn[2][q * 4]), be = *reinterpret_cast<const float4*>(&sBn[3][q * 4]);
      const float4 m1 = *reinterpret_cast<const float4*>(&sBn[4][q * 4]), m2 = *reinterpret_cast<const float4*>(&sBn[5][q * 4]);
#pragma unroll
      for (int u = 0; u < NL; ++u) {
        const int i = tid + u * 256;
        if (i >= TOT) break;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (xin[u]) {                       // same in-image predicate for all three tensors
          const float4 dv = *reinterpret_cast<const float4*>(sRd + i * 4), zv = *reinterpret_cast<const float4*>(sRz + i * 4);
#define YSP_DZ(f)                                                                                       \
  {                                                                                                     \
    const float zh = (zv.f - mu.f) * is.f;                                                              \
    const float dt = p.act ? dv.f * silu_grad(fmaf(ga.f, zh, be.f)) : dv.f;                             \
    o.f = ga.f * is.f * (dt - m1.f - zh * m2.f);                                                        \
  }
          YSP_DZ(x) YSP_DZ(y) YSP_DZ(z) YSP_DZ(w)
#undef YSP_DZ
        }
        *reinterpret_cast<float4*>(sDz + (i >> 2) * PS + q * 4) = o;
      }
    }
    {
      // ---- x' tile ----
      const float4 sc = *reinterpret_cast<const float4*>(&sPr[0][q * 4]), sh = *reinterpret_cast<const float4*>(&sPr[1][q * 4]);
#pragma unroll
      for (int u = 0; u < NL; ++u) {
        const int i = tid + u * 256;
        if (i >= TOT) break;
        float4 v = xv[u];
        if (tfon && xin[u]) {
          v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
          if (p.tf.act) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
        }
        *reinterpret_cast<float4*>(sIn + (i >> 2) * PS + q * 4) = v;
      }
    }
    __syncthreads();
    if (tile + (int)gridDim.x < ntiles) prefetch(tile + gridDim.x);
    const int y = ty0 + ty;
    const bool rowok = y < H && q < nq;
    // ---- input gradient (flipped taps over the dz tile) ----
    {
      // the producer's raw output at this thread's pixels (for its BN-backward sums): issued before the stencil
      float4 zr[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const int x = tx0 + txi * 4 + o;
        zr[o] = (p.psums && rowok && x < Wd) ? *reinterpret_cast<const float4*>(p.X + ((size_t)(n * H + y) * Wd + x) * p.ldx + cg + q * 4)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float4 acc[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < K; ++r) {
        const float* rp = sDz + ((ty + r) * IW + txi * 4) * PS + q * 4;
        float4 v[K + 3], w[K];
#pragma unroll
        for (int j = 0; j < K + 3; ++j) v[j] = *reinterpret_cast<const float4*>(rp + j * PS);
#pragma unroll
        for (int s2i = 0; s2i < K; ++s2i) w[s2i] = *reinterpret_cast<const float4*>(sW + (r * K + s2i) * 16 + q * 4);
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
          for (int s2i = 0; s2i < K; ++s2i) {
            const float2 lo = __ffma2_rn(make_float2(v[o + s2i].x, v[o + s2i].y), make_float2(w[s2i].x, w[s2i].y), make_float2(acc[o].x, acc[o].y));
            const float2 hi = __ffma2_rn(make_float2(v[o + s2i].z, v[o + s2i].w), make_float2(w[s2i].z, w[s2i].w), make_float2(acc[o].z, acc[o].w));
            acc[o] = make_float4(lo.x, lo.y, hi.x, hi.y);
          }
      }
      if (rowok) {
        const float4 pmu = *reinterpret_cast<const float4*>(&sPr[2][q * 4]), pis = *reinterpret_cast<const float4*>(&sPr[3][q * 4]);
        const float4 psc = *reinterpret_cast<const float4*>(&sPr[0][q * 4]), psh = *reinterpret_cast<const float4*>(&sPr[1][q * 4]);
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const int x = tx0 + txi * 4 + o;
          if (x >= Wd) break;
          const size_t pix = (size_t)(n * H + y) * Wd + x;
          const float4 a = acc[o];
          *reinterpret_cast<float4*>(p.DX + pix * p.lddx + cg + q * 4) = a;
          if (p.psums) {       // the producer's BatchNorm-backward sums
#define YSP_PS(f, j)                                                                                    \
  {                                                                                                     \
    const float zh = (zr[o].f - pmu.f) * pis.f;                                                         \
    const float dt = p.tf.act ? a.f * silu_grad(fmaf(zr[o].f, psc.f, psh.f)) : a.f;                     \
    s1[j] += dt; s2[j] = fmaf(dt, zh, s2[j]);                                                           \
  }
            YSP_PS(x, 0) YSP_PS(y, 1) YSP_PS(z, 2) YSP_PS(w, 3)
#undef YSP_PS
          }
        }
      }
    }
    // ---- weight gradient: dz at the centre pixels against the x' window ----
    {
      float4 d[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) d[o] = *reinterpret_cast<const float4*>(sDz + ((ty + PAD) * IW + txi * 4 + o + PAD) * PS + q * 4);
#pragma unroll
      for (int r = 0; r < K; ++r) {
        const float* rp = sIn + ((ty + r) * IW + txi * 4) * PS + q * 4;
        float4 v[K + 3];
#pragma unroll
        for (int j = 0; j < K + 3; ++j) v[j] = *reinterpret_cast<const float4*>(rp + j * PS);
#pragma unroll
        for (int s2i = 0; s2i < K; ++s2i)
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            wacc[r * K + s2i][0] = fmaf(d[o].x, v[o + s2i].x, wacc[r * K + s2i][0]);
            wacc[r * K + s2i][1] = fmaf(d[o].y, v[o + s2i].y, wacc[r * K + s2i][1]);
            wacc[r * K + s2i][2] = fmaf(d[o].z, v[o + s2i].z, wacc[r * K + s2i][2]);
            wacc[r * K + s2i][3] = fmaf(d[o].w, v[o + s2i].w, wacc[r * K + s2i][3]);
          }
      }
    }
  }
  // ---- block reductions (lanes with equal q: xor 4, 8, 16) ----
  __syncthreads();
  if (p.psums) {
    double a1[4], a2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a1[j] = s1[j]; a2[j] = s2[j];
#pragma unroll
      for (int m = 4; m < 32; m <<= 1) {
        a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], m);
        a2[j] += __shfl_xor_sync(0xffffffffu, a2[j], m);
      }
    }
    if ((tid & 31) < 4 && q < nq)
#pragma unroll
      for (int j = 0; j < 4; ++j) { atomicAdd(&sRed[0][q * 4 + j], a1[j]); atomicAdd(&sRed[1][q * 4 + j], a2[j]); }
  }
#pragma unroll
  for (int tp = 0; tp < KK; ++tp)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = wacc[tp][j];
#pragma unroll
      for (int m = 4; m < 32; m <<= 1) a += __shfl_xor_sync(0xffffffffu, a, m);
      if ((tid & 31) < 4) atomicAdd(&sWa[(q * 4 + j) * KK + tp], a);
    }
  __syncthreads();
  if (p.psums && tid < 32 && cg + (tid & 15) < C) atomicAdd(&p.psums[(tid >> 4) * C + cg + (tid & 15)], sRed[tid >> 4][tid & 15]);
  for (int e = tid; e < 16 * KK; e += 256)
    if (cg + e / KK < C) atomicAdd(&p.dW[(size_t)cg * KK + e], sWa[e]);
}

bool launch_dw_bwd_fused(const float* DY, int ldd, const float* Z, int ldz, const BnRef& bn, int act, const double* sums,
                         const float* X, int ldx, InTf tf, const float* W, float* DX, int lddx, float* dW, float* dgamma,
                         float* dbeta, double* psums, int N, int H, int Wd, int C, int k, cudaStream_t s) {
  static const bool off = getenv("YSP_TRAIN_NO_DWFUSE") != nullptr;
  if (off || k != 3 || !dw_tiled_shape(H, Wd, k) || (C & 3)) return false;
  constexpr int K = 3;
  constexpr size_t smem = sizeof(float) * (2 * (16 + K - 1) * (16 + K - 1) * 20 + 2 * K * K * 16 + 2 * (16 + K - 1) * (16 + K - 1) * 16);
  static unsigned long long attr_done = 0;
  ensure_dyn_smem(dw_bwd_fused_kernel<K>, smem, attr_done, "dw_bwd_fused_kernel");
  DwBwdP p = {DY, ldd, Z, ldz, bn, act, sums, X, ldx, tf, W, DX, lddx, dW, dgamma, dbeta, psums, N, H, Wd, C,
              1.0 / ((double)N * H * Wd)};
  const int ngrp = (C + 15) / 16;
  const int ntiles = N * ((H + 15) / 16) * ((Wd + 15) / 16);
  static const int occ = resident_ctas(dw_bwd_fused_kernel<K>, 256, smem);
  const int gx = std::min(ntiles, std::max(1, 148 * occ / ngrp));
  dw_bwd_fused_kernel<K><<<dim3(gx, ngrp), 256, smem, s>>>(p);
  return true;
}

// =====================================================================================================================
// Per-channel reductions in double precision.  sums layout: [segment][2][C]; segment = image (SEG=1) or whole batch.
//   MODE 0  BN statistics:    v1 = z            v2 = z*z
//   MODE 1  BN backward:      v1 = dt           v2 = dt*zhat     dt = dy * act'(gamma*zhat+beta)
//   MODE 2  column sums:      v1 = a                              (bias gradients, ECA average pool)
//   MODE 3  products:         v1 = a*b                            (ECA backward: d gate)
// =====================================================================================================================

template <int MODE>
__global__ void __launch_bounds__(256) col_reduce_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bv,
                                                         int ldb, BnRef bn, int act, double* sums, int C,
                                                         long long rows_per_seg, long long rows_per_cta) {
  extern __shared__ double sred[];   // [2][C]
  const int nthr = blockDim.x * blockDim.y, tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int e = tid; e < 2 * C; e += nthr) sred[e] = 0.0;
  __syncthreads();
  const int lanes = blockDim.y;
  const int q = threadIdx.x, lane = threadIdx.y;
  const long long seg = blockIdx.y;
  const long long r0 = seg * rows_per_seg + (long long)blockIdx.x * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > (seg + 1) * rows_per_seg) r1 = (seg + 1) * rows_per_seg;
  if (lane < lanes) {
    float mu[4], is[4], ga[4], be[4];
    if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c = q * 4 + j;
        mu[j] = bn.mean[c]; is[j] = bn.invstd[c]; ga[j] = bn.gamma[c]; be[j] = bn.beta[c];
      }
    }
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    float f1[4] = {0, 0, 0, 0}, f2[4] = {0, 0, 0, 0};
    int cnt = 0;
    // four rows per iteration: all eight 16-byte loads are issued before the arithmetic (one row at a time left a single load
    // pair in flight per thread: 3.2 TB/s where the apply kernel streams the same tensors at 5.7 TB/s)
    constexpr int UR = 4;
    for (long long m0 = r0 + lane; m0 < r1; m0 += (long long)UR * lanes) {
      float4 a4[UR], b4[UR];
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        const long long m = m0 + (long long)u * lanes;
        a4[u] = make_float4(0.f, 0.f, 0.f, 0.f); b4[u] = a4[u];
        if (m < r1) {
          a4[u] = *reinterpret_cast<const float4*>(A + m * lda + q * 4);
          if (MODE == 1 || MODE == 3) b4[u] = *reinterpret_cast<const float4*>(Bv + m * ldb + q * 4);
        }
      }
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        if (m0 + (long long)u * lanes >= r1) break;
        const float a[4] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w};
        const float b[4] = {b4[u].x, b4[u].y, b4[u].z, b4[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (MODE == 0) { f1[j] += a[j]; f2[j] = fmaf(a[j], a[j], f2[j]); }
          if (MODE == 1) {   // a = dy, b = z
            float zh = (b[j] - mu[j]) * is[j];
            float dt = act ? a[j] * silu_grad(fmaf(ga[j], zh, be[j])) : a[j];
            f1[j] += dt; f2[j] = fmaf(dt, zh, f2[j]);
          }
          if (MODE == 2) f1[j] += a[j];
          if (MODE == 3) f1[j] = fmaf(a[j], b[j], f1[j]);
        }
      }
      if (++cnt == 4) {   // short fp32 runs (16 rows), long double runs
#pragma unroll
        for (int j = 0; j < 4; ++j) { s1[j] += f1[j]; s2[j] += f2[j]; f1[j] = f2[j] = 0.f; }
        cnt = 0;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&sred[q * 4 + j], s1[j] + (double)f1[j]);
      if (MODE <= 1) atomicAdd(&sred[C + q * 4 + j], s2[j] + (double)f2[j]);
    }
  }
  __syncthreads();
  for (int e = tid; e < (MODE <= 1 ? 2 : 1) * C; e += nthr) atomicAdd(&sums[seg * 2 * C + e], sred[e]);
}

void launch_col_reduce(int mode, const float* A, int lda, const float* B, int ldb, const BnRef& bn, int act, double* sums,
                       int C, long long segs, long long rows_per_seg, cudaStream_t s) {
  // CTAs per SM of the reduction grid, measured on the whole step (ms): 1: 13.74, 2: 13.32, 3: 13.56, 4: 13.48, 8: 13.87, 16: 14.52 --
  // few CTAs with long contiguous row ranges and few double-precision atomics win
  static const int per_sm = getenv("YSP_COLRED_PER_SM") ? atoi(getenv("YSP_COLRED_PER_SM")) : 2;
  long long want = std::max<long long>(1, (148 * per_sm) / segs);
  long long per = std::max<long long>(cdivl(rows_per_seg, want), 64);
  dim3 grid(cdivl(rows_per_seg, per), (unsigned)segs);
  size_t sm = (size_t)2 * C * sizeof(double);
  switch (mode) {
    case 0: col_reduce_kernel<0><<<grid, quad_block(C), sm, s>>>(A, lda, B, ldb, bn, act, sums, C, rows_per_seg, per); break;
    case 1: col_reduce_kernel<1><<<grid, quad_block(C), sm, s>>>(A, lda, B, ldb, bn, act, sums, C, rows_per_seg, per); break;
    case 2: col_reduce_kernel<2><<<grid, quad_block(C), sm, s>>>(A, lda, B, ldb, bn, act, sums, C, rows_per_seg, per); break;
    default: col_reduce_kernel<3><<<grid, quad_block(C), sm, s>>>(A, lda, B, ldb, bn, act, sums, C, rows_per_seg, per); break;
  }
}

// BN: sums -> (mean, invstd) + running-statistics update (nn.BatchNorm2d train(): biased var normalises, unbiased var
// goes into running_var).  One thread per channel.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int C, double M, float eps, float momentum,
                                   float* mean, float* invstd, float* run_mean, float* run_var) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mu = sums[c] / M;
  double var = sums[C + c] / M - mu * mu;
  if (var < 0) var = 0;
  mean[c] = (float)mu;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (run_mean) {
    run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * (float)mu;
    run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)(var * (M / (M > 1 ? M - 1 : 1)));
  }
}

void launch_bn_finalize(const double* sums, int C, long long M, float eps, float momentum, float* mean, float* invstd,
                        float* run_mean, float* run_var, cudaStream_t s) {
  bn_finalize_kernel<<<cdivl(C, 128), 128, 0, s>>>(sums, C, (double)M, eps, momentum, mean, invstd, run_mean, run_var);
}

// y = act(gamma*(z-mean)*invstd + beta) (+ res)
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ Z, int ldz, BnRef bn, int act,
                                                       const float* __restrict__ R, int ldr, float* Y, int ldy, int C,
                                                       long long M) {
  const int q = threadIdx.x;
  float sc[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {            // y = z*sc + sh
    int c = q * 4 + j;
    sc[j] = bn.gamma[c] * bn.invstd[c];
    sh[j] = bn.beta[c] - bn.mean[c] * sc[j];
  }
  for (long long m = (long long)blockIdx.x * blockDim.y + threadIdx.y; m < M; m += (long long)gridDim.x * blockDim.y) {
    float4 z4 = *reinterpret_cast<const float4*>(Z + m * ldz + q * 4);
    float z[4] = {z4.x, z4.y, z4.z, z4.w}, r[4] = {0, 0, 0, 0}, o[4];
    if (R) { float4 r4 = *reinterpret_cast<const float4*>(R + m * ldr + q * 4); r[0] = r4.x; r[1] = r4.y; r[2] = r4.z; r[3] = r4.w; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = fmaf(z[j], sc[j], sh[j]);
      o[j] = (act ? silu_f(t) : t) + r[j];
    }
    *reinterpret_cast<float4*>(Y + m * ldy + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

void launch_bn_apply(const float* Z, int ldz, const BnRef& bn, int act, const float* R, int ldr, float* Y, int ldy, int C,
                     long long M, cudaStream_t s) {
  bn_apply_kernel<<<quad_grid(M, C), quad_block(C), 0, s>>>(Z, ldz, bn, act, R, ldr, Y, ldy, C, M);
}

// dz = gamma*invstd*(dt - S1/M - zhat*S2/M);  block 0 also accumulates dgamma += S2, dbeta += S1.
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ DY, int ldd, const float* __restrict__ Z,
                                                           int ldz, BnRef bn, int act, const double* __restrict__ sums,
                                                           float* DZ, int ldo, float* dgamma, float* dbeta, int C,
                                                           long long M) {
  const int q = threadIdx.x;
  const double invM = 1.0 / (double)M;
  if (blockIdx.x == 0 && threadIdx.y == 0)
#pragma unroll
    for (int j = 0; j < 4; ++j) { int c = q * 4 + j; dbeta[c] += (float)sums[c]; dgamma[c] += (float)sums[C + c]; }
  float mu[4], is[4], ga[4], be[4], m1[4], m2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = q * 4 + j;
    mu[j] = bn.mean[c]; is[j] = bn.invstd[c]; ga[j] = bn.gamma[c]; be[j] = bn.beta[c];
    m1[j] = (float)(sums[c] * invM); m2[j] = (float)(sums[C + c] * invM);
  }
  for (long long m = (long long)blockIdx.x * blockDim.y + threadIdx.y; m < M; m += (long long)gridDim.x * blockDim.y) {
    float4 d4 = *reinterpret_cast<const float4*>(DY + m * ldd + q * 4);
    float4 z4 = *reinterpret_cast<const float4*>(Z + m * ldz + q * 4);
    const float d[4] = {d4.x, d4.y, d4.z, d4.w}, z[4] = {z4.x, z4.y, z4.z, z4.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float zh = (z[j] - mu[j]) * is[j];
      float dt = act ? d[j] * silu_grad(fmaf(ga[j], zh, be[j])) : d[j];
      o[j] = ga[j] * is[j] * (dt - m1[j] - zh * m2[j]);
    }
    *reinterpret_cast<float4*>(DZ + m * ldo + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

void launch_bn_bwd_apply(const float* DY, int ldd, const float* Z, int ldz, const BnRef& bn, int act, const double* sums,
                         float* DZ, int ldo, float* dgamma, float* dbeta, int C, long long M, cudaStream_t s) {
  bn_bwd_apply_kernel<<<quad_grid(M, C), quad_block(C), 0, s>>>(DY, ldd, Z, ldz, bn, act, sums, DZ, ldo, dgamma, dbeta, C, M);
}

// g[j] += (float) sum_{f < fold} sums[j*fold + f]   (bias gradients from MODE-2 reductions; fold > 1 when a 1-channel
// tensor was reduced as [M/fold][fold])
__global__ void add_sums_kernel(const double* __restrict__ sums, float* g, int n, int fold) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a = 0;
  for (int f = 0; f < fold; ++f) a += sums[i * fold + f];
  g[i] += (float)a;
}
void launch_add_sums(const double* sums, float* g, int n, int fold, cudaStream_t s) {
  add_sums_kernel<<<cdivl(n, 128), 128, 0, s>>>(sums, g, n, fold);
}

// =====================================================================================================================
// Elementwise helpers over [M][C] views (C % 4 == 0 unless noted)
// =====================================================================================================================
// out = a (+ b)   (scalar version: any C)
__global__ void __launch_bounds__(256) add_copy_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B,
                                                       int ldb, float* O, int ldo, int C, long long M) {
  const long long total = M * C;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    int c = (int)(e % C);
    long long m = e / C;
    float v = A[m * lda + c];
    if (B) v += B[m * ldb + c];
    O[m * ldo + c] = v;
  }
}
// 16-byte version: C, the row strides and the base addresses are multiples of four floats
__global__ void __launch_bounds__(256) add_copy4_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B,
                                                        int ldb, float* O, int ldo, int C4, long long M) {
  const long long total = M * C4;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const int c = (int)(e % C4) * 4;
    const long long m = e / C4;
    float4 v = *reinterpret_cast<const float4*>(A + m * lda + c);
    if (B) { const float4 w = *reinterpret_cast<const float4*>(B + m * ldb + c); v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
    *reinterpret_cast<float4*>(O + m * ldo + c) = v;
  }
}
void launch_add_copy(const float* A, int lda, const float* B, int ldb, float* O, int ldo, int C, long long M, cudaStream_t s) {
  const bool v4 = ((C | lda | ldo | (B ? ldb : 0)) & 3) == 0 &&
                  ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(O) | reinterpret_cast<uintptr_t>(B)) & 15) == 0;
  if (v4) {
    add_copy4_kernel<<<(int)std::min<long long>(cdivl(M * (C / 4), 256), 148 * 16), 256, 0, s>>>(A, lda, B, ldb, O, ldo, C / 4, M);
    return;
  }
  int grid = (int)std::min<long long>(cdivl(M * C, 256), 148 * 16);
  add_copy_kernel<<<grid, 256, 0, s>>>(A, lda, B, ldb, O, ldo, C, M);
}

// torch Upsample(scale_factor=2, mode="bilinear", align_corners=False) (YOLOSegPlusPlus.py:155) and its adjoint.
// 1-D taps: out[2j] = .25 in[j-1] + .75 in[j], out[2j+1] = .75 in[j] + .25 in[j+1], indices clamped at the border;
// adjoint: din[j] = .25 do[2j-1] + .75 do[2j] + .75 do[2j+1] + .25 do[2j+2] with the OUTPUT indices clamped.
__global__ void __launch_bounds__(256) up2_kernel(const float* __restrict__ X, int ldx, float* Y, int ldy, int N, int h,
                                                  int w, int C) {
  const int H = 2 * h, Wd = 2 * w, q = threadIdx.x;
  const unsigned total = (unsigned)N * H * Wd;
  for (unsigned pix = blockIdx.x * blockDim.y + threadIdx.y; pix < total; pix += gridDim.x * blockDim.y) {
    const int ox = pix % (unsigned)Wd;
    const unsigned t = pix / (unsigned)Wd;
    const int oy = t % (unsigned)H;
    const unsigned n = t / (unsigned)H;
    int jy = oy >> 1, jx = ox >> 1;
    int y0 = (oy & 1) ? jy : max(jy - 1, 0), y1 = (oy & 1) ? min(jy + 1, h - 1) : jy;
    int x0 = (ox & 1) ? jx : max(jx - 1, 0), x1 = (ox & 1) ? min(jx + 1, w - 1) : jx;
    float wy1 = (oy & 1) ? 0.25f : 0.75f, wx1 = (ox & 1) ? 0.25f : 0.75f;   // weight of the higher index
    const float* b = X + (size_t)n * h * w * ldx + q * 4;
    float4 a00 = *reinterpret_cast<const float4*>(b + (size_t)(y0 * w + x0) * ldx);
    float4 a01 = *reinterpret_cast<const float4*>(b + (size_t)(y0 * w + x1) * ldx);
    float4 a10 = *reinterpret_cast<const float4*>(b + (size_t)(y1 * w + x0) * ldx);
    float4 a11 = *reinterpret_cast<const float4*>(b + (size_t)(y1 * w + x1) * ldx);
    float wy0 = 1.f - wy1, wx0 = 1.f - wx1;
    float4 o;
    // same association as ATen's upsample_bilinear2d: w_y0*(w_x0*a00 + w_x1*a01) + w_y1*(w_x0*a10 + w_x1*a11)
    o.x = wy0 * (wx0 * a00.x + wx1 * a01.x) + wy1 * (wx0 * a10.x + wx1 * a11.x);
    o.y = wy0 * (wx0 * a00.y + wx1 * a01.y) + wy1 * (wx0 * a10.y + wx1 * a11.y);
    o.z = wy0 * (wx0 * a00.z + wx1 * a01.z) + wy1 * (wx0 * a10.z + wx1 * a11.z);
    o.w = wy0 * (wx0 * a00.w + wx1 * a01.w) + wy1 * (wx0 * a10.w + wx1 * a11.w);
    *reinterpret_cast<float4*>(Y + (size_t)pix * ldy + q * 4) = o;
  }
}
void launch_up2(const float* X, int ldx, float* Y, int ldy, int N, int h, int w, int C, cudaStream_t s) {
  up2_kernel<<<quad_grid((long long)N * 4 * h * w, C), quad_block(C), 0, s>>>(X, ldx, Y, ldy, N, h, w, C);
}

__global__ void __launch_bounds__(256) up2_bwd_kernel(const float* __restrict__ DY, int ldd, float* DX, int ldx, int N,
                                                      int h, int w, int C) {
  const int H = 2 * h, Wd = 2 * w, q = threadIdx.x;
  const unsigned total = (unsigned)N * h * w;
  const float wt[4] = {0.25f, 0.75f, 0.75f, 0.25f};
  for (unsigned pix = blockIdx.x * blockDim.y + threadIdx.y; pix < total; pix += gridDim.x * blockDim.y) {
    const int jx = pix % (unsigned)w;
    const unsigned t = pix / (unsigned)w;
    const int jy = t % (unsigned)h;
    const unsigned n = t / (unsigned)h;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* b = DY + (size_t)n * H * Wd * ldd + q * 4;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      int oy = min(max(2 * jy - 1 + a, 0), H - 1);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int ox = min(max(2 * jx - 1 + c, 0), Wd - 1);
        float4 v = *reinterpret_cast<const float4*>(b + (size_t)(oy * Wd + ox) * ldd);
        float ww = wt[a] * wt[c];
        acc.x = fmaf(ww, v.x, acc.x); acc.y = fmaf(ww, v.y, acc.y);
        acc.z = fmaf(ww, v.z, acc.z); acc.w = fmaf(ww, v.w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(DX + (size_t)pix * ldx + q * 4) = acc;
  }
}
void launch_up2_bwd(const float* DY, int ldd, float* DX, int ldx, int N, int h, int w, int C, cudaStream_t s) {
  up2_bwd_kernel<<<quad_grid((long long)N * h * w, C), quad_block(C), 0, s>>>(DY, ldd, DX, ldx, N, h, w, C);
}

// The DoubleLightConv form of up2 (train.cu: dlc_fwd runs both 1x1 convs on the upsampled tensor at LOW resolution).
//   up2_split: X [N,h,w][C0 + C1] -> Y0 [N,2h,2w][C0] (= z of conv.0.conv1, with its BatchNorm statistics: column sum and sum
//              of squares into `sums`) and Y1 [N,2h,2w][C1] (= the residual branch): one read of X, no separate statistics pass.
// thread = one LOW-resolution pixel x channel quad: its 3 x 3 neighbourhood (clamped) gives the 2 x 2 output pixels it covers --
// 9 loads for 4 outputs where the per-output form needs 16, and a third of the index arithmetic
__global__ void __launch_bounds__(256) up2_split_kernel(const float* __restrict__ X, int ldx, float* Y0, int ldy0, int C0,
                                                        float* Y1, int ldy1, int N, int h, int w, int C, double* sums) {
  extern __shared__ double ssum[];      // [2][C0]
  const int Wd = 2 * w, q = threadIdx.x;
  const int nthr = blockDim.x * blockDim.y, tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int e = tid; e < 2 * C0; e += nthr) ssum[e] = 0.0;
  __syncthreads();
  const unsigned total = (unsigned)N * h * w;
  const bool first = q * 4 < C0;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  for (unsigned pix = blockIdx.x * blockDim.y + threadIdx.y; pix < total; pix += gridDim.x * blockDim.y) {
    const int jx = pix % (unsigned)w;
    const unsigned t = pix / (unsigned)w;
    const int jy = t % (unsigned)h;
    const unsigned n = t / (unsigned)h;
    const int ym = max(jy - 1, 0), yp = min(jy + 1, h - 1), xm = max(jx - 1, 0), xp = min(jx + 1, w - 1);
    const float* b = X + (size_t)n * h * w * ldx + q * 4;
    float4 a[3][3];
    const int ys[3] = {ym, jy, yp}, xs[3] = {xm, jx, xp};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) a[r][c] = *reinterpret_cast<const float4*>(b + (size_t)(ys[r] * w + xs[c]) * ldx);
    float* o0 = first ? Y0 + q * 4 : Y1 + (q * 4 - C0);
    const int ld = first ? ldy0 : ldy1;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      // output row 2 jy + dy: rows (jy - 1, jy) with weights (.25, .75) for dy = 0, rows (jy, jy + 1) with (.75, .25) for dy = 1
      const int r0 = dy, r1 = dy + 1;
      const float wy0 = dy ? 0.75f : 0.25f, wy1 = 1.f - wy0;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int c0 = dx, c1 = dx + 1;
        const float wx0 = dx ? 0.75f : 0.25f, wx1 = 1.f - wx0;
        float4 o;
        // same association as ATen's upsample_bilinear2d: w_y0*(w_x0*a00 + w_x1*a01) + w_y1*(w_x0*a10 + w_x1*a11)
        o.x = wy0 * (wx0 * a[r0][c0].x + wx1 * a[r0][c1].x) + wy1 * (wx0 * a[r1][c0].x + wx1 * a[r1][c1].x);
        o.y = wy0 * (wx0 * a[r0][c0].y + wx1 * a[r0][c1].y) + wy1 * (wx0 * a[r1][c0].y + wx1 * a[r1][c1].y);
        o.z = wy0 * (wx0 * a[r0][c0].z + wx1 * a[r0][c1].z) + wy1 * (wx0 * a[r1][c0].z + wx1 * a[r1][c1].z);
        o.w = wy0 * (wx0 * a[r0][c0].w + wx1 * a[r0][c1].w) + wy1 * (wx0 * a[r1][c0].w + wx1 * a[r1][c1].w);
        const size_t opix = ((size_t)n * 2 * h + 2 * jy + dy) * Wd + 2 * jx + dx;
        *reinterpret_cast<float4*>(o0 + opix * ld) = o;
        if (first) {
          s1[0] += o.x; s1[1] += o.y; s1[2] += o.z; s1[3] += o.w;
          s2[0] = fmaf(o.x, o.x, s2[0]); s2[1] = fmaf(o.y, o.y, s2[1]); s2[2] = fmaf(o.z, o.z, s2[2]); s2[3] = fmaf(o.w, o.w, s2[3]);
        }
      }
    }
  }
  if (first) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { atomicAdd(&ssum[q * 4 + j], (double)s1[j]); atomicAdd(&ssum[C0 + q * 4 + j], (double)s2[j]); }
  }
  __syncthreads();
  for (int e = tid; e < 2 * C0; e += nthr) atomicAdd(&sums[e], ssum[e]);
}
void launch_up2_split(const float* X, int ldx, float* Y0, int ldy0, int C0, float* Y1, int ldy1, int C1, int N, int h, int w,
                      double* sums, cudaStream_t s) {
  const int C = C0 + C1;
  up2_split_kernel<<<quad_grid((long long)N * h * w, C, 8), quad_block(C), (size_t)2 * C0 * sizeof(double), s>>>(
      X, ldx, Y0, ldy0, C0, Y1, ldy1, N, h, w, C, sums);
}

// up2^T for the DoubleLightConv backward, tiled: DX [N,h,w][C0 + C1] = up2^T of [ BatchNorm-backward(DY0, Z0) | DY1 ].
// The gather form above reads every full-resolution element four times (once per low-res pixel that uses it); here a CTA
// stages the 18x18 full-resolution patch of an 8x8 low-res tile ONCE in shared memory (border slots hold the clamped pixel, which
// is exactly the adjoint's clamped OUTPUT index), and for the first C0 channels forms dz = gamma*invstd*(dy - S1/M - zhat*S2/M) of
// conv.0.conv1 (no activation) while staging, so neither dz nor a second pass over dy / z exists.  CTA = 16-channel group x
// tile; thread = one low-res pixel x channel quad; pixel pitch 24 floats (stride-2 pixel reads of a quarter-warp hit distinct banks).
struct Up2BwdP {
  const float* DY0; int ldd0; const float* Z0; int ldz0; BnRef bn; const double* sums; double invM; float* dgamma; float* dbeta;
  int C0; const float* DY1; int ldd1; float* DX; int ldx; int N, h, w, tiles_x, tiles_y;
};
__global__ void __launch_bounds__(256) up2_bwd_tiled_kernel(const Up2BwdP p) {
  constexpr int TL = 8, TH = 2 * TL + 2, PS = 24;
  __shared__ __align__(16) float sT[TH * TH * PS];
  const int tid = threadIdx.x, q = tid & 3;
  const int H = 2 * p.h, Wd = 2 * p.w;
  int t = blockIdx.x;
  const int lx0 = (t % p.tiles_x) * TL; t /= p.tiles_x;
  const int ly0 = (t % p.tiles_y) * TL;
  const int n = t / p.tiles_y;
  const int cg = blockIdx.y * 16;                      // channel group inside [C0 + C1]
  const bool first = cg < p.C0;
  float k1[4] = {1.f, 1.f, 1.f, 1.f}, k0[4] = {0.f, 0.f, 0.f, 0.f}, kz[4] = {0.f, 0.f, 0.f, 0.f}, mus[4] = {0.f, 0.f, 0.f, 0.f};
  if (first) {                                         // dz = k1 * dy + kz * (z - mean) + k0
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = cg + q * 4 + j;
      const float mu = p.bn.mean[c], is = p.bn.invstd[c], ga = p.bn.gamma[c];
      const float m1 = (float)(p.sums[c] * p.invM), m2 = (float)(p.sums[p.C0 + c] * p.invM);
      k1[j] = ga * is; kz[j] = -ga * is * is * m2; k0[j] = -ga * is * m1; mus[j] = mu;
      if (blockIdx.x == 0 && tid < 4) { p.dbeta[c] += (float)p.sums[c]; p.dgamma[c] += (float)p.sums[p.C0 + c]; }
    }
  }
  const float* D = first ? p.DY0 + cg + q * 4 : p.DY1 + (cg - p.C0) + q * 4;
  const int ld = first ? p.ldd0 : p.ldd1;
  const size_t nb = (size_t)n * H * Wd;
  constexpr int TOT = TH * TH * 4, NL = (TOT + 255) / 256;
  float4 dv[NL], zv[NL];
  int pr = (tid >> 2) / TH, pc = (tid >> 2) % TH;      // patch row / column of item u: 64 pixels further each time, no division
#pragma unroll
  for (int u = 0; u < NL; ++u) {
    const int i = tid + u * 256;
    dv[u] = make_float4(0.f, 0.f, 0.f, 0.f); zv[u] = dv[u];
    if (i < TOT) {
      const int oy = min(max(2 * ly0 - 1 + pr, 0), H - 1), ox = min(max(2 * lx0 - 1 + pc, 0), Wd - 1);
      pr += 64 / TH; pc += 64 % TH;
      if (pc >= TH) { pc -= TH; ++pr; }
      const size_t pix = nb + (size_t)oy * Wd + ox;
      dv[u] = *reinterpret_cast<const float4*>(D + pix * ld);
      if (first) zv[u] = *reinterpret_cast<const float4*>(p.Z0 + pix * p.ldz0 + cg + q * 4);
    }
  }
#pragma unroll
  for (int u = 0; u < NL; ++u) {
    const int i = tid + u * 256;
    if (i >= TOT) break;
    float4 o = dv[u];
    if (first) {
      o.x = fmaf(k1[0], o.x, fmaf(kz[0], zv[u].x - mus[0], k0[0])); o.y = fmaf(k1[1], o.y, fmaf(kz[1], zv[u].y - mus[1], k0[1]));
      o.z = fmaf(k1[2], o.z, fmaf(kz[2], zv[u].z - mus[2], k0[2])); o.w = fmaf(k1[3], o.w, fmaf(kz[3], zv[u].w - mus[3], k0[3]));
    }
    *reinterpret_cast<float4*>(sT + (i >> 2) * PS + q * 4) = o;
  }
  __syncthreads();
  const int lx = (tid >> 2) & 7, ly = tid >> 5;
  const int jx = lx0 + lx, jy = ly0 + ly;
  if (jx >= p.w || jy >= p.h) return;
  const float wt[4] = {0.25f, 0.75f, 0.75f, 0.25f};
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(sT + ((2 * ly + a) * TH + 2 * lx + c) * PS + q * 4);
      row.x = fmaf(wt[c], v.x, row.x); row.y = fmaf(wt[c], v.y, row.y); row.z = fmaf(wt[c], v.z, row.z); row.w = fmaf(wt[c], v.w, row.w);
    }
    acc.x = fmaf(wt[a], row.x, acc.x); acc.y = fmaf(wt[a], row.y, acc.y); acc.z = fmaf(wt[a], row.z, acc.z); acc.w = fmaf(wt[a], row.w, acc.w);
  }
  *reinterpret_cast<float4*>(p.DX + ((size_t)(n * p.h + jy) * p.w + jx) * p.ldx + cg + q * 4) = acc;
}
// false when the channel counts are not multiples of 16 (the caller runs bn_bwd_apply + up2_bwd)
bool launch_up2_bwd_tiled(const float* DY0, int ldd0, const float* Z0, int ldz0, const BnRef& bn, const double* sums, long long M,
                          float* dgamma, float* dbeta, int C0, const float* DY1, int ldd1, int C1, float* DX, int ldx, int N,
                          int h, int w, cudaStream_t s) {
  static const bool off = getenv("YSP_TRAIN_NO_UP2T") != nullptr;
  if (off || (C0 & 15) || (C1 & 15) || ((ldd0 | ldz0 | ldd1 | ldx) & 3)) return false;
  Up2BwdP p = {DY0, ldd0, Z0, ldz0, bn, sums, 1.0 / (double)M, dgamma, dbeta, C0, DY1, ldd1, DX, ldx, N, h, w, (w + 7) / 8, (h + 7) / 8};
  up2_bwd_tiled_kernel<<<dim3((unsigned)(N * p.tiles_x * p.tiles_y), (unsigned)((C0 + C1) / 16)), 256, 0, s>>>(p);
  return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// ECA (YOLOSegPlusPlus.py:60-88): gate[n][c] = sigmoid(sum_j w3[j] * mean[n][c+j-1]);  y = x * gate.
// eca_gate: pooled sums (double, [n][2][C], slot 0) -> mean, gate.  eca_gate_bwd: dgate sums (slot 0 of dsum) ->
// dmean[n][c] (already divided by HW) and dw3 (atomics).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void eca_gate_kernel(const double* __restrict__ pool, const float* __restrict__ w3, float* mean, float* gate,
                                int N, int C, double HW) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N * C) return;
  int n = e / C, c = e % C;
  const double* p = pool + (size_t)n * 2 * C;
  float m0 = c > 0 ? (float)(p[c - 1] / HW) : 0.f, m1 = (float)(p[c] / HW), m2 = c + 1 < C ? (float)(p[c + 1] / HW) : 0.f;
  float v = w3[0] * m0 + w3[1] * m1 + w3[2] * m2;
  mean[e] = m1;
  gate[e] = 1.f / (1.f + expf(-v));
}
void launch_eca_gate(const double* pool, const float* w3, float* mean, float* gate, int N, int C, long long HW, cudaStream_t s) {
  eca_gate_kernel<<<cdivl((long long)N * C, 128), 128, 0, s>>>(pool, w3, mean, gate, N, C, (double)HW);
}

__global__ void eca_gate_bwd_kernel(const double* __restrict__ dsum, const float* __restrict__ w3,
                                    const float* __restrict__ mean, const float* __restrict__ gate, float* dmean,
                                    float* dw3, int N, int C, float invHW) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  float l0 = 0.f, l1 = 0.f, l2 = 0.f;
  if (e < N * C) {
    int n = e / C, c = e % C;
    const double* ds = dsum + (size_t)n * 2 * C;
    const float* g = gate + (size_t)n * C;
    const float* mu = mean + (size_t)n * C;
    // dv[c'] = dgate[c'] * g(1-g);  dmean[c] = sum_j w3[j] * dv[c - j + 1]
    auto dv = [&](int cc) -> float { return (cc < 0 || cc >= C) ? 0.f : (float)ds[cc] * g[cc] * (1.f - g[cc]); };
    dmean[e] = (w3[0] * dv(c + 1) + w3[1] * dv(c) + w3[2] * dv(c - 1)) * invHW;
    float d = dv(c);
    l0 = c > 0 ? d * mu[c - 1] : 0.f;
    l1 = d * mu[c];
    l2 = c + 1 < C ? d * mu[c + 1] : 0.f;
  }
  // block reduction of the three weight-gradient terms
  __shared__ float red[3][128];
  red[0][threadIdx.x] = l0; red[1][threadIdx.x] = l1; red[2][threadIdx.x] = l2;
  __syncthreads();
  for (int st = 64; st > 0; st >>= 1) {
    if (threadIdx.x < st)
      for (int j = 0; j < 3; ++j) red[j][threadIdx.x] += red[j][threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x < 3) atomicAdd(&dw3[threadIdx.x], red[threadIdx.x][0]);
}
void launch_eca_gate_bwd(const double* dsum, const float* w3, const float* mean, const float* gate, float* dmean, float* dw3,
                         int N, int C, long long HW, cudaStream_t s) {
  eca_gate_bwd_kernel<<<cdivl((long long)N * C, 128), 128, 0, s>>>(dsum, w3, mean, gate, dmean, dw3, N, C, 1.f / (float)HW);
}

// out[m][c] = a[m][c] * g[n][c] (+ add[n][c])      n = m / HW       (ECA scale forward; backward dx = dy*gate + dmean)
__global__ void __launch_bounds__(256) scale_rows_kernel(const float* __restrict__ A, int lda, const float* __restrict__ G,
                                                         const float* __restrict__ ADD, float* O, int ldo, int C,
                                                         long long M, long long HW) {
  const int Q = C >> 2;
  const long long total = M * Q;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    int q = (int)(e % Q);
    long long m = e / Q, n = m / HW;
    float4 a = *reinterpret_cast<const float4*>(A + m * lda + q * 4);
    float4 g = *reinterpret_cast<const float4*>(G + n * C + q * 4);
    float4 o = make_float4(a.x * g.x, a.y * g.y, a.z * g.z, a.w * g.w);
    if (ADD) { float4 d = *reinterpret_cast<const float4*>(ADD + n * C + q * 4); o.x += d.x; o.y += d.y; o.z += d.z; o.w += d.w; }
    *reinterpret_cast<float4*>(O + m * ldo + q * 4) = o;
  }
}
void launch_scale_rows(const float* A, int lda, const float* G, const float* ADD, float* O, int ldo, int C, long long M,
                       long long HW, cudaStream_t s) {
  int grid = (int)std::min<long long>(cdivl(M * (C / 4), 256), 148 * 16);
  scale_rows_kernel<<<grid, 256, 0, s>>>(A, lda, G, ADD, O, ldo, C, M, HW);
}

// ---------------------------------------------------------------------------------------------------------------------
// Loss (train.py:98-104): monai DiceLoss(sigmoid=True, soft_label=True, batch=True, smooth 1e-5) over the local batch:
//   p = sigmoid(x);  P = sum p, T = sum t, A = sum |p - t|;  tp = (P+T-A)/2;  L = 1 - (2tp + eps)/(P + T + eps)
// kind 1 adds mean BCE-with-logits (BASELINE cfg 4 wording "Dice+BCE"; not in the reference).
// acc (double[4]) = P, T, A, sum bce.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) loss_reduce_kernel(const float* __restrict__ X, const float* __restrict__ T,
                                                          long long n, double* acc) {
  double s[4] = {0, 0, 0, 0};
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n; e += (long long)gridDim.x * 256) {
    float x = X[e], t = T[e];
    float p = 1.f / (1.f + expf(-x));
    s[0] += p; s[1] += t; s[2] += fabsf(p - t);
    s[3] += fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
  }
  __shared__ double red[4][256];
  for (int j = 0; j < 4; ++j) red[j][threadIdx.x] = s[j];
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st)
      for (int j = 0; j < 4; ++j) red[j][threadIdx.x] += red[j][threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x < 4) atomicAdd(&acc[threadIdx.x], red[threadIdx.x][0]);
}

__global__ void __launch_bounds__(256) loss_grad_kernel(const float* __restrict__ X, const float* __restrict__ T,
                                                        long long n, const double* __restrict__ acc, int kind,
                                                        float grad_scale, float* DX, float* loss_out) {
  const double P = acc[0], Tt = acc[1], A = acc[2], eps = 1e-5;
  const double D = P + Tt + eps, num = (P + Tt - A) + eps;   // 2 tp + eps
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    loss_out[0] = (float)(1.0 - num / D + (kind == 1 ? acc[3] / (double)n : 0.0));
    loss_out[1] = (float)(1.0 - num / D);
    loss_out[2] = kind == 1 ? (float)(acc[3] / (double)n) : 0.f;
  }
  const float c1 = (float)(1.0 / D), c2 = (float)(num / (D * D)), invn = 1.f / (float)n;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n; e += (long long)gridDim.x * 256) {
    float x = X[e], t = T[e];
    float p = 1.f / (1.f + expf(-x));
    float sg = p > t ? 1.f : (p < t ? -1.f : 0.f);
    float dLdp = -((1.f - sg) * c1 - c2);
    float g = dLdp * p * (1.f - p);
    if (kind == 1) g += (p - t) * invn;
    DX[e] = g * grad_scale;
  }
}

__global__ void loss_value_kernel(const double* __restrict__ acc, long long n, int kind, float* loss_out) {
  const double P = acc[0], Tt = acc[1], A = acc[2], eps = 1e-5;
  const double dice = 1.0 - ((P + Tt - A) + eps) / (P + Tt + eps), bce = kind == 1 ? acc[3] / (double)n : 0.0;
  loss_out[0] = (float)(dice + bce); loss_out[1] = (float)dice; loss_out[2] = (float)bce;
}
// loss only (validation, train.py:356-357): same reductions as launch_loss, no gradient
void launch_loss_value(const float* X, const float* T, long long n, double* acc, int kind, float* loss_out, cudaStream_t s) {
  int grid = (int)std::min<long long>(cdivl(n, 256), 148 * 8);
  cudaMemsetAsync(acc, 0, 4 * sizeof(double), s);
  loss_reduce_kernel<<<grid, 256, 0, s>>>(X, T, n, acc);
  loss_value_kernel<<<1, 1, 0, s>>>(acc, n, kind, loss_out);
}

void launch_loss(const float* X, const float* T, long long n, double* acc, int kind, float grad_scale, float* DX,
                 float* loss_out, cudaStream_t s) {
  int grid = (int)std::min<long long>(cdivl(n, 256), 148 * 8);
  loss_reduce_kernel<<<grid, 256, 0, s>>>(X, T, n, acc);
  loss_grad_kernel<<<grid, 256, 0, s>>>(X, T, n, acc, kind, grad_scale, DX, loss_out);
}

// ---------------------------------------------------------------------------------------------------------------------
// Optimiser: torch.optim.AdamW (decoupled weight decay, bias-corrected), optional global-norm clip
// (clip_grad_norm_ semantics: scale = min(1, max_norm/(norm + 1e-6)); 0 disables -- what the reference effectively
// runs) and a gradient pre-scale (1/world_size after a SUM all-reduce).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ g, long long n, double* out) {
  double s = 0;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n; e += (long long)gridDim.x * 256) s += (double)g[e] * g[e];
  __shared__ double red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) { if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st]; __syncthreads(); }
  if (threadIdx.x == 0) atomicAdd(out, red[0]);
}

__global__ void __launch_bounds__(256) adamw_kernel(float* p, const float* __restrict__ g, float* m, float* v, long long n,
                                                    float lr, float b1, float b2, float eps, float wd, float bc1,
                                                    float bc2_sqrt, float gscale, float max_norm,
                                                    const double* __restrict__ sqn) {
  float clip = 1.f;
  if (max_norm > 0.f) {
    float nrm = (float)sqrt(*sqn) * gscale;
    clip = fminf(1.f, max_norm / (nrm + 1e-6f));
  }
  const float sc = gscale * clip;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n; e += (long long)gridDim.x * 256) {
    float gr = g[e] * sc;
    float pe = p[e] * (1.f - lr * wd);
    float me = b1 * m[e] + (1.f - b1) * gr;
    float ve = b2 * v[e] + (1.f - b2) * gr * gr;
    m[e] = me; v[e] = ve;
    float denom = sqrtf(ve) / bc2_sqrt + eps;
    p[e] = pe - (lr / bc1) * (me / denom);
  }
}

void launch_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                  float wd, int step, float gscale, float max_norm, double* sqn_ws, cudaStream_t s) {
  int grid = (int)std::min<long long>(cdivl(n, 256), 148 * 4);
  if (max_norm > 0.f) {
    cudaMemsetAsync(sqn_ws, 0, sizeof(double), s);
    sqnorm_kernel<<<grid, 256, 0, s>>>(g, n, sqn_ws);
  }
  double bc1 = 1.0 - std::pow((double)b1, (double)step), bc2 = 1.0 - std::pow((double)b2, (double)step);
  adamw_kernel<<<grid, 256, 0, s>>>(p, g, m, v, n, lr, b1, b2, eps, wd, (float)bc1, (float)std::sqrt(bc2), gscale, max_norm,
                                    sqn_ws);
}

void launch_sqnorm(const float* g, long long n, double* out, cudaStream_t s) {
  cudaMemsetAsync(out, 0, sizeof(double), s);
  sqnorm_kernel<<<(int)std::min<long long>(cdivl(n, 256), 148 * 4), 256, 0, s>>>(g, n, out);
}

// NHWC activation view (fp32 or bf16) -> dense fp32 [M][C]  (hands the frozen encoder's skips to the trainer)
template <typename T>
__global__ void __launch_bounds__(256) export_view_kernel(const T* __restrict__ in, int in_cs, float* out, int C, long long M) {
  const long long total = M * C;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    int c = (int)(e % C);
    long long m = e / C;
    out[e] = to_f<T>(in[m * in_cs + c]);
  }
}
void launch_export_view(const void* in, int in_cs, int dt, float* out, int C, long long M, cudaStream_t s) {
  if (dt == DT_F32 && ((C | in_cs) & 3) == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    launch_add_copy(reinterpret_cast<const float*>(in), in_cs, nullptr, 0, out, C, C, M, s);      // a strided fp32 copy
    return;
  }
  int grid = (int)std::min<long long>(cdivl(M * C, 256), 148 * 16);
  if (dt == DT_F32) export_view_kernel<float><<<grid, 256, 0, s>>>((const float*)in, in_cs, out, C, M);
  else export_view_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)in, in_cs, out, C, M);
}

}  // namespace ysp
