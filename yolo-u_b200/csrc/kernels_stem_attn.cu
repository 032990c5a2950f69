// kernels_stem_attn.cu -- (1) the 4-channel stem convolution fused with input normalisation/layout conversion,
//                         (2) the area-attention core on mma.sync bf16 tensor-core tiles (64-token areas, head_dim 32).
#include "kernels.h"

namespace ysp {

static inline int cdiv2(long long a, long long b) { return (int)((a + b - 1) / b); }

// =====================================================================================================================
// Stem: detector layer 0 / seg encoder.0 = Conv(4, 16, k=3, s=2) + folded BN + SiLU.  K = 36 is far
// too small for a tensor-core tile and Cin = 4 cannot feed TMA/UMMA chunks, so this is a direct FFMA kernel that
// reads the CALLER's tensor -- fp32 NCHW [B,4,H,W] (evaluate_model.py:136) or u8 HWC4 (a1: x/255 fused) -- and writes
// NHWC activations.  Reads beyond H/W are zero: conv padding and the bottom/right zero-pad of decision D1.
// One thread = 2 horizontally adjacent output pixels x 16 output channels (weights broadcast from shared memory).
// =====================================================================================================================
// byte / 255.f, correctly rounded (= ToTensor's true division) in three FP ops instead of an IEEE divide:
// q0 = x*r, q = fma(fma(-q0, 255, x), r, q0) with r = fl(1/255); equal to x / 255.f for all 256 byte values (checked
// exhaustively on the host and, end to end, by the u8-vs-fp32 input equality test).
__device__ __forceinline__ float div255(unsigned char b) {
  const float x = (float)b, r = 1.0f / 255.0f;
  const float q0 = x * r;
  return fmaf(fmaf(-q0, 255.0f, x), r, q0);
}

template <typename T, bool U8>
__global__ void __launch_bounds__(128) stem_conv_kernel(const void* __restrict__ in, T* __restrict__ out,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        int N, int H, int W, int OH, int OW, int out_cs, int wld) {
  __shared__ __align__(16) float sw[36 * 16];
  __shared__ float sb[16];
  pdl_sync();
  for (int i = threadIdx.x; i < 36 * 16; i += 128) sw[i] = w[(i >> 4) * wld + (i & 15)];
  if (threadIdx.x < 16) sb[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int OW2 = (OW + 1) >> 1;
  long long e = (long long)blockIdx.x * 128 + threadIdx.x;
  if (e >= (long long)N * OH * OW2) return;
  const int oxp = (int)(e % OW2);
  long long t = e / OW2;
  const int oy = (int)(t % OH), n = (int)(t / OH);
  const int ox0 = oxp * 2;
  // accumulators as fp32 pairs: one FFMA2 (sm_100 packed fp32 FMA, scalar x broadcast) does two of the 16 channels --
  // half the FMA issue slots of this issue-bound kernel, bit-identical to scalar fmaf
  float2 acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = make_float2(sb[2 * j], sb[2 * j + 1]); acc[1][j] = acc[0][j]; }
  const size_t HW = (size_t)H * W;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int iy = 2 * oy - 1 + r;
    if (iy < 0 || iy >= H) continue;
    // the two output pixels read input columns 2*ox0-1 .. 2*ox0+3 (5 columns x 4 channels)
    float x[5][4];
#pragma unroll
    for (int c5 = 0; c5 < 5; ++c5) {
      const int ix = 2 * ox0 - 1 + c5;
      if (ix >= 0 && ix < W) {
        if (U8) {
          uchar4 v = reinterpret_cast<const uchar4*>(in)[((size_t)n * H + iy) * W + ix];
          x[c5][0] = div255(v.x); x[c5][1] = div255(v.y); x[c5][2] = div255(v.z); x[c5][3] = div255(v.w);
        } else {
          const float* p = reinterpret_cast<const float*>(in) + (size_t)n * 4 * HW + (size_t)iy * W + ix;
          x[c5][0] = p[0]; x[c5][1] = p[HW]; x[c5][2] = p[2 * HW]; x[c5][3] = p[3 * HW];
        }
      } else {
        x[c5][0] = x[c5][1] = x[c5][2] = x[c5][3] = 0.f;
      }
    }
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const float* wr = sw + ((r * 3 + s) * 4 + ci) * 16;
        const float x0 = x[s][ci], x1 = x[s + 2][ci];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 wv = *reinterpret_cast<const float4*>(wr + j4 * 4);
          const float2 wlo = make_float2(wv.x, wv.y), whi = make_float2(wv.z, wv.w);
          acc[0][j4 * 2] = __ffma2_rn(make_float2(x0, x0), wlo, acc[0][j4 * 2]); acc[0][j4 * 2 + 1] = __ffma2_rn(make_float2(x0, x0), whi, acc[0][j4 * 2 + 1]);
          acc[1][j4 * 2] = __ffma2_rn(make_float2(x1, x1), wlo, acc[1][j4 * 2]); acc[1][j4 * 2 + 1] = __ffma2_rn(make_float2(x1, x1), whi, acc[1][j4 * 2 + 1]);
        }
      }
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int ox = ox0 + q;
    if (ox >= OW) break;
    T* o = out + (((size_t)n * OH + oy) * OW + ox) * out_cs;
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      F4 v;
#pragma unroll
      for (int j = 0; j < 2; ++j) { v.v[2 * j] = silu_for<T>(acc[q][j4 * 2 + j].x); v.v[2 * j + 1] = silu_for<T>(acc[q][j4 * 2 + j].y); }
      store4<T>(o + j4 * 4, v);
    }
  }
}

void launch_stem_conv(const void* in, int in_u8, void* out, const float* w, const float* bias, int N, int H, int W, int OH,
                      int OW, int out_cs, int wld, int dt, cudaStream_t s) {
  long long total = (long long)N * OH * ((OW + 1) / 2);
  int g = cdiv2(total, 128);
  if (dt == DT_F32) {
    if (in_u8) launch_pdl(stem_conv_kernel<float, true>, dim3(g), dim3(128), 0, s, in, (float*)out, w, bias, N, H, W, OH, OW, out_cs, wld);
    else launch_pdl(stem_conv_kernel<float, false>, dim3(g), dim3(128), 0, s, in, (float*)out, w, bias, N, H, W, OH, OW, out_cs, wld);
  } else {
    if (in_u8) launch_pdl(stem_conv_kernel<bf16, true>, dim3(g), dim3(128), 0, s, in, (bf16*)out, w, bias, N, H, W, OH, OW, out_cs, wld);
    else launch_pdl(stem_conv_kernel<bf16, false>, dim3(g), dim3(128), 0, s, in, (bf16*)out, w, bias, N, H, W, OH, OW, out_cs, wld);
  }
}

// =====================================================================================================================
// Area attention, bf16 tensor-core version for 64-token areas (the YOLOv12n shapes at 256x256: layer 6 = 256 tokens in
// 4 areas, layer 8 = 64 tokens in 1 area; head_dim 32).  One CTA = one (slice, area, head); 4 warps x 16 queries.
//   S = Q K^T (mma m16n8k16, fp32 accumulate) -> row softmax in registers (exp2, quad shuffles) -> O = P V with P
//   re-used from the S accumulator fragments as the A operand (flash-attention register trick) -> bf16 NHWC store.
// K is staged [token][dim] and V transposed [dim][token] in shared memory so both B operands are 32-bit pair loads.
// =====================================================================================================================
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(128) attention64_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int Ntok,
                                                          int area, int qkv_cs, int out_cs) {
  constexpr int NT = 64, HD = 32, KS = HD + 8, VS = NT + 8;
  pdl_sync();
  __shared__ __align__(16) bf16 sK[NT * KS];
  __shared__ __align__(16) bf16 sV[HD * VS];
  const int b = blockIdx.x / area, ar = blockIdx.x % area, h = blockIdx.y;
  const size_t tok0 = (size_t)b * Ntok + (size_t)ar * NT;
  const bf16* base = qkv + tok0 * qkv_cs + h * 3 * HD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tig = lane & 3;
  // stage K (row-major) and V (transposed): 64 tokens x 32 dims each, 8-byte vector loads
  {
    // all eight 8-byte loads of a thread are issued before its first shared-memory store
    constexpr int NL = NT * (HD / 4) / 128;
    uint2 kk[NL], vv[NL];
#pragma unroll
    for (int u = 0; u < NL; ++u) {
      const int i = tid + u * 128;
      const int tok = i / (HD / 4), d4 = (i % (HD / 4)) * 4;
      kk[u] = *reinterpret_cast<const uint2*>(base + (size_t)tok * qkv_cs + HD + d4);
      vv[u] = *reinterpret_cast<const uint2*>(base + (size_t)tok * qkv_cs + 2 * HD + d4);
    }
#pragma unroll
    for (int u = 0; u < NL; ++u) {
      const int i = tid + u * 128;
      const int tok = i / (HD / 4), d4 = (i % (HD / 4)) * 4;
      *reinterpret_cast<uint2*>(sK + tok * KS + d4) = kk[u];
      const bf16* ve = reinterpret_cast<const bf16*>(&vv[u]);
#pragma unroll
      for (int j = 0; j < 4; ++j) sV[(d4 + j) * VS + tok] = ve[j];
    }
  }
  // Q fragments straight from global: rows g / g+8 of this warp's 16 queries, 2 k-steps of 16 dims
  uint32_t qa[2][4];
  {
    const bf16* q0 = base + (size_t)(warp * 16 + g) * qkv_cs + tig * 2;
    const bf16* q1 = q0 + (size_t)8 * qkv_cs;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      qa[ks][0] = *reinterpret_cast<const uint32_t*>(q0 + ks * 16);
      qa[ks][1] = *reinterpret_cast<const uint32_t*>(q1 + ks * 16);
      qa[ks][2] = *reinterpret_cast<const uint32_t*>(q0 + ks * 16 + 8);
      qa[ks][3] = *reinterpret_cast<const uint32_t*>(q1 + ks * 16 + 8);
    }
  }
  __syncthreads();
  // S = Q K^T : 8 n-tiles (tokens) x 2 k-steps
  float sacc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const bf16* kp = sK + (nt * 8 + g) * KS + ks * 16 + tig * 2;
      mma16816(sacc[nt], qa[ks], *reinterpret_cast<const uint32_t*>(kp), *reinterpret_cast<const uint32_t*>(kp + 8));
    }
  }
  // softmax over 64 keys: rows g (regs 0,1) and g+8 (regs 2,3); a row lives in the 4 lanes of a quad
  const float sc = 0.17677669529663687f * 1.4426950408889634f;     // 32^-0.5 * log2(e)
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { m0 = fmaxf(m0, fmaxf(sacc[nt][0], sacc[nt][1])); m1 = fmaxf(m1, fmaxf(sacc[nt][2], sacc[nt][3])); }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    sacc[nt][0] = exp2f((sacc[nt][0] - m0) * sc); sacc[nt][1] = exp2f((sacc[nt][1] - m0) * sc);
    sacc[nt][2] = exp2f((sacc[nt][2] - m1) * sc); sacc[nt][3] = exp2f((sacc[nt][3] - m1) * sc);
    l0 += sacc[nt][0] + sacc[nt][1]; l1 += sacc[nt][2] + sacc[nt][3];
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  // O = P V : 4 k-steps of 16 tokens, 4 n-tiles of 8 dims
  float oacc[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) oacc[nt][0] = oacc[nt][1] = oacc[nt][2] = oacc[nt][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t pa[4];
    pa[0] = pack_bf16(sacc[2 * ks][0], sacc[2 * ks][1]);       pa[1] = pack_bf16(sacc[2 * ks][2], sacc[2 * ks][3]);
    pa[2] = pack_bf16(sacc[2 * ks + 1][0], sacc[2 * ks + 1][1]); pa[3] = pack_bf16(sacc[2 * ks + 1][2], sacc[2 * ks + 1][3]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const bf16* vp = sV + (nt * 8 + g) * VS + ks * 16 + tig * 2;
      mma16816(oacc[nt], pa, *reinterpret_cast<const uint32_t*>(vp), *reinterpret_cast<const uint32_t*>(vp + 8));
    }
  }
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  bf16* o0 = out + (tok0 + warp * 16 + g) * out_cs + h * HD + tig * 2;
  bf16* o1 = o0 + (size_t)8 * out_cs;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    *reinterpret_cast<uint32_t*>(o0 + nt * 8) = pack_bf16(oacc[nt][0] * i0, oacc[nt][1] * i0);
    *reinterpret_cast<uint32_t*>(o1 + nt * 8) = pack_bf16(oacc[nt][2] * i1, oacc[nt][3] * i1);
  }
}

void launch_attention64_bf16(const void* qkv, void* out, int B, int Ntok, int heads, int area, int qkv_cs, int out_cs,
                             cudaStream_t s) {
  dim3 g(B * area, heads);
  launch_pdl(attention64_kernel, g, dim3(128), 0, s, (const bf16*)qkv, (bf16*)out, Ntok, area, qkv_cs, out_cs);
}

}  // namespace ysp
