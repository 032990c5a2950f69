// attention_tc.cu -- YOLOv12 area attention (ultralytics AAttn.forward) on tcgen05 tensor cores, for the
// 64-token areas of the detector at 256 x 256 (layer 6: 256 tokens in 4 areas, layer 8: 64 tokens in 1 area; head_dim 32):
//     out[i, h*32 + d] = sum_j softmax_j(q_i . k_j * 32^-0.5) v_j[d],     qkv channel layout per head = [q32 | k32 | v32]
// One tile = TWO (slice, area, head) problems = heads (h, h+1) of one area, stacked on the M = 128 rows of the MMAs; thread t
// of the CTA's 128 owns row t (problem t / 64, token t % 64) from the global load to the final store:
//   1. stage: thread t loads its token's q | k | v (96 values) and writes them as tensor-core operands in the canonical
//      K-major NO-SWIZZLE layout ([8-element plane][row] x 16 B): Q and K row-wise, V transposed ([key plane][dim][8 keys]);
//   2. S = Q K^T : M = 128 (2 x 64 queries), N = 128 (2 x 64 keys), K = 32 -- one accumulator, only the two diagonal 64 x 64
//      blocks are used (the tensor pipe is far from being the bound; the alternative is two M = 64 instructions);
//   3. softmax of row t over ITS problem's 64 columns, straight out of TMEM into registers (tcgen05.ld), exp2 on the MUFU unit;
//   4. P (unnormalised, x 2^10 in parity mode so the fp16 lo parts stay normal) -> operand tile (over the dead Q / K tiles);
//      O_p = P[:, keys] V_p (accumulators over the dead S columns: 128 TMEM columns per CTA, four CTAs per SM) with
//      N = 32 per problem, rows of the other problem produce garbage in columns nobody reads;
//   5. row t reads O_{t/64}[t], scales by 1 / sum and stores 32 outputs.
// Parity mode (T = float): every operand x is split x = hi + lo (fp16) and each product is three kind::f16 MMAs
// (hi.hi + lo.hi + hi.lo, fp32 accumulation) like conv_tc32.cu.  Throughput mode (T = bf16): one bf16 MMA per product, the
// same rounding points as the mma.sync kernel this replaces (S from bf16 q/k, fp32 softmax, P rounded to bf16).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cmath>

#include "kernels.h"

namespace ysp {

namespace {

constexpr int HD = 32, NT = 64;
constexpr int QK_TILE = 4 * 128 * 16;          // 4 planes (8 dims) x 128 rows x 16 B
constexpr int V_LBO = 32 * 16 + 16;            // V^T plane pitch: 32 dims x 16 B, padded so the 2-byte transposing stores of
constexpr int V_TILE = 8 * V_LBO;              //   a warp (4 key planes x 8 keys) fall into different banks
constexpr int P_TILE = 8 * 128 * 16;           // 8 planes (8 keys) x 128 rows x 16 B

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWLA:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DNA;\n\tbra WLA;\n\tDNA:\n\t}"
               ::"r"(s32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) { split2_f16(a, b, hi, lo); }

}  // namespace

// SPLIT = true: T = float (parity mode, fp16 hi/lo);  SPLIT = false: T = bf16 (throughput mode)
template <typename T, bool SPLIT>
__global__ void __launch_bounds__(128) attention_tc_kernel(const T* __restrict__ qkv, T* __restrict__ out, int Ntok, int area,
                                                           int heads, int qkv_cs, int out_cs, int ntiles) {
  constexpr int NS = SPLIT ? 2 : 1;                      // operand tiles per matrix (hi, lo)
  extern __shared__ __align__(128) uint8_t asm_[];
  uint8_t* sQ = asm_;                                    // [NS][QK_TILE]
  uint8_t* sK = sQ + NS * QK_TILE;
  uint8_t* sV = sK + NS * QK_TILE;                       // [2 problems][NS][V_TILE]
  uint8_t* sP = sQ;                                      // [NS][P_TILE] ALIASES Q | K: both are dead once S = Q K^T has completed
  static_assert(P_TILE == 2 * QK_TILE, "P tile must fit exactly over the Q and K tiles");
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_s;
  const int t = threadIdx.x, warp = t >> 5;
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_s)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_sync();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  // instruction descriptors: D = F32; A = B = F16 (format 0) or BF16 (format 1); K-major; N >> 3 at bit 17, M >> 4 at bit 24
  const uint32_t fmt = SPLIT ? 0u : ((1u << 7) | (1u << 10));
  const uint32_t idesc_s = (1u << 4) | fmt | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc_o = (1u << 4) | fmt | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const int prob = t >> 6, tok = t & 63;
  const int hp = heads >> 1;                             // head pairs per (slice, area)
  uint32_t par = 0;

#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int h0 = (tile % hp) * 2, ba = tile / hp;      // ba = slice * area + area index
    const int b = ba / area, ar = ba - b * area;
    const size_t tok0 = (size_t)b * Ntok + (size_t)ar * NT;
    const T* row = qkv + (tok0 + tok) * qkv_cs + (h0 + prob) * 3 * HD;

    // ---- 1. stage q | k | v of (problem, token) = thread ----
    if (SPLIT) {
      const float4* r4 = reinterpret_cast<const float4*>(row);
      float4 v[24];
#pragma unroll
      for (int i = 0; i < 24; ++i) v[i] = r4[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {                       // Q, K: plane j = dims 8j .. 8j+7 of row t
        uint4 h, l;
        split2(v[2 * j].x, v[2 * j].y, h.x, l.x); split2(v[2 * j].z, v[2 * j].w, h.y, l.y);
        split2(v[2 * j + 1].x, v[2 * j + 1].y, h.z, l.z); split2(v[2 * j + 1].z, v[2 * j + 1].w, h.w, l.w);
        *reinterpret_cast<uint4*>(sQ + j * 2048 + t * 16) = h;
        *reinterpret_cast<uint4*>(sQ + QK_TILE + j * 2048 + t * 16) = l;
        split2(v[8 + 2 * j].x, v[8 + 2 * j].y, h.x, l.x); split2(v[8 + 2 * j].z, v[8 + 2 * j].w, h.y, l.y);
        split2(v[9 + 2 * j].x, v[9 + 2 * j].y, h.z, l.z); split2(v[9 + 2 * j].z, v[9 + 2 * j].w, h.w, l.w);
        *reinterpret_cast<uint4*>(sK + j * 2048 + t * 16) = h;
        *reinterpret_cast<uint4*>(sK + QK_TILE + j * 2048 + t * 16) = l;
      }
      uint8_t* vb = sV + prob * NS * V_TILE + (tok >> 3) * V_LBO + (tok & 7) * 2;   // V^T: [key plane][dim][key % 8]
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint32_t h0_, l0_, h1_, l1_;
        split2(v[16 + i].x, v[16 + i].y, h0_, l0_); split2(v[16 + i].z, v[16 + i].w, h1_, l1_);
        uint8_t* d = vb + (4 * i) * 16;
        *reinterpret_cast<uint16_t*>(d) = (uint16_t)h0_;            *reinterpret_cast<uint16_t*>(d + 16) = (uint16_t)(h0_ >> 16);
        *reinterpret_cast<uint16_t*>(d + 32) = (uint16_t)h1_;       *reinterpret_cast<uint16_t*>(d + 48) = (uint16_t)(h1_ >> 16);
        *reinterpret_cast<uint16_t*>(d + V_TILE) = (uint16_t)l0_;       *reinterpret_cast<uint16_t*>(d + V_TILE + 16) = (uint16_t)(l0_ >> 16);
        *reinterpret_cast<uint16_t*>(d + V_TILE + 32) = (uint16_t)l1_;  *reinterpret_cast<uint16_t*>(d + V_TILE + 48) = (uint16_t)(l1_ >> 16);
      }
    } else {
      const uint4* r4 = reinterpret_cast<const uint4*>(row);
      uint4 v[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) v[i] = r4[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        *reinterpret_cast<uint4*>(sQ + j * 2048 + t * 16) = v[j];
        *reinterpret_cast<uint4*>(sK + j * 2048 + t * 16) = v[4 + j];
      }
      uint8_t* vb = sV + prob * NS * V_TILE + (tok >> 3) * V_LBO + (tok & 7) * 2;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t w[4] = {v[8 + i].x, v[8 + i].y, v[8 + i].z, v[8 + i].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint8_t* d = vb + (8 * i + 2 * k) * 16;
          *reinterpret_cast<uint16_t*>(d) = (uint16_t)w[k];
          *reinterpret_cast<uint16_t*>(d + 16) = (uint16_t)(w[k] >> 16);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();

    // ---- 2. S = Q K^T ----
    if (t == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t q0 = s32(sQ), k0 = s32(sK);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint64_t ah = desc_nosw(q0 + ks * 4096, 2048u, 128u), bh = desc_nosw(k0 + ks * 4096, 2048u, 128u);
        umma(tmem, ah, bh, idesc_s, ks ? 1u : 0u);
        if (SPLIT) {
          const uint64_t al = desc_nosw(q0 + QK_TILE + ks * 4096, 2048u, 128u), bl = desc_nosw(k0 + QK_TILE + ks * 4096, 2048u, 128u);
          umma(tmem, al, bh, idesc_s, 1u);
          umma(tmem, ah, bl, idesc_s, 1u);
        }
      }
      commit(&bar[0]);
    }
    mbar_wait_parity(&bar[0], par);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- 3. softmax of row t over the 64 keys of its problem ----
    uint32_t sr[64];
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c = 0; c < 4; ++c) ld16(lane_addr + prob * 64 + c * 16, sr + c * 16);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const float sc = 0.17677669529663687f * 1.4426950408889634f;     // 32^-0.5 * log2(e)
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 64; ++j) mx = fmaxf(mx, __uint_as_float(sr[j]));
    float l = 0.f;
    const float pscale = SPLIT ? 1024.f : 1.f;             // keeps the fp16 lo parts of small probabilities normal
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      const float e = exp2f((__uint_as_float(sr[j]) - mx) * sc);
      l += e;
      sr[j] = __float_as_uint(e * pscale);
    }
    // ---- 4. P -> operand tile (row t, planes of 8 keys) ----
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (SPLIT) {
        uint4 h, lo;
        split2(__uint_as_float(sr[8 * j]), __uint_as_float(sr[8 * j + 1]), h.x, lo.x);
        split2(__uint_as_float(sr[8 * j + 2]), __uint_as_float(sr[8 * j + 3]), h.y, lo.y);
        split2(__uint_as_float(sr[8 * j + 4]), __uint_as_float(sr[8 * j + 5]), h.z, lo.z);
        split2(__uint_as_float(sr[8 * j + 6]), __uint_as_float(sr[8 * j + 7]), h.w, lo.w);
        *reinterpret_cast<uint4*>(sP + j * 2048 + t * 16) = h;
        *reinterpret_cast<uint4*>(sP + P_TILE + j * 2048 + t * 16) = lo;
      } else {
        uint4 h;
        uint32_t* hw = reinterpret_cast<uint32_t*>(&h);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __nv_bfloat162 pb = __floats2bfloat162_rn(__uint_as_float(sr[8 * j + 2 * k]), __uint_as_float(sr[8 * j + 2 * k + 1]));
          hw[k] = *reinterpret_cast<const uint32_t*>(&pb);
        }
        *reinterpret_cast<uint4*>(sP + j * 2048 + t * 16) = h;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (t == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t p0 = s32(sP);
#pragma unroll
      for (int pr = 0; pr < 2; ++pr) {
        const uint32_t v0 = s32(sV + pr * NS * V_TILE);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {                   // 16 keys per step
          const uint64_t ah = desc_nosw(p0 + ks * 4096, 2048u, 128u), bh = desc_nosw(v0 + ks * 2 * V_LBO, V_LBO, 128u);
          umma(tmem + pr * 32, ah, bh, idesc_o, ks ? 1u : 0u);
          if (SPLIT) {
            const uint64_t al = desc_nosw(p0 + P_TILE + ks * 4096, 2048u, 128u), bl = desc_nosw(v0 + V_TILE + ks * 2 * V_LBO, V_LBO, 128u);
            umma(tmem + pr * 32, al, bh, idesc_o, 1u);
            umma(tmem + pr * 32, ah, bl, idesc_o, 1u);
          }
        }
      }
      commit(&bar[1]);
    }
    mbar_wait_parity(&bar[1], par);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- 5. O_{prob}[t] / sum -> out ----
    uint32_t o[32];
    ld16(lane_addr + prob * 32, o);
    ld16(lane_addr + prob * 32 + 16, o + 16);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const float inv = 1.f / (l * pscale);
    T* op = out + (tok0 + tok) * out_cs + (h0 + prob) * HD;
    if (SPLIT) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(op)[j] = make_float4(__uint_as_float(o[4 * j]) * inv, __uint_as_float(o[4 * j + 1]) * inv,
                                                       __uint_as_float(o[4 * j + 2]) * inv, __uint_as_float(o[4 * j + 3]) * inv);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 w;
        uint32_t* ww = reinterpret_cast<uint32_t*>(&w);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __nv_bfloat162 pb = __floats2bfloat162_rn(__uint_as_float(o[8 * j + 2 * k]) * inv, __uint_as_float(o[8 * j + 2 * k + 1]) * inv);
          ww[k] = *reinterpret_cast<const uint32_t*>(&pb);
        }
        reinterpret_cast<uint4*>(op)[j] = w;
      }
    }
    // the next tile overwrites the operand tiles and both accumulators
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    par ^= 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
  }
}

bool attention_tc_supported(int Ntok, int heads, int area, int qkv_cs, int out_cs, int dt) {
  if (area < 1 || Ntok % area != 0 || Ntok / area != NT || (heads & 1)) return false;
  const int align = dt == DT_F32 ? 4 : 8;                // 16-byte row loads / stores
  return qkv_cs % align == 0 && out_cs % align == 0;
}

void launch_attention_tc(const void* qkv, void* out, int B, int Ntok, int heads, int area, int qkv_cs, int out_cs, int dt,
                         cudaStream_t s) {
  const int ntiles = B * area * (heads / 2);
  static int sms = 0;
  if (!sms) { int d = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d); if (sms <= 0) sms = 148; }
  const int grid = ntiles < 4 * sms ? ntiles : 4 * sms;
  if (dt == DT_F32) {
    constexpr size_t smem = 2 * (2 * QK_TILE + 2 * V_TILE);
    static unsigned long long done = 0;
    ensure_dyn_smem(attention_tc_kernel<float, true>, smem, done, "attention_tc_kernel<float>");
    launch_pdl(attention_tc_kernel<float, true>, dim3(grid), dim3(128), smem, s, (const float*)qkv, (float*)out, Ntok, area, heads,
               qkv_cs, out_cs, ntiles);
  } else {
    constexpr size_t smem = 2 * QK_TILE + 2 * V_TILE;
    static unsigned long long done = 0;
    ensure_dyn_smem(attention_tc_kernel<bf16, false>, smem, done, "attention_tc_kernel<bf16>");
    launch_pdl(attention_tc_kernel<bf16, false>, dim3(grid), dim3(128), smem, s, (const bf16*)qkv, (bf16*)out, Ntok, area, heads,
               qkv_cs, out_cs, ntiles);
  }
}

}  // namespace ysp
