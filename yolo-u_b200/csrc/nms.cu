// nms.cu -- batched box suppression for sm_100a (replaces reference nms.py:13-166 + torchvision.ops.nms / TorchNMS.nms).
//
// Semantics (SURVEY 8c rules, pinned by tests/golden/nms_golden.pt which holds outputs of the reference file itself):
//   candidate iff max class score > conf_thres (strict)            nms.py:76,:120-124
//   xywh -> xyxy: hw = w/2; x1 = cx - hw; x2 = cx + hw            nms.py:86
//   optional class filter                                          nms.py:127-131
//   n > max_nms: best max_nms by score                             nms.py:137-141
//   class offset cls*max_wh added in fp32 before IoU               nms.py:143,149
//   stable descending score order, ties -> lower anchor index      torchvision.ops.nms (nms.py:154)
//   j suppressed iff a kept earlier i has inter/(a_i+a_j-inter) > thr, every fp32 op rounded on its own
//   first max_det survivors                                        nms.py:157
//   no wall-clock limit                                            (nms.py:162-164 dropped: non-deterministic)
//
// Design (one CTA per image, two kernels so the scan runs at high occupancy):
//   K1 nms_sort_kernel : threshold filter -> 64-bit keys (orderable(score) desc | anchor idx asc) -> bitonic sort in
//                        shared memory (<= 16384 keys = 128 KB) or in the global workspace -> sorted anchor list.
//   K2 nms_scan_kernel : greedy scan against the KEPT list only (exactly equivalent to the all-pairs bitmask scan,
//                        but n*kept instead of n^2/2 IoUs and it stops at max_det).  Tiles of 256 candidates: each
//                        thread first tests its candidate against boxes kept from earlier tiles, then the tile is
//                        resolved in rounds (warp ballot picks the first survivor, everyone later tests against it).
// All IoU arithmetic uses __f*_rn intrinsics so nvcc can never contract it into FMAs (bit-exact vs the CPU oracle).
#include "kernels.h"

namespace ysp {

struct NmsP {
  const float* pred; int B, C, A, nc; float conf, iou; int max_det, max_nms; float max_wh; int agnostic;
  const int32_t* classes; int n_classes;
  const float* boxes; const float* scores;   // core mode (TorchNMS.nms): boxes [A,4] xyxy, scores [A]; pred == NULL
  int32_t* sorted; int32_t* ncand; unsigned long long* gkeys; int A_pad; int use_gkeys;
  float* kept; int kcap;
  float* out_boxes; int64_t* out_idx; int32_t* out_count; int out_row;
};

__device__ __forceinline__ uint32_t orderable_desc(float s) {
  if (s == 0.f) s = 0.f;                       // -0.0 == +0.0 must tie
  uint32_t u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // ascending order of floats
  return ~u;                                   // descending
}

__device__ __forceinline__ bool best_class(const NmsP& p, const float* pr, int a, float& best, int& bj) {
  best = pr[(size_t)4 * p.A + a]; bj = 0;
  for (int j = 1; j < p.nc; ++j) {
    float v = pr[(size_t)(4 + j) * p.A + a];
    if (v > best) { best = v; bj = j; }
  }
  if (!(best > p.conf)) return false;
  if (p.n_classes > 0) {
    bool ok = false;
    for (int i = 0; i < p.n_classes; ++i) ok |= (p.classes[i] == bj);
    if (!ok) return false;
  }
  return true;
}

__global__ void __launch_bounds__(1024) nms_sort_kernel(NmsP p) {
  extern __shared__ unsigned long long skeys[];
  __shared__ int s_n;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  unsigned long long* keys = p.use_gkeys ? p.gkeys + (size_t)b * p.A_pad : skeys;
  if (tid == 0) s_n = 0;
  __syncthreads();
  if (p.pred) {
    const float* pr = p.pred + (size_t)b * p.C * p.A;
    for (int a = tid; a < p.A; a += nt) {
      float best; int bj;
      if (best_class(p, pr, a, best, bj)) {
        int pos = atomicAdd(&s_n, 1);
        keys[pos] = ((unsigned long long)orderable_desc(best) << 32) | (uint32_t)a;
      }
    }
  } else {
    for (int a = tid; a < p.A; a += nt) {
      int pos = atomicAdd(&s_n, 1);
      keys[pos] = ((unsigned long long)orderable_desc(p.scores[a]) << 32) | (uint32_t)a;
    }
  }
  __syncthreads();
  const int n = s_n;
  int n_pad = 1;
  while (n_pad < n) n_pad <<= 1;
  for (int i = n + tid; i < n_pad; i += nt) keys[i] = ~0ull;
  __syncthreads();
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < n_pad; t += nt) {
        int x = t ^ j;
        if (x > t) {
          unsigned long long ka = keys[t], kb = keys[x];
          bool asc = (t & k) == 0;
          if ((ka > kb) == asc) { keys[t] = kb; keys[x] = ka; }
        }
      }
      __syncthreads();
    }
  }
  const int n_out = min(n, p.max_nms);
  for (int i = tid; i < n_out; i += nt) p.sorted[(size_t)b * p.A + i] = (int32_t)(keys[i] & 0xffffffffu);
  if (tid == 0) p.ncand[b] = n_out;
}

__device__ __forceinline__ bool iou_gt(float ax1, float ay1, float ax2, float ay2, float aa, float bx1, float by1,
                                       float bx2, float by2, float ba, float thr) {
  float xx1 = fmaxf(ax1, bx1), yy1 = fmaxf(ay1, by1), xx2 = fminf(ax2, bx2), yy2 = fminf(ay2, by2);
  float w = __fsub_rn(xx2, xx1), h = __fsub_rn(yy2, yy1);
  if (!(w > 0.f)) w = 0.f;
  if (!(h > 0.f)) h = 0.f;
  float inter = __fmul_rn(w, h);
  float uni = __fsub_rn(__fadd_rn(aa, ba), inter);
  return __fdiv_rn(inter, uni) > thr;
}

__global__ void __launch_bounds__(256) nms_scan_kernel(NmsP p) {
  __shared__ float s_tile[5][256];
  __shared__ int s_first[2][8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.ncand[b];
  const float* pr = p.pred ? p.pred + (size_t)b * p.C * p.A : nullptr;
  float* keptb = p.kept + (size_t)b * p.kcap * 5;
  int kept = 0, round = 0;
  for (int t0 = 0; t0 < n && kept < p.max_det; t0 += 256) {
    const int i = t0 + tid;
    bool alive = i < n;
    int idx = 0, cls = 0;
    float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f, conf = 0.f;
    float bx1 = 0.f, by1 = 0.f, bx2 = 0.f, by2 = 0.f, ar = 0.f;
    if (alive) {
      idx = p.sorted[(size_t)b * p.A + i];
      if (pr) {
        float cx = pr[idx], cy = pr[(size_t)p.A + idx], w = pr[(size_t)2 * p.A + idx], h = pr[(size_t)3 * p.A + idx];
        float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
        x1 = __fsub_rn(cx, hw); y1 = __fsub_rn(cy, hh); x2 = __fadd_rn(cx, hw); y2 = __fadd_rn(cy, hh);
        conf = pr[(size_t)4 * p.A + idx];
        for (int j = 1; j < p.nc; ++j) {
          float v = pr[(size_t)(4 + j) * p.A + idx];
          if (v > conf) { conf = v; cls = j; }
        }
        float off = p.agnostic ? 0.f : __fmul_rn((float)cls, p.max_wh);
        bx1 = __fadd_rn(x1, off); by1 = __fadd_rn(y1, off); bx2 = __fadd_rn(x2, off); by2 = __fadd_rn(y2, off);
      } else {
        const float* bb = p.boxes + (size_t)idx * 4;
        bx1 = x1 = bb[0]; by1 = y1 = bb[1]; bx2 = x2 = bb[2]; by2 = y2 = bb[3];
      }
      ar = __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(by2, by1));
    }
    s_tile[0][tid] = bx1; s_tile[1][tid] = by1; s_tile[2][tid] = bx2; s_tile[3][tid] = by2; s_tile[4][tid] = ar;
    // phase A: against boxes kept from earlier tiles (uniform addresses -> broadcast loads)
    if (alive) {
      for (int k = 0; k < kept; ++k) {
        const float* kb = keptb + (size_t)k * 5;
        if (iou_gt(kb[0], kb[1], kb[2], kb[3], kb[4], bx1, by1, bx2, by2, ar, p.iou)) { alive = false; break; }
      }
    }
    __syncthreads();
    // phase B: resolve the tile in rounds
    while (true) {
      unsigned bal = __ballot_sync(0xffffffffu, alive);
      if (lane == 0) s_first[round & 1][warp] = bal ? (warp * 32 + __ffs(bal) - 1) : (1 << 30);
      __syncthreads();
      int first = 1 << 30;
#pragma unroll
      for (int w = 0; w < 8; ++w) first = min(first, s_first[round & 1][w]);
      ++round;
      if (first >= (1 << 30)) break;
      if (tid == first) {
        alive = false;
        float* kb = keptb + (size_t)kept * 5;
        kb[0] = bx1; kb[1] = by1; kb[2] = bx2; kb[3] = by2; kb[4] = ar;
        p.out_idx[(size_t)b * p.max_det + kept] = idx;
        if (p.out_boxes) {
          float* o = p.out_boxes + ((size_t)b * p.max_det + kept) * p.out_row;
          o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2; o[4] = conf; o[5] = (float)cls;
          for (int e = 6; e < p.out_row; ++e) o[e] = pr[(size_t)(4 + p.nc + e - 6) * p.A + idx];
        }
      } else if (alive) {
        if (iou_gt(s_tile[0][first], s_tile[1][first], s_tile[2][first], s_tile[3][first], s_tile[4][first], bx1, by1,
                   bx2, by2, ar, p.iou))
          alive = false;
      }
      ++kept;
      if (kept >= p.max_det) break;
    }
    __syncthreads();
  }
  if (tid == 0) p.out_count[b] = kept;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int next_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }

size_t nms_workspace_bytes(int B, int C, int A, int max_det) {
  (void)C;
  int A_pad = next_pow2(A > 0 ? A : 1);
  int kcap = max_det < A ? max_det : A;
  if (kcap < 1) kcap = 1;
  size_t bytes = 0;
  bytes += align_up((size_t)B * (A > 0 ? A : 1) * 4, 256);       // sorted
  bytes += align_up((size_t)B * 4, 256);                         // ncand
  bytes += align_up((size_t)B * kcap * 5 * 4, 256);              // kept
  if (A_pad > 16384) bytes += align_up((size_t)B * A_pad * 8, 256);
  return bytes;
}

static int nms_common(NmsP p, void* ws, size_t ws_bytes, cudaStream_t s) {
  const int A1 = p.A > 0 ? p.A : 1;
  p.A_pad = next_pow2(A1);
  p.kcap = p.max_det < p.A ? p.max_det : p.A;
  if (p.kcap < 1) p.kcap = 1;
  if (ws_bytes < nms_workspace_bytes(p.B, p.C, p.A, p.max_det)) return -3;
  char* w = reinterpret_cast<char*>(ws);
  p.sorted = reinterpret_cast<int32_t*>(w); w += align_up((size_t)p.B * A1 * 4, 256);
  p.ncand = reinterpret_cast<int32_t*>(w); w += align_up((size_t)p.B * 4, 256);
  p.kept = reinterpret_cast<float*>(w); w += align_up((size_t)p.B * p.kcap * 5 * 4, 256);
  p.use_gkeys = p.A_pad > 16384;
  p.gkeys = p.use_gkeys ? reinterpret_cast<unsigned long long*>(w) : nullptr;
  if (p.B == 0) return 0;
  size_t smem = p.use_gkeys ? 0 : (size_t)p.A_pad * 8;
  static unsigned long long attr_done = 0;
  ensure_dyn_smem(nms_sort_kernel, 16384 * 8, attr_done, "nms_sort_kernel");
  int threads = p.A_pad >= 2048 ? 1024 : (p.A_pad >= 512 ? 256 : 64);
  nms_sort_kernel<<<p.B, threads, smem, s>>>(p);
  nms_scan_kernel<<<p.B, 256, 0, s>>>(p);
  return 0;
}

int launch_nms(const float* pred, int B, int C, int A, int nc, float conf, float iou, int max_det, int max_nms,
               float max_wh, int agnostic, const int32_t* classes, int n_classes, float* out_boxes, int64_t* out_idx,
               int32_t* out_count, void* ws, size_t ws_bytes, cudaStream_t s) {
  NmsP p = {};
  p.pred = pred; p.B = B; p.C = C; p.A = A; p.nc = nc; p.conf = conf; p.iou = iou; p.max_det = max_det;
  p.max_nms = max_nms; p.max_wh = max_wh; p.agnostic = agnostic; p.classes = classes; p.n_classes = n_classes;
  p.out_boxes = out_boxes; p.out_idx = out_idx; p.out_count = out_count; p.out_row = 6 + (C - 4 - nc);
  return nms_common(p, ws, ws_bytes, s);
}

int launch_nms_core(const float* boxes, const float* scores, int N, float iou, int64_t* keep, int32_t* count, void* ws,
                    size_t ws_bytes, cudaStream_t s) {
  NmsP p = {};
  p.pred = nullptr; p.boxes = boxes; p.scores = scores; p.B = 1; p.C = 5; p.A = N; p.nc = 1; p.conf = 0.f; p.iou = iou;
  p.max_det = N > 0 ? N : 1; p.max_nms = N; p.max_wh = 0.f; p.agnostic = 1;
  p.out_boxes = nullptr; p.out_idx = keep; p.out_count = count; p.out_row = 6;
  if (N == 0) { cudaMemsetAsync(count, 0, 4, s); return 0; }
  return nms_common(p, ws, ws_bytes, s);
}

// nms.py:84-86 side effect: the reference overwrites prediction[:, :4, :] with xyxy in place.
__global__ void __launch_bounds__(256) xywh2xyxy_inplace_kernel(float* pred, int C, int A, long long total) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  long long b = e / A;
  int a = (int)(e % A);
  float* pr = pred + (size_t)b * C * A;
  float cx = pr[a], cy = pr[(size_t)A + a], w = pr[(size_t)2 * A + a], h = pr[(size_t)3 * A + a];
  float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
  pr[a] = __fsub_rn(cx, hw); pr[(size_t)A + a] = __fsub_rn(cy, hh);
  pr[(size_t)2 * A + a] = __fadd_rn(cx, hw); pr[(size_t)3 * A + a] = __fadd_rn(cy, hh);
}
void launch_xywh2xyxy_inplace(float* pred, int B, int C, int A, cudaStream_t s) {
  long long total = (long long)B * A;
  if (total == 0) return;
  xywh2xyxy_inplace_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(pred, C, A, total);
}

}  // namespace ysp
