/* ORACLE (test infrastructure, NOT product code) -- plain-C restatement of the reference's box suppression.
 *
 * Follows /root/reference/nms.py:13-166 (non-rotated, single-label path) with the back-end of
 * nms.py:154 (`torchvision.ops.nms`, CPU kernel semantics restated from SURVEY App. A.4):
 *   - candidate iff max class score > conf_thres                      (nms.py:76, :120-121, strict >)
 *   - xywh -> xyxy: wh = x/2 ; xy - wh ; xy + wh                      (nms.py:86, App. A.4)
 *   - n > max_nms: keep the max_nms best by score                     (nms.py:137-141)
 *   - boxes offset by cls * max_wh unless agnostic                    (nms.py:143,149)
 *   - stable descending sort by score (ties: lower index first); j suppressed iff a kept earlier i has
 *     inter / (area_i + area_j - inter) > thr, every fp32 op rounded on its own (build with -ffp-contract=off)
 *   - first max_det survivors                                          (nms.py:157)
 *   - no wall-clock limit (nms.py:162-164 dropped: non-deterministic)
 * PINNED against outputs of the reference file itself: tests/golden/nms_*.pt (tests/test_oracle_pins.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -pthread -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float score; int32_t idx; } cand_t;

static int cand_cmp(const void* a, const void* b) {
    const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);      /* stable: lower original index first */
}

/* One image. pred: [C][A] (row = channel).  Returns number kept (<= max_det). */
static int nms_one(const float* pred, int C, int A, int nc, float conf_thres, float iou_thres, int max_det,
                   int max_nms, float max_wh, int agnostic, float* out_boxes, int64_t* out_idx) {
    cand_t* cand = (cand_t*)malloc(sizeof(cand_t) * (size_t)(A > 0 ? A : 1));
    int32_t* cls = (int32_t*)malloc(sizeof(int32_t) * (size_t)(A > 0 ? A : 1));
    int n = 0;
    for (int a = 0; a < A; ++a) {
        float best = pred[(size_t)4 * A + a]; int bj = 0;
        for (int j = 1; j < nc; ++j) { float v = pred[(size_t)(4 + j) * A + a]; if (v > best) { best = v; bj = j; } }
        if (best > conf_thres) { cand[n].score = best; cand[n].idx = a; cls[a] = bj; ++n; }
    }
    qsort(cand, (size_t)n, sizeof(cand_t), cand_cmp);
    if (n > max_nms) n = max_nms;
    float* bx = (float*)malloc(sizeof(float) * 5 * (size_t)(n > 0 ? n : 1));   /* offset boxes + area, sorted order */
    for (int i = 0; i < n; ++i) {
        int a = cand[i].idx;
        float cx = pred[a], cy = pred[(size_t)A + a], w = pred[(size_t)2 * A + a], h = pred[(size_t)3 * A + a];
        float hw = w / 2.0f, hh = h / 2.0f;
        float off = agnostic ? 0.0f : (float)cls[a] * max_wh;
        float x1 = (cx - hw) + off, y1 = (cy - hh) + off, x2 = (cx + hw) + off, y2 = (cy + hh) + off;
        bx[5 * i + 0] = x1; bx[5 * i + 1] = y1; bx[5 * i + 2] = x2; bx[5 * i + 3] = y2;
        bx[5 * i + 4] = (x2 - x1) * (y2 - y1);
    }
    unsigned char* sup = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
    int kept = 0;
    for (int i = 0; i < n && kept < max_det; ++i) {
        if (sup[i]) continue;
        int a = cand[i].idx;
        float cx = pred[a], cy = pred[(size_t)A + a], w = pred[(size_t)2 * A + a], h = pred[(size_t)3 * A + a];
        float hw = w / 2.0f, hh = h / 2.0f;
        float* o = out_boxes + (size_t)6 * kept;
        o[0] = cx - hw; o[1] = cy - hh; o[2] = cx + hw; o[3] = cy + hh; o[4] = cand[i].score; o[5] = (float)cls[a];
        out_idx[kept] = a;
        ++kept;
        const float ix1 = bx[5 * i], iy1 = bx[5 * i + 1], ix2 = bx[5 * i + 2], iy2 = bx[5 * i + 3], ia = bx[5 * i + 4];
        for (int j = i + 1; j < n; ++j) {
            if (sup[j]) continue;
            float xx1 = ix1 > bx[5 * j] ? ix1 : bx[5 * j];
            float yy1 = iy1 > bx[5 * j + 1] ? iy1 : bx[5 * j + 1];
            float xx2 = ix2 < bx[5 * j + 2] ? ix2 : bx[5 * j + 2];
            float yy2 = iy2 < bx[5 * j + 3] ? iy2 : bx[5 * j + 3];
            float ww = xx2 - xx1; if (!(ww > 0.0f)) ww = 0.0f;
            float hh2 = yy2 - yy1; if (!(hh2 > 0.0f)) hh2 = 0.0f;
            float inter = ww * hh2;
            float uni = ia + bx[5 * j + 4];
            uni = uni - inter;
            float ovr = inter / uni;
            if (ovr > iou_thres) sup[j] = 1;
        }
    }
    free(sup); free(bx); free(cls); free(cand);
    return kept;
}

/* Batched entry.  pred [B][C][A] fp32; out_boxes [B][max_det][6]; out_idx [B][max_det]; out_count [B].
 * Images are independent (the reference loops over them, nms.py:91); nthreads pthreads pull image indices. */
#include <pthread.h>
typedef struct {
    const float* pred; int B, C, A, nc; float conf, iou; int max_det, max_nms; float max_wh; int agnostic;
    float* out_boxes; int64_t* out_idx; int32_t* out_count; int next; pthread_mutex_t mu;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu); int b = j->next++; pthread_mutex_unlock(&j->mu);
        if (b >= j->B) break;
        j->out_count[b] = nms_one(j->pred + (size_t)b * j->C * j->A, j->C, j->A, j->nc, j->conf, j->iou, j->max_det,
                                  j->max_nms, j->max_wh, j->agnostic, j->out_boxes + (size_t)b * j->max_det * 6,
                                  j->out_idx + (size_t)b * j->max_det);
    }
    return 0;
}

int ysp_oracle_nms(const float* pred, int B, int C, int A, int nc, float conf_thres, float iou_thres, int max_det,
                   int max_nms, float max_wh, int agnostic, float* out_boxes, int64_t* out_idx,
                   int32_t* out_count, int nthreads) {
    if (nc <= 0) nc = C - 4;
    if (C < 4 + nc || max_det < 0) return -1;
    job_t j = {pred, B, C, A, nc, conf_thres, iou_thres, max_det, max_nms, max_wh, agnostic,
               out_boxes, out_idx, out_count, 0, PTHREAD_MUTEX_INITIALIZER};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads > B) nthreads = B > 0 ? B : 1;
    pthread_t th[256];
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], 0, worker, &j);
    worker(&j);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], 0);
    return 0;
}

/* Mask / Dice counters (evaluate_model.py:157-158,166-174): P = sigmoid(x) > 0.5 in fp32; counts per slice. */
#include <math.h>
int ysp_oracle_mask_counts(const float* logits, const float* target, int B, int HW, int32_t* counts /*[B][3]*/) {
    for (int b = 0; b < B; ++b) {
        int32_t inter = 0, p = 0, t = 0;
        for (int i = 0; i < HW; ++i) {
            float x = logits[(size_t)b * HW + i];
            float s = 1.0f / (1.0f + expf(-x));
            int pi = s > 0.5f, ti = target[(size_t)b * HW + i] > 0.5f;
            inter += pi & ti; p += pi; t += ti;
        }
        counts[3 * b] = inter; counts[3 * b + 1] = p; counts[3 * b + 2] = t;
    }
    return 0;
}
