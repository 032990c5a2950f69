/* ysp.h -- C ABI of the B200-native YOLO-Seg++ inference hot path (libysp.so).
 *
 * The reference (Jhewu/YOLO-U) is 100% Python: its "plugin interface" for this path is a set of Python call
 * signatures on torch.Tensors.  Each entry point below replaces one of them; the Python host side in
 * `yolo-u_b200/` keeps the reference signatures and hands raw device pointers to these functions through ctypes.
 *
 *   reference interface (file:line under the reference tree)                     entry point here
 *   ---------------------------------------------------------------------------  --------------------------------
 *   dataset.py:53-68 + evaluate_model.py:136 (u8 HWC -> float CHW /255)          ysp_normalize_u8
 *   evaluate_model.py:141  YOLO_predictor.model(img) -> [y, [P3,P4,P5]]          ysp_detector_forward
 *   evaluate_model.py:142-144  logits = sigmoid(P3[:, -1:])                      ysp_bottleneck (and inside ysp_pipeline)
 *   nms.py:13-166  non_max_suppression(...)                                      ysp_nms
 *   nms.py:239-296 TorchNMS.nms(boxes, scores, thr)                              ysp_nms_core
 *   YOLOSegPlusPlus.py:242-272  YOLOSegPlusPlus.forward(x, logits)               ysp_segpp_forward
 *   evaluate_model.py:157-158,166-174  sigmoid>0.5, TP/FP/FN, Dice counts        ysp_mask_dice
 *   evaluate_model.py:134-174  (the loop body = the de-facto predict())          ysp_pipeline
 *   evaluate_model.py:234-243 / YOLOSegPlusPlus.py:150 (state_dict tensors)      ysp_load_weight / ysp_finalize
 *   train.py:302-331  zero_grad / forward (train mode) / DiceLoss / backward      ysp_encoder_forward + ysp_train_step
 *   train.py:262,329  optim.AdamW(...).step()  (+ :328 clip_grad_norm_)           ysp_adamw
 *   train.py:266,325-340  GradScaler: unscale_ / inf check of scaler.step          ysp_grad_sqnorm (+ grad_scale of ysp_train_step / ysp_adamw)
 *   train.py:346-366  validation loss                                             ysp_seg_loss (+ ysp_mask_dice)
 *   dataset.py:59-70  cv2.resize (INTER_LINEAR / INTER_NEAREST) + ToTensor         ysp_resize_u8
 *   dataset.py:86-97  objectmap z-score + sigmoid                                  ysp_objectmap_transform
 *   custom_detseg_predictor.py:177  ops.scale_boxes                               ysp_scale_boxes
 *   evaluate_model.py:149-155  (commented-out) detection-confidence gate          ysp_conf_gate
 *
 * Conventions: every pointer named d_* is a DEVICE pointer owned by the caller; h_* is a HOST pointer.  Nothing
 * is allocated per call: scratch comes from the caller-provided workspace (`d_ws`, at least
 * ysp_workspace_bytes()).  All calls are asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant per
 * handle+workspace, and return 0 on success or a negative YSP_E* code; ysp_last_error() gives the message.
 * There is NO CPU fallback: without a CUDA device every compute entry returns YSP_ECUDA.
 */
#ifndef YSP_H_
#define YSP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YSP_OK 0
#define YSP_EINVAL (-1)   /* bad argument / shape (Python side raises ValueError / AssertionError) */
#define YSP_ECUDA (-2)    /* CUDA runtime error (message has the cudaError string) */
#define YSP_ESTATE (-3)   /* call order: weights missing, not finalized, workspace too small */
#define YSP_ENOWEIGHT (-4)/* a state_dict tensor the topology needs was never loaded (message names it) */

/* arithmetic / storage mode of the conv path */
#define YSP_MODE_FP32 0   /* "CUDA-core parity mode": fp32 activations, fp32 FFMA, bit-for-bit deterministic   */
#define YSP_MODE_BF16 1   /* "throughput mode": bf16 activations, tcgen05 bf16 MMA with fp32 TMEM accumulators */
#define YSP_MODE_TC32 2   /* "tensor-core parity mode": fp32 activations; every MMA operand split into fp16 hi + lo and
                             each product issued as 3 tcgen05 MMAs (hi.hi + lo.hi + hi.lo) with fp32 TMEM accumulators:
                             fp32-accurate results (1e-3 logit parity) on the tensor cores                             */

typedef struct ysp_handle ysp_handle;

int ysp_version(void);
const char* ysp_last_error(void);

/* -- lifetime / weights ------------------------------------------------------------------------------------------ */
int ysp_create(ysp_handle** out, int device, int mode);
void ysp_destroy(ysp_handle* h);
/* One state_dict tensor (HOST fp32, contiguous, PyTorch layout).  `name` = "det." + DetectionModel key
 * (e.g. det.model.0.conv.weight) or "seg." + YOLOSegPlusPlus key (seg.decoder.0.0.cv1.bn.running_var). */
int ysp_load_weight(ysp_handle* h, const char* name, const float* h_data, int ndim, const int64_t* shape);
/* Fold BatchNorm (eval; detector eps 1e-3 if it arrives unfused, decoder eps 1e-5), repack to the kernel layouts,
 * upload.  `which`: bit0 = detector, bit1 = seg head. */
int ysp_finalize(ysp_handle* h, int which);
/* scratch bytes needed for a batch of B slices of HxW through ysp_pipeline (>= every other entry's need) */
size_t ysp_workspace_bytes(ysp_handle* h, int B, int H, int W);
/* same for a caller-chosen max_det (the NMS scratch grows with it); ysp_workspace_bytes() == max_det 300 */
size_t ysp_pipeline_workspace_bytes(ysp_handle* h, int B, int H, int W, int max_det);

/* -- a1: input normalisation -------------------------------------------------------------------------------------- */
/* u8 [B,H,W,4] (channel order as stored) -> fp32 NCHW [B,4,H,W] = x/255 (ToTensor).  HBM-bound, 4 B in / 16 B out per pixel. */
int ysp_normalize_u8(const uint8_t* d_u8, float* d_out_nchw, int B, int H, int W, void* stream);

/* -- a2: detector ------------------------------------------------------------------------------------------------- */
/* img fp32 NCHW [B,4,H,W]; H,W are zero-padded bottom/right to S=ceil32 (decision D1).  Outputs (fp32):
 * y [B,5,A] (xywh px + sigmoid cls), raw maps NCHW [B,65,S/8,S/8], [B,65,S/16,..], [B,65,S/32,..] (any may be NULL). */
int ysp_detector_forward(ysp_handle* h, const float* d_img, int B, int H, int W, float* d_y, float* d_p3, float* d_p4,
                         float* d_p5, void* d_ws, size_t ws_bytes, void* stream);

/* -- a3: bottleneck ---------------------------------------------------------------------------------------------- */
/* logits[B,1,h,w] = sigmoid(P3_raw[:, C-1, :h, :w]) from NCHW raw map [B,C,Hs,Ws] */
int ysp_bottleneck(const float* d_p3, int B, int C, int Hs, int Ws, float* d_logits, int h, int w, void* stream);

/* -- a5..a8: seg head -------------------------------------------------------------------------------------------- */
/* x fp32 NCHW [B,4,H,W] (H,W % 8 == 0), logits fp32 [B,1,H/8,W/8] -> out fp32 [B,1,H,W] mask logits (no sigmoid) */
int ysp_segpp_forward(ysp_handle* h, const float* d_x, const float* d_logits, float* d_out, int B, int H, int W,
                      void* d_ws, size_t ws_bytes, void* stream);

/* -- a4: box suppression ------------------------------------------------------------------------------------------ */
/* pred fp32 [B,C,A] (rows 0..3 xywh, 4..4+nc-1 class scores, rest extra; NOT modified).  Outputs padded:
 * out_boxes [B,max_det,6+extra] (xyxy, conf, cls, extra), out_idx [B,max_det] int64 anchor indices, out_count [B] int32.
 * Stable descending score order (ties -> lower anchor index), suppress iff IoU > iou_thres with separately
 * rounded fp32 ops, conf strictly > conf_thres, class offset cls*max_wh unless agnostic, n > max_nms truncation,
 * max_det truncation; no wall-clock limit.  classes: optional DEVICE int32[n_classes] filter (nms.py:127-131). */
size_t ysp_nms_workspace_bytes(int B, int C, int A, int max_det);
int ysp_nms(const float* d_pred, int B, int C, int A, int nc, float conf_thres, float iou_thres, int max_det,
            int max_nms, float max_wh, int agnostic, const int32_t* d_classes, int n_classes, float* d_out_boxes,
            int64_t* d_out_idx, int32_t* d_out_count, void* d_ws, size_t ws_bytes, void* stream);
/* boxes [N,4] xyxy, scores [N] -> keep [N] int64 (score order), count [1] int32 */
int ysp_nms_core(const float* d_boxes, const float* d_scores, int N, float iou_thres, int64_t* d_keep,
                 int32_t* d_count, void* d_ws, size_t ws_bytes, void* stream);

/* nms.py:84-86 side effect (the reference overwrites prediction[:, :4, :] with xyxy in place) */
int ysp_xywh2xyxy_inplace(float* d_pred, int B, int C, int A, void* stream);

/* -- a9: mask + Dice counters ------------------------------------------------------------------------------------ */
/* counts[b] = (|P&T|, |P|, |T|), P = sigmoid(logit) > 0.5 (fp32), T = target > 0.5; optional u8 mask out [B,HW] */
int ysp_mask_dice(const float* d_logits, const float* d_target, int B, int HW, int32_t* d_counts, uint8_t* d_mask,
                  void* stream);

/* -- SURVEY 8(f) "next" rows ------------------------------------------------------------------------------------ */
/* dataset.py:86-97: per-map  sigmoid((x - mean) / std)  (torch.std, unbiased; std == 0 -> x - mean) on B maps of n
 * elements each (the raw P3 class-logit maps written by generate_objectmaps.py:91-106). */
int ysp_objectmap_transform(const float* d_maps, float* d_out, int B, int n, void* stream);
/* ultralytics ops.scale_boxes (custom_detseg_predictor.py:177): in place on n rows of `row` floats (xyxy first):
 * (b - pad) / gain, clipped to [0,w0] x [0,h0]. */
int ysp_scale_boxes(float* d_boxes, long long n, int row, float gain, float pad_x, float pad_y, float w0, float h0,
                    void* stream);

/* -- fused pipeline (evaluate_model.py:134-174 with decision D1) -------------------------------------------------- */
typedef struct ysp_pipeline_io {
  const float* d_img;        /* fp32 NCHW [B,4,H,W]   (or NULL if d_img_u8 given) */
  const uint8_t* d_img_u8;   /* u8 [B,H,W,4]          (normalised on the fly, a1) */
  const float* d_target;     /* fp32 [B,1,H,W] or NULL (then counts' |P&T|,|T| are 0) */
  float* d_mask_logits;      /* fp32 [B,1,H,W] */
  float* d_y;                /* fp32 [B,5,A] or NULL */
  float* d_bottleneck;       /* fp32 [B,1,H/8,W/8] or NULL */
  float* d_det_boxes;        /* fp32 [B,max_det,6] */
  int64_t* d_det_idx;        /* int64 [B,max_det] */
  int32_t* d_det_count;      /* int32 [B] */
  int32_t* d_counts;         /* int32 [B,3] */
  uint8_t* d_mask;           /* u8 [B,H,W] or NULL */
  float conf_thres, iou_thres;
  int max_det;
  const uint8_t* d_target_u8;/* u8 [B,H,W] ground-truth mask as stored in the PNG (dataset.py:55, before ToTensor): T = v/255 > 0.5
                                <=> v >= 128; used when d_target is NULL (a quarter of the upload).  NULL = not given */
  uint32_t* d_mask_bits;     /* bit-packed predicted mask, uint32 [B, H*W/32] (pixel i -> bit i%32 of word i/32), or NULL;
                                needs H*W % 128 == 0.  1/32 of the logits: what a predict() caller copies back */
} ysp_pipeline_io;
int ysp_pipeline(ysp_handle* h, const ysp_pipeline_io* io, int B, int H, int W, void* d_ws, size_t ws_bytes,
                 void* stream);

/* The detection-confidence gate sketched (commented out) at evaluate_model.py:149-155, applied to the outputs of
 * ysp_pipeline / ysp_nms + ysp_mask_dice: a slice with no detection, or whose best detection has conf <= conf_gate, gets an
 * all-zero predicted mask: counts[b] = (0, 0, |T|), d_mask[b] zeroed (optional), d_gated[b] = 1 (optional). */
int ysp_conf_gate(const float* d_det_boxes, const int32_t* d_det_count, int B, int max_det, int row, float conf_gate,
                  int32_t* d_counts, uint8_t* d_mask, int HW, uint8_t* d_gated, void* stream);

/* -- slice ingest (dataset.py:59-70) ---------------------------------------------------------------------
 * cv2.resize on decoded uint8 slices + transforms.ToTensor, on device and bit-exact with OpenCV's generic uint8 path:
 * interp 1 = INTER_LINEAR (image, dataset.py:63), 0 = INTER_NEAREST (mask, dataset.py:65).  src u8 [B,h,w,C], C = 4
 * (cv2 BGRA order, IMREAD_UNCHANGED) or 1; outputs (either may be NULL): dst_u8 [B,dh,dw,C] (feeds ysp_pipeline's
 * uint8 input) and dst_f32 [B,C,dh,dw] = value / 255 (what ToTensor returns, dataset.py:68-70). */
int ysp_resize_u8(const uint8_t* d_src, int B, int h, int w, int C, int dh, int dw, int interp, uint8_t* d_dst_u8,
                  float* d_dst_f32, void* stream);

/* -- seg-head training step (reference train.py:241-360, non-AMP branch) -------------------------
 * Trainable = everything of YOLOSegPlusPlus outside `encoder.*` and the unused `param` scalar (train.py:256-267).
 * The trainer owns no tensors: the caller (yolo_u_b200/trainer.py) allocates flat fp32 DEVICE buffers
 *   params / grads / adam_m / adam_v  [ysp_train_param_count]   and   stats (BN running mean/var) [ysp_train_stat_count]
 * whose sub-tensors are named by the reference's state_dict keys in PyTorch's own layouts (ysp_train_tensor_info), so
 * a `best.pth` (train.py:428) round-trips by plain slicing, and `grads` can be all-reduced with ONE collective. */
typedef struct ysp_trainer ysp_trainer;
int ysp_train_create(ysp_trainer** out, int device, int B, int H, int W);
void ysp_train_destroy(ysp_trainer* t);
int ysp_train_num_tensors(const ysp_trainer* t);
/* kind 0 = parameter (offset into params/grads/adam buffers), 1 = BN running statistic (offset into stats) */
int ysp_train_tensor_info(const ysp_trainer* t, int i, char* name, int name_cap, int* kind, int64_t* offset,
                          int64_t* numel);
int64_t ysp_train_param_count(const ysp_trainer* t);
int64_t ysp_train_stat_count(const ysp_trainer* t);
size_t ysp_train_workspace_bytes(const ysp_trainer* t);
int ysp_train_last_launch_count(const ysp_trainer* t);
/* ALGORITHMIC HBM bytes (compulsory fp32 reads + writes of every kernel, unfused) of the last ysp_train_step */
double ysp_train_last_step_bytes(const ysp_trainer* t);
/* frozen encoder (YOLOSegPlusPlus.py:255-259) through the inference engine: x [B,4,H,W] fp32 ->
 * skipA [B,H/4,W/4,64], skipB [B,H/8,W/8,128] dense NHWC fp32.  Workspace: ysp_workspace_bytes. */
int ysp_encoder_forward(ysp_handle* h, const float* d_x, float* d_skipA, float* d_skipB, int B, int H, int W, void* d_ws,
                        size_t ws_bytes, void* stream);
/* optimizer.zero_grad(); pred = decoder(skips, logits) in train() mode (BN batch statistics; running statistics in
 * `d_stats` updated with `momentum` unless d_stats is NULL); loss = monai DiceLoss(sigmoid, soft_label, batch=True)
 * (loss_kind 0, train.py:98-104) or Dice + mean BCE-with-logits (loss_kind 1); loss.backward() -> d_grads, every
 * gradient multiplied by grad_scale (AMP loss scale / 1 / 1/world_size).  d_loss3 = {total, dice, bce}.
 * d_mask_logits (optional) receives pred [B,1,H,W].  d_logits [B,1,H/8,W/8], d_target [B,1,H,W]. */
int ysp_train_step(ysp_trainer* t, const float* d_skipA, const float* d_skipB, const float* d_logits,
                   const float* d_target, const float* d_params, float* d_grads, float* d_stats, float momentum,
                   int loss_kind, float grad_scale, float* d_loss3, float* d_mask_logits, void* d_ws, size_t ws_bytes,
                   void* stream);
/* The loss alone, over n = B*H*W mask logits (validation half of the epoch, train.py:356-357): d_loss3 = {total, dice, bce}
 * with the same reductions as ysp_train_step.  d_ws32 = 32 bytes of device scratch. */
int ysp_seg_loss(const float* d_logits, const float* d_target, int64_t n, int loss_kind, float* d_loss3, void* d_ws32,
                 void* stream);
/* torch.optim.AdamW step over a flat buffer (train.py:262, :325): grads are multiplied by grad_scale first; max_norm > 0
 * applies clip_grad_norm_ semantics (train.py:324 -- a no-op in the reference, its parameter generator is already
 * exhausted; so callers pass 0).  d_ws8 = 8 bytes of device scratch (only read when clipping). */
int ysp_adamw(float* d_params, const float* d_grads, float* d_m, float* d_v, int64_t n, float lr, float beta1, float beta2,
              float eps, float weight_decay, int step, float grad_scale, float max_norm, void* d_ws8, void* stream);
/* Sum of squares (double) of a flat gradient buffer -> d_out8[0].  The mixed-precision branch (train.py:325-340) needs it
 * for GradScaler's inf/NaN check before scaler.step(): a non-finite entry makes the sum non-finite. */
int ysp_grad_sqnorm(const float* d_grads, int64_t n, void* d_out8, void* stream);

/* -- introspection (tests / profiling) ---------------------------------------------------------------------------- */
/* keep every intermediate alive (no workspace reuse) so ysp_debug_tensor can read them; affects plans built later */
int ysp_set_keep_intermediates(ysp_handle* h, int on);
/* number of kernels the last ysp_* call on this handle launched */
int ysp_last_launch_count(ysp_handle* h);
/* per-step device timing: enable=1 start/continue, 2 = clear and start, 0 = stop.  While on, every plan step is
 * bracketed by CUDA events on the launching stream and the call synchronises at its end (NOT for timed regions).
 * ysp_profile_report writes a JSON array [{name,kind,ms,calls,launches,bytes,flops}] (bytes/flops = ALGORITHMIC
 * per-step totals accumulated over calls) and returns its length, or <0. */
int ysp_profile(ysp_handle* h, int enable);
int ysp_profile_report(ysp_handle* h, char* buf, size_t cap);
/* copy a named intermediate of the last plan run ("det:model.6", "seg:decoder.0", ...) to fp32 NCHW host-visible
 * device buffer; returns dims via shape[4] = N,C,H,W.  d_out may be NULL to query the shape only. */
int ysp_debug_tensor(ysp_handle* h, const char* name, void* d_ws, float* d_out, int64_t* shape, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YSP_H_ */
