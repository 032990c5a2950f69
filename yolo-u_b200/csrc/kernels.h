// kernels.h -- host-callable launchers (one per kernel family).  dtype: 0 = fp32 activations, 1 = bf16 activations.
#pragma once
#include "common.cuh"

namespace ysp {

enum { DT_F32 = 0, DT_BF16 = 1 };

// kernels_simt.cu
void launch_conv_dense(const ConvP& p, int in_dt, int out_dt, cudaStream_t s);
void launch_conv_dw(const DwP& p, int dt, cudaStream_t s);
void launch_add(const EwP& p, int dt, cudaStream_t s);                       // out = a + b
void launch_up_nearest2(const EwP& p, int dt, cudaStream_t s);               // out[oy,ox] = a[oy/2,ox/2]
void launch_up_bilinear2(const EwP& p, int dt, cudaStream_t s);              // torch bilinear x2, align_corners=False
void launch_eca(void* x, int N, int HW, int C, int cs, const float* w3, float* mean_ws, int dt, cudaStream_t s);
// ECA gate only (fp32): gate[n][c] = sigmoid(conv1d_k3(mean_hw(x)));  the consumer applies it (ConvP::in_scale)
bool eca_gate_f32_supported(int C, int cs);
void launch_eca_gate_f32(const void* x, int N, int HW, int C, int cs, const float* w3, float* gate, cudaStream_t s);
void launch_attention(const void* qkv, void* out, int B, int Ntok, int C, int heads, int area, int qkv_cs,
                      int out_cs, int dt, cudaStream_t s);
void launch_nchw_to_nhwc(const float* in, void* out, int N, int C, int H, int W, int out_cs, int dt, cudaStream_t s);
void launch_u8_to_nhwc(const uint8_t* in, void* out, int N, int H, int W, int out_cs, int dt, cudaStream_t s);
void launch_normalize_u8(const uint8_t* in, float* out_nchw, int N, int H, int W, cudaStream_t s);
// writes logits (fp32 [B,1,h,w]) into channel slice of an NHWC view and zero-fills `zero_pad` following channels
void launch_logits_to_nhwc(const float* logits, void* out, int N, int h, int w, int out_cs, int zero_pad, int dt,
                           cudaStream_t s);
// raw head maps NHWC fp32 [B,hw,cs] x3 -> y [B,4+nc,A] + NCHW raw copies (nullable) + optional bottleneck logits
struct DecodeP {
  const float* raw[3]; int h[3], w[3]; int cs; int nc; int B; int A;
  float stride[3];
  float* y; float* p[3];
  float* bott; int bh, bw;     // sigmoid(raw0[..., 64+nc-1]) cropped to bh x bw (fp32 [B,1,bh,bw]) or NULL
};
void launch_detect_decode(const DecodeP& p, cudaStream_t s);
void launch_bottleneck_nhwc(const float* raw, int cs, int ch, int B, int Hs, int Ws, float* logits, int h, int w, cudaStream_t s);
void launch_bottleneck(const float* p3, int B, int C, int Hs, int Ws, float* logits, int h, int w, cudaStream_t s);
void launch_mask_dice(const float* logits, const float* target, const uint8_t* target_u8, int B, int HW, int32_t* counts,
                      uint8_t* mask, uint32_t* bits, cudaStream_t s);
void launch_conf_gate(const float* det_boxes, const int32_t* det_count, int B, int max_det, int row, float thres, int32_t* counts,
                      uint8_t* mask, int HW, uint8_t* gated, cudaStream_t s);
void launch_nhwc_to_nchw_f32(const void* in, float* out, int N, int H, int W, int C, int in_cs, int dt, cudaStream_t s);

void launch_objectmap_transform(const float* in, float* out, int B, int n, cudaStream_t s);
void launch_resize_u8(const uint8_t* src, int B, int sh, int sw, int C, int dh, int dw, int interp, uint8_t* dst_u8,
                      float* dst_f32, cudaStream_t s);
void launch_scale_boxes(float* boxes, long long n, int row, float gain, float pad_x, float pad_y, float w0, float h0, cudaStream_t s);

// kernels_stem_attn.cu
void launch_stem_conv(const void* in, int in_u8, void* out, const float* w, const float* bias, int N, int H, int W, int OH,
                      int OW, int out_cs, int wld, int dt, cudaStream_t s);
void launch_attention64_bf16(const void* qkv, void* out, int B, int Ntok, int heads, int area, int qkv_cs, int out_cs,
                             cudaStream_t s);

// attention_tc.cu (tcgen05: both tensor-core modes, 64-token areas, head_dim 32, even head count)
bool attention_tc_supported(int Ntok, int heads, int area, int qkv_cs, int out_cs, int dt);
void launch_attention_tc(const void* qkv, void* out, int B, int Ntok, int heads, int area, int qkv_cs, int out_cs, int dt,
                         cudaStream_t s);

// nms.cu
size_t nms_workspace_bytes(int B, int C, int A, int max_det);
int launch_nms(const float* pred, int B, int C, int A, int nc, float conf, float iou, int max_det, int max_nms,
               float max_wh, int agnostic, const int32_t* classes, int n_classes, float* out_boxes, int64_t* out_idx,
               int32_t* out_count, void* ws, size_t ws_bytes, cudaStream_t s);
int launch_nms_core(const float* boxes, const float* scores, int N, float iou, int64_t* keep, int32_t* count, void* ws,
                    size_t ws_bytes, cudaStream_t s);
void launch_xywh2xyxy_inplace(float* pred, int B, int C, int A, cudaStream_t s);

// conv_tc.cu (tcgen05 + TMA implicit GEMM, bf16)
struct TcConvPlan;   // opaque: tensor maps + tiling, built once per (layer, shape)
TcConvPlan* tc_conv_plan_create(const ConvP& p, const void* w_bf16_kmajor, int out_dt);
void tc_conv_plan_destroy(TcConvPlan*);
void launch_conv_tc(const TcConvPlan* plan, const ConvP& p, cudaStream_t s);
bool tc_conv_supported(const ConvP& p);

// conv_tc32.cu (parity mode: fp32 activations, fp16 hi/lo operand split, 3 tcgen05 MMAs per product)
struct Tc32Tiling { int cin_pad, cout_pad, Kc, kchunks, n_tiles_n, N_tile; uint32_t b_bytes; size_t total_bytes; };
Tc32Tiling tc32_tiling(int Cin, int Cout, int taps);   // block shape of the packed weights (engine.cu packs, the kernel reads)
struct Tc32ConvPlan;
Tc32ConvPlan* tc32_conv_plan_create(const ConvP& p, const void* wpack, float w_unscale);
void tc32_conv_plan_destroy(Tc32ConvPlan*);
bool tc32_conv_plan_flat(const Tc32ConvPlan*);          // flat 1x1 path (the only one that applies ConvP::in_scale)
void launch_conv_tc32(const Tc32ConvPlan* plan, const ConvP& p, cudaStream_t s);
bool tc32_conv_supported(const ConvP& p);

// conv_halo.cu (tcgen05, halo tile + shifted descriptor views: small-channel 3x3 stride-1 convs on maps <= 64 wide)
bool conv_halo_supported(const ConvP& p, int in_dt, int out_dt);
void launch_conv_halo(const ConvP& p, const void* w_bf16_kmajor, int Ktc, cudaStream_t s);

// conv_halo32.cu (parity mode: halo tile split once + shifted descriptor views, Cin 16 / 32, 3x3 stride 1, maps <= 64 wide)
bool conv_halo32_supported(const ConvP& p);
void launch_conv_halo32(const ConvP& p, const void* wpack, float w_unscale, cudaStream_t s);

}  // namespace ysp

namespace ysp {
// kernels_fused.cu -- fused DoubleLightConv tail (everything after the two low-resolution 1x1 convs)
struct DlcP {
  const void* P;            // low-res NHWC [N, h, w, 2C] : [0,C) = conv.0.conv1 (BN folded, linear), [C,2C) = residual_conv
  void* out;                // hi-res NHWC [N, 2h, 2w, C] (activation dtype) or, with head, fp32 [N, 2h, 2w]
  const float *dw1, *b1;    // conv.0.conv2 depthwise 3x3 [9][C] + bias (SiLU)
  const float *w2, *b2;     // conv.1.conv1 1x1 [K=C][C] + bias (linear)
  const float *dw2, *b3;    // conv.1.conv2 depthwise 3x3 [9][C] + bias (SiLU)
  const float *wo, *bo;     // optional head: 1x1 C -> 1 (+bias); NULL otherwise
  int N, h, w, C, p_cs, out_cs, w2ld, wo_ld;
};
void launch_dlc_fused(const DlcP& p, int dt, cudaStream_t s);
}  // namespace ysp

namespace ysp {
// kernels_dlc32.cu -- fused DoubleLightConv tail of the PARITY mode (fp32 activations; pointwise GEMM on tcgen05 with fp16
// hi/lo operand splits, composite up2 o depthwise on CUDA cores): see the file header
struct Dlc32P {
  const void* P;            // low-res NHWC fp32 [N, h, w, 2C] : [0,C) = conv.0.conv1 (BN folded, linear), [C,2C) = residual_conv
  void* out;                // hi-res NHWC fp32 [N, 2h, 2w, C] or, with head, fp32 [N, 2h, 2w]
  const float *dw1, *b1;    // conv.0.conv2 depthwise 3x3 [9][C] + bias (SiLU)
  const void* wpack;        // conv.1.conv1 1x1 in the tc32 weight pack (engine.cu pack_conv_cat, tc32_tiling(C, C, 1))
  float w_unscale;          // the pack is scaled by a power of two; acc * w_unscale is exact
  const float* b2;          // conv.1.conv1 bias (linear)
  const float *dw2, *b3;    // conv.1.conv2 depthwise 3x3 [9][C] + bias (SiLU)
  const float *wo, *bo;     // optional head: 1x1 C -> 1 (+bias); NULL otherwise
  int N, h, w, C, p_cs, out_cs, wo_ld;
  unsigned magic_x = 0, magic_y = 0;   // filled by launch_dlc32
};
bool dlc32_supported(int C, bool head);
void launch_dlc32(const Dlc32P& p, cudaStream_t s);
}  // namespace ysp

namespace ysp {
// kernels_dlc_tc.cu -- decoder stage on tcgen05 (bf16 mode): see the file header
struct DlcTcP {
  const void* x;            // low-res NHWC bf16 [N, h, w, Cin] (pixel stride x_cs)
  void* out;                // HEAD: fp32 [N, 2h, 2w]; else NHWC bf16 [N, 2h, 2w, C] (pixel stride out_cs)
  const uint8_t* wpack;     // dlc_tc_prepare_kernel output
  const float* bo;          // head bias (device) or NULL
  int N, h, w, Cin, C, x_cs, out_cs;
  int probe = 0;            // YSP_DLC_PROBE timing experiments (0 in production)
  unsigned magic_x = 0, magic_y = 0;   // ceil(2^32 / tiles_{x,y}) (0: divisor 1), filled by launch_dlc_tc: exact tile decode by __umulhi
};
struct DlcTcPrep {          // fp32 folded weights in the engine's layouts ([K][ld] dense, [9][C] depthwise)
  const float *w1, *c1, *dw1, *b1, *w2, *c2, *dw2, *b3, *wr, *cr, *wo;
  int w1ld, w2ld, wrld, wold, Cin, C;
};
size_t dlc_tc_pack_bytes(int Cin, int C);
void launch_dlc_tc_prepare(const DlcTcPrep& q, void* out, cudaStream_t s);
bool dlc_tc_supported(int Cin, int C, bool head);
void launch_dlc_tc(const DlcTcP& p, cudaStream_t s);
}  // namespace ysp

namespace ysp {
// kernels_ghost.cu -- GhostBottleneck(c, c, s=1) of the decoder's C3Ghost blocks as one kernel (bf16 mode, c = 32 or 48)
struct GhostP {
  const bf16* a; bf16* out;                 // NHWC views [N,H,W,c] (pixel strides a_cs / out_cs)
  const float *w1, *b1, *dw1, *bd1;         // conv.0.cv1 dense [c][ld] (c -> c/4, SiLU), conv.0.cv2 depthwise [25][c/4] (SiLU)
  const float *w3, *b3, *dw2, *bd2;         // conv.2.cv1 dense [c/2][ld] (c/2 -> c/2), conv.2.cv2 depthwise [25][c/2] (linear)
  int w1ld, w3ld, N, H, W, a_cs, out_cs;
};
bool ghost_fused_supported(int c, int a_cs, int out_cs);
void launch_ghost_fused(const GhostP& p, int c, cudaStream_t s);
}  // namespace ysp
