// kernels_train.h -- launchers of the seg-head training kernels (kernels_train.cu).  fp32, NHWC views (ptr + row stride).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ysp {

struct BnRef { const float *gamma, *beta, *mean, *invstd; };
// Optional transform applied to a kernel's INPUT as it is loaded: x = act(gamma*(z - mean)*invstd + beta).  It lets a
// consumer read its producer's raw conv output z, so the producer's normalised tensor is never written (gamma == NULL: off).
struct InTf { const float *gamma = nullptr, *beta = nullptr, *mean = nullptr, *invstd = nullptr; int act = 0; };

// Optional second operand set of launch_pw_gemm:
//   A2/W2/I2 : a second reduction range, C += A2 * W2op  (the two input-gradient GEMMs that feed one tensor as ONE kernel);
//   Jsplit   : output columns >= Jsplit use weights Wb (row j - Jsplit), bias biasb and go to Cb (two 1x1 convs that read
//              the same input as ONE kernel); BN statistics are then taken over the first Jsplit columns only.
struct PwDual {
  const float* A2 = nullptr; int lda2 = 0; const float* W2 = nullptr; int ldw2 = 0; int I2 = 0;
  int Jsplit = 0; const float* Wb = nullptr; int ldwb = 0; const float* biasb = nullptr; float* Cb = nullptr; int ldcb = 0;
};
// `sums` (optional, [2][J] doubles, pre-zeroed): per-column sum and sum of squares of the product (BN statistics)
void launch_pw_gemm(const float* A, int lda, const float* W, int ldw, int trans, const float* bias, float* C, int ldc,
                    long long M, int I, int J, int beta, cudaStream_t s, double* sums = nullptr, InTf tf = InTf(),
                    PwDual du = PwDual());
// tcgen05 version (kernels_train_tc.cu): returns false when the shape is outside its envelope (the caller runs the FFMA kernel)
bool launch_pw_gemm_tc(const float* A, int lda, const float* W, int ldw, int trans, const float* bias, float* C, int ldc,
                       long long M, int I, int J, int beta, cudaStream_t s, double* sums, InTf tf, PwDual du);
bool launch_pw_wgrad_tc(const float* D, int ldd, const float* X, int ldx, float* dW, int ldw, long long M, int I, int J,
                        cudaStream_t s, InTf tf);
// the C -> 1 mask head as streaming kernels (false: shape outside their envelope, the caller uses launch_pw_gemm)
bool launch_lin1_fwd(const float* X, int ldx, const float* w, const float* b, float* Y, int C, long long M, cudaStream_t s);
bool launch_lin1_dgrad(const float* DY, const float* w, float* DX, int ldx, int C, long long M, cudaStream_t s);
void launch_pw_wgrad(const float* D, int ldd, const float* X, int ldx, float* dW, int ldw, long long M, int I, int J,
                     cudaStream_t s, InTf tf = InTf());
void launch_dw_conv(const float* X, int ldx, const float* W, float* Y, int ldy, int N, int H, int Wd, int C, int k,
                    int flip, int beta, cudaStream_t s);
bool launch_dw_fwd_stats(const float* X, int ldx, const float* W, float* Y, int ldy, double* sums, int N, int H, int Wd,
                         int C, int k, cudaStream_t s, InTf tf = InTf());
void launch_dw_wgrad(const float* D, int ldd, const float* X, int ldx, float* dW, int N, int H, int Wd, int C, int k,
                     cudaStream_t s, InTf tf = InTf());
// BatchNorm backward + depthwise input gradient + weight gradient (+ the producer's BN-backward sums into psums, optional) in
// one pass; false when the shape is outside its envelope (k == 3, tiled maps) -- the caller then runs the unfused chain
bool launch_dw_bwd_fused(const float* DY, int ldd, const float* Z, int ldz, const BnRef& bn, int act, const double* sums,
                         const float* X, int ldx, InTf tf, const float* W, float* DX, int lddx, float* dW, float* dgamma,
                         float* dbeta, double* psums, int N, int H, int Wd, int C, int k, cudaStream_t s);
bool dw_tiled_shape(int H, int Wd, int k);   // true when the tiled depthwise family (the one that takes an InTf) handles it
void launch_col_reduce(int mode, const float* A, int lda, const float* B, int ldb, const BnRef& bn, int act, double* sums,
                       int C, long long segs, long long rows_per_seg, cudaStream_t s);
void launch_bn_finalize(const double* sums, int C, long long M, float eps, float momentum, float* mean, float* invstd,
                        float* run_mean, float* run_var, cudaStream_t s);
void launch_bn_apply(const float* Z, int ldz, const BnRef& bn, int act, const float* R, int ldr, float* Y, int ldy, int C,
                     long long M, cudaStream_t s);
void launch_bn_bwd_apply(const float* DY, int ldd, const float* Z, int ldz, const BnRef& bn, int act, const double* sums,
                         float* DZ, int ldo, float* dgamma, float* dbeta, int C, long long M, cudaStream_t s);
void launch_add_sums(const double* sums, float* g, int n, int fold, cudaStream_t s);
void launch_add_copy(const float* A, int lda, const float* B, int ldb, float* O, int ldo, int C, long long M, cudaStream_t s);
void launch_up2(const float* X, int ldx, float* Y, int ldy, int N, int h, int w, int C, cudaStream_t s);
void launch_up2_bwd(const float* DY, int ldd, float* DX, int ldx, int N, int h, int w, int C, cudaStream_t s);
void launch_up2_split(const float* X, int ldx, float* Y0, int ldy0, int C0, float* Y1, int ldy1, int C1, int N, int h, int w,
                      double* sums, cudaStream_t s);
bool launch_up2_bwd_tiled(const float* DY0, int ldd0, const float* Z0, int ldz0, const BnRef& bn, const double* sums, long long M,
                          float* dgamma, float* dbeta, int C0, const float* DY1, int ldd1, int C1, float* DX, int ldx, int N,
                          int h, int w, cudaStream_t s);
void launch_eca_gate(const double* pool, const float* w3, float* mean, float* gate, int N, int C, long long HW, cudaStream_t s);
void launch_eca_gate_bwd(const double* dsum, const float* w3, const float* mean, const float* gate, float* dmean, float* dw3,
                         int N, int C, long long HW, cudaStream_t s);
void launch_scale_rows(const float* A, int lda, const float* G, const float* ADD, float* O, int ldo, int C, long long M,
                       long long HW, cudaStream_t s);
void launch_loss(const float* X, const float* T, long long n, double* acc, int kind, float grad_scale, float* DX,
                 float* loss_out, cudaStream_t s);
void launch_loss_value(const float* X, const float* T, long long n, double* acc, int kind, float* loss_out, cudaStream_t s);
void launch_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                  float wd, int step, float gscale, float max_norm, double* sqn_ws, cudaStream_t s);
void launch_sqnorm(const float* g, long long n, double* out, cudaStream_t s);
void launch_export_view(const void* in, int in_cs, int dt, float* out, int C, long long M, cudaStream_t s);

}  // namespace ysp
