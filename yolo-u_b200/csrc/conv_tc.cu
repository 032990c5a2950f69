// conv_tc.cu -- tcgen05 + TMA implicit-GEMM convolution (bf16 in, fp32 TMEM accumulate).  STUB until the kernel lands.
#include "kernels.h"
namespace ysp {
struct TcConvPlan { int dummy; };
bool tc_conv_supported(const ConvP&) { return false; }
TcConvPlan* tc_conv_plan_create(const ConvP&, const void*, int) { return nullptr; }
void tc_conv_plan_destroy(TcConvPlan* p) { delete p; }
void launch_conv_tc(const TcConvPlan*, const ConvP&, cudaStream_t) {}
}  // namespace ysp
