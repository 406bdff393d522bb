// tcgen05 / TMEM / TMA implicit-GEMM modulated convolution (bf16 operands, fp32 accumulation).
// Placeholder until the kernel lands: reports "unsupported" so the launcher uses the CUDA-core path.
#include "conv_common.cuh"

namespace l2i {

bool conv_tc_supported(const ConvGeom&, const EpiParams&) { return false; }

int launch_conv_tc(const void*, const __nv_bfloat16*, const ConvGeom&, const EpiParams&, cudaStream_t) {
  set_error("conv_tc: kernel not built");
  return L2I_ERR_UNSUPPORTED;
}

}  // namespace l2i
