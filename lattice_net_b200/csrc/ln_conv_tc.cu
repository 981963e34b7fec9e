// tcgen05 tensor-core lattice convolution (placeholder until the UMMA kernel lands).
#include "ln_common.cuh"
namespace ln {
int conv_fwd_tc(const float*, const int*, const float*, const float*, int, int, int, int, int, int, float*, cudaStream_t) {
    set_error("ln_conv_fwd: tensor-core precision modes are not built yet");
    return LN_ERR_UNSUPPORTED;
}
}  // namespace ln
