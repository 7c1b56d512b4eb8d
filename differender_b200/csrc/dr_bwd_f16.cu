// dr_bwd_f16.cu -- instantiations of the backward march kernel (dr_kernels.cuh) for fp16-stored volumes.
#include "dr_kernels.cuh"

namespace dr {
int launch_backward_f16(const BwdArgs& a) { return dispatch_bwd_layout<__half>(a); }
}  // namespace dr
