// dr_bwd_f32.cu -- instantiations of the backward march kernel (dr_kernels.cuh) for fp32-stored volumes.
#include "dr_kernels.cuh"

namespace dr {
int launch_backward_f32(const BwdArgs& a) { return dispatch_bwd_layout<float>(a); }
}  // namespace dr
