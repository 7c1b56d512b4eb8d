// dr_fwd_f32.cu -- instantiations of the forward march kernel (dr_kernels.cuh) for fp32-stored volumes.
#include "dr_kernels.cuh"

namespace dr {
int launch_forward_f32(const FwdArgs& a) { return forward_vt<float>(a); }
}  // namespace dr
