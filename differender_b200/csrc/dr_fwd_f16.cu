// dr_fwd_f16.cu -- instantiations of the forward march kernel (dr_kernels.cuh) for fp16-stored volumes.
#include "dr_kernels.cuh"

namespace dr {
int launch_forward_f16(const FwdArgs& a) { return forward_vt<__half>(a); }
}  // namespace dr
