// dr_fwd_u8.cu -- instantiations of the forward march kernel (dr_kernels.cuh) for uint8-stored volumes (cell-major copy only).
#include "dr_kernels.cuh"

namespace dr {
int launch_forward_u8(const FwdArgs& a) { return forward_cell8<u8vox>(a); }
}  // namespace dr
