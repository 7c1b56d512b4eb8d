// dr_bwd_u8.cu -- instantiations of the backward march kernel (dr_kernels.cuh) for uint8-stored volumes (cell-major copy only).
#include "dr_kernels.cuh"

namespace dr {
int launch_backward_u8(const BwdArgs& a) { return dispatch_bwd_cell8<u8vox>(a); }
}  // namespace dr
