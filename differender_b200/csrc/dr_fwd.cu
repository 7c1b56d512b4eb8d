// dr_fwd.cu -- instantiations of the forward march kernel (dr_kernels.cuh) and their dispatch.
#include "dr_kernels.cuh"

namespace dr {

// The SR1 = false kernels are correct for any sampling rate (powf(x, 1) == x); the generic-tap path (volumes > ~2000 voxels
// per axis, linear layout only) only has those.
template <typename VT>
static int forward_vt(const FwdArgs& a)
{
    const DrDesc* d = a.d;
    const bool nd = d->flags & DR_F_NONDIFF, sr1 = d->inv_sr == 1.0f;
    const int taps = tap_mode(*d);
#define DR_FWD_ND(LAY, TAPS, SR1) (nd ? launch_fwd<VT, LAY, true, TAPS, SR1>(a) : launch_fwd<VT, LAY, false, TAPS, SR1>(a))
#define DR_FWD_SR(LAY, TAPS) (sr1 ? DR_FWD_ND(LAY, TAPS, true) : DR_FWD_ND(LAY, TAPS, false))
    if (d->flags & DR_F_LAYOUT_CELL8) return taps == TAPS_ONE ? DR_FWD_SR(LAYOUT_CELL8, TAPS_ONE) : DR_FWD_SR(LAYOUT_CELL8, TAPS_TWO);
    if (d->flags & DR_F_LAYOUT_BRICK8) return taps == TAPS_ONE ? DR_FWD_SR(LAYOUT_BRICK8, TAPS_ONE) : DR_FWD_SR(LAYOUT_BRICK8, TAPS_TWO);
    if (taps == TAPS_GENERIC) return DR_FWD_ND(LAYOUT_LINEAR, TAPS_GENERIC, false);
    return taps == TAPS_ONE ? DR_FWD_SR(LAYOUT_LINEAR, TAPS_ONE) : DR_FWD_SR(LAYOUT_LINEAR, TAPS_TWO);
#undef DR_FWD_SR
#undef DR_FWD_ND
}

int launch_forward(const FwdArgs& a)
{
    return a.d->vox_dtype == DR_VOX_F32 ? forward_vt<float>(a) : forward_vt<__half>(a);
}

}  // namespace dr
