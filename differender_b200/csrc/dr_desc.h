// dr_desc.h -- host-side construction and validation of DrDesc (shared by the C-ABI and tests/hostsim).
// Constants are folded in double and rounded to fp32, as Python scalars captured by the reference's Taichi
// kernels are (volume_raycaster.py:75-78, 146-147, 165, 215, 248-249).
#pragma once
#include <math.h>
#include <string.h>

#include "diffrender.h"

namespace dr {

inline const char* desc_init(DrDesc* d, int X, int Y, int Z, int W, int H, int R, int M, int BS, int Bvol, int Btf,
                             int vox_dtype, unsigned flags, double sr, double fov_deg, double near_plane)
{
    if (!d) return "null descriptor";
    memset(d, 0, sizeof(*d));
    if (X < 2 || Y < 2 || Z < 2) return "volume dims must be >= 2";
    if (W < 1 || H < 1) return "render resolution must be >= 1";
    if (R < 2) return "tf resolution must be >= 2";
    if (R > (1 << 20)) return "tf resolution too large";
    if (M < 1) return "max_samples must be >= 1";
    if (BS < 1) return "need at least one view";
    if (!(Bvol == 1 || Bvol == BS)) return "Bvol must be 1 or BS";
    if (!(Btf == 1 || Btf == BS)) return "Btf must be 1 or BS";
    if (!(sr > 0.0)) return "sampling_rate must be > 0";
    if (vox_dtype != DR_VOX_F32 && vox_dtype != DR_VOX_F16 && vox_dtype != DR_VOX_U8) return "unsupported voxel dtype";
    const long long nbx = (X + 7) / 8, nby = (Y + 7) / 8, nbz = (Z + 7) / 8;
    if (nbx * nby * nbz * 512LL >= (1LL << 31)) return "volume too large (>= 2^31 bricked elements)";
    if ((long long)W * H >= (1LL << 30)) return "image too large";
    d->X = X; d->Y = Y; d->Z = Z; d->W = W; d->H = H; d->R = R; d->M = M;
    d->BS = BS; d->Bvol = Bvol; d->Btf = Btf; d->vox_dtype = vox_dtype; d->flags = flags;
    const double pi = 3.14159265358979323846;
    const double near_h = 2.0 * tan(fov_deg * (pi / 180.0)) * near_plane;    // :146 (tan of the full angle)
    const double near_w = near_h * ((double)W / (double)H);                   // :147, :75
    d->sr = (float)sr; d->inv_sr = (float)(1.0 / sr);
    d->near_ = (float)near_plane; d->near_h = (float)near_h; d->near_w = (float)near_w;
    const int dim[3] = { X, Y, Z };
    double diag2 = 0.0;
    int generic = 0;
    for (int a = 0; a < 3; ++a) {
        d->scale[a] = (float)((double)dim[a] - 1.0 - 1e-4);                   // :165
        diag2 += ((double)dim[a] - 1.0) * ((double)dim[a] - 1.0);
        // a normal tap moves 0.5*delta*scale voxels; the corner-reuse path needs that to stay below one cell
        if (0.5 * 1e-3 * ((double)dim[a] - 1.0) >= 0.999) generic = 1;
    }
    d->vol_diag = (float)sqrt(diag2);                                          // :248-249
    d->tf_len = (float)(R - 1);                                                // :215
    d->ambient = 0.4f; d->diffuse = 0.8f; d->specular = 0.3f;                  // :91-93
    d->ert = 0.99f; d->delta = 1e-3f; d->alpha_skip = 1e-3f;                   // :267, :193, :334
    d->nbx = (int)nbx; d->nby = (int)nby; d->nbz = (int)nbz;
    d->tap_generic = (generic || (flags & DR_F_GENERIC_TAPS)) ? 1 : 0;
    return nullptr;
}

}  // namespace dr
