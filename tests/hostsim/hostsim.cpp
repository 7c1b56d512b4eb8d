// tests/hostsim/hostsim.cpp -- TEST-ONLY host build of the kernels' per-ray arithmetic.
//
// Compiles differender_b200/csrc/dr_math.cuh (the very file the CUDA kernels are built from) with g++ so that
// the "not gpu" test suite can compare the device math against the oracle without a GPU: corner-reuse taps vs
// full taps, tape-free reverse march vs the oracle's taped adjoint, merged 32-voxel scatter vs 56-weight scatter.
// It is NOT a CPU fallback: nothing under differender_b200/ builds, loads or calls it.
#include <stdlib.h>
#include <string.h>

#include "dr_desc.h"
#include "dr_math.cuh"

using namespace dr;

namespace {
struct HostVolSink {          // cell-major [cell][8]; same interface as the device CellVolSink, accumulates in place
    float* g;
    float* open(int cell) { return g + (size_t)cell * 8; }
    void close() {}
    void direct(int cell, const float* v) { for (int q = 0; q < 8; ++q) g[(size_t)cell * 8 + q] += v[q]; }
    void flush() {}
};
struct HostTfSink {
    float* g;   // [R][4]
    int Rm1;
    void add(int lo, float f, F4 dc)
    {
        const int hi = lo + 1 < Rm1 ? lo + 1 : Rm1;
        const float w0 = 1.0f - f, w1 = f;
        g[4 * lo + 0] += dc.x * w0; g[4 * lo + 1] += dc.y * w0; g[4 * lo + 2] += dc.z * w0; g[4 * lo + 3] += dc.w * w0;
        g[4 * hi + 0] += dc.x * w1; g[4 * hi + 1] += dc.y * w1; g[4 * hi + 2] += dc.z * w1; g[4 * hi + 3] += dc.w * w1;
    }
    void add_alpha(int lo, float f, float dcw) { add(lo, f, F4 { 0.0f, 0.0f, 0.0f, dcw }); }
    void flush() {}
};
// the staged table the kernels build in shared memory (stage_tf in dr_kernels.cuh): bin r holds tf[r] and tf[min(r+1, R-1)]
TfBin* make_tf_table(const DrDesc& d, const float* tf)
{
    const F4* t4 = (const F4*)tf;
    TfBin* tab = (TfBin*)aligned_alloc(16, sizeof(TfBin) * (size_t)d.R);
    for (int r = 0; r < d.R; ++r) tab[r] = make_tf_bin(t4[r], t4[r + 1 < d.R ? r + 1 : d.R - 1]);
    return tab;
}
}  // namespace

extern "C" {

int sim_desc_init(DrDesc* d, int X, int Y, int Z, int W, int H, int R, int M, unsigned flags, double sr, double fov,
                  double near_plane)
{
    return desc_init(d, X, Y, Z, W, H, R, M, 1, 1, 1, DR_VOX_F32, flags, sr, fov, near_plane) ? -1 : 0;
}

size_t sim_bricked_elems(const DrDesc* d) { return (size_t)d->nbx * d->nby * d->nbz * 512; }
// macro-cells of the skip grid per axis (y, z, x) and their edge length in cells
void sim_macro_dims(const DrDesc* d, int* out4) { out4[0] = macro_ny(*d); out4[1] = macro_nz(*d); out4[2] = macro_nx(*d); out4[3] = kMacro; }

// linear [Y][Z][X] -> bricked
void sim_brick(const DrDesc* d, const float* lin, float* bricked)
{
    Layout L = make_layout(*d, layout_consts(*d));
    memset(bricked, 0, sizeof(float) * sim_bricked_elems(d));
    for (int y = 0; y < d->Y; ++y) for (int z = 0; z < d->Z; ++z) for (int x = 0; x < d->X; ++x)
        bricked[offx(x) + offy(y, L.sY) + offz(z, L.sZ)] = lin[((size_t)y * d->Z + z) * d->X + x];
}
// linear [Y][Z][X] -> cell-major records [cell][8] (what dr_expand_cells builds)
void sim_expand(const DrDesc* d, const float* lin, float* cells)
{
    for (int y = 0; y < d->Y; ++y) for (int z = 0; z < d->Z; ++z) for (int x = 0; x < d->X; ++x) {
        const size_t cell = ((size_t)y * d->Z + z) * d->X + x;
        for (int q = 0; q < 8; ++q) {
            const int a = (q >> 1) & 1, b = q >> 2, c = q & 1;                    // slot = c + 2a + 4b
            const int xx = x + a < d->X ? x + a : d->X - 1, yy = y + b < d->Y ? y + b : d->Y - 1, zz = z + c < d->Z ? z + c : d->Z - 1;
            cells[cell * 8 + q] = lin[((size_t)yy * d->Z + zz) * d->X + xx];
        }
    }
}
// skip grid (what dr_build_skip_grid builds): per-macro-cell voxel min / max, then skip_classify; tf is [R][4]
void sim_skip_grid(const DrDesc* d, const float* lin, const float* tf, unsigned char* grid)
{
    const int E = kMacro + 1;
    for (int my = 0; my < macro_ny(*d); ++my) for (int mz = 0; mz < macro_nz(*d); ++mz) for (int mx = 0; mx < macro_nx(*d); ++mx) {
        float mn = 3.4e38f, mxv = -3.4e38f;
        bool bad = false;
        for (int e = 0; e < E * E * E; ++e) {
            const int x = mx * kMacro + e % E < d->X ? mx * kMacro + e % E : d->X - 1, z = mz * kMacro + (e / E) % E < d->Z ? mz * kMacro + (e / E) % E : d->Z - 1;
            const int y = my * kMacro + e / (E * E) < d->Y ? my * kMacro + e / (E * E) : d->Y - 1;
            const float f = lin[((size_t)y * d->Z + z) * d->X + x];
            bad |= (f != f);
            mn = f < mn ? f : mn; mxv = f > mxv ? f : mxv;
        }
        grid[((size_t)my * macro_nz(*d) + mz) * macro_nx(*d) + mx] = bad ? 0 : skip_classify(*d, mn, mxv, tf + 3, 4);
    }
}
// cell-major gradient [Y*Z*X][8] -> linear [Y][Z][X]
void sim_gather(const DrDesc* d, const float* gcell, float* lin)
{
    for (int y = 0; y < d->Y; ++y) for (int z = 0; z < d->Z; ++z) for (int x = 0; x < d->X; ++x)
        lin[((size_t)y * d->Z + z) * d->X + x] = gather_voxel(*d, gcell, x, y, z);
}

// one view; vol_data is linear [Y][Z][X] or bricked (DR_F_LAYOUT_BRICK8); tf is [R][4]; jitter/out_K/out_Tprev are [H][W] image orientation; out is [4][H][W]
void sim_forward(const DrDesc* d, const float* vol_data, const float* tf, const float* cam3, const float* jitter,
                 float* out, int* out_K, float* out_Tprev, int* out_n, const unsigned char* skip_grid)
{
    Layout L = make_layout(*d, layout_consts(*d));
    VolView<float> vol { vol_data };          // bricked copy or the linear volume, per DR_F_LAYOUT_BRICK8
    TfBin* tab = make_tf_table(*d, tf);
    const TfTable tf4 { tab };
    F3 cam = { cam3[0], cam3[1], cam3[2] };
    const size_t plane = (size_t)d->W * d->H;
    const bool sr1 = d->inv_sr == 1.0f;
    const int taps = tap_mode(*d);
    for (int j = 0; j < d->H; ++j) for (int i = 0; i < d->W; ++i) {
        const size_t pix = (size_t)(d->H - 1 - j) * d->W + i;
        Ray r;
        setup_ray(*d, cam, i, j, jitter ? jitter[pix] : 0.0f, r);
        F4 A; int K; float Tp;
        const bool nd = d->flags & DR_F_NONDIFF;
#define FWD2(LAY, ND, TAPS) do { if (sr1 && TAPS != TAPS_GENERIC) march_forward<float, LAY, ND, TAPS, TAPS != TAPS_GENERIC>(*d, vol, L, tf4, cam, r, A, K, Tp, skip_grid); \
                                else march_forward<float, LAY, ND, TAPS, false>(*d, vol, L, tf4, cam, r, A, K, Tp, skip_grid); } while (0)
#define FWD(LAY, ND) do { if (taps == TAPS_ONE) FWD2(LAY, ND, TAPS_ONE); else FWD2(LAY, ND, TAPS_TWO); } while (0)
        if (d->flags & DR_F_LAYOUT_CELL8) {
            if (nd) FWD(LAYOUT_CELL8, true); else FWD(LAYOUT_CELL8, false);
        } else if (d->flags & DR_F_LAYOUT_BRICK8) {
            if (nd) FWD(LAYOUT_BRICK8, true); else FWD(LAYOUT_BRICK8, false);
        } else if (taps == TAPS_GENERIC) {
            if (nd) FWD2(LAYOUT_LINEAR, true, TAPS_GENERIC); else FWD2(LAYOUT_LINEAR, false, TAPS_GENERIC);
        } else {
            if (nd) FWD(LAYOUT_LINEAR, true); else FWD(LAYOUT_LINEAR, false);
        }
#undef FWD2
#undef FWD
        out[pix] = A.x; out[plane + pix] = A.y; out[2 * plane + pix] = A.z; out[3 * plane + pix] = A.w;
        if (out_K) out_K[pix] = K;
        if (out_Tprev) out_Tprev[pix] = Tp;
        if (out_n) out_n[pix] = r.n;
    }
    free(tab);
}

void sim_backward(const DrDesc* d, const float* vol_data, const float* tf, const float* cam3, const float* jitter,
                  const float* grad_out, const float* out, const int* Kin, const float* Tprev,
                  float* gvol_cells, float* gtf, const unsigned char* skip_grid)
{
    Layout L = make_layout(*d, layout_consts(*d));
    VolView<float> vol { vol_data };          // bricked copy or the linear volume, per DR_F_LAYOUT_BRICK8
    TfBin* tab = make_tf_table(*d, tf);
    const TfTable tf4 { tab };
    F3 cam = { cam3[0], cam3[1], cam3[2] };
    const size_t plane = (size_t)d->W * d->H;
    HostVolSink vs { gvol_cells };
    HostTfSink ts { gtf, d->R - 1 };
    const bool sr1 = d->inv_sr == 1.0f;
    const bool wv = d->flags & DR_F_NEEDS_VOL_GRAD, wt = d->flags & DR_F_NEEDS_TF_GRAD;
    const int taps = tap_mode(*d);
    for (int j = 0; j < d->H; ++j) for (int i = 0; i < d->W; ++i) {
        const size_t pix = (size_t)(d->H - 1 - j) * d->W + i;
        Ray r;
        setup_ray(*d, cam, i, j, jitter ? jitter[pix] : 0.0f, r);
        F4 A = { out[pix], out[plane + pix], out[2 * plane + pix], out[3 * plane + pix] };
        F4 g = { grad_out[pix], grad_out[plane + pix], grad_out[2 * plane + pix], grad_out[3 * plane + pix] };
        const int K = Kin[pix];
        const float Tp = Tprev[pix];
#define CALL(LAY, TAPS, WV, WT) do { if (sr1 && TAPS != TAPS_GENERIC) march_backward<float, LAY, TAPS, WV, WT, TAPS != TAPS_GENERIC, true>(*d, vol, L, tf4, cam, r, A, K, Tp, g, vs, ts, skip_grid); \
                                     else march_backward<float, LAY, TAPS, WV, WT, false, true>(*d, vol, L, tf4, cam, r, A, K, Tp, g, vs, ts, skip_grid); } while (0)
#define CALL3(LAY, TAPS) do { if (wv && wt) CALL(LAY, TAPS, true, true); else if (wv) CALL(LAY, TAPS, true, false); else if (wt) CALL(LAY, TAPS, false, true); } while (0)
        if (d->flags & DR_F_LAYOUT_CELL8) { if (taps == TAPS_ONE) CALL3(LAYOUT_CELL8, TAPS_ONE); else CALL3(LAYOUT_CELL8, TAPS_TWO); }
        else if (d->flags & DR_F_LAYOUT_BRICK8) { if (taps == TAPS_ONE) CALL3(LAYOUT_BRICK8, TAPS_ONE); else CALL3(LAYOUT_BRICK8, TAPS_TWO); }
        else if (taps == TAPS_GENERIC) CALL3(LAYOUT_LINEAR, TAPS_GENERIC);
        else if (taps == TAPS_ONE) CALL3(LAYOUT_LINEAR, TAPS_ONE);
        else CALL3(LAYOUT_LINEAR, TAPS_TWO);
#undef CALL3
#undef CALL
    }
    free(tab);
}

}  // extern "C"

// the fp32 value the kernels give the 256 possible uint8 voxels (DR_VOX_U8): must equal fl(b / 255) for every b
extern "C" void sim_u8_values(float* out256)
{
    for (int b = 0; b < 256; ++b) {
        out256[b] = u8_unit((unsigned)b);
        const F2 p = u8_unit2((unsigned)b, (unsigned)(255 - b));
        if (p.x != out256[b] || p.y != u8_unit((unsigned)(255 - b))) out256[b] = -1.0f;     // the packed form must agree with the scalar one
    }
}
