// diffrender.cu -- the C ABI (include/diffrender.h) of the differentiable ray-march and the small kernels around it.
//
// Kernels (DESIGN.md has the roofline of each):
//   brick_kernel     linear [Y][Z][X] volume -> 8x8x8 bricks                 (set_volume, reference :118-119)
//   fwd_kernel       ray set-up + march + compositing + final image          (:221-372)   dr_kernels.cuh / dr_fwd_f32.cu, dr_fwd_f16.cu
//   bwd_kernel       tape-free reverse march, TF + volume gradient scatter   (raycast.grad, :460-461)   dr_bwd_f32.cu / dr_bwd_f16.cu
//   tf_reduce_kernel sums the privatised TF-gradient copies                  (tf_tex.grad.to_torch, :464,475)
//   gather_grad_kernel cell-major fp32 gradient -> linear, nan_to_num        (volume.grad.to_torch, :463,474)
//   gather_step_kernel gather + momentum/projection step + cell-major volume refresh, fused   (examples :375-381, test_opt_tf.py:86-88)
//   momentum_step_kernel, ingest_u8_kernel                                   (the caller's steps either side of the march)
//   l2_read_probe_kernel L2 -> SM read bandwidth of the device (bench.py's L2 roof)
//
// The march kernels are instantiated in four other translation units so that the library builds in parallel; the
// bounds-checking debug build (-DDR_BOUNDS_CHECK -DDR_UNITY_BUILD) compiles everything as one unit instead, because its
// device-side violation counter is one __device__ variable.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "diffrender.h"
#include "dr_desc.h"
#include "dr_host.h"
#include "dr_kernels.cuh"
#include "dr_math.cuh"

#if defined(DR_BOUNDS_CHECK)
__device__ unsigned long long dr_oob_counter = 0ULL;
#endif
#if defined(DR_UNITY_BUILD)
#include "dr_fwd_f32.cu"
#include "dr_fwd_f16.cu"
#include "dr_fwd_u8.cu"
#include "dr_bwd_f32.cu"
#include "dr_bwd_f16.cu"
#include "dr_bwd_u8.cu"
#endif

using namespace dr;

namespace dr {
static thread_local char g_err[256] = "";

int fail(int code, const char* msg)
{
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
int fail_cuda(cudaError_t e, const char* where)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return DR_ECUDA;
}
}  // namespace dr

namespace {

// the fp32 value of a stored voxel
__device__ __forceinline__ float vox_value(float v) { return v; }
__device__ __forceinline__ float vox_value(__half v) { return __half2float(v); }
__device__ __forceinline__ float vox_value(u8vox v) { return u8_unit(v); }

// ---------------------------------------------------------------------------------------------------------
// bricking / un-bricking  (HBM-bound: 2 * sizeof(voxel) bytes per voxel)
// ---------------------------------------------------------------------------------------------------------
template <typename VT>
__global__ void __launch_bounds__(256) brick_kernel(DrDesc d, const VT* __restrict__ lin, VT* __restrict__ br, size_t elems)
{
    // one thread per bricked element: writes are fully coalesced, reads are 8-voxel (32/16-byte) row segments
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= elems) return;
    const int b = blockIdx.y;
    const int inb = (int)(e & 511);
    size_t brick = e >> 9;
    const int bx = (int)(brick % d.nbx); brick /= d.nbx;
    const int by = (int)(brick % d.nby);
    const int bz = (int)(brick / d.nby);
    const int x = bx * 8 + (inb & 7), y = by * 8 + ((inb >> 3) & 7), z = bz * 8 + (inb >> 6);
    VT v = VT(0.0f);
    if (x < d.X && y < d.Y && z < d.Z)
        v = lin[(size_t)b * d.X * d.Y * d.Z + ((size_t)y * d.Z + z) * d.X + x];
    br[(size_t)b * elems + e] = v;
}

// linear [Y][Z][X] volume -> cell-major records [cell][8] (LAYOUT_CELL8): slot c + 2a + 4b of cell (x,y,z) = voxel
// (min(x+a,X-1), min(y+b,Y-1), min(z+c,Z-1)).  Write-bound: 8 * sizeof(voxel) bytes per voxel; the 8x re-read of the
// input is served by L1/L2.
template <typename VT>
__global__ void __launch_bounds__(256) expand_cells_kernel(DrDesc d, const VT* __restrict__ lin, VT* __restrict__ cells)
{
    const size_t n = (size_t)d.X * d.Y * d.Z;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int b = blockIdx.y;
    const int x = (int)(e % d.X);
    const size_t r = e / d.X;
    const int z = (int)(r % d.Z), y = (int)(r / d.Z);
    const size_t dx = x + 1 < d.X ? 1 : 0, dz = z + 1 < d.Z ? (size_t)d.X : 0, dy = y + 1 < d.Y ? (size_t)d.X * d.Z : 0;
    const VT* p = lin + (size_t)b * n + e;
    struct alignas(sizeof(VT) * 8 < 16 ? sizeof(VT) * 8 : 16) Rec { VT v[8]; } rec;      // 32 / 16 / 8 bytes
    rec.v[0] = p[0]; rec.v[1] = p[dz]; rec.v[2] = p[dx]; rec.v[3] = p[dx + dz];
    rec.v[4] = p[dy]; rec.v[5] = p[dy + dz]; rec.v[6] = p[dy + dx]; rec.v[7] = p[dy + dx + dz];
    reinterpret_cast<Rec*>(cells)[(size_t)b * n + e] = rec;
}

// skip grid, step 1: voxel min / max of every macro-cell (8x8x8 cells = the 9x9x9 voxels its cells' corners touch).
// One warp per macro-cell; depends on the volume only.
template <typename VT>
__global__ void __launch_bounds__(256) skip_minmax_kernel(DrDesc d, const VT* __restrict__ lin, float2* __restrict__ mm)
{
    const int nx = macro_nx(d), nz = macro_nz(d);
    const size_t cells = (size_t)nx * macro_ny(d) * nz;
    const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= cells) return;
    const int lane = threadIdx.x & 31, b = blockIdx.y;
    const int mx_ = (int)(w % nx), mz_ = (int)((w / nx) % nz), my_ = (int)(w / ((size_t)nx * nz));
    constexpr int E = kMacro + 1;                                  // the (kMacro + 1)^3 voxels the macro-cell's cells touch
    const VT* v = lin + (size_t)b * d.X * d.Y * d.Z;
    float mn = 3.4e38f, mx = -3.4e38f;
    bool bad = false;
    for (int e = lane; e < E * E * E; e += 32) {
        const int x = min(mx_ * kMacro + e % E, d.X - 1), z = min(mz_ * kMacro + (e / E) % E, d.Z - 1), y = min(my_ * kMacro + e / (E * E), d.Y - 1);
        const float f = vox_value(v[((size_t)y * d.Z + z) * d.X + x]);
        bad |= (f != f);
        mn = fminf(mn, f); mx = fmaxf(mx, f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) mm[(size_t)b * cells + w] = bad ? make_float2(1.0f, -1.0f) : make_float2(mn, mx);     // NaN inside: mn > mx = never skip
}

// skip grid, step 2: classify every macro-cell against the view's transfer function (dr_math.cuh skip_classify)
__global__ void __launch_bounds__(256) skip_classify_kernel(DrDesc d, const float2* __restrict__ mm, const float* __restrict__ tf,
                                                            unsigned char* __restrict__ grid, int views)
{
    const size_t cells = skip_cells(&d);
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    const bool live = e < cells;
    const float2 r = live ? mm[(size_t)(d.Bvol == 1 ? 0 : b) * cells + e] : make_float2(1.0f, -1.0f);
    const float* t = tf + (size_t)(d.Btf == 1 ? 0 : b) * d.R * 4;
    const bool t4r = d.flags & DR_F_TF_4R;
    const unsigned char c = !live ? 0 : skip_classify(d, r.x, r.y, t4r ? t + 3 * (size_t)d.R : t + 3, t4r ? 1 : 4);
    if (live) grid[kSkipHeader + (size_t)b * cells + e] = c;
    const int n = __syncthreads_count(c);                       // header: total number of empty macro-cells
    if (threadIdx.x == 0 && n) atomicAdd(reinterpret_cast<unsigned*>(grid), (unsigned)n);
}

// ---------------------------------------------------------------------------------------------------------
// Gather of the cell-major gradient, and the fused optimiser step on top of it.
//
// The gradient of voxel (x,y,z) is the sum of the (up to) 8 slots that alias it: cell (x-a, y-b, z-c), slot a + 2b + 4c.
// Both kernels below load whole records, each as two coalesced 16-byte loads, into shared memory transposed to 8 slot planes
// (consecutive lanes -> consecutive cells -> conflict-free), and every voxel then sums its 8 slots from shared memory in a fixed
// order.  Round 1's kernel read 8 scattered 4-byte words per voxel through L1 out of local index arrays (51 LDL/STL) and reached
// 36 % of the HBM roof.  gather_step_kernel works on GT_X x GT_Y x GT_Z tiles (records of the cells [origin - 1, origin + T): the
// tile plus a one-cell halo, 1.45x, served by L2); gather_grad_kernel marches along y instead (below).
// Voxels on the last plane of an axis also receive the clamped slot of their own cell (:170-172); the march never writes
// such cells, but dr_gather_grad defines it, so those voxels take gather_voxel()'s general path from global memory.
// ---------------------------------------------------------------------------------------------------------
#ifndef DR_GT_Y
#define DR_GT_Y 4
#endif
#ifndef DR_GT_LOAD
#define DR_GT_LOAD __ldg
#endif
constexpr int GT_X = 32, GT_Y = DR_GT_Y, GT_Z = 8, GT_THREADS = 512;
constexpr int GR_X = GT_X + 1, GR_Y = GT_Y + 1, GR_Z = GT_Z + 1;       // records per axis: the tile and its -1 halo
constexpr int GR_XP = GR_X | 1;                                          // odd row pitch
constexpr int GR_PLANE = GR_Y * GR_Z * GR_XP;
constexpr size_t kGatherSmem = (size_t)8 * GR_PLANE * sizeof(float);    // 47.5 KB

__device__ __forceinline__ float nan_to_num(float v)           // torch.nan_to_num (:463, :474)
{
    if (v != v) return 0.0f;
    return fminf(fmaxf(v, -3.4028234663852886e38f), 3.4028234663852886e38f);
}

__device__ __forceinline__ void gather_tile_origin(const DrDesc& d, int& ox, int& oy, int& oz)
{
    const int ntx = (d.X + GT_X - 1) / GT_X, ntz = (d.Z + GT_Z - 1) / GT_Z;
    ox = (blockIdx.x % ntx) * GT_X; oz = ((blockIdx.x / ntx) % ntz) * GT_Z; oy = (blockIdx.x / (ntx * ntz)) * GT_Y;
}

// loads the tile's gradient records into shared memory (slot planes); cells outside the volume read as zero.  All of a thread's
// records (three at most) are requested before the first one is stored, so that a CTA has its whole tile in flight at once.
__device__ __forceinline__ void gather_load_tile(const DrDesc& d, const float* __restrict__ gcell, int ox, int oy, int oz, float* s)
{
    constexpr int kRecs = GR_X * GR_Y * GR_Z, kIter = (kRecs + GT_THREADS - 1) / GT_THREADS;
    float4 a[kIter], b[kIter];
#pragma unroll
    for (int k = 0; k < kIter; ++k) {
        const int r = threadIdx.x + k * GT_THREADS;
        const int rx = r % GR_X, rz = (r / GR_X) % GR_Z, ry = r / (GR_X * GR_Z);
        const int cx = ox - 1 + rx, cy = oy - 1 + ry, cz = oz - 1 + rz;
        a[k] = make_float4(0.f, 0.f, 0.f, 0.f); b[k] = a[k];
        if (r < kRecs && cx >= 0 && cy >= 0 && cz >= 0 && cx < d.X && cy < d.Y && cz < d.Z) {
            const float4* p = reinterpret_cast<const float4*>(gcell) + (((size_t)cy * d.Z + cz) * d.X + cx) * 2;
            a[k] = DR_GT_LOAD(p); b[k] = DR_GT_LOAD(p + 1);
        }
    }
#pragma unroll
    for (int k = 0; k < kIter; ++k) {
        const int r = threadIdx.x + k * GT_THREADS;
        if (r >= kRecs) break;
        const int rx = r % GR_X, rz = (r / GR_X) % GR_Z, ry = r / (GR_X * GR_Z);
        float* q = s + (ry * GR_Z + rz) * GR_XP + rx;
        q[0 * GR_PLANE] = a[k].x; q[1 * GR_PLANE] = a[k].y; q[2 * GR_PLANE] = a[k].z; q[3 * GR_PLANE] = a[k].w;
        q[4 * GR_PLANE] = b[k].x; q[5 * GR_PLANE] = b[k].y; q[6 * GR_PLANE] = b[k].z; q[7 * GR_PLANE] = b[k].w;
    }
}

// gradient of the voxel at tile-local coordinates (vx, vy, vz) = global (x, y, z), summed in gather_voxel()'s order
__device__ __forceinline__ float gather_from_tile(const DrDesc& d, const float* __restrict__ gcell, const float* s, int vx, int vy, int vz,
                                                  int x, int y, int z)
{
    if (x == d.X - 1 || y == d.Y - 1 || z == d.Z - 1) return gather_voxel(d, gcell, x, y, z);     // clamped slots: general path
    float sum = 0.0f;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int a = 0; a < 2; ++a)      // record local index = voxel local index + 1 - offset; cells outside the volume hold +0
                sum += s[(a + 2 * b + 4 * c) * GR_PLANE + ((vy + 1 - b) * GR_Z + (vz + 1 - c)) * GR_XP + (vx + 1 - a)];
    return sum;
}

// dr_gather_grad: cell-major gradient [cell][8] -> linear [Y][Z][X] fp32 with nan_to_num, marching along y.  A CTA owns a 32 (x) x
// GM_Z (z) column of voxels and walks GM_CH planes of y.  Per step it loads ONE plane of records (33 x (GM_Z + 1)) into shared-memory
// slot planes and keeps the previous plane, so there is no y halo (the re-read factor of the 32x4x8 tile above, 1.45, falls to
// 33/32 * (GM_Z+1)/GM_Z * (GM_CH+1)/GM_CH = 1.19), and the records of plane y + 1 are already in flight in registers while plane y
// is summed and written (software pipelining across the loop).  Measured: 60 % (256^3) / 74 % (512^3) of the HBM copy peak on
// compulsory bytes, against 53 / 59 % for the tile version and 36 % for round 1's scattered loads (profiles/r02_experiments.md).
constexpr int GM_X = 32, GM_Z = 8, GM_CH = 32, GM_THREADS = GM_X * GM_Z;
constexpr int GM_RX = GM_X + 1, GM_RZ = GM_Z + 1, GM_RXP = GM_RX | 1, GM_PLANE = GM_RZ * GM_RXP;      // one slot plane of one y plane
constexpr int GM_RECS = GM_RX * GM_RZ, GM_ITER = (GM_RECS + GM_THREADS - 1) / GM_THREADS;

__global__ void __launch_bounds__(GM_THREADS) gather_grad_kernel(DrDesc d, const float* __restrict__ gcell, float* __restrict__ lin, int accumulate)
{
    __shared__ float s[2][8][GM_PLANE];
    const int ntx = (d.X + GM_X - 1) / GM_X, ntz = (d.Z + GM_Z - 1) / GM_Z;
    const int ox = (blockIdx.x % ntx) * GM_X, oz = ((blockIdx.x / ntx) % ntz) * GM_Z, y0 = (blockIdx.x / (ntx * ntz)) * GM_CH;
    const size_t n = (size_t)d.X * d.Y * d.Z;
    const float* gc = gcell + (size_t)blockIdx.y * n * 8;
    float* out = lin + (size_t)blockIdx.y * n;
    const int tx = threadIdx.x % GM_X, tz = threadIdx.x / GM_X;
    const int y1 = min(y0 + GM_CH, d.Y);
    float4 a[GM_ITER], b[GM_ITER];
    auto fetch = [&](int cy) {                  // the records of cell plane cy -> registers (zeros outside the volume)
#pragma unroll
        for (int k = 0; k < GM_ITER; ++k) {
            const int r = threadIdx.x + k * GM_THREADS;
            const int cx = ox - 1 + r % GM_RX, cz = oz - 1 + r / GM_RX;
            a[k] = make_float4(0.f, 0.f, 0.f, 0.f); b[k] = a[k];
            if (r < GM_RECS && cy >= 0 && cy < d.Y && cx >= 0 && cz >= 0 && cx < d.X && cz < d.Z) {
                const float4* p = reinterpret_cast<const float4*>(gc) + (((size_t)cy * d.Z + cz) * d.X + cx) * 2;
                a[k] = __ldg(p); b[k] = __ldg(p + 1);
            }
        }
    };
    auto stash = [&](int buf) {                 // registers -> slot planes of buffer `buf`
#pragma unroll
        for (int k = 0; k < GM_ITER; ++k) {
            const int r = threadIdx.x + k * GM_THREADS;
            if (r >= GM_RECS) break;
            float* q = &s[buf][0][(r / GM_RX) * GM_RXP + r % GM_RX];
            q[0 * GM_PLANE] = a[k].x; q[1 * GM_PLANE] = a[k].y; q[2 * GM_PLANE] = a[k].z; q[3 * GM_PLANE] = a[k].w;
            q[4 * GM_PLANE] = b[k].x; q[5 * GM_PLANE] = b[k].y; q[6 * GM_PLANE] = b[k].z; q[7 * GM_PLANE] = b[k].w;
        }
    };
    fetch(y0 - 1);
    stash(1);                                   // plane y0 - 1 plays "previous" for the first step
    fetch(y0);
    for (int y = y0; y < y1; ++y) {
        const int cur = (y - y0) & 1;
        __syncthreads();                        // everyone is done reading buffer `cur` (it held plane y - 2)
        stash(cur);
        if (y + 1 < y1) fetch(y + 1);           // in flight while this plane is summed
        __syncthreads();
        const int x = ox + tx, z = oz + tz;
        if (x < d.X && z < d.Z) {
            float g;
            if (x == d.X - 1 || y == d.Y - 1 || z == d.Z - 1) g = gather_voxel(d, gc, x, y, z);       // clamped slots: general path
            else {
                g = 0.0f;
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int bb = 0; bb < 2; ++bb)
#pragma unroll
                        for (int aa = 0; aa < 2; ++aa)      // cell (x - aa, y - bb, z - c), slot aa + 2 bb + 4 c; bb = 1 reads the previous plane
                            g += s[bb ? cur ^ 1 : cur][aa + 2 * bb + 4 * c][(tz + 1 - c) * GM_RXP + (tx + 1 - aa)];
            }
            g = nan_to_num(g);
            float* o = out + ((size_t)y * d.Z + z) * d.X + x;
            *o = accumulate ? (*o + g) : g;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// dr_gather_step: [gather + nan_to_num] + momentum-SGD step with clipping and projection + refresh of the cell-major
// VOLUME copy, in ONE kernel (replaces gather_grad_kernel + momentum_step_kernel + expand_cells_kernel).
//   m = gamma*m + lr*clamp(g, -max_grad, max_grad);  p = clamp(p - m, lo, hi)    one rounding per operator (numpy float32)
//   (examples/taichi_volume_raycaster.py:375-381 `apply_grad`; examples/test_opt_tf.py:86-88 vol.clamp_(0, 1))
// Every (record, slot) of the volume copy has exactly one source voxel -- slot c + 2a + 4b of cell q is voxel min(q + (a,b,c),
// dim - 1) -- so a CTA writes the slots fed by ITS OWN voxels and nothing else: no halo of parameters is read (the update is in
// place and race-free), records inside the tile go out as whole 32-byte (fp16: 16-byte) stores, and the records on the tile's low
// faces are completed by the neighbouring CTAs' partial stores (merged in L2).
// ---------------------------------------------------------------------------------------------------------
struct StepArgs { float lr, gamma, max_grad, lo, hi; };

template <typename VT, bool FROM_CELLS>
__global__ void __launch_bounds__(GT_THREADS) gather_step_kernel(DrDesc d, const float* __restrict__ gcell, const float* __restrict__ glin,
                                                                  float* __restrict__ param, float* __restrict__ mom, VT* __restrict__ vcells,
                                                                  float* __restrict__ gout, StepArgs sa)
{
    extern __shared__ __align__(16) float s_rec[];
    float (*s_p)[GT_Z][GT_X] = reinterpret_cast<float (*)[GT_Z][GT_X]>(s_rec + (FROM_CELLS ? 8 * GR_PLANE : 0));   // the tile's new parameter values
    int ox, oy, oz;
    gather_tile_origin(d, ox, oy, oz);
    if (FROM_CELLS) {
        gather_load_tile(d, gcell, ox, oy, oz, s_rec);
        __syncthreads();
    }
    for (int v = threadIdx.x; v < GT_X * GT_Y * GT_Z; v += GT_THREADS) {
        const int vx = v % GT_X, vz = (v / GT_X) % GT_Z, vy = v / (GT_X * GT_Z);
        const int x = ox + vx, y = oy + vy, z = oz + vz;
        if (x >= d.X || y >= d.Y || z >= d.Z) continue;
        const size_t e = ((size_t)y * d.Z + z) * d.X + x;
        const float g = FROM_CELLS ? nan_to_num(gather_from_tile(d, gcell, s_rec, vx, vy, vz, x, y, z)) : __ldg(glin + e);
        const float gc = fminf(fmaxf(g, -sa.max_grad), sa.max_grad);
        const float m = __fadd_rn(__fmul_rn(sa.gamma, mom[e]), __fmul_rn(sa.lr, gc));
        const float p = fminf(fmaxf(__fsub_rn(param[e], m), sa.lo), sa.hi);
        mom[e] = m; param[e] = p;
        if (gout) gout[e] = g;
        s_p[vy][vz][vx] = p;
    }
    if (!vcells) return;
    __syncthreads();
    // records of the cells [origin - 1, origin + T): slot (a,b,c) comes from voxel min(cell + (a,b,c), dim - 1) if that voxel is ours
    for (int r = threadIdx.x; r < GR_X * GR_Y * GR_Z; r += GT_THREADS) {
        const int rx = r % GR_X, rz = (r / GR_X) % GR_Z, ry = r / (GR_X * GR_Z);
        const int cx = ox - 1 + rx, cy = oy - 1 + ry, cz = oz - 1 + rz;
        if (cx < 0 || cy < 0 || cz < 0 || cx >= d.X || cy >= d.Y || cz >= d.Z) continue;
        struct alignas(16) Rec { VT v[8]; } rec;
        unsigned mine = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {                                              // slot q = c + 2a + 4b
            const int a = (q >> 1) & 1, b = q >> 2, c = q & 1;
            const int lx = min(cx + a, d.X - 1) - ox, ly = min(cy + b, d.Y - 1) - oy, lz = min(cz + c, d.Z - 1) - oz;
            const bool in = lx >= 0 && ly >= 0 && lz >= 0 && lx < GT_X && ly < GT_Y && lz < GT_Z;
            rec.v[q] = VT(in ? s_p[in ? ly : 0][in ? lz : 0][in ? lx : 0] : 0.0f);
            mine |= in ? (1u << q) : 0u;
        }
        VT* dst = vcells + (((size_t)cy * d.Z + cz) * d.X + cx) * 8;
        if (mine == 0xFFu) *reinterpret_cast<Rec*>(dst) = rec;
        else
#pragma unroll
            for (int q = 0; q < 8; ++q) if (mine & (1u << q)) dst[q] = rec.v[q];
    }
}

// L2 -> SM read-bandwidth probe (bench.py's L2 roof; SURVEY 8(d): MEASURED_PEAKS.json has no L2 figure): every CTA reads the whole
// buffer `reps` times with 16-byte ld.global.cg loads (L1 bypassed), CTA b starting at a different offset so that the CTAs do not
// march through the same lines in lock-step.  With a buffer that fits L2 (<= 64 MiB of the 126 MB) all passes after the first are L2 hits.
__global__ void __launch_bounds__(512) l2_read_probe_kernel(const uint4* __restrict__ buf, size_t n16, int reps, unsigned* __restrict__ sink)
{
    unsigned acc = 0;
    const size_t start = ((size_t)blockIdx.x * 8191u * 512u) % n16;
    for (int r = 0; r < reps; ++r) {
        size_t i = start + threadIdx.x;
#pragma unroll 8
        for (size_t k = threadIdx.x; k < n16; k += 512) {
            if (i >= n16) i -= n16;
            const uint4 v = __ldcg(buf + i);
            acc += v.x ^ v.y ^ v.z ^ v.w;
            i += 512;
        }
    }
    if (acc == 0x9E3779B9u) *sink = acc;          // keeps the loads alive; practically never taken
}

// sums the kTfSlots privatised copies: one WARP per (tf, bin, channel) -- each lane adds every 32nd copy (independent loads), then a
// shuffle tree -- and adds the total into grad_tf in the caller's layout.  (Round 1 used one THREAD per output walking all 1024
// copies serially: 48 us, 1.4 % of a C2 iteration.)
__global__ void __launch_bounds__(256) tf_reduce_kernel(DrDesc d, const float* __restrict__ slots, float* __restrict__ grad_tf)
{
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;       // r*4 + c
    const int lane = threadIdx.x & 31, tb = blockIdx.y;
    if (e >= d.R * 4) return;
    // a copy has R + 1 bins: the march writes the pair (lo, lo + 1) without clamping, and what lands in the pad bin R belongs to
    // bin R - 1 (the reference clamps the upper bin, :216-218)
    const size_t copy = (size_t)(d.R + 1) * 4;
    const float* p = slots + (size_t)tb * kTfSlots * copy + e;
    const bool last = (e >> 2) == d.R - 1;
    float acc = 0.0f;
#pragma unroll 8
    for (int s = lane; s < kTfSlots; s += 32) {
        acc += __ldg(p + (size_t)s * copy);
        if (last) acc += __ldg(p + (size_t)s * copy + 4);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane) return;
    float v = acc;
    if (v != v) v = 0.0f;                                      // torch.nan_to_num (:464, :475)
    const int r = e >> 2, c = e & 3;
    float* o = grad_tf + (size_t)tb * d.R * 4 + ((d.flags & DR_F_TF_4R) ? (c * d.R + r) : e);
    *o += v;
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
int check_desc(const DrDesc* d)
{
    if (!d) return fail(DR_EINVAL, "null descriptor");
    if (d->X < 2 || d->Y < 2 || d->Z < 2 || d->W < 1 || d->H < 1 || d->R < 2 || d->M < 1 || d->BS < 1)
        return fail(DR_EINVAL, "descriptor not initialised (use dr_desc_init)");
    if (d->nbx != (d->X + 7) / 8 || d->nby != (d->Y + 7) / 8 || d->nbz != (d->Z + 7) / 8)
        return fail(DR_EINVAL, "descriptor brick counts inconsistent (use dr_desc_init)");
    if (d->vox_dtype != DR_VOX_F32 && d->vox_dtype != DR_VOX_F16 && d->vox_dtype != DR_VOX_U8) return fail(DR_EDTYPE, "unsupported voxel dtype");
    if (d->BS > 65535) return fail(DR_EINVAL, "more than 65535 views in one call");
    if (d->tap_generic && (d->flags & (DR_F_LAYOUT_BRICK8 | DR_F_LAYOUT_CELL8)))
        return fail(DR_EINVAL, "the generic tap path (volumes > ~2000 voxels per axis) needs the linear layout");
    if ((d->flags & DR_F_LAYOUT_BRICK8) && (d->flags & DR_F_LAYOUT_CELL8)) return fail(DR_EINVAL, "two volume layouts selected");
    return DR_OK;
}

bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

unsigned gather_tiles(const DrDesc* d)
{
    return (unsigned)(((d->X + GT_X - 1) / GT_X) * ((d->Z + GT_Z - 1) / GT_Z) * ((d->Y + GT_Y - 1) / GT_Y));
}

// ---------------------------------------------------------------------------------------------------------
// caller-side steps either side of the march: optimiser update and raw-volume ingest.  Elementwise, HBM-bound.
// ---------------------------------------------------------------------------------------------------------
// momentum-SGD step of the reference's TF optimisation demo (examples/taichi_volume_raycaster.py:375-381):
//   m = gamma*m + lr*clamp(g, -max_grad, max_grad);  p -= m;  p = clamp(p, lo, hi)
// one rounding per operator (matches numpy float32 bit for bit)
__global__ void __launch_bounds__(256) momentum_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                             size_t n, float lr, float gamma, float max_grad, float lo, float hi)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const float gc = fminf(fmaxf(g[e], -max_grad), max_grad);
    const float mv = __fadd_rn(__fmul_rn(gamma, m[e]), __fmul_rn(lr, gc));
    m[e] = mv;
    p[e] = fminf(fmaxf(__fsub_rn(p[e], mv), lo), hi);
}

// raw uint8 volume [A][B][C] -> linear voxel volume [Y][Z][X], value = u8 / 255 (fp32 division), optionally with the
// reference's np.swapaxes(raw, 0, 1) (examples/taichi_volume_raycaster.py:548-550) fused into the read
template <typename VT>
__global__ void __launch_bounds__(256) ingest_u8_kernel(DrDesc d, const uint8_t* __restrict__ src, VT* __restrict__ dst, int swap01)
{
    const size_t n = (size_t)d.X * d.Y * d.Z;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int x = (int)(e % d.X);
    const size_t r = e / d.X;
    const int z = (int)(r % d.Z), y = (int)(r / d.Z);
    // dst[y][z][x] = swap01 ? src[z][y][x] (src dims [Z][Y][X]) : src[y][z][x]
    const size_t s = swap01 ? ((size_t)z * d.Y + y) * d.X + x : e;
    dst[e] = VT(__fdiv_rn((float)src[s], 255.0f));
}

int forward_impl(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter, float* out_rgba,
                 int32_t* out_K, float* out_Tprev, const float* target, float* loss_sum, const unsigned char* skip_grid, void* stream)
{
    if (int rc = check_desc(d)) return rc;
    if (!vol || !tf || !cam || !out_rgba) return fail(DR_EINVAL, "dr_forward: null pointer");
    if ((d->flags & DR_F_HAS_JITTER) && !jitter) return fail(DR_EINVAL, "dr_forward: DR_F_HAS_JITTER set but jitter is null");
    if (!aligned(tf, 16) || !aligned(out_rgba, 16)) return fail(DR_EALIGN, "dr_forward: tf and out_rgba must be 16-byte aligned");
    if (target && (!loss_sum || !aligned(target, 16))) return fail(DR_EINVAL, "dr_forward_mse: loss_sum is null or target is not 16-byte aligned");
    if ((size_t)d->R * 32 > kMaxTfSmem) return fail(DR_EINVAL, "tf resolution too large for shared memory staging (R <= 6400)");
    if (skip_grid && d->tap_generic) skip_grid = nullptr;          // the generic tap path marches every sample
    const FwdArgs a { d, vol, tf, cam, jitter, out_rgba, out_K, out_Tprev, static_cast<cudaStream_t>(stream), target, loss_sum, skip_grid };
    if (d->vox_dtype == DR_VOX_U8) {
        if (!(d->flags & DR_F_LAYOUT_CELL8)) return fail(DR_EDTYPE, "uint8 volumes are marched from their cell-major copy only (dr_expand_cells, DR_F_LAYOUT_CELL8)");
        return launch_forward_u8(a);
    }
    return d->vox_dtype == DR_VOX_F32 ? launch_forward_f32(a) : launch_forward_f16(a);
}

int backward_impl(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter,
                  const float* grad_out, const float* out_rgba, const int32_t* K, const float* Tprev, float* grad_vol_cells,
                  float* grad_tf, void* workspace, size_t workspace_bytes, float mse_scale, void* stream,
                  const unsigned char* skip_grid = nullptr, const float* scale_dev = nullptr)
{
    if (int rc = check_desc(d)) return rc;
    if (d->flags & DR_F_NONDIFF) return fail(DR_EINVAL, "dr_backward: the non-differentiable march has no backward");
    const bool wv = d->flags & DR_F_NEEDS_VOL_GRAD, wt = d->flags & DR_F_NEEDS_TF_GRAD;
    if (!wv && !wt) return DR_OK;
    if (!vol || !tf || !cam || !grad_out || !out_rgba || !K || !Tprev) return fail(DR_EINVAL, "dr_backward: null pointer");
    if ((d->flags & DR_F_HAS_JITTER) && !jitter) return fail(DR_EINVAL, "dr_backward: DR_F_HAS_JITTER set but jitter is null");
    if (wv && !grad_vol_cells) return fail(DR_EINVAL, "dr_backward: grad_vol_cells is null");
    if (wv && !aligned(grad_vol_cells, 32)) return fail(DR_EALIGN, "dr_backward: grad_vol_cells must be 32-byte aligned");
    if (wt && !grad_tf) return fail(DR_EINVAL, "dr_backward: grad_tf is null");
    if (!aligned(tf, 16) || !aligned(out_rgba, 16) || !aligned(grad_out, 16))
        return fail(DR_EALIGN, "dr_backward: tf, out_rgba and grad_out must be 16-byte aligned");
    const size_t need = dr_workspace_bytes(d);
    if (wt) {
        if (!workspace || workspace_bytes < need) return fail(DR_EWORKSPACE, "dr_backward: workspace too small (dr_workspace_bytes)");
        if (!aligned(workspace, 16)) return fail(DR_EALIGN, "dr_backward: workspace must be 16-byte aligned");
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (wt) {
        cudaError_t e = cudaMemsetAsync(workspace, 0, need, st);
        if (e != cudaSuccess) return fail_cuda(e, "cudaMemsetAsync(workspace)");
    }
    if ((size_t)d->R * 32 > kMaxTfSmem) return fail(DR_EINVAL, "tf resolution too large for shared memory staging (R <= 6400)");
    if (skip_grid && (wt || d->tap_generic)) skip_grid = nullptr;         // transparent samples feed the TF gradient: nothing to skip then
    if (skip_grid && !aligned(skip_grid, 4)) return fail(DR_EALIGN, "dr_backward_ex: skip_grid must be 4-byte aligned");
    const BwdArgs a { d, vol, tf, cam, jitter, grad_out, out_rgba, K, Tprev, reinterpret_cast<float4*>(grad_vol_cells),
                      static_cast<float4*>(workspace), st, mse_scale, skip_grid, scale_dev };
    if (d->vox_dtype == DR_VOX_U8 && !(d->flags & DR_F_LAYOUT_CELL8))
        return fail(DR_EDTYPE, "uint8 volumes are marched from their cell-major copy only (dr_expand_cells, DR_F_LAYOUT_CELL8)");
    const int rc = d->vox_dtype == DR_VOX_U8 ? launch_backward_u8(a) : d->vox_dtype == DR_VOX_F32 ? launch_backward_f32(a) : launch_backward_f16(a);
    if (rc) return rc;
    if (wt) {
        dim3 grid((d->R * 4 * 32 + 255) / 256, d->Btf);
        tf_reduce_kernel<<<grid, 256, 0, st>>>(*d, static_cast<const float*>(workspace), grad_tf);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail_cuda(e, "tf_reduce_kernel launch");
    }
    return DR_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int dr_version(void) { return DR_VERSION; }

// number of out-of-range volume loads / gradient reductions seen so far; always 0 unless built with -DDR_BOUNDS_CHECK
long long dr_debug_oob_count(void)
{
#if defined(DR_BOUNDS_CHECK)
    unsigned long long v = 0;
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(&v, dr_oob_counter, sizeof(v)) != cudaSuccess) return -1;
    return (long long)v;
#else
    return 0;
#endif
}

const char* dr_last_error(void) { return g_err; }

int dr_desc_init(DrDesc* d, int32_t X, int32_t Y, int32_t Z, int32_t W, int32_t H, int32_t R, int32_t M, int32_t BS,
                 int32_t Bvol, int32_t Btf, int32_t vox_dtype, uint32_t flags, double sampling_rate, double fov_deg,
                 double near_plane)
{
    const char* msg = desc_init(d, X, Y, Z, W, H, R, M, BS, Bvol, Btf, vox_dtype, flags, sampling_rate, fov_deg, near_plane);
    if (msg) return fail(strstr(msg, "dtype") ? DR_EDTYPE : DR_EINVAL, msg);
    return DR_OK;
}

size_t dr_bricked_elems(const DrDesc* d) { return d ? (size_t)d->nbx * d->nby * d->nbz * 512 : 0; }

size_t dr_workspace_bytes(const DrDesc* d)
{
    if (!d || !(d->flags & DR_F_NEEDS_TF_GRAD)) return 0;
    return (size_t)d->Btf * kTfSlots * (d->R + 1) * sizeof(float4);          // R bins + one pad bin per privatised copy
}

int dr_brick_volume(const DrDesc* d, const void* vol_linear, void* vol_bricked, void* stream)
{
    if (int rc = check_desc(d)) return rc;
    if (!vol_linear || !vol_bricked) return fail(DR_EINVAL, "dr_brick_volume: null pointer");
    if (d->vox_dtype == DR_VOX_U8) return fail(DR_EDTYPE, "dr_brick_volume: uint8 volumes use the cell-major copy (dr_expand_cells)");
    const size_t elems = dr_bricked_elems(d);
    dim3 grid((unsigned)((elems + 255) / 256), d->Bvol);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d->vox_dtype == DR_VOX_F32)
        brick_kernel<float><<<grid, 256, 0, st>>>(*d, static_cast<const float*>(vol_linear), static_cast<float*>(vol_bricked), elems);
    else
        brick_kernel<__half><<<grid, 256, 0, st>>>(*d, static_cast<const __half*>(vol_linear), static_cast<__half*>(vol_bricked), elems);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DR_OK : fail_cuda(e, "brick_kernel launch");
}

int dr_expand_cells(const DrDesc* d, const void* vol_linear, void* vol_cells, void* stream)
{
    if (int rc = check_desc(d)) return rc;
    if (!vol_linear || !vol_cells) return fail(DR_EINVAL, "dr_expand_cells: null pointer");
    if ((reinterpret_cast<uintptr_t>(vol_cells) & 31) != 0) return fail(DR_EALIGN, "dr_expand_cells: vol_cells must be 32-byte aligned");
    const size_t n = (size_t)d->X * d->Y * d->Z;
    dim3 grid((unsigned)((n + 255) / 256), d->Bvol);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d->vox_dtype == DR_VOX_F32)
        expand_cells_kernel<float><<<grid, 256, 0, st>>>(*d, static_cast<const float*>(vol_linear), static_cast<float*>(vol_cells));
    else if (d->vox_dtype == DR_VOX_F16)
        expand_cells_kernel<__half><<<grid, 256, 0, st>>>(*d, static_cast<const __half*>(vol_linear), static_cast<__half*>(vol_cells));
    else
        expand_cells_kernel<u8vox><<<grid, 256, 0, st>>>(*d, static_cast<const u8vox*>(vol_linear), static_cast<u8vox*>(vol_cells));
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DR_OK : fail_cuda(e, "expand_cells_kernel launch");
}

int dr_forward(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter,
               float* out_rgba, int32_t* out_K, float* out_Tprev, void* stream)
{
    return forward_impl(d, vol, tf, cam, jitter, out_rgba, out_K, out_Tprev, nullptr, nullptr, nullptr, stream);
}

int dr_forward_mse(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter,
                   const float* target, float* out_rgba, int32_t* out_K, float* out_Tprev, float* loss_sum, void* stream)
{
    if (!target || !loss_sum) return fail(DR_EINVAL, "dr_forward_mse: target or loss_sum is null");
    return forward_impl(d, vol, tf, cam, jitter, out_rgba, out_K, out_Tprev, target, loss_sum, nullptr, stream);
}

int dr_forward_ex(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter, const float* target,
                  const uint8_t* skip_grid, float* out_rgba, int32_t* out_K, float* out_Tprev, float* loss_sum, void* stream)
{
    if (target && !loss_sum) return fail(DR_EINVAL, "dr_forward_ex: target given but loss_sum is null");
    return forward_impl(d, vol, tf, cam, jitter, out_rgba, out_K, out_Tprev, target, target ? loss_sum : nullptr, skip_grid, stream);
}

size_t dr_skip_minmax_bytes(const DrDesc* d) { return d ? (size_t)d->Bvol * skip_cells(d) * sizeof(float2) : 0; }

size_t dr_skip_grid_bytes(const DrDesc* d) { return d ? kSkipHeader + (size_t)skip_views(d) * skip_cells(d) : 0; }

int dr_build_skip_grid(const DrDesc* d, const void* vol_linear, const float* tf, void* minmax, int minmax_valid, uint8_t* skip_grid,
                       void* stream)
{
    if (int rc = check_desc(d)) return rc;
    if (!tf || !minmax || !skip_grid || (!minmax_valid && !vol_linear)) return fail(DR_EINVAL, "dr_build_skip_grid: null pointer");
    if (!aligned(minmax, 8)) return fail(DR_EALIGN, "dr_build_skip_grid: minmax must be 8-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t cells = skip_cells(d);
    if (!minmax_valid) {
        dim3 grid((unsigned)((cells * 32 + 255) / 256), d->Bvol);
        if (d->vox_dtype == DR_VOX_F32)
            skip_minmax_kernel<float><<<grid, 256, 0, st>>>(*d, static_cast<const float*>(vol_linear), static_cast<float2*>(minmax));
        else if (d->vox_dtype == DR_VOX_F16)
            skip_minmax_kernel<__half><<<grid, 256, 0, st>>>(*d, static_cast<const __half*>(vol_linear), static_cast<float2*>(minmax));
        else
            skip_minmax_kernel<u8vox><<<grid, 256, 0, st>>>(*d, static_cast<const u8vox*>(vol_linear), static_cast<float2*>(minmax));
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail_cuda(e, "skip_minmax_kernel launch");
    }
    if (!aligned(skip_grid, 4)) return fail(DR_EALIGN, "dr_build_skip_grid: skip_grid must be 4-byte aligned");
    cudaError_t em = cudaMemsetAsync(skip_grid, 0, kSkipHeader, st);
    if (em != cudaSuccess) return fail_cuda(em, "cudaMemsetAsync(skip grid header)");
    dim3 grid((unsigned)((cells + 255) / 256), skip_views(d));
    skip_classify_kernel<<<grid, 256, 0, st>>>(*d, static_cast<const float2*>(minmax), tf, skip_grid, skip_views(d));
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DR_OK : fail_cuda(e, "skip_classify_kernel launch");
}

int dr_backward(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter,
                const float* grad_out, const float* out_rgba, const int32_t* K, const float* Tprev,
                float* grad_vol_cells, float* grad_tf, void* workspace, size_t workspace_bytes, void* stream)
{
    if (d && (d->flags & DR_F_FUSED_MSE)) return fail(DR_EINVAL, "dr_backward: DR_F_FUSED_MSE is set by dr_backward_mse only");
    return backward_impl(d, vol, tf, cam, jitter, grad_out, out_rgba, K, Tprev, grad_vol_cells, grad_tf, workspace,
                         workspace_bytes, 0.0f, stream);
}

int dr_backward_mse(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter,
                    const float* target, float scale, const float* out_rgba, const int32_t* K, const float* Tprev,
                    float* grad_vol_cells, float* grad_tf, void* workspace, size_t workspace_bytes, void* stream)
{
    if (!d) return fail(DR_EINVAL, "null descriptor");
    DrDesc dd = *d;
    dd.flags |= DR_F_FUSED_MSE;
    return backward_impl(&dd, vol, tf, cam, jitter, target, out_rgba, K, Tprev, grad_vol_cells, grad_tf, workspace,
                         workspace_bytes, scale, stream);
}

int dr_backward_ex(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter, const float* grad_out,
                   const float* target, float scale, const float* scale_dev, const uint8_t* skip_grid, const float* out_rgba, const int32_t* K,
                   const float* Tprev, float* grad_vol_cells, float* grad_tf, void* workspace, size_t workspace_bytes, void* stream)
{
    if (!d) return fail(DR_EINVAL, "null descriptor");
    if ((grad_out == nullptr) == (target == nullptr)) return fail(DR_EINVAL, "dr_backward_ex: give exactly one of grad_out and target");
    if (d->flags & DR_F_FUSED_MSE) return fail(DR_EINVAL, "dr_backward_ex: DR_F_FUSED_MSE is internal (pass target instead of grad_out)");
    DrDesc dd = *d;
    if (target) dd.flags |= DR_F_FUSED_MSE;
    return backward_impl(&dd, vol, tf, cam, jitter, target ? target : grad_out, out_rgba, K, Tprev, grad_vol_cells, grad_tf, workspace,
                         workspace_bytes, target ? scale : 0.0f, stream, skip_grid, target ? scale_dev : nullptr);
}

int dr_momentum_step(float* param, const float* grad, float* momentum, size_t n, float lr, float gamma, float max_grad,
                     float lo, float hi, void* stream)
{
    if (!param || !grad || !momentum) return fail(DR_EINVAL, "dr_momentum_step: null pointer");
    if (n == 0) return DR_OK;
    momentum_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, momentum, n, lr, gamma,
                                                                                                   max_grad, lo, hi);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DR_OK : fail_cuda(e, "momentum_step_kernel launch");
}

int dr_ingest_u8(const DrDesc* d, const uint8_t* src, void* vol_linear, int swap_axes01, void* stream)
{
    if (int rc = check_desc(d)) return rc;
    if (!src || !vol_linear) return fail(DR_EINVAL, "dr_ingest_u8: null pointer");
    if (d->vox_dtype == DR_VOX_U8) return fail(DR_EDTYPE, "dr_ingest_u8 converts TO fp32 / fp16; a uint8 volume is marched as it is (DR_VOX_U8)");
    const size_t n = (size_t)d->X * d->Y * d->Z;
    const unsigned grid = (unsigned)((n + 255) / 256);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d->vox_dtype == DR_VOX_F32) ingest_u8_kernel<float><<<grid, 256, 0, st>>>(*d, src, static_cast<float*>(vol_linear), swap_axes01);
    else ingest_u8_kernel<__half><<<grid, 256, 0, st>>>(*d, src, static_cast<__half*>(vol_linear), swap_axes01);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DR_OK : fail_cuda(e, "ingest_u8_kernel launch");
}

size_t dr_grad_cells_elems(const DrDesc* d) { return d ? (size_t)d->X * d->Y * d->Z * 8 : 0; }

int dr_gather_grad(const DrDesc* d, const float* grad_vol_cells, float* grad_linear, int accumulate, void* stream)
{
    if (int rc = check_desc(d)) return rc;
    if (!grad_vol_cells || !grad_linear) return fail(DR_EINVAL, "dr_gather_grad: null pointer");
    if (!aligned(grad_vol_cells, 16)) return fail(DR_EALIGN, "dr_gather_grad: grad_vol_cells must be 16-byte aligned");
    const unsigned tiles = (unsigned)(((d->X + GM_X - 1) / GM_X) * ((d->Z + GM_Z - 1) / GM_Z) * ((d->Y + GM_CH - 1) / GM_CH));
    gather_grad_kernel<<<dim3(tiles, d->Bvol), GM_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(*d, grad_vol_cells, grad_linear, accumulate);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DR_OK : fail_cuda(e, "gather_grad_kernel launch");
}

int dr_gather_step(const DrDesc* d, const float* grad_vol_cells, const float* grad_linear, float* param, float* momentum, void* vol_cells,
                   float* grad_out, float lr, float gamma, float max_grad, float lo, float hi, void* stream)
{
    if (int rc = check_desc(d)) return rc;
    if (!param || !momentum) return fail(DR_EINVAL, "dr_gather_step: null pointer");
    if ((grad_vol_cells == nullptr) == (grad_linear == nullptr))
        return fail(DR_EINVAL, "dr_gather_step: give exactly one of grad_vol_cells (cell-major) and grad_linear (already gathered)");
    if (d->Bvol != 1) return fail(DR_EINVAL, "dr_gather_step: one volume per call (Bvol == 1)");
    if (d->vox_dtype == DR_VOX_U8) return fail(DR_EDTYPE, "dr_gather_step: the refreshed volume copy is fp32 or fp16 (a uint8 volume is not a parameter)");
    if (grad_vol_cells && !aligned(grad_vol_cells, 16)) return fail(DR_EALIGN, "dr_gather_step: grad_vol_cells must be 16-byte aligned");
    if (vol_cells && !aligned(vol_cells, 32)) return fail(DR_EALIGN, "dr_gather_step: vol_cells must be 32-byte aligned");
    const StepArgs sa { lr, gamma, max_grad, lo, hi };
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid(gather_tiles(d));
    constexpr size_t kTileSmem = (size_t)GT_X * GT_Y * GT_Z * sizeof(float);
#define DR_GS(VT, FC) do { \
        const size_t smem = (FC ? kGatherSmem : 0) + kTileSmem; \
        cudaError_t ea = cudaFuncSetAttribute(gather_step_kernel<VT, FC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (ea != cudaSuccess) return fail_cuda(ea, "cudaFuncSetAttribute(gather_step_kernel)"); \
        gather_step_kernel<VT, FC><<<grid, GT_THREADS, smem, st>>>(*d, grad_vol_cells, grad_linear, param, momentum, \
                                                                   static_cast<VT*>(vol_cells), grad_out, sa); } while (0)
    if (d->vox_dtype == DR_VOX_F32) { if (grad_vol_cells) DR_GS(float, true); else DR_GS(float, false); }
    else { if (grad_vol_cells) DR_GS(__half, true); else DR_GS(__half, false); }
#undef DR_GS
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DR_OK : fail_cuda(e, "gather_step_kernel launch");
}

long long dr_probe_l2_read(const void* buf, size_t bytes, int reps, void* sink, void* stream)
{
    if (!buf || !sink || bytes < 16 * 512 * 8 || reps < 1) return fail(DR_EINVAL, "dr_probe_l2_read: bad arguments");
    if (!aligned(buf, 16) || !aligned(sink, 4)) return fail(DR_EALIGN, "dr_probe_l2_read: buffer must be 16-byte aligned");
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    l2_read_probe_kernel<<<sms * 2, 512, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(buf), bytes / 16, reps,
                                                                                 static_cast<unsigned*>(sink));
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? (long long)(sms * 2) * reps * (long long)(bytes / 16 * 16) : fail_cuda(e, "l2_read_probe_kernel launch");
}

}  // extern "C"
