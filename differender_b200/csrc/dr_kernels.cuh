// dr_kernels.cuh -- the two march kernels (templates) and their launchers.  Included by dr_fwd_f32.cu, dr_fwd_f16.cu,
// dr_bwd_f32.cu and dr_bwd_f16.cu, which instantiate them in separate translation units so the library builds in parallel.
//
//   fwd_kernel       ray set-up + march + compositing + final image          (reference :221-372)
//   bwd_kernel       tape-free reverse march, TF + volume gradient scatter   (raycast.grad, :460-461)
//
// Thread mapping: one ray per thread; a warp is an 8x4 pixel tile, a CTA (4 warps) a 16x8 tile, so the 32 rays of
// a warp traverse neighbouring voxels and their corner fetches fall into a few 128-byte lines of the same rows/bricks.
// No tensor cores: nothing here is a dense contraction (north_star).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "diffrender.h"
#include "dr_host.h"
#include "dr_math.cuh"

namespace dr {

// CTA = DR_CTA_WARPS warps, each an 8x4 pixel tile, laid out kWarpsX x kWarpsY
#ifndef DR_CTA_WARPS
#define DR_CTA_WARPS 4
#endif
constexpr int kWarpsX = DR_CTA_WARPS >= 2 ? 2 : 1, kWarpsY = DR_CTA_WARPS / kWarpsX;
constexpr int kTileW = 8 * kWarpsX, kTileH = 4 * kWarpsY, kThreads = 32 * DR_CTA_WARPS;
// minimum resident CTAs per SM the compiler must allow (caps registers); tuned on B200 (profiles/r01_experiments.md):
// forward 5 CTAs/SM = 96 registers without spills (6 = 80 registers spills 40 bytes once the taps are branch-free);
// backward 5 CTAs/SM for the linear layout (96 registers, no spills, +3 %), 4 for brick8 (its addressing needs the
// registers: 5 loses 6 % at C5).
#ifndef DR_FWD_MIN_BLOCKS
#define DR_FWD_MIN_BLOCKS 5
#endif
#ifndef DR_FWD_MIN_BLOCKS_SKIP      // cell-major differentiable forward WITH the skip grid (it is only used while the TF leaves >= 5 % of the macro-cells
#define DR_FWD_MIN_BLOCKS_SKIP 6    // empty): mostly short centre-only samples, latency-bound -- 6 CTAs/SM (80 registers, a few spilled bytes in the shaded
#endif                              // path) gives C3 +2 %, C4 +6 %, C5 +4.5 % forward; the every-sample-shaded kernels (no grid, nondiff, sr != 1) lose 1-5 % at 6 and stay at 5
#ifndef DR_BWD_MIN_BLOCKS_LINEAR
#define DR_BWD_MIN_BLOCKS_LINEAR 5
#endif
#ifndef DR_BWD_MIN_BLOCKS_BRICK
#define DR_BWD_MIN_BLOCKS_BRICK 4
#endif
#ifndef DR_BWD_MIN_BLOCKS_TFONLY    // TF-only backward (C2): 80 registers without the volume sink
#define DR_BWD_MIN_BLOCKS_TFONLY 5
#endif
#ifndef DR_BWD_MIN_BLOCKS_SR        // sampling rate != 1 with both gradients (the powf path): spills 72 / 96 bytes at 96 registers
#define DR_BWD_MIN_BLOCKS_SR 4
#endif
#ifndef DR_BWD_MIN_BLOCKS_TWO       // two-neighbour taps (an axis of 1001..2000 voxels) with both gradients, corner-reuse path (linear layout): 114 registers, spills at 96.
                                    // (The cell-major layout uses direct taps there -- 94 registers, no spills at 5 CTAs/SM: C5 backward 60.2 -> 62.6.)
#define DR_BWD_MIN_BLOCKS_TWO 4
#endif

// skip grid: one byte per macro-cell; shared by all views when the volume and the TF are, else one grid per view
constexpr size_t kSkipHeader = 16;      // bytes before the grid: uint32 number of empty macro-cells (+ padding)
__host__ __device__ inline size_t skip_cells(const DrDesc* d) { return (size_t)macro_nx(*d) * macro_ny(*d) * macro_nz(*d); }
inline int skip_views(const DrDesc* d) { return (d->Bvol == 1 && d->Btf == 1) ? 1 : d->BS; }

// ---------------------------------------------------------------------------------------------------------
// shared prologue: stage the view's transfer function in shared memory as R TfBin entries (32 bytes each: tf[r] and
// what the lookup needs of tf[min(r+1, R-1)], see dr_math.cuh).  Returns the table accessor.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ TfTable stage_tf(const DrDesc& d, const float* __restrict__ tf, int tb, TfBin* s_tf)
{
    const float* src = tf + (size_t)tb * d.R * 4;
    const int R = d.R;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const int r1 = min(r + 1, R - 1);
        F4 a, b;
        if (d.flags & DR_F_TF_4R) {                              // [4][R]: a warp reads 32 consecutive bins of one channel
            a = F4 { __ldg(src + r), __ldg(src + R + r), __ldg(src + 2 * R + r), __ldg(src + 3 * R + r) };
            b = F4 { __ldg(src + r1), __ldg(src + R + r1), __ldg(src + 2 * R + r1), __ldg(src + 3 * R + r1) };
        } else {
            const float4 va = __ldg(reinterpret_cast<const float4*>(src) + r), vb = __ldg(reinterpret_cast<const float4*>(src) + r1);
            a = F4 { va.x, va.y, va.z, va.w };
            b = F4 { vb.x, vb.y, vb.z, vb.w };
        }
        s_tf[r] = make_tf_bin(a, b);
    }
    __syncthreads();
    TfTable t;
    t.base = (unsigned)__cvta_generic_to_shared(s_tf);
    return t;
}

// Tile rows are handed out from the image centre outwards (mid, mid-1, mid+1, ...): the hardware dispatches CTAs in blockIdx
// order, the rays through the middle of the volume are the longest, and a launch of only a few waves (C1, C2: one 512^2 view)
// ends when its last long ray does -- so the long ones start first.
__device__ __forceinline__ bool pixel_of_thread(const DrDesc& d, int& i, int& j)
{
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int k = blockIdx.y, mid = gridDim.y >> 1;
    const int ty = (k & 1) ? mid - ((k + 1) >> 1) : mid + (k >> 1);
    i = blockIdx.x * kTileW + (w % kWarpsX) * 8 + (l & 7);
    j = ty * kTileH + (w / kWarpsX) * 4 + (l >> 3);
    return i < d.W && j < d.H;
}

#if defined(DR_PERSIST)
// EXPERIMENT: persistent warps and a Morton-ordered tile queue.  One counter per translation unit, zeroed by the
// launcher before every launch (single-stream use only: this is a measurement variant, not the product path).
static __device__ unsigned dr_tile_counter;
__device__ __forceinline__ unsigned compact1by1(unsigned x)
{
    x &= 0x55555555u; x = (x | (x >> 1)) & 0x33333333u; x = (x | (x >> 2)) & 0x0f0f0f0fu; x = (x | (x >> 4)) & 0x00ff00ffu;
    return (x | (x >> 8)) & 0x0000ffffu;
}
__device__ __forceinline__ bool next_tile(const DrDesc& d, unsigned* counter, int& b, int& i, int& j, bool& valid)
{
    const int twx = (d.W + 7) >> 3, twy = (d.H + 3) >> 2;
    unsigned P = 1;
    while ((int)P < max(twx, twy)) P <<= 1;
    const unsigned per_view = P * P, total = per_view * (unsigned)d.BS;
    const int l = threadIdx.x & 31;
    for (;;) {
        unsigned t = 0;
        if (l == 0) t = atomicAdd(counter, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= total) return false;
        const unsigned m = t % per_view;
        const int wx = (int)compact1by1(m), wy = (int)compact1by1(m >> 1);
        if (wx >= twx || wy >= twy) continue;
        b = (int)(t / per_view);
        i = wx * 8 + (l & 7); j = wy * 4 + (l >> 3);
        valid = i < d.W && j < d.H;
        return true;
    }
}
#endif

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
template <typename VT, int LAYOUT, bool NONDIFF, int TAPS, bool SR1, bool SKIP>
__global__ void __launch_bounds__(kThreads, (LAYOUT == LAYOUT_CELL8 && SKIP && SR1 && !NONDIFF) ? DR_FWD_MIN_BLOCKS_SKIP : DR_FWD_MIN_BLOCKS)
fwd_kernel(DrDesc d, const VT* __restrict__ volp, const float* __restrict__ tf, const float* __restrict__ camp,
           const float* __restrict__ jitter, float* __restrict__ out, int32_t* __restrict__ outK,
           float* __restrict__ outT, size_t vol_elems, const float* __restrict__ target, float* __restrict__ loss_sum, LayoutConsts lc,
           const unsigned char* __restrict__ skip_grid, size_t skip_stride)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
#if defined(DR_PERSIST)
    // EXPERIMENT (profiles/r02_experiments.md): persistent warps pulling 8x4-pixel tiles of all views from a Morton-ordered queue
    const TfTable s_tf = stage_tf(d, tf, 0, reinterpret_cast<TfBin*>(s_raw));       // the launcher guarantees Btf == 1
    for (;;) {
    int b, i, j;
    bool valid;
    if (!next_tile(d, &dr_tile_counter, b, i, j, valid)) break;
    if (!valid && !target) continue;
#else
    const int b = blockIdx.z;
    const TfTable s_tf = stage_tf(d, tf, d.Btf == 1 ? 0 : b, reinterpret_cast<TfBin*>(s_raw));
    int i, j;
    const bool valid = pixel_of_thread(d, i, j);
    if (!valid && !target) return;
#endif
    float sq = 0.0f;
    if (valid) {
    const size_t pix = (size_t)b * d.W * d.H + (size_t)(d.H - 1 - j) * d.W + i;      // image orientation
    const F3 cam = { __ldg(camp + 3 * b), __ldg(camp + 3 * b + 1), __ldg(camp + 3 * b + 2) };
    const float jit = (d.flags & DR_F_HAS_JITTER) ? __ldg(jitter + pix) : 0.0f;
    Ray r;
    setup_ray(d, cam, i, j, jit, r);
    const VolView<VT> vol { volp + (d.Bvol == 1 ? 0 : (size_t)b * vol_elems) };
    const Layout L = make_layout(d, lc, SKIP && !NONDIFF);
    F4 A; int K; float Tp;
    // skip grid: a 16-byte header (the number of empty macro-cells: nothing to skip -> march as if there were no grid) + the bytes
    const unsigned char* grid = nullptr;
    if (SKIP && __ldg(reinterpret_cast<const unsigned*>(skip_grid)) != 0u) grid = skip_grid + kSkipHeader + (size_t)b * skip_stride;
    march_forward<VT, LAYOUT, NONDIFF, TAPS, SR1, SKIP>(d, vol, L, s_tf, cam, r, A, K, Tp, grid);
    if (d.flags & DR_F_OUT_IMAGE) {
        const size_t plane = (size_t)d.W * d.H;
        const size_t o0 = (size_t)b * 4 * plane + (size_t)(d.H - 1 - j) * d.W + i;
        float* o = out + o0;
        o[0] = A.x; o[plane] = A.y; o[2 * plane] = A.z; o[3 * plane] = A.w;
        if (target) {
            const float ex = A.x - __ldg(target + o0), ey = A.y - __ldg(target + o0 + plane);
            const float ez = A.z - __ldg(target + o0 + 2 * plane), ew = A.w - __ldg(target + o0 + 3 * plane);
            sq = ex * ex + ey * ey + ez * ez + ew * ew;
        }
    } else {
        const size_t o0 = ((size_t)b * d.W + i) * d.H + j;
        reinterpret_cast<float4*>(out)[o0] = make_float4(A.x, A.y, A.z, A.w);
        if (target) {
            const float4 tg = __ldg(reinterpret_cast<const float4*>(target) + o0);
            const float ex = A.x - tg.x, ey = A.y - tg.y, ez = A.z - tg.z, ew = A.w - tg.w;
            sq = ex * ex + ey * ey + ez * ez + ew * ew;
        }
    }
    if (outK) outK[pix] = K;
    if (outT) outT[pix] = Tp;
    }
    if (target) {
        // fused loss (reference examples: torch mse_loss on output_rgba): one atomic per warp into this view's sum
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_down_sync(0xffffffffu, sq, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(loss_sum + b, sq);
    }
#if defined(DR_PERSIST)
    }
#endif
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
// Volume gradient: one cell = one 32-byte sector = two RED.E.ADD.F32x4.  The centre cell's 8-vector stays in registers
// while consecutive samples of the ray fall into the same cell (about 3 samples per cell at sampling rate 1); the
// per-sample weights are accumulated into it with FMAs (dr_math.cuh scatter_volume_grad).
struct CellVolSink {
    float4* g;
    int cur;
    float acc[8];
#if defined(DR_BOUNDS_CHECK)
    long long n_cells;
#endif
    __device__ __forceinline__ void red(int cell, const float* v)
    {
#if defined(DR_BOUNDS_CHECK)
        DR_OOB_IF(cell < 0 || cell >= n_cells);
#endif
        float4* p = g + (size_t)cell * 2;
        atomicAdd(p, make_float4(v[0], v[1], v[2], v[3]));
        atomicAdd(p + 1, make_float4(v[4], v[5], v[6], v[7]));
    }
    __device__ __forceinline__ float* open(int cell)
    {
        if (cell != cur) {
            if (cur >= 0) red(cur, acc);
            cur = cell;
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = 0.0f;
        }
        return acc;
    }
    __device__ __forceinline__ void close() {}
    __device__ __forceinline__ void direct(int cell, const float* v) { red(cell, v); }
    __device__ __forceinline__ void flush() { if (cur >= 0) red(cur, acc); }
};
// TF gradient: two 16-byte vector reductions (bins lo, lo+1) into one of kTfSlots privatised copies of the table, summed by
// tf_reduce_kernel.  Shared-memory fp32 atomicAdd is a CAS spin loop on sm_100a (ATOMS.CAST.SPIN) and measured 4x slower
// than RED.F32x4 into L2 (profiles/r01_atomic_microbench.txt), so the privatised copies live in L2, not in shared memory.
// While consecutive samples of the ray hit the same bin, S = sum dc and S1 = sum f*dc stay in registers (8 FP ops per
// sample); the bin pair receives (S - S1, S1) when the bin changes.
struct RedTfSink {
    float4* g;      // [R] of this CTA's slot
    int cur;
    float4 s, s1;
    bool rgb;       // a shaded sample went into the current bin; otherwise the colour sums are exactly zero and stay home
    __device__ __forceinline__ void flush()
    {
        if (cur >= 0) {
            float4* p = g + cur;
            float4* q = p + 1;                           // bin R is a pad bin that tf_reduce_kernel folds into bin R - 1 (the reference clamps the upper bin, :216-218)
            if (rgb) {
                atomicAdd(p, make_float4(s.x - s1.x, s.y - s1.y, s.z - s1.z, s.w - s1.w));
                atomicAdd(q, s1);
            } else {                                     // a run of transparent samples only moves alpha: two scalar REDs
                atomicAdd(&p->w, s.w - s1.w);
                atomicAdd(&q->w, s1.w);
            }
        }
    }
    __device__ __forceinline__ void next(int lo)
    {
        flush();
        cur = lo;
        s.w = 0.f; s1.w = 0.f;
        if (rgb) { s.x = s.y = s.z = 0.f; s1.x = s1.y = s1.z = 0.f; rgb = false; }
    }
    __device__ __forceinline__ void add(int lo, float f, F4 dc)
    {
        if (lo != cur) next(lo);
        rgb = true;
        s.x += dc.x; s.y += dc.y; s.z += dc.z; s.w += dc.w;
        s1.x += f * dc.x; s1.y += f * dc.y; s1.z += f * dc.z; s1.w += f * dc.w;
    }
    __device__ __forceinline__ void add_alpha(int lo, float f, float dcw)       // dc = (0, 0, 0, dcw)
    {
        if (lo != cur) next(lo);
        s.w += dcw; s1.w += f * dcw;
    }
    __device__ __forceinline__ void init() { cur = -1; rgb = false; s = make_float4(0.f, 0.f, 0.f, 0.f); s1 = s; }
};

template <typename VT, int LAYOUT, int TAPS, bool WANT_VOL, bool WANT_TF, bool SR1, bool SKIP>
__global__ void __launch_bounds__(kThreads, LAYOUT == LAYOUT_BRICK8 ? DR_BWD_MIN_BLOCKS_BRICK
                                              : (TAPS == TAPS_TWO && WANT_VOL && WANT_TF && LAYOUT != LAYOUT_CELL8) ? DR_BWD_MIN_BLOCKS_TWO
                                              : (!SR1 && WANT_VOL && WANT_TF) ? DR_BWD_MIN_BLOCKS_SR
                                              : (SR1 && !WANT_VOL && LAYOUT == LAYOUT_CELL8 && TAPS == TAPS_ONE) ? DR_BWD_MIN_BLOCKS_TFONLY : DR_BWD_MIN_BLOCKS_LINEAR)
bwd_kernel(DrDesc d, const VT* __restrict__ volp, const float* __restrict__ tf, const float* __restrict__ camp,
           const float* __restrict__ jitter, const float* __restrict__ gout, const float* __restrict__ outp,
           const int32_t* __restrict__ Kp, const float* __restrict__ Tp, float4* __restrict__ gcell,
           float4* __restrict__ tf_slots, size_t vol_elems, float mse_scale, LayoutConsts lc,
           const unsigned char* __restrict__ skip_grid, size_t skip_stride, const float* __restrict__ scale_dev)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
#if defined(DR_PERSIST)
    const int tb = 0;
    const TfTable s_tf = stage_tf(d, tf, 0, reinterpret_cast<TfBin*>(s_raw));
    for (;;) {
    int b, i, j;
    bool valid;
    if (!next_tile(d, &dr_tile_counter, b, i, j, valid)) break;
    if (!valid) continue;
    const size_t pix = (size_t)b * d.W * d.H + (size_t)(d.H - 1 - j) * d.W + i;
    const int K = __ldg(Kp + pix);
    if (K <= 0) continue;
#define DR_RAY_DONE continue
#else
    const int b = blockIdx.z;
    const int tb = d.Btf == 1 ? 0 : b;
    const TfTable s_tf = stage_tf(d, tf, tb, reinterpret_cast<TfBin*>(s_raw));
    int i, j;
    if (!pixel_of_thread(d, i, j)) return;
    const size_t pix = (size_t)b * d.W * d.H + (size_t)(d.H - 1 - j) * d.W + i;
    const int K = __ldg(Kp + pix);
    if (K <= 0) return;
#define DR_RAY_DONE return
#endif
    const F3 cam = { __ldg(camp + 3 * b), __ldg(camp + 3 * b + 1), __ldg(camp + 3 * b + 2) };
    const float jit = (d.flags & DR_F_HAS_JITTER) ? __ldg(jitter + pix) : 0.0f;
    Ray r;
    setup_ray(d, cam, i, j, jit, r);
    F4 A, g;
    if (d.flags & DR_F_OUT_IMAGE) {
        const size_t plane = (size_t)d.W * d.H;
        const size_t o = (size_t)b * 4 * plane + (size_t)(d.H - 1 - j) * d.W + i;
        A = F4 { __ldg(outp + o), __ldg(outp + o + plane), __ldg(outp + o + 2 * plane), __ldg(outp + o + 3 * plane) };
        g = F4 { __ldg(gout + o), __ldg(gout + o + plane), __ldg(gout + o + 2 * plane), __ldg(gout + o + 3 * plane) };
    } else {
        const size_t o = ((size_t)b * d.W + i) * d.H + j;
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(outp) + o);
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gout) + o);
        A = F4 { a4.x, a4.y, a4.z, a4.w };
        g = F4 { g4.x, g4.y, g4.z, g4.w };
    }
    if (d.flags & DR_F_FUSED_MSE) {
        // `gout` holds the TARGET image: dL/dA = mse_scale * (A - target), never materialised in HBM.  The upstream gradient of the
        // loss may stay on the device (scale_dev): reading it on the host would stall the host on the whole forward
        const float ms = scale_dev ? mse_scale * __ldg(scale_dev) : mse_scale;
        g.x = ms * (A.x - g.x); g.y = ms * (A.y - g.y); g.z = ms * (A.z - g.z); g.w = ms * (A.w - g.w);
    }
    if (g.x == 0.0f && g.y == 0.0f && g.z == 0.0f && g.w == 0.0f) DR_RAY_DONE;       // this ray's gradient is exactly zero
    const size_t voff = d.Bvol == 1 ? 0 : (size_t)b * vol_elems;
    const VolView<VT> vol { volp + voff };
    const Layout L = make_layout(d, lc, WANT_VOL);
    CellVolSink vs;
    vs.g = WANT_VOL ? gcell + (d.Bvol == 1 ? 0 : (size_t)b * d.X * d.Y * d.Z * 2) : nullptr;
    vs.cur = -1;
#if defined(DR_BOUNDS_CHECK)
    vs.n_cells = (long long)d.X * d.Y * d.Z;
#endif
    const int slot = (blockIdx.y * gridDim.x + blockIdx.x) & (kTfSlots - 1);      // (persistent launch: blockIdx.y == 0)
    RedTfSink ts;
    ts.g = WANT_TF ? tf_slots + ((size_t)tb * kTfSlots + slot) * (d.R + 1) : nullptr;
    ts.init();
    const unsigned char* grid = nullptr;
    if (SKIP && __ldg(reinterpret_cast<const unsigned*>(skip_grid)) != 0u) grid = skip_grid + kSkipHeader + (size_t)b * skip_stride;
    march_backward<VT, LAYOUT, TAPS, WANT_VOL, WANT_TF, SR1, SKIP>(d, vol, L, s_tf, cam, r, A, K, __ldg(Tp + pix), g, vs, ts, grid);
#if defined(DR_PERSIST)
    }
#endif
#undef DR_RAY_DONE
}

// ---------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------
template <typename K> int set_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024) {
        if (bytes > kMaxTfSmem) return fail(DR_EINVAL, "tf resolution too large for shared memory staging (R <= 6400)");
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute");
    }
    return DR_OK;
}

#if defined(DR_PERSIST)
template <typename K> int persistent_grid(K kernel, size_t smem, const DrDesc* d, dim3& grid, cudaStream_t st)
{
    if (d->Btf != 1) return fail(DR_EINVAL, "DR_PERSIST experiment: shared transfer function only");
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem);
    void* ctr = nullptr;
    cudaGetSymbolAddress(&ctr, dr_tile_counter);
    cudaMemsetAsync(ctr, 0, sizeof(unsigned), st);
    const unsigned tiles = grid.x * grid.y * grid.z;
    grid = dim3(min((unsigned)(sms * max(per_sm, 1)), tiles), 1, 1);
    return DR_OK;
}
#endif

inline size_t vol_stride(const DrDesc* d)
{
    if (d->flags & DR_F_LAYOUT_CELL8) return (size_t)d->X * d->Y * d->Z * 8;
    return (d->flags & DR_F_LAYOUT_BRICK8) ? (size_t)d->nbx * d->nby * d->nbz * 512 : (size_t)d->X * d->Y * d->Z;
}

template <typename VT, int LAYOUT, bool NONDIFF, int TAPS, bool SR1, bool SKIP>
int launch_fwd_skip(const FwdArgs& a)
{
    const DrDesc* d = a.d;
    const size_t smem = (size_t)d->R * sizeof(TfBin);
    auto kern = fwd_kernel<VT, LAYOUT, NONDIFF, TAPS, SR1, SKIP>;
    if (int rc = set_smem(kern, smem)) return rc;
    dim3 grid((d->W + kTileW - 1) / kTileW, (d->H + kTileH - 1) / kTileH, d->BS);
#if defined(DR_PERSIST)
    if (int rc = persistent_grid(kern, smem, d, grid, a.st)) return rc;
#endif
    kern<<<grid, kThreads, smem, a.st>>>(*d, static_cast<const VT*>(a.vol), a.tf, a.cam, a.jitter, a.out, a.K, a.T, vol_stride(d),
                                         a.target, a.loss_sum, layout_consts(*d), a.skip_grid, skip_views(d) == 1 ? 0 : skip_cells(d));
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DR_OK : fail_cuda(e, "fwd_kernel launch");
}

template <typename VT, int LAYOUT, int TAPS, bool WV, bool WT, bool SR1, bool SKIP>
int launch_bwd_skip(const BwdArgs& a)
{
    const DrDesc* d = a.d;
    const size_t smem = (size_t)d->R * sizeof(TfBin);
    auto kern = bwd_kernel<VT, LAYOUT, TAPS, WV, WT, SR1, SKIP>;
    if (int rc = set_smem(kern, smem)) return rc;
    dim3 grid((d->W + kTileW - 1) / kTileW, (d->H + kTileH - 1) / kTileH, d->BS);
#if defined(DR_PERSIST)
    if (int rc = persistent_grid(kern, smem, d, grid, a.st)) return rc;
#endif
    kern<<<grid, kThreads, smem, a.st>>>(*d, static_cast<const VT*>(a.vol), a.tf, a.cam, a.jitter, a.gout, a.out, a.K, a.T, a.gvol,
                                         a.slots, vol_stride(d), a.mse_scale, layout_consts(*d), a.skip_grid, skip_views(d) == 1 ? 0 : skip_cells(d),
                                         a.scale_dev);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DR_OK : fail_cuda(e, "bwd_kernel launch");
}
// the skip grid only serves the volume-only backward (no TF gradient) of the corner-reuse tap paths
template <typename VT, int LAYOUT, int TAPS, bool WV, bool WT, bool SR1>
int launch_bwd(const BwdArgs& a)
{
    constexpr bool kCanSkip = WV && !WT && TAPS != TAPS_GENERIC;
    if (kCanSkip && a.skip_grid) return launch_bwd_skip<VT, LAYOUT, TAPS, WV, WT, SR1, kCanSkip>(a);
    return launch_bwd_skip<VT, LAYOUT, TAPS, WV, WT, SR1, false>(a);
}

// with or without the skip grid (the kernels without it carry none of the skip code); the generic tap path never skips
template <typename VT, int LAYOUT, bool NONDIFF, int TAPS, bool SR1>
int launch_fwd(const FwdArgs& a)
{
    constexpr bool kCanSkip = TAPS != TAPS_GENERIC;
    if (kCanSkip && a.skip_grid) return launch_fwd_skip<VT, LAYOUT, NONDIFF, TAPS, SR1, kCanSkip>(a);
    return launch_fwd_skip<VT, LAYOUT, NONDIFF, TAPS, SR1, false>(a);
}

// WANT_VOL / WANT_TF / sampling-rate dispatch of one (voxel type, layout, tap mode).  The SR1 = false kernels are correct
// for any sampling rate (powf(x, 1) == x); the generic-tap path (volumes > ~2000 voxels per axis) only has those.
template <typename VT, int LAYOUT, int TAPS>
int dispatch_bwd(const BwdArgs& a)
{
    const bool wv = a.d->flags & DR_F_NEEDS_VOL_GRAD, wt = a.d->flags & DR_F_NEEDS_TF_GRAD;
    constexpr bool kHasSr1 = TAPS != TAPS_GENERIC;
    const bool sr1 = kHasSr1 && a.d->inv_sr == 1.0f;
#define DR_BWD(WV, WT) (sr1 ? launch_bwd<VT, LAYOUT, TAPS, WV, WT, kHasSr1>(a) : launch_bwd<VT, LAYOUT, TAPS, WV, WT, false>(a))
    if (wv && wt) return DR_BWD(true, true);
    if (wv) return DR_BWD(true, false);
    return DR_BWD(false, true);
#undef DR_BWD
}

// uint8 volumes are marched from their cell-major copy only (8-byte records)
template <typename VT>
int dispatch_bwd_cell8(const BwdArgs& a)
{
    return tap_mode(*a.d) == TAPS_ONE ? dispatch_bwd<VT, LAYOUT_CELL8, TAPS_ONE>(a) : dispatch_bwd<VT, LAYOUT_CELL8, TAPS_TWO>(a);
}

template <typename VT>
int dispatch_bwd_layout(const BwdArgs& a)
{
    const int taps = tap_mode(*a.d);
    if (a.d->flags & DR_F_LAYOUT_CELL8) return taps == TAPS_ONE ? dispatch_bwd<VT, LAYOUT_CELL8, TAPS_ONE>(a) : dispatch_bwd<VT, LAYOUT_CELL8, TAPS_TWO>(a);
    if (a.d->flags & DR_F_LAYOUT_BRICK8) return taps == TAPS_ONE ? dispatch_bwd<VT, LAYOUT_BRICK8, TAPS_ONE>(a) : dispatch_bwd<VT, LAYOUT_BRICK8, TAPS_TWO>(a);
    if (taps == TAPS_GENERIC) return dispatch_bwd<VT, LAYOUT_LINEAR, TAPS_GENERIC>(a);
    return taps == TAPS_ONE ? dispatch_bwd<VT, LAYOUT_LINEAR, TAPS_ONE>(a) : dispatch_bwd<VT, LAYOUT_LINEAR, TAPS_TWO>(a);
}

// The SR1 = false kernels are correct for any sampling rate (powf(x, 1) == x); the generic-tap path (volumes > ~2000 voxels
// per axis, linear layout only) only has those.
template <typename VT>
int forward_vt(const FwdArgs& a)
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

template <typename VT>
int forward_cell8(const FwdArgs& a)
{
    const DrDesc* d = a.d;
    const bool nd = d->flags & DR_F_NONDIFF, sr1 = d->inv_sr == 1.0f;
#define DR_FWD_ND(TAPS, SR1) (nd ? launch_fwd<VT, LAYOUT_CELL8, true, TAPS, SR1>(a) : launch_fwd<VT, LAYOUT_CELL8, false, TAPS, SR1>(a))
#define DR_FWD_SR(TAPS) (sr1 ? DR_FWD_ND(TAPS, true) : DR_FWD_ND(TAPS, false))
    return tap_mode(*d) == TAPS_ONE ? DR_FWD_SR(TAPS_ONE) : DR_FWD_SR(TAPS_TWO);
#undef DR_FWD_SR
#undef DR_FWD_ND
}


}  // namespace dr
