// diffrender.cu -- sm_100a kernels and the C ABI (include/diffrender.h) of the differentiable ray-march.
//
// Kernels (DESIGN.md has the roofline of each):
//   brick_kernel     linear [Y][Z][X] volume -> 8x8x8 bricks                 (set_volume, reference :118-119)
//   fwd_kernel       ray set-up + march + compositing + final image          (:221-372)
//   bwd_kernel       tape-free reverse march, TF + volume gradient scatter   (raycast.grad, :460-461)
//   tf_reduce_kernel sums the privatised TF-gradient copies                  (tf_tex.grad.to_torch, :464,475)
//   gather_grad_kernel cell-major fp32 gradient -> linear, nan_to_num        (volume.grad.to_torch, :463,474)
//
// Thread mapping: one ray per thread; a warp is an 8x4 pixel tile, a CTA (4 warps) a 16x8 tile, so the 32 rays of
// a warp traverse neighbouring voxels and their corner fetches fall into a few 128-byte lines of the same bricks.
// No tensor cores: nothing here is a dense contraction (north_star).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "diffrender.h"
#include "dr_desc.h"
#include "dr_math.cuh"

#if defined(DR_BOUNDS_CHECK)
__device__ unsigned long long dr_oob_counter = 0ULL;
#endif

using namespace dr;

namespace {

// CTA = DR_CTA_WARPS warps, each an 8x4 pixel tile, laid out kWarpsX x kWarpsY
#ifndef DR_CTA_WARPS
#define DR_CTA_WARPS 4
#endif
constexpr int kWarpsX = DR_CTA_WARPS >= 2 ? 2 : 1, kWarpsY = DR_CTA_WARPS / kWarpsX;
constexpr int kTileW = 8 * kWarpsX, kTileH = 4 * kWarpsY, kThreads = 32 * DR_CTA_WARPS;
// minimum resident CTAs per SM the compiler must allow (caps registers); tuned on B200, see DESIGN.md
// (B200, C3: forward 6 CTAs/SM = 80 regs, no spills: +2 %; backward 5 CTAs/SM = 96 regs spills and loses 5 %, so 4.)
#ifndef DR_FWD_MIN_BLOCKS
#define DR_FWD_MIN_BLOCKS 6
#endif
#ifndef DR_BWD_MIN_BLOCKS
#define DR_BWD_MIN_BLOCKS 4
#endif
constexpr int kTfSlots = 1024;          // privatised TF-gradient copies (power of two)

thread_local char g_err[256] = "";

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

__host__ __device__ inline Layout make_layout(const DrDesc& d)
{
    Layout L;
    L.sY = d.nbx * 512; L.sZ = d.nbx * d.nby * 512;
    L.mx = d.X - 1; L.my = d.Y - 1; L.mz = d.Z - 1;
    return L;
}

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

// cell-major gradient [cell][8] -> linear [Y][Z][X] fp32 (HBM-bound: reads 32 B per voxel once, L2 serves the 8x reuse)
__global__ void __launch_bounds__(256) gather_grad_kernel(DrDesc d, const float* __restrict__ gcell, float* __restrict__ lin, int accumulate)
{
    const size_t n = (size_t)d.X * d.Y * d.Z;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int b = blockIdx.y;
    const int x = (int)(e % d.X);
    const size_t r = e / d.X;
    const int z = (int)(r % d.Z), y = (int)(r / d.Z);
    float v = gather_voxel(d, gcell + (size_t)b * n * 8, x, y, z);
    // torch.nan_to_num (:463, :474)
    if (v != v) v = 0.0f;
    else if (v > 3.4028234663852886e38f) v = 3.4028234663852886e38f;
    else if (v < -3.4028234663852886e38f) v = -3.4028234663852886e38f;
    float* o = lin + (size_t)b * n + e;
    *o = accumulate ? (*o + v) : v;
}

// ---------------------------------------------------------------------------------------------------------
// shared prologue: stage the view's transfer function in shared memory as R float4 (RGBA) entries
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_tf(const DrDesc& d, const float* __restrict__ tf, int tb, F4* s_tf)
{
    const float* src = tf + (size_t)tb * d.R * 4;
    if (d.flags & DR_F_TF_4R) {
        for (int e = threadIdx.x; e < d.R * 4; e += blockDim.x) {
            const int c = e / d.R, r = e - c * d.R;            // coalesced read of [4][R]
            reinterpret_cast<float*>(s_tf)[r * 4 + c] = __ldg(src + e);
        }
    } else {
        for (int e = threadIdx.x; e < d.R; e += blockDim.x) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src) + e);
            s_tf[e] = F4 { v.x, v.y, v.z, v.w };
        }
    }
    __syncthreads();
}

__device__ __forceinline__ bool pixel_of_thread(const DrDesc& d, int& i, int& j)
{
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    i = blockIdx.x * kTileW + (w % kWarpsX) * 8 + (l & 7);
    j = blockIdx.y * kTileH + (w / kWarpsX) * 4 + (l >> 3);
    return i < d.W && j < d.H;
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
template <typename VT, int LAYOUT, bool NONDIFF, bool GENERIC>
__global__ void __launch_bounds__(kThreads, DR_FWD_MIN_BLOCKS)
fwd_kernel(DrDesc d, const VT* __restrict__ volp, const float* __restrict__ tf, const float* __restrict__ camp,
           const float* __restrict__ jitter, float* __restrict__ out, int32_t* __restrict__ outK,
           float* __restrict__ outT, size_t vol_elems, const float* __restrict__ target, float* __restrict__ loss_sum)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    F4* s_tf = reinterpret_cast<F4*>(s_raw);
    const int b = blockIdx.z;
    stage_tf(d, tf, d.Btf == 1 ? 0 : b, s_tf);
    int i, j;
    const bool valid = pixel_of_thread(d, i, j);
    if (!valid && !target) return;
    float sq = 0.0f;
    if (valid) {
    const size_t pix = (size_t)b * d.W * d.H + (size_t)(d.H - 1 - j) * d.W + i;      // image orientation
    const F3 cam = { __ldg(camp + 3 * b), __ldg(camp + 3 * b + 1), __ldg(camp + 3 * b + 2) };
    const float jit = (d.flags & DR_F_HAS_JITTER) ? __ldg(jitter + pix) : 0.0f;
    Ray r;
    setup_ray(d, cam, i, j, jit, r);
    const VolView<VT> vol { volp + (d.Bvol == 1 ? 0 : (size_t)b * vol_elems) };
    const Layout L = make_layout(d);
    F4 A; int K; float Tp;
    march_forward<VT, LAYOUT, NONDIFF, GENERIC>(d, vol, L, s_tf, cam, r, A, K, Tp);
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
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
// Volume gradient: one cell = one 32-byte sector = two RED.E.ADD.F32x4.  With ACCUM the centre cell's 8-vector stays in
// registers while consecutive samples of the ray fall into the same cell (about 3 samples per cell at sampling rate 1).
template <bool ACCUM> struct CellVolSink {
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
    __device__ __forceinline__ void centre(int cell, const float* v)
    {
        if (!ACCUM) { red(cell, v); return; }
        if (cell == cur) {
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] += v[q];
        } else {
            if (cur >= 0) red(cur, acc);
            cur = cell;
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = v[q];
        }
    }
    __device__ __forceinline__ void direct(int cell, const float* v) { red(cell, v); }
    __device__ __forceinline__ void flush() { if (ACCUM && cur >= 0) red(cur, acc); }
};
// TF gradient: two 16-byte vector reductions (bins lo, hi) into one of kTfSlots privatised copies of the table, summed by
// tf_reduce_kernel.  Shared-memory fp32 atomicAdd is a CAS spin loop on sm_100a (ATOMS.CAST.SPIN) and measured 4x slower
// than RED.F32x4 into L2 (profiles/r01_atomic_microbench.txt), so the privatised copies live in L2, not in shared memory.
// With ACCUM the two bins stay in registers while consecutive samples of the ray hit the same bin.
template <bool ACCUM> struct RedTfSink {
    float4* g;      // [R] of this CTA's slot
    int cur, cur_hi;
    float4 a0, a1;
    __device__ __forceinline__ void add(int lo, int hi, float f, F4 dc)
    {
        const float w0 = 1.0f - f;
        const float4 v0 = make_float4(dc.x * w0, dc.y * w0, dc.z * w0, dc.w * w0);
        const float4 v1 = make_float4(dc.x * f, dc.y * f, dc.z * f, dc.w * f);
        if (!ACCUM) { atomicAdd(g + lo, v0); atomicAdd(g + hi, v1); return; }
        if (lo == cur) {
            a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
            a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
        } else {
            if (cur >= 0) { atomicAdd(g + cur, a0); atomicAdd(g + cur_hi, a1); }
            cur = lo; cur_hi = hi; a0 = v0; a1 = v1;
        }
    }
    __device__ __forceinline__ void flush() { if (ACCUM && cur >= 0) { atomicAdd(g + cur, a0); atomicAdd(g + cur_hi, a1); } }
};

template <typename VT, int LAYOUT, bool GENERIC, bool WANT_VOL, bool WANT_TF, bool ACCUM>
__global__ void __launch_bounds__(kThreads, DR_BWD_MIN_BLOCKS)
bwd_kernel(DrDesc d, const VT* __restrict__ volp, const float* __restrict__ tf, const float* __restrict__ camp,
           const float* __restrict__ jitter, const float* __restrict__ gout, const float* __restrict__ outp,
           const int32_t* __restrict__ Kp, const float* __restrict__ Tp, float4* __restrict__ gcell,
           float4* __restrict__ tf_slots, size_t vol_elems, float mse_scale)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    F4* s_tf = reinterpret_cast<F4*>(s_raw);
    const int b = blockIdx.z;
    const int tb = d.Btf == 1 ? 0 : b;
    stage_tf(d, tf, tb, s_tf);
    int i, j;
    if (!pixel_of_thread(d, i, j)) return;
    const size_t pix = (size_t)b * d.W * d.H + (size_t)(d.H - 1 - j) * d.W + i;
    const int K = __ldg(Kp + pix);
    if (K <= 0) return;
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
        // `gout` holds the TARGET image: dL/dA = mse_scale * (A - target), never materialised in HBM
        g.x = mse_scale * (A.x - g.x); g.y = mse_scale * (A.y - g.y); g.z = mse_scale * (A.z - g.z); g.w = mse_scale * (A.w - g.w);
    }
    if (g.x == 0.0f && g.y == 0.0f && g.z == 0.0f && g.w == 0.0f) return;       // this ray's gradient is exactly zero
    const size_t voff = d.Bvol == 1 ? 0 : (size_t)b * vol_elems;
    const VolView<VT> vol { volp + voff };
    const Layout L = make_layout(d);
    CellVolSink<ACCUM> vs;
    vs.g = WANT_VOL ? gcell + (d.Bvol == 1 ? 0 : (size_t)b * d.X * d.Y * d.Z * 2) : nullptr;
    vs.cur = -1;
#if defined(DR_BOUNDS_CHECK)
    vs.n_cells = (long long)d.X * d.Y * d.Z;
#endif
    const int slot = (blockIdx.y * gridDim.x + blockIdx.x) & (kTfSlots - 1);
    RedTfSink<ACCUM> ts;
    ts.g = WANT_TF ? tf_slots + ((size_t)tb * kTfSlots + slot) * d.R : nullptr;
    ts.cur = -1;
    march_backward<VT, LAYOUT, GENERIC, WANT_VOL, WANT_TF>(d, vol, L, s_tf, cam, r, A, K, __ldg(Tp + pix), g, vs, ts);
}

// sums the kTfSlots privatised copies; one thread per (tf, bin, channel); adds into grad_tf in the caller's layout
__global__ void __launch_bounds__(256) tf_reduce_kernel(DrDesc d, const float* __restrict__ slots, float* __restrict__ grad_tf)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;       // r*4 + c
    const int tb = blockIdx.y;
    if (e >= d.R * 4) return;
    const float* p = slots + (size_t)tb * kTfSlots * d.R * 4 + e;
    float acc[4] = { 0.f, 0.f, 0.f, 0.f };
    for (int s = 0; s < kTfSlots; s += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += __ldg(p + (size_t)(s + u) * d.R * 4);
    }
    float v = (acc[0] + acc[1]) + (acc[2] + acc[3]);
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
    if (d->vox_dtype != DR_VOX_F32 && d->vox_dtype != DR_VOX_F16) return fail(DR_EDTYPE, "unsupported voxel dtype");
    if (d->BS > 65535) return fail(DR_EINVAL, "more than 65535 views in one call");
    if (d->tap_generic && (d->flags & DR_F_LAYOUT_BRICK8))
        return fail(DR_EINVAL, "the generic tap path (volumes > ~2000 voxels per axis) needs the linear layout");
    return DR_OK;
}

bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

template <typename K> int set_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024) {
        if (bytes > 200 * 1024) return fail(DR_EINVAL, "tf resolution too large for shared memory staging");
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute");
    }
    return DR_OK;
}

size_t vol_stride(const DrDesc* d)
{
    return (d->flags & DR_F_LAYOUT_BRICK8) ? dr_bricked_elems(d) : (size_t)d->X * d->Y * d->Z;
}

template <typename VT, int LAYOUT, bool NONDIFF, bool GENERIC>
int launch_fwd(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter, float* out,
               int32_t* K, float* T, cudaStream_t st, const float* target, float* loss_sum)
{
    const size_t smem = (size_t)d->R * sizeof(F4);
    auto kern = fwd_kernel<VT, LAYOUT, NONDIFF, GENERIC>;
    if (int rc = set_smem(kern, smem)) return rc;
    dim3 grid((d->W + kTileW - 1) / kTileW, (d->H + kTileH - 1) / kTileH, d->BS);
    kern<<<grid, kThreads, smem, st>>>(*d, static_cast<const VT*>(vol), tf, cam, jitter, out, K, T, vol_stride(d), target, loss_sum);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DR_OK : fail_cuda(e, "fwd_kernel launch");
}

template <typename VT, int LAYOUT, bool GENERIC, bool WV, bool WT, bool ACC>
int launch_bwd(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter,
               const float* gout, const float* out, const int32_t* K, const float* T, float4* gvol, float4* slots,
               cudaStream_t st, float mse_scale)
{
    const size_t smem = (size_t)d->R * sizeof(F4);
    auto kern = bwd_kernel<VT, LAYOUT, GENERIC, WV, WT, ACC>;
    if (int rc = set_smem(kern, smem)) return rc;
    dim3 grid((d->W + kTileW - 1) / kTileW, (d->H + kTileH - 1) / kTileH, d->BS);
    kern<<<grid, kThreads, smem, st>>>(*d, static_cast<const VT*>(vol), tf, cam, jitter, gout, out, K, T, gvol, slots,
                                       vol_stride(d), mse_scale);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DR_OK : fail_cuda(e, "bwd_kernel launch");
}

template <typename VT, int LAYOUT, bool GENERIC>
int dispatch_bwd(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter,
                 const float* gout, const float* out, const int32_t* K, const float* T, float4* gvol, float4* slots,
                 cudaStream_t st, float mse_scale)
{
    const bool wv = d->flags & DR_F_NEEDS_VOL_GRAD, wt = d->flags & DR_F_NEEDS_TF_GRAD;
    const bool acc = !(d->flags & DR_F_NO_REG_ACCUM);
#define DR_BWD(WV, WT)                                                                                              \
    (acc ? launch_bwd<VT, LAYOUT, GENERIC, WV, WT, true>(d, vol, tf, cam, jitter, gout, out, K, T, gvol, slots, st, mse_scale) \
         : launch_bwd<VT, LAYOUT, GENERIC, WV, WT, false>(d, vol, tf, cam, jitter, gout, out, K, T, gvol, slots, st, mse_scale))
    if (wv && wt) return DR_BWD(true, true);
    if (wv) return DR_BWD(true, false);
    return DR_BWD(false, true);
#undef DR_BWD
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
                 int32_t* out_K, float* out_Tprev, const float* target, float* loss_sum, void* stream)
{
    if (int rc = check_desc(d)) return rc;
    if (!vol || !tf || !cam || !out_rgba) return fail(DR_EINVAL, "dr_forward: null pointer");
    if ((d->flags & DR_F_HAS_JITTER) && !jitter) return fail(DR_EINVAL, "dr_forward: DR_F_HAS_JITTER set but jitter is null");
    if (!aligned(tf, 16) || !aligned(out_rgba, 16)) return fail(DR_EALIGN, "dr_forward: tf and out_rgba must be 16-byte aligned");
    if (target && (!loss_sum || !aligned(target, 16))) return fail(DR_EINVAL, "dr_forward_mse: loss_sum is null or target is not 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool nd = d->flags & DR_F_NONDIFF, gen = d->tap_generic, brick = d->flags & DR_F_LAYOUT_BRICK8;
#define DR_FWD1(VT, LAY, ND, GEN) launch_fwd<VT, LAY, ND, GEN>(d, vol, tf, cam, jitter, out_rgba, out_K, out_Tprev, st, target, loss_sum)
#define DR_FWD(VT)                                                                                                  \
    (brick ? (nd ? DR_FWD1(VT, LAYOUT_BRICK8, true, false) : DR_FWD1(VT, LAYOUT_BRICK8, false, false))              \
           : (nd ? (gen ? DR_FWD1(VT, LAYOUT_LINEAR, true, true) : DR_FWD1(VT, LAYOUT_LINEAR, true, false))         \
                 : (gen ? DR_FWD1(VT, LAYOUT_LINEAR, false, true) : DR_FWD1(VT, LAYOUT_LINEAR, false, false))))
    return d->vox_dtype == DR_VOX_F32 ? DR_FWD(float) : DR_FWD(__half);
#undef DR_FWD1
#undef DR_FWD
}

int backward_impl(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter,
                  const float* grad_out, const float* out_rgba, const int32_t* K, const float* Tprev, float* grad_vol_cells,
                  float* grad_tf, void* workspace, size_t workspace_bytes, float mse_scale, void* stream)
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
    float4* slots = static_cast<float4*>(workspace);
    float4* gcells = reinterpret_cast<float4*>(grad_vol_cells);
    int rc;
    const bool brick = d->flags & DR_F_LAYOUT_BRICK8;
#define DR_BWDL(VT, LAY, GEN) dispatch_bwd<VT, LAY, GEN>(d, vol, tf, cam, jitter, grad_out, out_rgba, K, Tprev, gcells, slots, st, mse_scale)
#define DR_BWDV(VT) (brick ? DR_BWDL(VT, LAYOUT_BRICK8, false) : (d->tap_generic ? DR_BWDL(VT, LAYOUT_LINEAR, true) : DR_BWDL(VT, LAYOUT_LINEAR, false)))
    rc = d->vox_dtype == DR_VOX_F32 ? DR_BWDV(float) : DR_BWDV(__half);
#undef DR_BWDV
#undef DR_BWDL
    if (rc) return rc;
    if (wt) {
        dim3 grid((d->R * 4 + 255) / 256, d->Btf);
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
    return (size_t)d->Btf * kTfSlots * d->R * sizeof(float4);
}

int dr_brick_volume(const DrDesc* d, const void* vol_linear, void* vol_bricked, void* stream)
{
    if (int rc = check_desc(d)) return rc;
    if (!vol_linear || !vol_bricked) return fail(DR_EINVAL, "dr_brick_volume: null pointer");
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

int dr_forward(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter,
               float* out_rgba, int32_t* out_K, float* out_Tprev, void* stream)
{
    return forward_impl(d, vol, tf, cam, jitter, out_rgba, out_K, out_Tprev, nullptr, nullptr, stream);
}

int dr_forward_mse(const DrDesc* d, const void* vol, const float* tf, const float* cam, const float* jitter,
                   const float* target, float* out_rgba, int32_t* out_K, float* out_Tprev, float* loss_sum, void* stream)
{
    if (!target || !loss_sum) return fail(DR_EINVAL, "dr_forward_mse: target or loss_sum is null");
    return forward_impl(d, vol, tf, cam, jitter, out_rgba, out_K, out_Tprev, target, loss_sum, stream);
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
    const size_t n = (size_t)d->X * d->Y * d->Z;
    dim3 grid((unsigned)((n + 255) / 256), d->Bvol);
    gather_grad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*d, grad_vol_cells, grad_linear, accumulate);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DR_OK : fail_cuda(e, "gather_grad_kernel launch");
}

}  // extern "C"
