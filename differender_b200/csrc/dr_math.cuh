// dr_math.cuh -- per-ray arithmetic of the differentiable ray-march, shared by the forward and backward kernels.
//
// Follows the reference's differender/volume_raycaster.py (cited per function as :line).  Written from the
// behavioural spec, not translated from the Taichi source.
//
// Rounding contract (DESIGN.md "Numerics"): the alpha path (ray set-up -> sample position -> trilinear taps
// -> TF alpha -> opacity -> accumulated alpha) uses explicitly rounded operations (DR_MUL/DR_ADD/DR_FMA ...,
// never contracted by the compiler), so the discrete decisions (sample count n, early termination K) and the
// ill-conditioned central-difference normal are reproducible.  Shading and all adjoint arithmetic use plain
// operators (the compiler may fuse them).
//
// The functions are __host__ __device__ so that tests/hostsim can compile this very file with g++ and compare
// it to the oracle without a GPU.  That harness is test-only; the product never runs this code on the CPU.
#pragma once

#include <math.h>
#include <stdint.h>

#include "diffrender.h"
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

#if defined(__CUDACC__)
#define DR_HD __host__ __device__ __forceinline__
#else
#define DR_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define DR_MUL(a, b) __fmul_rn((a), (b))
#define DR_ADD(a, b) __fadd_rn((a), (b))
#define DR_SUB(a, b) __fsub_rn((a), (b))
#define DR_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define DR_DIV(a, b) __fdiv_rn((a), (b))
#define DR_SQRT(a) __fsqrt_rn((a))
#define DR_RSQRT(a) dr::rsqrt_fast((a))
#define DR_SAT(a) __saturatef((a))
#define DR_ALIGN16 __align__(16)
#else
// host build (tests/hostsim): compiled with -ffp-contract=off, so each operator is one IEEE operation
#define DR_MUL(a, b) ((a) * (b))
#define DR_ADD(a, b) ((a) + (b))
#define DR_SUB(a, b) ((a) - (b))
#define DR_FMA(a, b, c) fmaf((a), (b), (c))
#define DR_DIV(a, b) ((a) / (b))
#define DR_SQRT(a) sqrtf((a))
#define DR_RSQRT(a) (1.0f / sqrtf((a)))
#define DR_SAT(a) fminf(1.0f, fmaxf(0.0f, (a)))
#define DR_ALIGN16 alignas(16)
#endif

// Debug build (-DDR_BOUNDS_CHECK): every volume load and gradient reduction checks its index and counts violations
// (compute-sanitizer is not available on the B200 pool).  Compiled out otherwise.
#if defined(DR_BOUNDS_CHECK) && defined(__CUDA_ARCH__)
extern __device__ unsigned long long dr_oob_counter;
#define DR_OOB_IF(cond) do { if (cond) atomicAdd(&dr_oob_counter, 1ULL); } while (0)
#else
#define DR_OOB_IF(cond) do { } while (0)
#endif

namespace dr {

DR_HD int imin(int a, int b) { return a < b ? a : b; }

// 1/sqrt(x) off the exact path as FMUL + MUFU.RSQ + FMUL, branch-free: rsqrtf() wraps the MUFU in compare / predicate /
// rescale code for denormal arguments (11 issue slots once if-converted).  Scaling the argument by 2^40 (exact) keeps
// every positive fp32 up to 2^87 inside MUFU.RSQ's normal range, and the result is scaled back by 2^20 (exact).
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float rsqrt_fast(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1099511627776.0f));
    return r * 1048576.0f;
}
#endif
// 1/x for x well inside the normal range, off the exact path: one MUFU.RCP (no denormal/overflow fix-up code)
DR_HD float fast_rcp(float x)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

struct F3 { float x, y, z; };
struct F4 { float x, y, z, w; };

// ---------------------------------------------------------------------------------------------------------
// exact helpers
// ---------------------------------------------------------------------------------------------------------
// taichi_glsl.mix(x, y, a) = x*(1-a) + y*a, rounded as fma(x, 1-a, y*a)  (omt = 1-a precomputed)
DR_HD float mix_e(float a, float b, float omt, float t) { return DR_FMA(a, omt, DR_MUL(b, t)); }

// ---------------------------------------------------------------------------------------------------------
// Packed pairs.  sm_100a has two-wide fp32 instructions (FMUL2 / FFMA2 / FADD2 on 64-bit register pairs: PTX
// mul/fma/add .f32x2): each lane is rounded exactly like the scalar instruction (IEEE rn, no ftz), the FMA pipe is busy
// for two cycles, but the pair costs ONE issue slot -- and issue slots are what bounds these kernels
// (profiles/r01_f32x2_microbench.txt: 8 FFMA + 16 integer instructions 16.6 ms, 4 FFMA2 + the same integer work 13.6 ms).
// The trilinear mixes come in natural pairs (the same fraction applied to two rows of voxels), so the sample evaluation
// is written on F2.  On the host (tests/hostsim) an F2 operation is two scalar operations with the same rounding.
// ---------------------------------------------------------------------------------------------------------
struct F2 { float x, y; };
DR_HD F2 f2(float x, float y) { F2 r = { x, y }; return r; }
DR_HD F2 splat(float a) { F2 r = { a, a }; return r; }
// (x != y) ? a : b on a register pair: the predicate is formed inside the asm (no boolean materialised) and the pair is
// selected as one 64-bit selp, which ptxas lowers to two SELs (two scalar selects written into an aligned pair became
// four predicated MOVs)
DR_HD F2 sel2_ne(int x, int y, F2 a, F2 b)
{
#if defined(__CUDA_ARCH__)
    F2 r;
    asm("{\n\t.reg .pred q;\n\t.reg .b64 pa, pb;\n\tsetp.ne.s32 q, %6, %7;\n\tmov.b64 pa, {%2, %3};\n\tmov.b64 pb, {%4, %5};\n\t"
        "selp.b64 pa, pa, pb, q;\n\tmov.b64 {%0, %1}, pa;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "r"(x), "r"(y));
    return r;
#else
    return x != y ? a : b;
#endif
}
DR_HD F2 mul2(F2 a, F2 b)
{
#if defined(__CUDA_ARCH__)
    F2 r;
    asm("{\n\t.reg .b64 pa, pb;\n\tmov.b64 pa, {%2, %3};\n\tmov.b64 pb, {%4, %5};\n\tmul.rn.f32x2 pa, pa, pb;\n\tmov.b64 {%0, %1}, pa;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
#else
    return f2(DR_MUL(a.x, b.x), DR_MUL(a.y, b.y));
#endif
}
DR_HD F2 fma2(F2 a, F2 b, F2 c)
{
#if defined(__CUDA_ARCH__)
    F2 r;
    asm("{\n\t.reg .b64 pa, pb, pc;\n\tmov.b64 pa, {%2, %3};\n\tmov.b64 pb, {%4, %5};\n\tmov.b64 pc, {%6, %7};\n\t"
        "fma.rn.f32x2 pa, pa, pb, pc;\n\tmov.b64 {%0, %1}, pa;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
#else
    return f2(DR_FMA(a.x, b.x, c.x), DR_FMA(a.y, b.y, c.y));
#endif
}
DR_HD F2 sub2(F2 a, F2 b)
{
#if defined(__CUDA_ARCH__)
    F2 r;
    asm("{\n\t.reg .b64 pa, pb;\n\tmov.b64 pa, {%2, %3};\n\tmov.b64 pb, {%4, %5};\n\tsub.rn.f32x2 pa, pa, pb;\n\tmov.b64 {%0, %1}, pa;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
#else
    return f2(DR_SUB(a.x, b.x), DR_SUB(a.y, b.y));
#endif
}
// two mixes with per-lane fractions: (a.x*(1-t.x) + b.x*t.x, a.y*(1-t.y) + b.y*t.y), each rounded like mix_e
DR_HD F2 mix2(F2 a, F2 b, F2 omt, F2 t) { return fma2(a, omt, mul2(b, t)); }

DR_HD float dot_e(F3 a, F3 b) { return DR_ADD(DR_ADD(DR_MUL(a.x, b.x), DR_MUL(a.y, b.y)), DR_MUL(a.z, b.z)); }
DR_HD F3 cross_e(F3 a, F3 b)
{
    F3 r;
    r.x = DR_SUB(DR_MUL(a.y, b.z), DR_MUL(a.z, b.y));
    r.y = DR_SUB(DR_MUL(a.z, b.x), DR_MUL(a.x, b.z));
    r.z = DR_SUB(DR_MUL(a.x, b.y), DR_MUL(a.y, b.x));
    return r;
}
// Taichi Vector.normalized(): v * (1 / sqrt(v.v))
DR_HD F3 normalized_e(F3 a)
{
    float inv = DR_DIV(1.0f, DR_SQRT(dot_e(a, a)));
    F3 r = { DR_MUL(a.x, inv), DR_MUL(a.y, inv), DR_MUL(a.z, inv) };
    return r;
}

// floor of a value in [0, 2^22): returns floor as float and writes the BIASED integer b = floor + kFloorBias (the bit
// pattern of the float 2^23 + floor).  Cell comparisons use b directly; lo_of() gives the index.      low_high_frac :7-21
constexpr int kFloorBias = 0x4B000000;
DR_HD float floor_pos(float p, int& b)
{
#if defined(__CUDA_ARCH__)
    // p + 2^23 rounded toward -inf is exactly 2^23 + floor(p); all on the FMA/ALU pipes (no F2I/I2F)
    float r = __fadd_rd(p, 8388608.0f);
    b = __float_as_int(r);
    return __fsub_rn(r, 8388608.0f);
#else
    float l = floorf(p);
    b = (int)l + kFloorBias;
    return l;
#endif
}

struct Loc { int b; float f; };      // b: biased floor index (see floor_pos), f: fraction
DR_HD int lo_of(Loc q) { return q.b - kFloorBias; }

// address part of sample_volume_trilinear for one axis                                   :163-172
DR_HD Loc locate(float pos, float scale)
{
    float p = DR_MUL(DR_SAT(DR_FMA(0.5f, pos, 0.5f)), scale);
    Loc r;
    float l = floor_pos(p, r.b);
    r.f = DR_SUB(p, l);
    return r;
}

// locate() of the two taps pos + delta and pos - delta of one axis, on packed pairs (same operations per lane)
DR_HD void locate_pair(float pos, float delta, float scale, Loc& plus, Loc& minus)
{
    const F2 p = mul2(f2(DR_SAT(DR_FMA(0.5f, DR_ADD(pos, delta), 0.5f)), DR_SAT(DR_FMA(0.5f, DR_SUB(pos, delta), 0.5f))), splat(scale));
#if defined(__CUDA_ARCH__)
    F2 r;
    asm("{\n\t.reg .b64 pa, pb;\n\tmov.b64 pa, {%2, %3};\n\tmov.b64 pb, {%4, %4};\n\tadd.rm.f32x2 pa, pa, pb;\n\tmov.b64 {%0, %1}, pa;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(p.x), "f"(p.y), "f"(8388608.0f));
    plus.b = __float_as_int(r.x); minus.b = __float_as_int(r.y);
    const F2 f = sub2(p, sub2(r, splat(8388608.0f)));
    plus.f = f.x; minus.f = f.y;
#else
    float l = floor_pos(p.x, plus.b);
    plus.f = DR_SUB(p.x, l);
    l = floor_pos(p.y, minus.b);
    minus.f = DR_SUB(p.y, l);
#endif
}

// ---------------------------------------------------------------------------------------------------------
// bricked volume addressing: 8x8x8 bricks, x fastest inside a brick, bricks in x,y,z raster order.
// The element offset is separable: off(x,y,z) = offx(x) + offy(y) + offz(z).
// ---------------------------------------------------------------------------------------------------------
struct Layout {
    int sY, sZ;           // element strides between bricks along y and z: nbx*512, nbx*nby*512
    int mx, my, mz;       // dim - 1 (index clamps)
    unsigned cbias;       // (kFloorBias*Z + kFloorBias)*X + kFloorBias mod 2^32: turns biased floor indices into a cell index
    unsigned ycells;      // X*Z: cells (= voxels) between y-neighbours
};
// the two layout constants the kernels take as an ARGUMENT (constant-bank operands; derived on the device from d.X / d.Z the
// compiler re-computed them inside the march loop with vector-register multiplies)
struct LayoutConsts { unsigned cbias, ycells; };
// Offsets are UNSIGNED so that `pointer + offset` is one IMAD.WIDE.U32 (a signed int needs LEA + LEA.HI.X.SX32).
typedef unsigned int uoff;
DR_HD uoff offx(int x) { return (uoff)(((x >> 3) << 9) | (x & 7)); }
DR_HD uoff offy(int y, int sY) { return (uoff)((y >> 3) * sY + ((y & 7) << 3)); }
DR_HD uoff offz(int z, int sZ) { return (uoff)((z >> 3) * sZ + ((z & 7) << 6)); }

// pointer + unsigned element offset as ONE instruction (IMAD.WIDE.U32, FMA pipe).  Left to the compiler, a 2-byte element
// type becomes four 64-bit IADD3/IADD3.X per address (base + off + off), which made integer adds 25 % of the fp16 kernels.
// The element size comes from constant memory so that ptxas cannot strength-reduce the multiply back into LEA + LEA.HI.X.
#if defined(__CUDACC__)
static __constant__ unsigned dr_elem_size[3] = { 1u, 2u, 4u };     // indexed by sizeof(VT) / 2
#endif
template <typename VT>
DR_HD const VT* ptr_add(const VT* p, uoff off)
{
#if defined(__CUDA_ARCH__)
    unsigned long long a;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(off), "r"(dr_elem_size[sizeof(VT) / 2]), "l"((unsigned long long)p));
    return reinterpret_cast<const VT*>(a);
#else
    return p + off;
#endif
}
DR_HD float load_vox(const float* p, int off)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p + off);
#else
    return p[off];
#endif
}
DR_HD float load_vox(const float* p, uoff off) { return load_vox(ptr_add(p, off), 0); }
#if defined(__CUDACC__)
DR_HD float load_vox(const __half* p, int off)
{
#if defined(__CUDA_ARCH__)
    return __half2float(__ldg(p + off));
#else
    return __half2float(p[off]);
#endif
}
DR_HD float load_vox(const __half* p, uoff off) { return load_vox(ptr_add(p, off), 0); }
#endif
// predicated load: ONE predicated LDG instead of a divergent branch around the load.  When !pred the result is UNDEFINED on
// the device (the register is left as it is -- a zero-initialising MOV per value cost 25 issue slots per sample); every
// caller discards it with a select.  The host build returns 0.
DR_HD float load_vox_if(const float* p, bool pred)
{
#if defined(__CUDA_ARCH__)
    float v;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q ld.global.nc.f32 %0, [%1];\n\t}" : "=f"(v) : "l"(p), "r"((int)pred));
    return v;
#else
    return pred ? *p : 0.0f;
#endif
}
#if defined(__CUDACC__)
DR_HD float load_vox_if(const __half* p, bool pred)
{
#if defined(__CUDA_ARCH__)
    unsigned short h;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q ld.global.nc.b16 %0, [%1];\n\t}" : "=h"(h) : "l"(p), "r"((int)pred));
    return __half2float(__ushort_as_half(h));
#else
    return pred ? __half2float(*p) : 0.0f;
#endif
}
#endif
// ---- cell-major records (LAYOUT_CELL8): 8 consecutive elements per cell ---------------------------------
#if defined(__CUDACC__)
static __constant__ unsigned dr_rec_size[3] = { 8u, 16u, 32u };    // bytes per record, indexed by sizeof(VT) / 2
#endif
template <typename VT>
DR_HD const VT* rec_add(const VT* p, uoff cell)                     // p + 8*cell as ONE IMAD.WIDE.U32 (64-bit result)
{
#if defined(__CUDA_ARCH__)
    unsigned long long a;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(cell), "r"(dr_rec_size[sizeof(VT) / 2]), "l"((unsigned long long)p));
    return reinterpret_cast<const VT*>(a);
#else
    return p + (size_t)cell * 8;
#endif
}
DR_HD void load_vox8(const float* r, float v[8])
{
#if defined(__CUDA_ARCH__)
    const float4 a = __ldg(reinterpret_cast<const float4*>(r)), b = __ldg(reinterpret_cast<const float4*>(r) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#else
    for (int q = 0; q < 8; ++q) v[q] = r[q];
#endif
}
DR_HD void load_vox4_if(const float* r, bool pred, float n[4])
{
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %5, 0;\n\t"
        "@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "=f"(n[0]), "=f"(n[1]), "=f"(n[2]), "=f"(n[3]) : "l"(r), "r"((int)pred));
#else
    for (int q = 0; q < 4; ++q) n[q] = pred ? r[q] : 0.0f;
#endif
}
DR_HD void load_vox2_if(const float* r, bool pred, float n[2])
{
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t"
        "@q ld.global.nc.v2.f32 {%0, %1}, [%2];\n\t}" : "=f"(n[0]), "=f"(n[1]) : "l"(r), "r"((int)pred));
#else
    for (int q = 0; q < 2; ++q) n[q] = pred ? r[q] : 0.0f;
#endif
}
#if defined(__CUDACC__)
DR_HD void load_vox8(const __half* r, float v[8])
{
#if defined(__CUDA_ARCH__)
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(r));
    const float2 p0 = __half22float2(*reinterpret_cast<const __half2*>(&a.x)), p1 = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
    const float2 p2 = __half22float2(*reinterpret_cast<const __half2*>(&a.z)), p3 = __half22float2(*reinterpret_cast<const __half2*>(&a.w));
    v[0] = p0.x; v[1] = p0.y; v[2] = p1.x; v[3] = p1.y; v[4] = p2.x; v[5] = p2.y; v[6] = p3.x; v[7] = p3.y;
#else
    for (int q = 0; q < 8; ++q) v[q] = __half2float(r[q]);
#endif
}
DR_HD void load_vox4_if(const __half* r, bool pred, float n[4])
{
#if defined(__CUDA_ARCH__)
    unsigned lo, hi;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q ld.global.nc.v2.b32 {%0, %1}, [%2];\n\t}"
        : "=r"(lo), "=r"(hi) : "l"(r), "r"((int)pred));
    const float2 p0 = __half22float2(*reinterpret_cast<const __half2*>(&lo)), p1 = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    n[0] = p0.x; n[1] = p0.y; n[2] = p1.x; n[3] = p1.y;
#else
    for (int q = 0; q < 4; ++q) n[q] = pred ? __half2float(r[q]) : 0.0f;
#endif
}
DR_HD void load_vox2_if(const __half* r, bool pred, float n[2])
{
#if defined(__CUDA_ARCH__)
    unsigned w;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q ld.global.nc.b32 %0, [%1];\n\t}" : "=r"(w) : "l"(r), "r"((int)pred));
    const float2 p0 = __half22float2(*reinterpret_cast<const __half2*>(&w));
    n[0] = p0.x; n[1] = p0.y;
#else
    for (int q = 0; q < 2; ++q) n[q] = pred ? __half2float(r[q]) : 0.0f;
#endif
}
#endif
// Plane beyond the centre cell from the face neighbours' records (LAYOUT_CELL8), as TWO loads under complementary-use
// predicates: `rm` (the - neighbour) under bm != bc, then `rp` (the + neighbour) under bp != bc, into the same registers
// (when both taps crossed the + plane wins; TAPS_TWO fetches the other one separately).  The predicates are formed inside
// the asm from the biased floor indices, so no boolean is materialised in a register and no address is selected: the
// ALU pipe, the busiest of the backward, does nothing for the fetch but the two address IMADs.  Undefined when neither
// tap crossed (every caller discards the values with a select); the host build returns 0 then.
#define DR_PM_PRED4 "{\n\t.reg .pred p, m;\n\tsetp.ne.s32 p, %6, %8;\n\tsetp.ne.s32 m, %7, %8;\n\t"     /* 4 outputs: rp %4, rm %5, bp %6, bm %7, bc %8 */
#define DR_PM_PRED2 "{\n\t.reg .pred p, m;\n\tsetp.ne.s32 p, %4, %6;\n\tsetp.ne.s32 m, %5, %6;\n\t"     /* 2 outputs: rp %2, rm %3, bp %4, bm %5, bc %6 */
DR_HD void cell_plane_y(const float* rp, const float* rm, int bp, int bm, int bc, float n[4])
{
#if defined(__CUDA_ARCH__)
    asm(DR_PM_PRED4 "@m ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%5];\n\t@p ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4+16];\n\t}"
        : "=f"(n[0]), "=f"(n[1]), "=f"(n[2]), "=f"(n[3]) : "l"(rp), "l"(rm), "r"(bp), "r"(bm), "r"(bc));
#else
    const float* r = bp != bc ? rp + 4 : rm;
    for (int q = 0; q < 4; ++q) n[q] = (bp != bc || bm != bc) ? r[q] : 0.0f;
#endif
}
// the x neighbours are the adjacent records: rc is the CENTRE record and the loads use immediate offsets from it (+ neighbour
// at +32 bytes: quarters {2,3} {6,7}; - neighbour at -32 bytes: quarters {0,1} {4,5}) -- no address arithmetic at all
DR_HD void cell_plane_x(const float* rc, int bp, int bm, int bc, float n[4])
{
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .pred p, m;\n\tsetp.ne.s32 p, %5, %7;\n\tsetp.ne.s32 m, %6, %7;\n\t"
        "@m ld.global.nc.v2.f32 {%0, %1}, [%4+-32];\n\t@m ld.global.nc.v2.f32 {%2, %3}, [%4+-16];\n\t"
        "@p ld.global.nc.v2.f32 {%0, %1}, [%4+40];\n\t@p ld.global.nc.v2.f32 {%2, %3}, [%4+56];\n\t}"
        : "=f"(n[0]), "=f"(n[1]), "=f"(n[2]), "=f"(n[3]) : "l"(rc), "r"(bp), "r"(bm), "r"(bc));
#else
    const float* r = bp != bc ? rc + 8 + 2 : rc - 8;
    const bool any = bp != bc || bm != bc;
    n[0] = any ? r[0] : 0.0f; n[1] = any ? r[1] : 0.0f; n[2] = any ? r[4] : 0.0f; n[3] = any ? r[5] : 0.0f;
#endif
}
// The corners of an x tap's OWN cell: the centre's (already in a0..b1) unless the tap crossed (bt != bc), in which case the
// face neighbour's whole record -- which is exactly the tap cell's (A0, B0, A1, B1) in the same slot order -- overwrites them
// with two predicated 16-byte loads at an immediate offset from the centre record.  No selects at all, and both taps of
// an axis may cross independently (no TAPS_TWO case).  fp32 records only.
template <bool PLUS>
DR_HD void cell_tap_x(const float* rc, int bt, int bc, F2& a0, F2& b0, F2& a1, F2& b1)
{
#if defined(__CUDA_ARCH__)
    if (PLUS)
        asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %9, %10;\n\t@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%8+32];\n\t"
            "@q ld.global.nc.v4.f32 {%4, %5, %6, %7}, [%8+48];\n\t}"
            : "+f"(a0.x), "+f"(a0.y), "+f"(b0.x), "+f"(b0.y), "+f"(a1.x), "+f"(a1.y), "+f"(b1.x), "+f"(b1.y) : "l"(rc), "r"(bt), "r"(bc));
    else
        asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %9, %10;\n\t@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%8+-32];\n\t"
            "@q ld.global.nc.v4.f32 {%4, %5, %6, %7}, [%8+-16];\n\t}"
            : "+f"(a0.x), "+f"(a0.y), "+f"(b0.x), "+f"(b0.y), "+f"(a1.x), "+f"(a1.y), "+f"(b1.x), "+f"(b1.y) : "l"(rc), "r"(bt), "r"(bc));
#else
    if (bt != bc) {
        const float* r = PLUS ? rc + 8 : rc - 8;
        a0 = f2(r[0], r[1]); b0 = f2(r[2], r[3]); a1 = f2(r[4], r[5]); b1 = f2(r[6], r[7]);
    }
#endif
}
// n = slots (c), (c+4), (c+2), (c+6) with c = 1 for the + neighbour, 0 for the - neighbour: (x0,y0) (x0,y1) (x1,y0) (x1,y1)
DR_HD void cell_plane_z(const float* rp, const float* rm, int bp, int bm, int bc, float n[4])
{
#if defined(__CUDA_ARCH__)
    asm(DR_PM_PRED4 "@m ld.global.nc.f32 %0, [%5];\n\t@m ld.global.nc.f32 %1, [%5+16];\n\t@m ld.global.nc.f32 %2, [%5+8];\n\t"
        "@m ld.global.nc.f32 %3, [%5+24];\n\t@p ld.global.nc.f32 %0, [%4+4];\n\t@p ld.global.nc.f32 %1, [%4+20];\n\t"
        "@p ld.global.nc.f32 %2, [%4+12];\n\t@p ld.global.nc.f32 %3, [%4+28];\n\t}"
        : "=f"(n[0]), "=f"(n[1]), "=f"(n[2]), "=f"(n[3]) : "l"(rp), "l"(rm), "r"(bp), "r"(bm), "r"(bc));
#else
    const float* r = bp != bc ? rp + 1 : rm;
    const bool any = bp != bc || bm != bc;
    n[0] = any ? r[0] : 0.0f; n[1] = any ? r[4] : 0.0f; n[2] = any ? r[2] : 0.0f; n[3] = any ? r[6] : 0.0f;
#endif
}
#if defined(__CUDACC__)
DR_HD void cell_plane_y(const __half* rp, const __half* rm, int bp, int bm, int bc, float n[4])
{
#if defined(__CUDA_ARCH__)
    unsigned lo, hi;
    asm(DR_PM_PRED2 "@m ld.global.nc.v2.b32 {%0, %1}, [%3];\n\t@p ld.global.nc.v2.b32 {%0, %1}, [%2+8];\n\t}"
        : "=r"(lo), "=r"(hi) : "l"(rp), "l"(rm), "r"(bp), "r"(bm), "r"(bc));
    const float2 p0 = __half22float2(*reinterpret_cast<const __half2*>(&lo)), p1 = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    n[0] = p0.x; n[1] = p0.y; n[2] = p1.x; n[3] = p1.y;
#else
    const __half* r = bp != bc ? rp + 4 : rm;
    for (int q = 0; q < 4; ++q) n[q] = (bp != bc || bm != bc) ? __half2float(r[q]) : 0.0f;
#endif
}
DR_HD void cell_plane_x(const __half* rc, int bp, int bm, int bc, float n[4])      // records are 16 bytes
{
#if defined(__CUDA_ARCH__)
    unsigned lo, hi;
    asm("{\n\t.reg .pred p, m;\n\tsetp.ne.s32 p, %3, %5;\n\tsetp.ne.s32 m, %4, %5;\n\t"
        "@m ld.global.nc.b32 %0, [%2+-16];\n\t@m ld.global.nc.b32 %1, [%2+-8];\n\t"
        "@p ld.global.nc.b32 %0, [%2+20];\n\t@p ld.global.nc.b32 %1, [%2+28];\n\t}"
        : "=r"(lo), "=r"(hi) : "l"(rc), "r"(bp), "r"(bm), "r"(bc));
    const float2 p0 = __half22float2(*reinterpret_cast<const __half2*>(&lo)), p1 = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    n[0] = p0.x; n[1] = p0.y; n[2] = p1.x; n[3] = p1.y;
#else
    const __half* r = bp != bc ? rc + 8 + 2 : rc - 8;
    const bool any = bp != bc || bm != bc;
    n[0] = any ? __half2float(r[0]) : 0.0f; n[1] = any ? __half2float(r[1]) : 0.0f;
    n[2] = any ? __half2float(r[4]) : 0.0f; n[3] = any ? __half2float(r[5]) : 0.0f;
#endif
}
DR_HD void cell_plane_z(const __half* rp, const __half* rm, int bp, int bm, int bc, float n[4])
{
#if defined(__CUDA_ARCH__)
    unsigned short h0, h1, h2, h3;
    asm(DR_PM_PRED4 "@m ld.global.nc.b16 %0, [%5];\n\t@m ld.global.nc.b16 %1, [%5+8];\n\t@m ld.global.nc.b16 %2, [%5+4];\n\t"
        "@m ld.global.nc.b16 %3, [%5+12];\n\t@p ld.global.nc.b16 %0, [%4+2];\n\t@p ld.global.nc.b16 %1, [%4+10];\n\t"
        "@p ld.global.nc.b16 %2, [%4+6];\n\t@p ld.global.nc.b16 %3, [%4+14];\n\t}"
        : "=h"(h0), "=h"(h1), "=h"(h2), "=h"(h3) : "l"(rp), "l"(rm), "r"(bp), "r"(bm), "r"(bc));
    n[0] = __half2float(__ushort_as_half(h0)); n[1] = __half2float(__ushort_as_half(h1));
    n[2] = __half2float(__ushort_as_half(h2)); n[3] = __half2float(__ushort_as_half(h3));
#else
    const __half* r = bp != bc ? rp + 1 : rm;
    const bool any = bp != bc || bm != bc;
    n[0] = any ? __half2float(r[0]) : 0.0f; n[1] = any ? __half2float(r[4]) : 0.0f;
    n[2] = any ? __half2float(r[2]) : 0.0f; n[3] = any ? __half2float(r[6]) : 0.0f;
#endif
}
#endif
// ---- uint8-stored volumes (DR_VOX_U8; the reference's skull.raw, examples/taichi_volume_raycaster.py:548-550) -----------
// A voxel is u8 / 255 in fp32, exactly as dr_ingest_u8 (and numpy's float32 division) rounds it:
//   fl(x / 255) == fma(x, r, fl(x * r_lo))   for every x in 0..255, with r = fl(1/255), r_lo = fl(1/255 - r)
// (checked exhaustively: tests/test_host_logic.py; a plain multiply by r is off by one ulp for 126 of the 256 values).
// Two FMA-pipe instructions per voxel (one FMUL2 + one FFMA2 per pair) instead of a division.  A cell record is 8 bytes:
// ONE 8-byte load per cell, an eighth of the fp32 record.
typedef unsigned char u8vox;
constexpr float kU8r = 1.0f / 255.0f;
constexpr float kU8rlo = (float)(1.0 / 255.0 - (double)kU8r);
DR_HD float u8_unit(unsigned b)
{
    const float x = (float)b;
    return DR_FMA(x, kU8r, DR_MUL(x, kU8rlo));
}
DR_HD F2 u8_unit2(unsigned b0, unsigned b1)
{
    const F2 x = f2((float)b0, (float)b1);
    return fma2(x, splat(kU8r), mul2(x, splat(kU8rlo)));
}
DR_HD void u8_unpack4(unsigned w, float n[4])       // bytes 0..3 of w -> n[0..3]
{
    const F2 a = u8_unit2(w & 0xffu, (w >> 8) & 0xffu), b = u8_unit2((w >> 16) & 0xffu, w >> 24);
    n[0] = a.x; n[1] = a.y; n[2] = b.x; n[3] = b.y;
}
DR_HD float load_vox(const u8vox* p, int off)
{
#if defined(__CUDA_ARCH__)
    return u8_unit(__ldg(p + off));
#else
    return u8_unit(p[off]);
#endif
}
DR_HD float load_vox(const u8vox* p, uoff off) { return load_vox(ptr_add(p, off), 0); }
DR_HD float load_vox_if(const u8vox* p, bool pred)
{
#if defined(__CUDA_ARCH__)
    unsigned b;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q ld.global.nc.u8 %0, [%1];\n\t}" : "=r"(b) : "l"(p), "r"((int)pred));
    return u8_unit(b & 0xffu);
#else
    return pred ? u8_unit(*p) : 0.0f;
#endif
}
DR_HD void load_vox8(const u8vox* r, float v[8])
{
#if defined(__CUDA_ARCH__)
    const uint2 a = __ldg(reinterpret_cast<const uint2*>(r));
    u8_unpack4(a.x, v); u8_unpack4(a.y, v + 4);
#else
    for (int q = 0; q < 8; ++q) v[q] = u8_unit(r[q]);
#endif
}
DR_HD void load_vox4_if(const u8vox* r, bool pred, float n[4])
{
#if defined(__CUDA_ARCH__)
    unsigned w;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q ld.global.nc.b32 %0, [%1];\n\t}" : "=r"(w) : "l"(r), "r"((int)pred));
    u8_unpack4(w, n);
#else
    for (int q = 0; q < 4; ++q) n[q] = pred ? u8_unit(r[q]) : 0.0f;
#endif
}
DR_HD void load_vox2_if(const u8vox* r, bool pred, float n[2])
{
#if defined(__CUDA_ARCH__)
    unsigned short h;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q ld.global.nc.b16 %0, [%1];\n\t}" : "=h"(h) : "l"(r), "r"((int)pred));
    const F2 a = u8_unit2((unsigned)h & 0xffu, ((unsigned)h >> 8) & 0xffu);
    n[0] = a.x; n[1] = a.y;
#else
    for (int q = 0; q < 2; ++q) n[q] = pred ? u8_unit(r[q]) : 0.0f;
#endif
}
// DUAL plane fetches (one predicated load per neighbour, see the fp32 versions above); records are 8 bytes
DR_HD void cell_plane_y(const u8vox* rp, const u8vox* rm, int bp, int bm, int bc, float n[4])
{
#if defined(__CUDA_ARCH__)
    unsigned w;
    asm("{\n\t.reg .pred p, m;\n\tsetp.ne.s32 p, %3, %5;\n\tsetp.ne.s32 m, %4, %5;\n\t"
        "@m ld.global.nc.b32 %0, [%2];\n\t@p ld.global.nc.b32 %0, [%1+4];\n\t}" : "=r"(w) : "l"(rp), "l"(rm), "r"(bp), "r"(bm), "r"(bc));
    u8_unpack4(w, n);
#else
    const u8vox* r = bp != bc ? rp + 4 : rm;
    for (int q = 0; q < 4; ++q) n[q] = (bp != bc || bm != bc) ? u8_unit(r[q]) : 0.0f;
#endif
}
DR_HD void cell_plane_x(const u8vox* rc, int bp, int bm, int bc, float n[4])
{
#if defined(__CUDA_ARCH__)
    unsigned short lo, hi;
    asm("{\n\t.reg .pred p, m;\n\tsetp.ne.s32 p, %3, %5;\n\tsetp.ne.s32 m, %4, %5;\n\t"
        "@m ld.global.nc.b16 %0, [%2+-8];\n\t@m ld.global.nc.b16 %1, [%2+-4];\n\t"
        "@p ld.global.nc.b16 %0, [%2+10];\n\t@p ld.global.nc.b16 %1, [%2+14];\n\t}"
        : "=h"(lo), "=h"(hi) : "l"(rc), "r"(bp), "r"(bm), "r"(bc));
    const F2 a = u8_unit2((unsigned)lo & 0xffu, ((unsigned)lo >> 8) & 0xffu), b = u8_unit2((unsigned)hi & 0xffu, ((unsigned)hi >> 8) & 0xffu);
    n[0] = a.x; n[1] = a.y; n[2] = b.x; n[3] = b.y;
#else
    const u8vox* r = bp != bc ? rc + 8 + 2 : rc - 8;
    const bool any = bp != bc || bm != bc;
    n[0] = any ? u8_unit(r[0]) : 0.0f; n[1] = any ? u8_unit(r[1]) : 0.0f; n[2] = any ? u8_unit(r[4]) : 0.0f; n[3] = any ? u8_unit(r[5]) : 0.0f;
#endif
}
DR_HD void cell_plane_z(const u8vox* rp, const u8vox* rm, int bp, int bm, int bc, float n[4])
{
#if defined(__CUDA_ARCH__)
    unsigned b0, b1, b2, b3;
    asm(DR_PM_PRED4 "@m ld.global.nc.u8 %0, [%5];\n\t@m ld.global.nc.u8 %1, [%5+4];\n\t@m ld.global.nc.u8 %2, [%5+2];\n\t"
        "@m ld.global.nc.u8 %3, [%5+6];\n\t@p ld.global.nc.u8 %0, [%4+1];\n\t@p ld.global.nc.u8 %1, [%4+5];\n\t"
        "@p ld.global.nc.u8 %2, [%4+3];\n\t@p ld.global.nc.u8 %3, [%4+7];\n\t}"
        : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "l"(rp), "l"(rm), "r"(bp), "r"(bm), "r"(bc));
    const F2 a = u8_unit2(b0 & 0xffu, b1 & 0xffu), b = u8_unit2(b2 & 0xffu, b3 & 0xffu);
    n[0] = a.x; n[1] = a.y; n[2] = b.x; n[3] = b.y;
#else
    const u8vox* r = bp != bc ? rp + 1 : rm;
    const bool any = bp != bc || bm != bc;
    n[0] = any ? u8_unit(r[0]) : 0.0f; n[1] = any ? u8_unit(r[4]) : 0.0f; n[2] = any ? u8_unit(r[2]) : 0.0f; n[3] = any ? u8_unit(r[6]) : 0.0f;
#endif
}
template <typename VT> struct VolView {
    const VT* p;
    DR_HD float ld(uoff off) const { return load_vox(p, off); }
};

// ---------------------------------------------------------------------------------------------------------
// ray set-up: compute_entry_exit :221-259, get_ray_direction :127-151, get_entry_exit_points :28-53
// ---------------------------------------------------------------------------------------------------------
struct Ray {
    F3 dir;
    float t0;       // entry + 0.5*len/n                                                     :273-275
    float texit;    // exit
    float inv_nm1;  // fl(1/(n-1)); 0 when n <= 1
    int n;          // sample_step_nums
};

DR_HD void setup_ray(const DrDesc& d, F3 cam, int i, int j, float jit, Ray& r)
{
    float cn = DR_DIV(1.0f, DR_SQRT(dot_e(cam, cam)));
    F3 view = { DR_MUL(-cam.x, cn), DR_MUL(-cam.y, cn), DR_MUL(-cam.z, cn) };          // :233
    float x = DR_DIV(DR_ADD((float)i, 0.5f), (float)d.W);                               // :239
    float y = DR_DIV(DR_ADD((float)j, 0.5f), (float)d.H);                               // :240
    float u = DR_SUB(x, 0.5f), v = DR_SUB(y, 0.5f);                                     // :140-141
    F3 up0 = { 0.0f, 1.0f, 0.0f };
    F3 right = normalized_e(cross_e(view, up0));                                        // :144
    F3 up = normalized_e(cross_e(right, view));                                         // :145
    float a = DR_MUL(u, d.near_w), b = DR_MUL(v, d.near_h);
    F3 nm = { DR_ADD(cam.x, DR_MUL(d.near_, view.x)), DR_ADD(cam.y, DR_MUL(d.near_, view.y)),
              DR_ADD(cam.z, DR_MUL(d.near_, view.z)) };                                 // :148
    F3 np = { DR_ADD(DR_ADD(nm.x, DR_MUL(a, right.x)), DR_MUL(b, up.x)),
              DR_ADD(DR_ADD(nm.y, DR_MUL(a, right.y)), DR_MUL(b, up.y)),
              DR_ADD(DR_ADD(nm.z, DR_MUL(a, right.z)), DR_MUL(b, up.z)) };              // :149
    F3 dd = { DR_SUB(np.x, cam.x), DR_SUB(np.y, cam.y), DR_SUB(np.z, cam.z) };
    r.dir = normalized_e(dd);                                                           // :151
    // slab test against [-1,1]^3                                                          :41-52
    float ix = DR_DIV(1.0f, r.dir.x), iy = DR_DIV(1.0f, r.dir.y), iz = DR_DIV(1.0f, r.dir.z);
    float t1 = DR_MUL(DR_SUB(-1.0f, cam.x), ix), t2 = DR_MUL(DR_SUB(1.0f, cam.x), ix);
    float t3 = DR_MUL(DR_SUB(-1.0f, cam.y), iy), t4 = DR_MUL(DR_SUB(1.0f, cam.y), iy);
    float t5 = DR_MUL(DR_SUB(-1.0f, cam.z), iz), t6 = DR_MUL(DR_SUB(1.0f, cam.z), iz);
    float tmin = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6));
    float tmax = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6));
    bool hit = !(tmax < 0.0f || tmin > tmax);
    float len = DR_SUB(tmax, tmin);                                                     // :250
    float nf = floorf(DR_MUL(DR_MUL(d.sr, len), d.vol_diag)) + 1.0f;                    // :251-253
    int n = hit ? (int)nf : 0;
    float entry = tmin;
    if ((d.flags & DR_F_HAS_JITTER) && n > 0) entry = DR_ADD(tmin, DR_DIV(DR_MUL(jit, len), nf));   // :254-255
    float len2 = DR_SUB(tmax, entry);                                                   // :272 (jittered entry)
    r.n = n;
    r.texit = tmax;
    r.t0 = (n > 0) ? DR_ADD(entry, DR_DIV(DR_MUL(0.5f, len2), (float)n)) : entry;       // :273-275
    r.inv_nm1 = (n > 1) ? DR_DIV(1.0f, (float)(n - 1)) : 0.0f;
}

// sample position :277-280.  float(s)/float(n-1) is evaluated as s * fl(1/(n-1)) (reciprocal-multiply, what a
// fast-math compiler emits for the reference's division; the oracle defines it the same way).  H3: n == 1 -> t = t0.
DR_HD F3 sample_pos(const Ray& r, F3 cam, int s, float& t)
{
    // n <= 1: inv_nm1 = 0, so q = 0 and mix_e(t0, texit, 1, 0) = fma(t0, 1, texit*0) = t0 exactly (no branch)
    const float q = DR_MUL((float)s, r.inv_nm1);
    t = mix_e(r.t0, r.texit, DR_SUB(1.0f, q), q);
    F3 p = { DR_FMA(t, r.dir.x, cam.x), DR_FMA(t, r.dir.y, cam.y), DR_FMA(t, r.dir.z, cam.z) };
    return p;
}

// ---------------------------------------------------------------------------------------------------------
// One sample: centre tap + the six normal taps of get_volume_normal (:191-203) with corner reuse.
// A tap whose cell equals the centre cell is the same trilinear polynomial at a shifted fraction, so only the
// mixes downstream of the shifted axis are redone (x: 7, y: 3, z: 1) on register-held values.  A tap that
// crosses a cell face (|dlo| == 1) fetches the 4 corners of the one new voxel plane (cell8 fp32, x axis: the 8
// corners of the tap's own cell record, one predicated 32-byte fetch -- cell_tap_x).  All give exactly the value
// a full 8-load trilinear evaluation would give (same operations, same order).
// The evaluation is split in two phases so that the forward can stop after the centre tap when the sample turns
// out to be exactly transparent: eval_centre (8 corner loads, 7 mixes) and eval_normals (the six taps).
// ---------------------------------------------------------------------------------------------------------
struct Centre {
    Loc cx, cy, cz;               // centre cell
    int cidx;                     // torch-linear index (cy*Z + cz)*X + cx of the centre cell's low corner
    // the 8 corners as pairs over z: A = voxels at x0, B = voxels at x1; suffix = y;  A0 = (v[x0,y0,z0], v[x0,y0,z1]) ...
    F2 A0, B0, A1, B1;
    F2 xm0, xm1;                  // x mixes at y0 and y1, each a pair over (z0, z1)
    F2 ym;                        // y mixes at (z0, z1)
    float I;                      // centre intensity
};
struct Taps {
    Loc cx, cy, cz;               // centre cell
    Loc xp, xm, yp, ym, zp, zm;   // tap cells (only the shifted axis differs from the centre)
    int cidx;
    float I;
    F3 g;                         // (f(x+d)-f(x-d), ...), un-normalised                     :197-202
};

// Volume storage layouts, as fetch policies.  A policy is initialised for one centre cell and provides, as pairs over z
// (x = the voxel at z0, y = the voxel at z1) unless stated:
//   centre(A0, B0, A1, B1)        the cell's 8 corners: A = x0, B = x1, suffix = y0 / y1
//   plane_x(plus, pred, N0, N1)   the voxel plane beyond the cell along x (x0+2 if plus else x0-1), predicated: N0 = row y0, N1 = row y1
//   plane_y(plus, pred, N0, N1)   likewise along y (y0+2 / y0-1): N0 = x0, N1 = x1
//   plane_z(plus, pred, N0, N1)   likewise along z (z0+2 / z0-1): N0 = x0, N1 = x1, each a pair over (y0, y1)
// No index clamps: in the corner-reuse path lo <= dim-2 on every axis (scale < dim-1 for dims <= 2000) and a plane beyond
// the cell is only fetched (pred) by a tap that crossed into it, so it is in range by construction.
//   LAYOUT_LINEAR reads the caller's contiguous torch tensor [y][z][x] in place (zero copy): a row is a pointer
//     (one IMAD.WIDE.U32), x-neighbours are immediate offsets.  8 scalar loads per cell.
//   LAYOUT_BRICK8 reads the 8x8x8-bricked copy made by dr_brick_volume (separable offsets offx/offy/offz).
//   LAYOUT_CELL8  reads the cell-major copy made by dr_expand_cells: record `cell` (the torch-linear index of the cell's
//     low corner) holds the cell's 8 corners contiguously (32 bytes fp32 = one sector, 16 bytes fp16), slot c + 2a + 4b =
//     voxel (x0+a, y0+b, z0+c), so a cell is ONE address and two (fp16: one) 16-byte loads that land directly in the
//     register pairs above -- instead of four rows and eight 4-byte loads that touch ~5 sectors each across a warp: the
//     L1 data pipe, which bounds the forward in the linear layout, sees a quarter of the wavefronts.  A plane beyond the
//     cell is half (y) / two quarters (x) / four eighths (z) of the neighbour's record.  Costs 8x the volume's bytes.
enum { LAYOUT_LINEAR = 0, LAYOUT_BRICK8 = 1, LAYOUT_CELL8 = 2 };
// How the six normal taps are evaluated (chosen per call from the volume dims, tap_mode() below):
//   TAPS_ONE      corner reuse; at most one tap of an axis can leave the centre cell (tap offset < 1/2 voxel: dims <= ~1000)
//   TAPS_TWO      corner reuse; both taps of an axis can leave it (1/2 <= offset < 1 voxel: dims <= ~2000)
//   TAPS_GENERIC  a tap can skip a whole cell: seven full 8-load trilinear evaluations (linear layout only)
enum { TAPS_ONE = 0, TAPS_TWO = 1, TAPS_GENERIC = 2 };
inline int tap_mode(const DrDesc& d)
{
    if (d.tap_generic) return TAPS_GENERIC;
    const float m = fmaxf(d.scale[0], fmaxf(d.scale[1], d.scale[2]));
    return (0.5f * d.delta * m >= 0.499f) ? TAPS_TWO : TAPS_ONE;
}

// rows of voxels: what the linear and bricked layouts have in common
template <typename Derived> struct RowFetch {
    static constexpr bool kRecordTapsX = false;
    DR_HD const Derived& self() const { return *static_cast<const Derived*>(this); }
    DR_HD void centre(F2& A0, F2& B0, F2& A1, F2& B1) const
    {
        const Derived& t = self();
        const typename Derived::Row r00 = t.row(0, 0), r10 = t.row(1, 0), r01 = t.row(0, 1), r11 = t.row(1, 1);
        A0 = f2(t.ld(r00, 0), t.ld(r01, 0)); B0 = f2(t.ld(r00, 1), t.ld(r01, 1));
        A1 = f2(t.ld(r10, 0), t.ld(r11, 0)); B1 = f2(t.ld(r10, 1), t.ld(r11, 1));
    }
    DR_HD void plane_x(int bp, int bm, int bc, F2& N0, F2& N1) const
    {
        const bool plus = bp != bc, pred = plus | (bm != bc);
        const Derived& t = self();
        const typename Derived::Row r00 = t.row(0, 0), r10 = t.row(1, 0), r01 = t.row(0, 1), r11 = t.row(1, 1);
        N0 = f2(t.ld_sel(r00, plus, pred), t.ld_sel(r01, plus, pred));
        N1 = f2(t.ld_sel(r10, plus, pred), t.ld_sel(r11, plus, pred));
    }
    DR_HD void plane_y(int bp, int bm, int bc, F2& N0, F2& N1) const
    {
        const bool plus = bp != bc, pred = plus | (bm != bc);
        const Derived& t = self();
        const typename Derived::Row n0 = t.row_y(plus, 0), n1 = t.row_y(plus, 1);
        N0 = f2(t.ld_if(n0, 0, pred), t.ld_if(n1, 0, pred)); N1 = f2(t.ld_if(n0, 1, pred), t.ld_if(n1, 1, pred));
    }
    DR_HD void plane_z(int bp, int bm, int bc, F2& N0, F2& N1) const
    {
        const bool plus = bp != bc, pred = plus | (bm != bc);
        const Derived& t = self();
        const typename Derived::Row n0 = t.row_z(0, plus), n1 = t.row_z(1, plus);
        N0 = f2(t.ld_if(n0, 0, pred), t.ld_if(n1, 0, pred)); N1 = f2(t.ld_if(n0, 1, pred), t.ld_if(n1, 1, pred));
    }
};
template <typename VT> struct LinearAddr : RowFetch<LinearAddr<VT> > {
    typedef const VT* Row;
    const VT* vp; uoff sy, sz, i00;
#if defined(DR_BOUNDS_CHECK)
    long long n_elems;
#endif
    DR_HD void init(const DrDesc& d, const VT* p, const Layout& L, const Centre& c)
    {
        vp = p; sz = (uoff)d.X; sy = (uoff)L.ycells; i00 = (uoff)c.cidx;
#if defined(DR_BOUNDS_CHECK)
        n_elems = (long long)d.X * d.Y * d.Z;
#endif
    }
    DR_HD Row row(int yi, int zi) const { return ptr_add(vp, i00 + (uoff)yi * sy + (uoff)zi * sz); }   // wraps correctly for -1
    DR_HD Row row_y(bool plus, int zi) const { return ptr_add(vp, i00 + (plus ? (uoff)2 : (uoff)-1) * sy + (uoff)zi * sz); }
    DR_HD Row row_z(int yi, bool plus) const { return ptr_add(vp, i00 + (uoff)yi * sy + (plus ? (uoff)2 : (uoff)-1) * sz); }
    DR_HD float ld(Row r, int xi) const
    {
#if defined(DR_BOUNDS_CHECK)
        DR_OOB_IF((r + xi) - vp < 0 || (r + xi) - vp >= n_elems);
#endif
        return load_vox(r, xi);
    }
    DR_HD float ld_if(Row r, int xi, bool pred) const
    {
#if defined(DR_BOUNDS_CHECK)
        DR_OOB_IF(pred && ((r + xi) - vp < 0 || (r + xi) - vp >= n_elems));
#endif
        return load_vox_if(r + xi, pred);
    }
    DR_HD float ld_sel(Row r, bool plus, bool pred) const { return ld_if(r + (plus ? 2 : -1), 0, pred); }
};
template <typename VT> struct BrickAddr : RowFetch<BrickAddr<VT> > {
    typedef uoff Row;
    const VT* vp; Layout L; int lx, ly, lz;
    DR_HD void init(const DrDesc&, const VT* p, const Layout& L_, const Centre& c)
    {
        vp = p; L = L_; lx = lo_of(c.cx); ly = lo_of(c.cy); lz = lo_of(c.cz);
    }
    DR_HD Row row(int yi, int zi) const { return offy(ly + yi, L.sY) + offz(lz + zi, L.sZ); }
    DR_HD Row row_y(bool plus, int zi) const { return offy(ly + (plus ? 2 : -1), L.sY) + offz(lz + zi, L.sZ); }
    DR_HD Row row_z(int yi, bool plus) const { return offy(ly + yi, L.sY) + offz(lz + (plus ? 2 : -1), L.sZ); }
    DR_HD float ld(Row r, int xi) const
    {
        DR_OOB_IF(lx + xi < 0 || lx + xi > L.mx || (long long)(r + offx(lx + xi)) >= (long long)L.sZ * (((L.mz + 8) >> 3)));
        return load_vox(vp, r + offx(lx + xi));
    }
    DR_HD float ld_if(Row r, int xi, bool pred) const
    {
        DR_OOB_IF(pred && (lx + xi < 0 || lx + xi > L.mx || (long long)(r + offx(lx + xi)) >= (long long)L.sZ * (((L.mz + 8) >> 3))));
        return load_vox_if(ptr_add(vp, r + offx(lx + xi)), pred);
    }
    DR_HD float ld_sel(Row r, bool plus, bool pred) const { return ld_if(r, plus ? 2 : -1, pred); }
};
// DUAL: fetch a plane beyond the cell with one load per neighbour (cell_plane_*: no address select, no materialised
// predicate) instead of one load from a selected address.  Twice the load instructions for fewer ALU instructions: it pays
// in the ALU-bound backward while at most one tap of an axis crosses (C3: backward +7 %), and costs where loads matter
// (forward -6 %) or crossings are frequent (C5, TAPS_TWO: backward -15 %), so only the TAPS_ONE backward uses it.
template <typename VT, int DUAL> struct CellAddr {
    static constexpr bool kRecordTapsX = sizeof(VT) == 4;      // x taps read their own cell's record (cell_tap_x)
    const VT* vp; uoff cell, sy, sz;
    template <bool PLUS> DR_HD void tap_x(int bt, int bc, F2& a0, F2& b0, F2& a1, F2& b1) const
    {
#if defined(DR_BOUNDS_CHECK)
        DR_OOB_IF(bt != bc && (long long)(PLUS ? cell + 1 : cell - 1) >= n_cells);
#endif
        cell_tap_x<PLUS>(reinterpret_cast<const float*>(rec_add(vp, cell)), bt, bc, a0, b0, a1, b1);
    }
#if defined(DR_BOUNDS_CHECK)
    long long n_cells;
#endif
    DR_HD void init(const DrDesc& d, const VT* p, const Layout& L, const Centre& c)
    {
        vp = p; cell = (uoff)c.cidx; sz = (uoff)d.X; sy = (uoff)L.ycells;
#if defined(DR_BOUNDS_CHECK)
        n_cells = (long long)d.X * d.Y * d.Z;
#endif
    }
    DR_HD const VT* rec(uoff c, bool pred) const
    {
#if defined(DR_BOUNDS_CHECK)
        DR_OOB_IF(pred && (long long)c >= n_cells);
#endif
        return rec_add(vp, c);
    }
    DR_HD void centre(F2& A0, F2& B0, F2& A1, F2& B1) const
    {
        float v[8];
        load_vox8(rec(cell, true), v);
        A0 = f2(v[0], v[1]); B0 = f2(v[2], v[3]); A1 = f2(v[4], v[5]); B1 = f2(v[6], v[7]);
    }
    // the records of the two face neighbours along an axis with cell stride `st`
    DR_HD void nbr(uoff st, int bp, int bm, int bc, const VT*& rp, const VT*& rm) const
    {
#if defined(DR_BOUNDS_CHECK)
        DR_OOB_IF(bp != bc && (long long)(cell + st) >= n_cells);
        DR_OOB_IF(bm != bc && (long long)(cell - st) >= n_cells);
#endif
        if (DUAL == 2) {
            // centre record +- the axis' byte stride (a 64-bit uniform operand): two adds per neighbour instead of the index add,
            // the widening multiply and the base add of rec_add(vp, cell +- st)  (7 issue slots per sample in the backward with a
            // volume gradient, +1 %; the TF-only backward lost 2.7 % with it and keeps the index form, DUAL == 1)
            const char* rc = reinterpret_cast<const char*>(rec_add(vp, cell));
            const long long off = (long long)st * (long long)(8 * sizeof(VT));
            rp = reinterpret_cast<const VT*>(rc + off); rm = reinterpret_cast<const VT*>(rc - off);
        } else {
            rp = rec_add(vp, cell + st); rm = rec_add(vp, cell - st);
        }
    }
    DR_HD void plane_x(int bp, int bm, int bc, F2& N0, F2& N1) const
    {
        // a = 1 quarters {2,3} {6,7} of the + neighbour, a = 0 quarters {0,1} {4,5} of the - neighbour
        float n[4];
        if (DUAL) {
#if defined(DR_BOUNDS_CHECK)
            DR_OOB_IF(bp != bc && (long long)(cell + 1) >= n_cells);
            DR_OOB_IF(bm != bc && (long long)(cell - 1) >= n_cells);
#endif
            cell_plane_x(rec_add(vp, cell), bp, bm, bc, n);
        } else {
            const bool plus = bp != bc, pred = plus | (bm != bc);
            const VT* r = rec(plus ? cell + 1 : cell - 1, pred) + (plus ? 2 : 0);
            load_vox2_if(r, pred, n); load_vox2_if(r + 4, pred, n + 2);
        }
        N0 = f2(n[0], n[1]); N1 = f2(n[2], n[3]);
    }
    DR_HD void plane_y(int bp, int bm, int bc, F2& N0, F2& N1) const
    {
        // the + neighbour's b = 1 half (slots 4..7 = row y0+2) or the - neighbour's b = 0 half (slots 0..3 = row y0-1)
        float n[4];
        if (DUAL) {
            const VT *rp, *rm;
            nbr(sy, bp, bm, bc, rp, rm);
            cell_plane_y(rp, rm, bp, bm, bc, n);
        } else {
            const bool plus = bp != bc, pred = plus | (bm != bc);
            load_vox4_if(rec(plus ? cell + sy : cell - sy, pred) + (plus ? 4 : 0), pred, n);
        }
        N0 = f2(n[0], n[1]); N1 = f2(n[2], n[3]);
    }
    DR_HD void plane_z(int bp, int bm, int bc, F2& N0, F2& N1) const
    {
        // c = 1 slots {1,5} {3,7} of the + neighbour, c = 0 slots {0,4} {2,6} of the - neighbour
        float n[4];
        if (DUAL) {
            const VT *rp, *rm;
            nbr(sz, bp, bm, bc, rp, rm);
            cell_plane_z(rp, rm, bp, bm, bc, n);
        } else {
            const bool plus = bp != bc, pred = plus | (bm != bc);
            const VT* r = rec(plus ? cell + sz : cell - sz, pred) + (plus ? 1 : 0);
            n[0] = load_vox_if(r, pred); n[1] = load_vox_if(r + 4, pred); n[2] = load_vox_if(r + 2, pred); n[3] = load_vox_if(r + 6, pred);
        }
        N0 = f2(n[0], n[1]); N1 = f2(n[2], n[3]);
    }
};
template <typename VT, int LAYOUT, int DUAL> struct AddrOf { typedef LinearAddr<VT> type; };
template <typename VT, int DUAL> struct AddrOf<VT, LAYOUT_BRICK8, DUAL> { typedef BrickAddr<VT> type; };
template <typename VT, int DUAL> struct AddrOf<VT, LAYOUT_CELL8, DUAL> { typedef CellAddr<VT, DUAL> type; };

// cell-index bias of locate_centre; computed on the host and passed to the kernels as an argument (a constant-bank operand:
// left to the device, the compiler re-derived it inside the march loop, 5 issue slots per sample)
inline LayoutConsts layout_consts(const DrDesc& d)
{
    LayoutConsts c;
    c.cbias = ((unsigned)kFloorBias * (unsigned)d.Z + (unsigned)kFloorBias) * (unsigned)d.X + (unsigned)kFloorBias;
    c.ycells = (unsigned)d.X * (unsigned)d.Z;
    return c;
}
// `ycells_from_args` = false derives X*Z on the device as before: measured on B200 (profiles/r02_experiments.md) the argument helps
// the backward (+1 %) and the forward kernels that carry the skip grid (+3..4 %: fewer spilled bytes at 80 registers) and costs the
// every-sample-shaded forward 4.5 % (same instruction counts, another schedule), so each kernel picks its own
DR_HD Layout make_layout(const DrDesc& d, LayoutConsts lc, bool ycells_from_args = true)
{
    Layout L;
    L.sY = d.nbx * 512; L.sZ = d.nbx * d.nby * 512;
    L.mx = d.X - 1; L.my = d.Y - 1; L.mz = d.Z - 1;
    L.cbias = lc.cbias; L.ycells = ycells_from_args ? lc.ycells : (unsigned)(d.X * d.Z);
    return L;
}

DR_HD void locate_centre(const DrDesc& d, const Layout& L, F3 pos, Centre& c)
{
    // x and y as one packed pair (same operations per lane as locate()), z scalar
    const F2 p = mul2(f2(DR_SAT(DR_FMA(0.5f, pos.x, 0.5f)), DR_SAT(DR_FMA(0.5f, pos.y, 0.5f))), f2(d.scale[0], d.scale[1]));
#if defined(__CUDA_ARCH__)
    F2 r;
    asm("{\n\t.reg .b64 pa, pb;\n\tmov.b64 pa, {%2, %3};\n\tmov.b64 pb, {%4, %4};\n\tadd.rm.f32x2 pa, pa, pb;\n\tmov.b64 {%0, %1}, pa;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(p.x), "f"(p.y), "f"(8388608.0f));
    c.cx.b = __float_as_int(r.x); c.cy.b = __float_as_int(r.y);
    const F2 f = sub2(p, sub2(r, splat(8388608.0f)));
    c.cx.f = f.x; c.cy.f = f.y;
#else
    float l = floor_pos(p.x, c.cx.b);
    c.cx.f = DR_SUB(p.x, l);
    l = floor_pos(p.y, c.cy.b);
    c.cy.f = DR_SUB(p.y, l);
#endif
    c.cz = locate(pos.z, d.scale[2]);
    // (lo_y*Z + lo_z)*X + lo_x from the biased indices: the three bias subtractions fold into one constant (mod 2^32)
    c.cidx = (int)(((unsigned)c.cy.b * (unsigned)d.Z + (unsigned)c.cz.b) * (unsigned)d.X + (unsigned)c.cx.b - L.cbias);
}

// centre tap: x mixes, y mixes, z mix                                                      :173-189
template <typename A>
DR_HD void eval_centre(const A& ad, Centre& c)
{
    ad.centre(c.A0, c.B0, c.A1, c.B1);
    const float fx = c.cx.f, fy = c.cy.f, fz = c.cz.f;
    const F2 fx2 = splat(fx), ox2 = splat(DR_SUB(1.0f, fx));
    c.xm0 = mix2(c.A0, c.B0, ox2, fx2);
    c.xm1 = mix2(c.A1, c.B1, ox2, fx2);
    c.ym = mix2(c.xm0, c.xm1, splat(DR_SUB(1.0f, fy)), splat(fy));
    c.I = mix_e(c.ym.x, c.ym.y, DR_SUB(1.0f, fz), fz);
}

// The six normal taps on top of an evaluated centre, branch-free in the common case: per axis the ONE voxel plane beyond
// the centre cell that a crossed tap needs (lo+2 for the + tap, lo-1 for the - tap) is fetched with predicated loads and
// both taps pick their operands with selects -- the same mixes on the same values a branch per tap would do, but a warp
// does not serialise through six divergent branches per sample (almost every warp has a lane that crosses:
// P = 1-(1-0.13)^32 at 256^3).  Both taps of an axis leave the cell only when the tap offset exceeds half a voxel
// (dims > 1000, TAPS_TWO); that case takes a branch for the second plane.  Mixes run as packed pairs (F2).
template <int TAPS, typename A>
DR_HD void eval_normals(const DrDesc& d, const A& ad, F3 pos, const Centre& c, Taps& t)
{
    t.cx = c.cx; t.cy = c.cy; t.cz = c.cz; t.cidx = c.cidx; t.I = c.I;
    locate_pair(pos.x, d.delta, d.scale[0], t.xp, t.xm);
    locate_pair(pos.y, d.delta, d.scale[1], t.yp, t.ym);
    locate_pair(pos.z, d.delta, d.scale[2], t.zp, t.zm);
    const float fx = c.cx.f, fy = c.cy.f, fz = c.cz.f;
    const float oz = DR_SUB(1.0f, fz);
    const F2 fx2 = splat(fx), ox2 = splat(DR_SUB(1.0f, fx)), fy2 = splat(fy), oy2 = splat(DR_SUB(1.0f, fy));
    {   // ---- z taps: only the last mix changes; both taps in one packed mix
        const bool cp = t.zp.b != c.cz.b, cm = t.zm.b != c.cz.b;
        F2 N0, N1;
        ad.plane_z(t.zp.b, t.zm.b, c.cz.b, N0, N1);
        const F2 xn = mix2(N0, N1, ox2, fx2);                               // x mixes of the new plane at (y0, y1)
        const float yn = mix_e(xn.x, xn.y, oy2.x, fy);
        const F2 f = f2(t.zp.f, t.zm.f), o = sub2(splat(1.0f), f);
        F2 v = mix2(f2(cp ? c.ym.y : c.ym.x, cm ? yn : c.ym.x), f2(cp ? yn : c.ym.y, cm ? c.ym.x : c.ym.y), o, f);
        if (TAPS == TAPS_TWO && (cp & cm)) {    // yn is plane lo+2; the - tap needs plane lo-1
            ad.plane_z(c.cz.b, t.zm.b, c.cz.b, N0, N1);
            const F2 xq = mix2(N0, N1, ox2, fx2);
            v.y = mix_e(mix_e(xq.x, xq.y, oy2.x, fy), c.ym.x, o.y, f.y);
        }
        t.g.z = DR_SUB(v.x, v.y);
    }
    {   // ---- y taps: y mixes (a pair over z per tap) and the z mix change
        const bool cp = t.yp.b != c.cy.b, cm = t.ym.b != c.cy.b;
        F2 N0, N1;
        ad.plane_y(t.yp.b, t.ym.b, c.cy.b, N0, N1);
        const F2 m = mix2(N0, N1, ox2, fx2);                                // x mixes of the new row at (z0, z1)
        const F2 fp = splat(t.yp.f), op = splat(DR_SUB(1.0f, t.yp.f)), fm = splat(t.ym.f), om = splat(DR_SUB(1.0f, t.ym.f));
        const F2 yp = mix2(sel2_ne(t.yp.b, c.cy.b, c.xm1, c.xm0), sel2_ne(t.yp.b, c.cy.b, m, c.xm1), op, fp);
        F2 ym = mix2(sel2_ne(t.ym.b, c.cy.b, m, c.xm0), sel2_ne(t.ym.b, c.cy.b, c.xm0, c.xm1), om, fm);
        if (TAPS == TAPS_TWO && (cp & cm)) {    // m is row lo+2; the - tap needs row lo-1
            ad.plane_y(c.cy.b, t.ym.b, c.cy.b, N0, N1);
            ym = mix2(mix2(N0, N1, ox2, fx2), c.xm0, om, fm);
        }
        t.g.y = DR_SUB(mix_e(yp.x, yp.y, oz, fz), mix_e(ym.x, ym.y, oz, fz));
    }
    if constexpr (A::kRecordTapsX) {   // ---- x taps, cell-major fp32 records: each tap mixes the corners of its own cell
        const F2 fp = splat(t.xp.f), op = splat(DR_SUB(1.0f, t.xp.f)), fm = splat(t.xm.f), om = splat(DR_SUB(1.0f, t.xm.f));
        F2 a0 = c.A0, b0 = c.B0, a1 = c.A1, b1 = c.B1;
        ad.template tap_x<false>(t.xm.b, c.cx.b, a0, b0, a1, b1);
        const F2 m0 = mix2(a0, b0, om, fm), m1 = mix2(a1, b1, om, fm);
        a0 = c.A0; b0 = c.B0; a1 = c.A1; b1 = c.B1;
        ad.template tap_x<true>(t.xp.b, c.cx.b, a0, b0, a1, b1);
        const F2 p0 = mix2(a0, b0, op, fp), p1 = mix2(a1, b1, op, fp);
        const F2 yp = mix2(p0, p1, oy2, fy2), ym = mix2(m0, m1, oy2, fy2);
        t.g.x = DR_SUB(mix_e(yp.x, yp.y, oz, fz), mix_e(ym.x, ym.y, oz, fz));
    } else {   // ---- x taps: everything downstream of the corners changes
        const bool cp = t.xp.b != c.cx.b, cm = t.xm.b != c.cx.b;
        F2 N0, N1;
        ad.plane_x(t.xp.b, t.xm.b, c.cx.b, N0, N1);
        const F2 fp = splat(t.xp.f), op = splat(DR_SUB(1.0f, t.xp.f)), fm = splat(t.xm.f), om = splat(DR_SUB(1.0f, t.xm.f));
        const int bp = t.xp.b, bm = t.xm.b, bc = c.cx.b;
        const F2 p0 = mix2(sel2_ne(bp, bc, c.B0, c.A0), sel2_ne(bp, bc, N0, c.B0), op, fp);       // row y0, pair over z
        const F2 p1 = mix2(sel2_ne(bp, bc, c.B1, c.A1), sel2_ne(bp, bc, N1, c.B1), op, fp);       // row y1
        F2 m0 = mix2(sel2_ne(bm, bc, N0, c.A0), sel2_ne(bm, bc, c.A0, c.B0), om, fm);
        F2 m1 = mix2(sel2_ne(bm, bc, N1, c.A1), sel2_ne(bm, bc, c.A1, c.B1), om, fm);
        if (TAPS == TAPS_TWO && (cp & cm)) {    // N0, N1 are plane lo+2; the - tap needs plane lo-1
            ad.plane_x(c.cx.b, t.xm.b, c.cx.b, N0, N1);
            m0 = mix2(N0, c.A0, om, fm);
            m1 = mix2(N1, c.A1, om, fm);
        }
        const F2 yp = mix2(p0, p1, oy2, fy2), ym = mix2(m0, m1, oy2, fy2);
        t.g.x = DR_SUB(mix_e(yp.x, yp.y, oz, fz), mix_e(ym.x, ym.y, oz, fz));
    }
}

// Direct taps (cell-major records in the two-neighbour regime, 1000 < dim <= 2000: C5).  With the tap offset above half a voxel
// a tap leaves the centre cell about every other time, on every axis, so the corner-reuse path above spends its time on
// predicated plane fetches, operand selects and the divergent "both taps crossed" branches.  Here each of the six taps simply
// evaluates the trilinear polynomial of ITS OWN cell -- the record at cell + (lo_tap - lo_centre) * stride: one 16-byte (fp16) or
// two 16-byte (fp32) loads that mostly hit L1 -- with the same seven mixes in the same order on the same eight corners, so the
// values are bit-identical to the other paths; there is no branch, no select and no second-plane case.
template <typename VT>
DR_HD F2 tap_rows(const VT* rec, float fx, float fy)      // x then y mixes of one cell record: the pair (value at z0, at z1)
{
    float v[8];
    load_vox8(rec, v);
    const F2 fx2 = splat(fx), ox2 = splat(DR_SUB(1.0f, fx));
    const F2 m0 = mix2(f2(v[0], v[1]), f2(v[2], v[3]), ox2, fx2), m1 = mix2(f2(v[4], v[5]), f2(v[6], v[7]), ox2, fx2);
    return mix2(m0, m1, splat(DR_SUB(1.0f, fy)), splat(fy));
}
template <typename VT>
DR_HD void eval_normals_direct(const DrDesc& d, const Layout& L, const VT* vp, F3 pos, const Centre& c, Taps& t)
{
    t.cx = c.cx; t.cy = c.cy; t.cz = c.cz; t.cidx = c.cidx; t.I = c.I;
    locate_pair(pos.x, d.delta, d.scale[0], t.xp, t.xm);
    locate_pair(pos.y, d.delta, d.scale[1], t.yp, t.ym);
    locate_pair(pos.z, d.delta, d.scale[2], t.zp, t.zm);
    // a tap's record = the centre record + (lo_tap - lo_centre) x the axis' byte stride: one widening multiply-add per tap on the centre
    // POINTER (the strides fit 32 bits: X*Z*32 < 2^31 for axes <= 2000) instead of an index multiply-add, a widening multiply and the
    // base add
    const char* rc = reinterpret_cast<const char*>(rec_add(vp, (uoff)c.cidx));
    const int rb = (int)(8 * sizeof(VT)), zb = d.X * rb, yb = (int)L.ycells * rb;
#define DR_TAP_REC(db, stride) reinterpret_cast<const VT*>(rc + (long long)(db) * (long long)(stride))
    const float fx = c.cx.f, fy = c.cy.f, fz = c.cz.f, oz = DR_SUB(1.0f, fz);
    {   // x taps: cells cell + (b_tap - b_centre), x fraction of the tap
        const F2 p = tap_rows(DR_TAP_REC(t.xp.b - c.cx.b, rb), t.xp.f, fy), m = tap_rows(DR_TAP_REC(t.xm.b - c.cx.b, rb), t.xm.f, fy);
        const F2 v = mix2(f2(p.x, m.x), f2(p.y, m.y), splat(oz), splat(fz));
        t.g.x = DR_SUB(v.x, v.y);
    }
    {   // y taps
        const F2 p = tap_rows(DR_TAP_REC(t.yp.b - c.cy.b, yb), fx, t.yp.f), m = tap_rows(DR_TAP_REC(t.ym.b - c.cy.b, yb), fx, t.ym.f);
        const F2 v = mix2(f2(p.x, m.x), f2(p.y, m.y), splat(oz), splat(fz));
        t.g.y = DR_SUB(v.x, v.y);
    }
    {   // z taps: z fraction of the tap
        const F2 p = tap_rows(DR_TAP_REC(t.zp.b - c.cz.b, zb), fx, fy), m = tap_rows(DR_TAP_REC(t.zm.b - c.cz.b, zb), fx, fy);
        const F2 f = f2(t.zp.f, t.zm.f);
        const F2 v = mix2(f2(p.x, m.x), f2(p.y, m.y), sub2(splat(1.0f), f), f);
        t.g.z = DR_SUB(v.x, v.y);
    }
#undef DR_TAP_REC
}

// Generic path (a normal tap can skip a whole cell: dims > ~2000; linear layout only): every tap is a full 8-load
// trilinear evaluation with the reference's clamps hi = min(lo+1, dim-1)                    :170-172
template <typename VT>
DR_HD float trilinear_full_linear(const DrDesc& d, const VT* p, Loc ax, Loc ay, Loc az)
{
    const int xl = lo_of(ax), yl = lo_of(ay), zl = lo_of(az);
    const int x0 = xl, x1 = imin(xl + 1, d.X - 1);
    const int r00 = (yl * d.Z + zl) * d.X, r10 = (imin(yl + 1, d.Y - 1) * d.Z + zl) * d.X;
    const int r01 = (yl * d.Z + imin(zl + 1, d.Z - 1)) * d.X, r11 = (imin(yl + 1, d.Y - 1) * d.Z + imin(zl + 1, d.Z - 1)) * d.X;
    const float ox = DR_SUB(1.0f, ax.f), oy = DR_SUB(1.0f, ay.f), oz = DR_SUB(1.0f, az.f);
    float a = mix_e(load_vox(p, r00 + x0), load_vox(p, r00 + x1), ox, ax.f);
    float b = mix_e(load_vox(p, r10 + x0), load_vox(p, r10 + x1), ox, ax.f);
    const float lo = mix_e(a, b, oy, ay.f);
    a = mix_e(load_vox(p, r01 + x0), load_vox(p, r01 + x1), ox, ax.f);
    b = mix_e(load_vox(p, r11 + x0), load_vox(p, r11 + x1), ox, ax.f);
    const float hi = mix_e(a, b, oy, ay.f);
    return mix_e(lo, hi, oz, az.f);
}
template <typename VT>
DR_HD void eval_normals_generic(const DrDesc& d, const VT* vp, F3 pos, const Centre& c, Taps& t)
{
    t.cx = c.cx; t.cy = c.cy; t.cz = c.cz; t.cidx = c.cidx; t.I = c.I;
    const float dl = d.delta;
    t.xp = locate(DR_ADD(pos.x, dl), d.scale[0]); t.xm = locate(DR_SUB(pos.x, dl), d.scale[0]);
    t.yp = locate(DR_ADD(pos.y, dl), d.scale[1]); t.ym = locate(DR_SUB(pos.y, dl), d.scale[1]);
    t.zp = locate(DR_ADD(pos.z, dl), d.scale[2]); t.zm = locate(DR_SUB(pos.z, dl), d.scale[2]);
    t.g.x = DR_SUB(trilinear_full_linear(d, vp, t.xp, t.cy, t.cz), trilinear_full_linear(d, vp, t.xm, t.cy, t.cz));
    t.g.y = DR_SUB(trilinear_full_linear(d, vp, t.cx, t.yp, t.cz), trilinear_full_linear(d, vp, t.cx, t.ym, t.cz));
    t.g.z = DR_SUB(trilinear_full_linear(d, vp, t.cx, t.cy, t.zp), trilinear_full_linear(d, vp, t.cx, t.cy, t.zm));
}

// phase 1 / phase 2 dispatch on layout and tap path
template <typename VT, int LAYOUT, int TAPS, int DUAL = 0>
DR_HD void sample_centre(const DrDesc& d, const VolView<VT>& vol, const Layout& L, F3 pos, Centre& c)
{
    locate_centre(d, L, pos, c);
    if (TAPS == TAPS_GENERIC) { c.I = trilinear_full_linear(d, vol.p, c.cx, c.cy, c.cz); return; }
    typename AddrOf<VT, LAYOUT, DUAL>::type ad;
    ad.init(d, vol.p, L, c);
    eval_centre(ad, c);
}
template <typename VT, int LAYOUT, int TAPS, int DUAL>
DR_HD void sample_normals(const DrDesc& d, const VolView<VT>& vol, const Layout& L, F3 pos, const Centre& c, Taps& t)
{
    if (TAPS == TAPS_GENERIC) { eval_normals_generic(d, vol.p, pos, c, t); return; }
    if (TAPS == TAPS_TWO && LAYOUT == LAYOUT_CELL8) { eval_normals_direct(d, L, vol.p, pos, c, t); return; }     // C5: fwd +6.6 %, bwd +3.3 %
    typename AddrOf<VT, LAYOUT, DUAL>::type ad;
    ad.init(d, vol.p, L, c);
    eval_normals<TAPS>(d, ad, pos, c, t);
}

// ---------------------------------------------------------------------------------------------------------
// transfer function lookup :205-219.  The table is staged (shared memory on the device) as one 32-byte TfBin per
// bin r: a = tf[r], b = tf[min(r+1, R-1)] stored as (b.rgb - a.rgb, b.w).  Alpha is on the exact path and is evaluated
// as mix_e(a.w, b.w, 1-f, f) like the reference; the colour only feeds the shading and is a.rgb + f*(b.rgb - a.rgb)
// (one FMA per channel; the differences are what the adjoint needs anyway).
// ---------------------------------------------------------------------------------------------------------
struct DR_ALIGN16 TfBin { F4 a; float dx, dy, dz, bw; };

DR_HD TfBin make_tf_bin(F4 a, F4 b)
{
    TfBin t;
    t.a = a; t.dx = b.x - a.x; t.dy = b.y - a.y; t.dz = b.z - a.z; t.bw = b.w;
    return t;
}

// table accessor: a plain pointer (host harness) or, on the device, a 32-bit shared-memory address kept in a register
// (one LEA + two LDS.128 per lookup; a generic pointer costs three uniform-datapath instructions per lookup to re-form
// the shared window base)
struct TfTable {
#if defined(__CUDACC__)
    unsigned base;
    DR_HD TfBin get(int r) const
    {
        TfBin t;
#if defined(__CUDA_ARCH__)
        const unsigned a = base + ((unsigned)r << 5);
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.a.x), "=f"(t.a.y), "=f"(t.a.z), "=f"(t.a.w) : "r"(a));
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+16];" : "=f"(t.dx), "=f"(t.dy), "=f"(t.dz), "=f"(t.bw) : "r"(a));
#else
        t = TfBin();        // nvcc's host pass only: the kernels are the sole callers
#endif
        return t;
    }
#else
    const TfBin* p;
    DR_HD TfBin get(int r) const { return p[r]; }
#endif
};

struct TfHit { int lo; float f, x; F4 c; F4 d; TfBin t; };   // c = colour, d = tf[hi] - tf[lo] (for dI), t = the bin

// alpha only (the exact path); tf_colour() adds the colour and, for the adjoint, the bin differences
DR_HD void apply_tf(const DrDesc& d, const TfTable& tf, float intensity, TfHit& h)
{
    float x = fmaxf(DR_MUL(intensity, d.tf_len), 0.0f);
    h.x = x;
    int b;
    float l = floor_pos(x, b);
    h.f = DR_SUB(x, l);
    h.lo = imin(b - kFloorBias, d.R - 1);        // H9
    h.t = tf.get(h.lo);
    h.c.w = mix_e(h.t.a.w, h.t.bw, DR_SUB(1.0f, h.f), h.f);
}
DR_HD void tf_colour(TfHit& h, bool want_diff)
{
    const TfBin& t = h.t;
    h.c.x = t.a.x + h.f * t.dx; h.c.y = t.a.y + h.f * t.dy; h.c.z = t.a.z + h.f * t.dz;
    if (want_diff) { h.d.x = t.dx; h.d.y = t.dy; h.d.z = t.dz; h.d.w = t.bw - t.a.w; }
}

// opacity = 1 - pow(1 - alpha, 1/sr)                                                       :284-285
template <bool SR1>
DR_HD float opacity(const DrDesc& d, float alpha)
{
    float base = DR_SUB(1.0f, alpha);
    if (SR1) return DR_SUB(1.0f, base);
    return DR_SUB(1.0f, powf(base, d.inv_sr));
}

// ---------------------------------------------------------------------------------------------------------
// Phong factor :287-298.  Not on the exact path (only rgb depends on it).
// ---------------------------------------------------------------------------------------------------------
struct Shade {
    float inv_g;       // 1/|g| (0 when flat)
    float inv_l;       // 1/|pos - light|
    float nl, rv, p32, kraw, k;       // N.l, r.(-dir), max(rv,0)^32, un-clamped and clamped Phong factor
};

// Only scalars are formed: with L = pos - (cam + (0,1,0)) = t*dir - (0,1,0) and |dir| = 1,
//   L.dir = t - dir.y,  |L|^2 = t*(t - 2*dir.y) + 1,  g.L = t*(g.dir) - g.y,
//   N.l = (g.L) / (|g| |L|),   r.(-dir) = (l - 2 (N.l) N).(-dir) = -(L.dir)/|L| + 2 (N.l) (g.dir)/|g|,
// i.e. the normal N, the light direction l and the reflected vector r of :287-295 are never materialised (32 instead of 45
// instructions per sample; every sample of the backward needs k, only the shaded ones need N and l -- shade_vectors()).
DR_HD void shade(const DrDesc& d, F3 dir, float t, F3 g, bool clamp_k, Shade& s)
{
    const float g2 = g.x * g.x + g.y * g.y + g.z * g.z;
    const bool flat = !(g2 > 0.0f);               // H4: 0/0 normal -> ambient only
    const float ig = flat ? 0.0f : DR_RSQRT(g2);
    s.inv_g = ig;
    const float gd = g.x * dir.x + g.y * dir.y + g.z * dir.z;
    const float u = t - dir.y;                                                     // L.dir
    const float il = DR_RSQRT(t * (u - dir.y) + 1.0f);                             // light at cam + (0,1,0) :281
    s.inv_l = il;
    s.nl = (t * gd - g.y) * (ig * il);
    const float ndl = fmaxf(s.nl, 0.0f);                                           // :291
    s.rv = flat ? 0.0f : (2.0f * s.nl) * (ig * gd) - il * u;                       // :293-294; NaN normal -> max(NaN,0) = 0
    const float rdv = fmaxf(s.rv, 0.0f);                                           // :295
    float p2 = rdv * rdv, p4 = p2 * p2, p8 = p4 * p4, p16 = p8 * p8;
    s.p32 = p16 * p16;                                                             // shininess 32 :296
    s.kraw = d.diffuse * ndl + d.specular * s.p32 + d.ambient;
    s.k = clamp_k ? fminf(1.0f, s.kraw) : s.kraw;                                  // :298 vs :345
}
// the unit normal and light direction behind a Shade (the normal-path adjoint of a shaded sample needs them)
DR_HD void shade_vectors(const Shade& s, F3 dir, float t, F3 g, F3& N, F3& l)
{
    N.x = g.x * s.inv_g; N.y = g.y * s.inv_g; N.z = g.z * s.inv_g;
    const float ti = t * s.inv_l;
    l.x = ti * dir.x; l.y = ti * dir.y - s.inv_l; l.z = ti * dir.z;
}

// ---------------------------------------------------------------------------------------------------------
// Adjoint of one sample (Taichi reverse-mode of :281-302).
//   in : T = transmittance before the sample, g = dL/dA_s (rgb constant along the ray, w evolves)
//   out: dc (adjoint of the TF sample), dI (centre intensity), dgr (un-normalised gradient g)
// returns C.g so that the caller can update g.w
// ---------------------------------------------------------------------------------------------------------
struct SampleAdj { F4 dc; float dI; F3 dg; bool has_dg; };

template <bool SR1>
DR_HD float sample_adjoint(const DrDesc& d, F3 dir, float t, F3 grad, const TfHit& h, float o, const Shade& s, float T, F4 g,
                           bool want_vol, SampleAdj& a)
{
    const float ko = s.k * o;
    const float Cg = ko * (h.c.x * g.x + h.c.y * g.y + h.c.z * g.z) + o * g.w;     // C . g
    const float cd = T * (h.c.x * g.x + h.c.y * g.y + h.c.z * g.z);                // c.rgb . dC.rgb
    const float d_o = s.k * cd + T * g.w;
    const float dk = o * cd;
    float dpow = 1.0f;
    if (!SR1) dpow = d.inv_sr * powf(fmaxf(1.0f - h.c.w, 1e-12f), d.inv_sr - 1.0f);   // H8
    const float kT = ko * T;
    a.dc.x = kT * g.x; a.dc.y = kT * g.y; a.dc.z = kT * g.z; a.dc.w = d_o * dpow;
    a.has_dg = false;
    if (want_vol) {
        const float df = a.dc.x * h.d.x + a.dc.y * h.d.y + a.dc.z * h.d.z + a.dc.w * h.d.w;
        a.dI = (h.x > 0.0f) ? df * d.tf_len : 0.0f;
        a.dg.x = a.dg.y = a.dg.z = 0.0f;
        // dk == 0 (exactly transparent sample, or zero incoming gradient) makes every normal-path term exactly zero
        if (dk != 0.0f && s.inv_g > 0.0f && !(1.0f < s.kraw)) {
            const float d_ndl = (s.nl > 0.0f) ? d.diffuse * dk : 0.0f;
            // d/d(rdv) rdv^32 = 32 rdv^31 = 32 p32 / rdv; below 1e-6 rdv^31 is exactly 0 in fp32 anyway
            const float d_rdv = (s.rv > 1e-6f) ? d.specular * 32.0f * (s.p32 * fast_rcp(s.rv)) * dk : 0.0f;
            F3 N, l;
            shade_vectors(s, dir, t, grad, N, l);
            const float drx = -dir.x * d_rdv, dry = -dir.y * d_rdv, drz = -dir.z * d_rdv;
            const float drN = drx * N.x + dry * N.y + drz * N.z;
            const float cN = d_ndl - 2.0f * drN, c2 = 2.0f * s.nl;
            const float dNx = cN * l.x - c2 * drx, dNy = cN * l.y - c2 * dry, dNz = cN * l.z - c2 * drz;
            const float NdN = N.x * dNx + N.y * dNy + N.z * dNz;
            a.dg.x = (dNx - N.x * NdN) * s.inv_g;
            a.dg.y = (dNy - N.y * NdN) * s.inv_g;
            a.dg.z = (dNz - N.z * NdN) * s.inv_g;
            a.has_dg = true;
        }
    }
    return Cg;
}

// ---------------------------------------------------------------------------------------------------------
// Volume-gradient scatter into a CELL-MAJOR buffer: gcell[cell][8], cell = (cy*Z + cz)*X + cx (the torch linear
// index of the cell's low corner), slot = a + 2b + 4c for corner (cx+a, cy+b, cz+c).  One cell is one 32-byte
// sector, so a tap's 8 trilinear weights go out as two 16-byte vector reductions (RED.E.ADD.F32x4) instead of 8
// scalar ones -- measured 3.5x faster on B200 (profiles/r01_atomic_microbench.txt).  A gather pass
// (gather_grad_kernel) sums the 8 slots that alias each voxel.
//   * taps that stay in the centre cell are merged with the centre tap into one 8-vector, added into the vector that
//     Sink::open(cell) returns and committed by Sink::close() -- a sink may keep that vector in registers while
//     consecutive samples stay in the same cell;
//   * a tap that crossed a face goes to its own (neighbour) cell (Sink::direct).
// ---------------------------------------------------------------------------------------------------------
DR_HD int cell_index(const DrDesc& d, int cx, int cy, int cz) { return (cy * d.Z + cz) * d.X + cx; }

DR_HD void tap_weights(float adj, Loc ax, Loc ay, Loc az, float v[8])
{
    const float z0 = adj * (1.0f - az.f), z1 = adj * az.f;
    const float wx0 = 1.0f - ax.f, wx1 = ax.f, wy0 = 1.0f - ay.f, wy1 = ay.f;
    const float a00 = z0 * wy0, a10 = z0 * wy1, a01 = z1 * wy0, a11 = z1 * wy1;
    v[0] = a00 * wx0; v[1] = a00 * wx1; v[2] = a10 * wx0; v[3] = a10 * wx1;
    v[4] = a01 * wx0; v[5] = a01 * wx1; v[6] = a11 * wx0; v[7] = a11 * wx1;
}

template <typename Sink, bool GENERIC, bool TWO>
DR_HD void scatter_volume_grad(const DrDesc& d, const Layout& L, Sink& sink, const Taps& t, const SampleAdj& a)
{
    float v[8];
    const int cc = t.cidx;
    if (GENERIC) {
        const int cx = lo_of(t.cx), cy = lo_of(t.cy), cz = lo_of(t.cz);
        tap_weights(a.dI, t.cx, t.cy, t.cz, v); sink.direct(cc, v);
        if (a.has_dg) {
            tap_weights(a.dg.x, t.xp, t.cy, t.cz, v); sink.direct(cell_index(d, lo_of(t.xp), cy, cz), v);
            tap_weights(-a.dg.x, t.xm, t.cy, t.cz, v); sink.direct(cell_index(d, lo_of(t.xm), cy, cz), v);
            tap_weights(a.dg.y, t.cx, t.yp, t.cz, v); sink.direct(cell_index(d, cx, lo_of(t.yp), cz), v);
            tap_weights(-a.dg.y, t.cx, t.ym, t.cz, v); sink.direct(cell_index(d, cx, lo_of(t.ym), cz), v);
            tap_weights(a.dg.z, t.cx, t.cy, t.zp, v); sink.direct(cell_index(d, cx, cy, lo_of(t.zp)), v);
            tap_weights(-a.dg.z, t.cx, t.cy, t.zm, v); sink.direct(cell_index(d, cx, cy, lo_of(t.zm)), v);
        }
        return;
    }
    const float wx0 = 1.0f - t.cx.f, wx1 = t.cx.f, wy0 = 1.0f - t.cy.f, wy1 = t.cy.f;
    const float wz0 = 1.0f - t.cz.f, wz1 = t.cz.f;
    // per axis: coefficients of the two voxel planes of the centre cell contributed by the taps that did NOT cross
    const bool xpc = t.xp.b != t.cx.b, xmc = t.xm.b != t.cx.b;
    const bool ypc = t.yp.b != t.cy.b, ymc = t.ym.b != t.cy.b;
    const bool zpc = t.zp.b != t.cz.b, zmc = t.zm.b != t.cz.b;
    const float ex0 = a.dg.x * ((xpc ? 0.0f : 1.0f - t.xp.f) - (xmc ? 0.0f : 1.0f - t.xm.f));
    const float ex1 = a.dg.x * ((xpc ? 0.0f : t.xp.f) - (xmc ? 0.0f : t.xm.f));
    const float ey0 = a.dg.y * ((ypc ? 0.0f : 1.0f - t.yp.f) - (ymc ? 0.0f : 1.0f - t.ym.f));
    const float ey1 = a.dg.y * ((ypc ? 0.0f : t.yp.f) - (ymc ? 0.0f : t.ym.f));
    const float ez0 = a.dg.z * ((zpc ? 0.0f : 1.0f - t.zp.f) - (zmc ? 0.0f : 1.0f - t.zm.f));
    const float ez1 = a.dg.z * ((zpc ? 0.0f : t.zp.f) - (zmc ? 0.0f : t.zm.f));
    const float X0 = a.dI * wx0 + ex0, X1 = a.dI * wx1 + ex1;
    const float yz00 = wy0 * wz0, yz10 = wy1 * wz0, yz01 = wy0 * wz1, yz11 = wy1 * wz1;
    const float xz00 = wx0 * wz0, xz10 = wx1 * wz0, xz01 = wx0 * wz1, xz11 = wx1 * wz1;
    const float xy00 = wx0 * wy0, xy10 = wx1 * wy0, xy01 = wx0 * wy1, xy11 = wx1 * wy1;
    // G[a][b][c] = wyz[b][c]*X_a + wxz[a][c]*ey_b + wxy[a][b]*ez_c
    float* acc = sink.open(cc);         // three FMAs per slot, straight into the sink's (register-held) vector
    acc[0] = yz00 * X0 + (xz00 * ey0 + (xy00 * ez0 + acc[0]));
    acc[1] = yz00 * X1 + (xz10 * ey0 + (xy10 * ez0 + acc[1]));
    acc[2] = yz10 * X0 + (xz00 * ey1 + (xy01 * ez0 + acc[2]));
    acc[3] = yz10 * X1 + (xz10 * ey1 + (xy11 * ez0 + acc[3]));
    acc[4] = yz01 * X0 + (xz01 * ey0 + (xy00 * ez1 + acc[4]));
    acc[5] = yz01 * X1 + (xz11 * ey0 + (xy10 * ez1 + acc[5]));
    acc[6] = yz11 * X0 + (xz01 * ey1 + (xy01 * ez1 + acc[6]));
    acc[7] = yz11 * X1 + (xz11 * ey1 + (xy11 * ez1 + acc[7]));
    sink.close();
    if (!a.has_dg) return;
    // a crossed tap lives in the face-neighbour cell: +-1 (x), +-X (z), +-X*Z (y) in the torch-linear cell order.  Its 8
    // weights reuse the centre's pair products: only the weight pair of the shifted axis differs.  One block per AXIS (the
    // crossed tap is picked with selects; in a shaded warp some lane nearly always crossed on each axis, so a block per
    // tap would cost twice the issue slots); both taps of an axis cross only under TAPS_TWO (second pass of the loop).
    const int sz = d.X, sy = (int)L.ycells;
#pragma unroll
    for (int pass = 0; pass < (TWO ? 2 : 1); ++pass) {
        const bool fx = pass ? (xpc & xmc) : (xpc | xmc), fy = pass ? (ypc & ymc) : (ypc | ymc), fz = pass ? (zpc & zmc) : (zpc | zmc);
        if (fx) {
            const bool m = pass ? true : !xpc;          // first pass: the + tap if it crossed, else the - tap
            const float qf = m ? t.xm.f : t.xp.f, dg = m ? -a.dg.x : a.dg.x, q0 = dg * (1.0f - qf), q1 = dg * qf;
            v[0] = q0 * yz00; v[1] = q1 * yz00; v[2] = q0 * yz10; v[3] = q1 * yz10;
            v[4] = q0 * yz01; v[5] = q1 * yz01; v[6] = q0 * yz11; v[7] = q1 * yz11;
            sink.direct(m ? cc - 1 : cc + 1, v);
        }
        if (fy) {
            const bool m = pass ? true : !ypc;
            const float qf = m ? t.ym.f : t.yp.f, dg = m ? -a.dg.y : a.dg.y, q0 = dg * (1.0f - qf), q1 = dg * qf;
            v[0] = q0 * xz00; v[1] = q0 * xz10; v[2] = q1 * xz00; v[3] = q1 * xz10;
            v[4] = q0 * xz01; v[5] = q0 * xz11; v[6] = q1 * xz01; v[7] = q1 * xz11;
            sink.direct(m ? cc - sy : cc + sy, v);
        }
        if (fz) {
            const bool m = pass ? true : !zpc;
            const float qf = m ? t.zm.f : t.zp.f, dg = m ? -a.dg.z : a.dg.z, q0 = dg * (1.0f - qf), q1 = dg * qf;
            v[0] = q0 * xy00; v[1] = q0 * xy10; v[2] = q0 * xy01; v[3] = q0 * xy11;
            v[4] = q1 * xy00; v[5] = q1 * xy10; v[6] = q1 * xy01; v[7] = q1 * xy11;
            sink.direct(m ? cc - sz : cc + sz, v);
        }
    }
}

// Gather pass: the gradient of voxel (x,y,z) is the sum of every (cell, slot) that aliases it, i.e. all (c, a) per axis
// with min(c + a, dim-1) == v: (v,0), (v-1,1) and, on the last plane only, the clamped (dim-1,1)  (:170-172).
DR_HD float gather_voxel(const DrDesc& d, const float* gcell, int x, int y, int z)
{
    // per axis the candidates are, in this order: (cell v, slot 0), (cell v-1, slot 1) if v > 0, (cell v, slot 1) if v is the last plane
    float s = 0.0f;
    for (int k = 0; k < 3; ++k) {
        const int cz = k == 1 ? z - 1 : z, az = k == 0 ? 0 : 1;
        if (k == 1 ? z == 0 : (k == 2 && z != d.Z - 1)) continue;
        for (int j = 0; j < 3; ++j) {
            const int cy = j == 1 ? y - 1 : y, ay = j == 0 ? 0 : 1;
            if (j == 1 ? y == 0 : (j == 2 && y != d.Y - 1)) continue;
            for (int i = 0; i < 3; ++i) {
                const int cx = i == 1 ? x - 1 : x, ax = i == 0 ? 0 : 1;
                if (i == 1 ? x == 0 : (i == 2 && x != d.X - 1)) continue;
                s += gcell[(size_t)cell_index(d, cx, cy, cz) * 8 + (ax + 2 * ay + 4 * az)];
            }
        }
    }
    return s;
}

// ---------------------------------------------------------------------------------------------------------
// Exact empty-space skipping (forward).  The skip grid has one byte per MACRO-CELL (8x8x8 cells, the brick counts nbx,
// nby, nbz of the descriptor): 1 when every cell whose low corner lies in the macro-cell is exactly transparent under the
// call's transfer function, i.e. every TF bin that a trilinear value of the cell's corners can select has alpha 0 (built
// by skip_classify below from per-macro-cell voxel min/max with a rounding margin).  A sample in such a cell has TF alpha
// fma(0, 1-f, 0*f) = 0, hence o = 0 and A_s = A_{s-1} bit for bit: it only counts (K, :303) -- so a run of samples that
// stays inside one empty macro-cell is replaced by K += m.  m is conservative: the ray's exit from the macro-cell is
// computed in voxel space against a box shrunk by kSkipMargin voxels (fp32 position error is < 2e-4 voxel at 1024^3) and
// one more sample is left to the normal path.  Images, K and Tprev are bit-identical with and without the grid.
// ---------------------------------------------------------------------------------------------------------
constexpr float kSkipMargin = 0.02f;
#ifndef DR_MACRO_SHIFT
#define DR_MACRO_SHIFT 3            // macro-cells of 8 x 8 x 8 cells
#endif
constexpr int kMacroShift = DR_MACRO_SHIFT, kMacro = 1 << kMacroShift;
DR_HD int macro_nx(const DrDesc& d) { return (d.X + kMacro - 1) >> kMacroShift; }
DR_HD int macro_ny(const DrDesc& d) { return (d.Y + kMacro - 1) >> kMacroShift; }
DR_HD int macro_nz(const DrDesc& d) { return (d.Z + kMacro - 1) >> kMacroShift; }

DR_HD unsigned char load_u8(const unsigned char* p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

DR_HD size_t macro_index(const DrDesc& d, int cx, int cy, int cz)       // cell low corner -> macro-cell, [y][z][x] order
{
    return ((size_t)(cy >> kMacroShift) * macro_nz(d) + (cz >> kMacroShift)) * macro_nx(d) + (cx >> kMacroShift);
}

// number of consecutive samples s, s+1, ... (at least 1: sample s itself is known to be inside) whose cells stay in the
// macro-cell of sample s's cell (cx, cy, cz); never more than nn - s
DR_HD int skip_run(const DrDesc& d, const Ray& r, F3 cam, int s, int nn, int cx, int cy, int cz)
{
    if (r.n < 2) return 1;
    // voxel-space position along the ray: p_a(t) = P_a + t * D_a   (locate() without the clamp, which only acts on the faces)
    const float Px = (0.5f * cam.x + 0.5f) * d.scale[0], Dx = 0.5f * r.dir.x * d.scale[0];
    const float Py = (0.5f * cam.y + 0.5f) * d.scale[1], Dy = 0.5f * r.dir.y * d.scale[1];
    const float Pz = (0.5f * cam.z + 0.5f) * d.scale[2], Dz = 0.5f * r.dir.z * d.scale[2];
    const float bx = (float)(cx & ~(kMacro - 1)), by = (float)(cy & ~(kMacro - 1)), bz = (float)(cz & ~(kMacro - 1));
    // exit parameter per axis (the face the ray moves towards, shrunk by the margin); an axis the ray does not move along never exits
    float tout = 3.0e38f;
    // (approximate reciprocals: their 1e-7 relative error moves a position by far less than the margin)
    if (fabsf(Dx) > 1e-20f) tout = fminf(tout, ((Dx > 0.0f ? bx + ((float)kMacro - kSkipMargin) : bx + kSkipMargin) - Px) * fast_rcp(Dx));
    if (fabsf(Dy) > 1e-20f) tout = fminf(tout, ((Dy > 0.0f ? by + ((float)kMacro - kSkipMargin) : by + kSkipMargin) - Py) * fast_rcp(Dy));
    if (fabsf(Dz) > 1e-20f) tout = fminf(tout, ((Dz > 0.0f ? bz + ((float)kMacro - kSkipMargin) : bz + kSkipMargin) - Pz) * fast_rcp(Dz));
    // samples sit at t(s') = t0 + (texit - t0) * s'/(n-1): the last one strictly before tout, minus one for safety
    const float dt = (r.texit - r.t0) * r.inv_nm1;
    if (!(dt > 1e-20f)) return 1;
    const float last = floorf((tout - r.t0) * fast_rcp(dt)) - 1.0f;
    int m = (last < (float)s) ? 1 : (last > (float)(nn - 1) ? nn - s : (int)last - s + 1);
    return m < 1 ? 1 : m;
}

// the same for the backward march, which walks s, s-1, ...: number of consecutive samples (at least 1) down from s whose cells
// stay in the macro-cell of sample s's cell; never more than s + 1.  The ray ENTERS the macro-cell at the latest of the per-axis
// entry faces (shrunk by the margin); one more sample than necessary is left to the normal path.
DR_HD int skip_run_back(const DrDesc& d, const Ray& r, F3 cam, int s, int cx, int cy, int cz)
{
    if (r.n < 2) return 1;
    const float Px = (0.5f * cam.x + 0.5f) * d.scale[0], Dx = 0.5f * r.dir.x * d.scale[0];
    const float Py = (0.5f * cam.y + 0.5f) * d.scale[1], Dy = 0.5f * r.dir.y * d.scale[1];
    const float Pz = (0.5f * cam.z + 0.5f) * d.scale[2], Dz = 0.5f * r.dir.z * d.scale[2];
    const float bx = (float)(cx & ~(kMacro - 1)), by = (float)(cy & ~(kMacro - 1)), bz = (float)(cz & ~(kMacro - 1));
    float tin = -3.0e38f;
    if (fabsf(Dx) > 1e-20f) tin = fmaxf(tin, ((Dx > 0.0f ? bx + kSkipMargin : bx + ((float)kMacro - kSkipMargin)) - Px) * fast_rcp(Dx));
    if (fabsf(Dy) > 1e-20f) tin = fmaxf(tin, ((Dy > 0.0f ? by + kSkipMargin : by + ((float)kMacro - kSkipMargin)) - Py) * fast_rcp(Dy));
    if (fabsf(Dz) > 1e-20f) tin = fmaxf(tin, ((Dz > 0.0f ? bz + kSkipMargin : bz + ((float)kMacro - kSkipMargin)) - Pz) * fast_rcp(Dz));
    const float dt = (r.texit - r.t0) * r.inv_nm1;
    if (!(dt > 1e-20f)) return 1;
    // samples sit at t(s') = t0 + dt * s': the first one strictly after tin, plus one for safety
    const float first = fmaxf(ceilf((tin - r.t0) * fast_rcp(dt)) + 1.0f, 0.0f);
    if (!(first <= (float)s)) return 1;                       // (also NaN)
    return s - (int)first + 1;
}

// 1 when every TF bin reachable from voxel values in [mn, mx] has alpha 0 (tf is the staged bin table's source: alpha of bin
// r is tf_alpha[r * tf_stride]).  The range is widened by the rounding slack of three fp32 mix levels and of the bin scale.
DR_HD unsigned char skip_classify(const DrDesc& d, float mn, float mx, const float* tf_alpha, int tf_stride)
{
    if (!(mn <= mx)) return 0;                                   // NaN anywhere: never skip
    const float slack = 4e-6f * fmaxf(fabsf(mn), fabsf(mx)) + 1e-30f;
    const float lo = fmaxf((mn - slack) * d.tf_len, 0.0f), hi = fmaxf((mx + slack) * d.tf_len, 0.0f);
    if (!(hi < 16777216.0f)) return 0;
    int b0 = (int)floorf(lo), b1 = (int)floorf(hi) + 1;          // + 1: the lookup also reads bin lo + 1
    if (b0 > d.R - 1) b0 = d.R - 1;
    if (b1 > d.R - 1) b1 = d.R - 1;
    for (int b = b0; b <= b1; ++b)
        if (tf_alpha[(size_t)b * tf_stride] != 0.0f) return 0;   // (also false for NaN alpha: NaN != 0)
    return 1;
}

// ---------------------------------------------------------------------------------------------------------
// Per-ray forward march: raycast :261-306 / raycast_nondiff :308-351 + get_final_image(_nondiff) :353-372.
// State per ray is O(1): A (accumulated premultiplied RGBA), K (active samples), Tprev (transmittance before
// the last active sample).  Nothing per sample is stored (the reference stores 16*M bytes per ray, :82,102-103).
// ---------------------------------------------------------------------------------------------------------
template <typename VT, int LAYOUT, bool NONDIFF, int TAPS, bool SR1, bool SKIP = false>
DR_HD void march_forward(const DrDesc& d, const VolView<VT>& vol, const Layout& L, const TfTable& tf, F3 cam,
                         const Ray& r, F4& A, int& K, float& Tprev, const unsigned char* skip_grid = nullptr)
{
    A.x = A.y = A.z = A.w = 0.0f;                       // H1: tape[-1] = 0
    K = 0;
    int Kshaded = 0;                                    // diagnostic (DR_F_COUNT_SHADED): samples with non-zero opacity
    Tprev = 1.0f;
    const int nn = NONDIFF ? r.n : imin(r.n, d.M);      // :267-269 (s < max_samples only in the diff kernel)
    // SKIP: whether the sample's cell left the macro-cell of the previous sample costs two LOP3 and a compare on the
    // biased floor indices (the bias is a multiple of 8); the grid byte is loaded and a run skipped only on such a change
    int pbx = -1, pby = -1, pbz = -1;                   // biased cell indices of the sample examined last
    bool in_empty = false;                              // ... and whether its macro-cell is empty
    for (int s = 0; s < nn; ++s) {
        if (!(A.w < d.ert)) break;                      // :267 / :318; later iterations only copy A forward :304-306
        float ts;
        const F3 pos = sample_pos(r, cam, s, ts);
        Centre c;
        if (SKIP && TAPS != TAPS_GENERIC && skip_grid) {
            locate_centre(d, L, pos, c);
            if ((((c.cx.b ^ pbx) | (c.cy.b ^ pby) | (c.cz.b ^ pbz)) & ~(kMacro - 1)) != 0) {
                pbx = c.cx.b; pby = c.cy.b; pbz = c.cz.b;
                const int cx = lo_of(c.cx), cy = lo_of(c.cy), cz = lo_of(c.cz);
                in_empty = load_u8(skip_grid + macro_index(d, cx, cy, cz)) != 0;
                if (in_empty) {
                    // an exactly transparent run: the samples only count (diff march) / are skipped (nondiff, alpha <= 1e-3)
                    const int m = skip_run(d, r, cam, s, nn, cx, cy, cz);
                    if (!NONDIFF) { K += m; Tprev = DR_SUB(1.0f, A.w); }
                    s += m - 1;
                    continue;
                }
            } else if (in_empty) {                      // the conservative run ended a sample or two before the macro-cell does
                if (!NONDIFF) { ++K; Tprev = DR_SUB(1.0f, A.w); }
                continue;
            }
            typename AddrOf<VT, LAYOUT, false>::type ad;
            ad.init(d, vol.p, L, c);
            eval_centre(ad, c);
        } else {
            sample_centre<VT, LAYOUT, TAPS>(d, vol, L, pos, c);
        }
        TfHit h;
        apply_tf(d, tf, c.I, h);
        if (NONDIFF && !(h.c.w > d.alpha_skip)) continue;      // :334: skipped samples never evaluate the normal
        const float o = opacity<SR1>(d, h.c.w);
        const float T = DR_SUB(1.0f, A.w);
        if (o == 0.0f) {
            // Exactly transparent sample (TF alpha 0, the empty space between a transfer function's bumps):
            // C = (k*c*0, 0), so A_s = fma(T, 0, A_{s-1}) = A_{s-1} bit for bit whatever the Phong factor k is.  The
            // sample still counts as active (:303) but its six normal taps and shading are never evaluated.
            Tprev = T;
            ++K;
            continue;
        }
        tf_colour(h, false);
        Taps t;
        sample_normals<VT, LAYOUT, TAPS, 0>(d, vol, L, pos, c, t);
        Shade sh;
        shade(d, r.dir, ts, t.g, !NONDIFF, sh);
        const float ko = sh.k * o;
        Tprev = T;
        A.x = DR_FMA(T, ko * h.c.x, A.x);
        A.y = DR_FMA(T, ko * h.c.y, A.y);
        A.z = DR_FMA(T, ko * h.c.z, A.z);
        A.w = DR_FMA(T, o, A.w);                        // :300-302
        ++K;                                            // :303
        ++Kshaded;
    }
    if (d.flags & DR_F_COUNT_SHADED) K = Kshaded;
    if (NONDIFF) { A.x = fminf(1.0f, A.x); A.y = fminf(1.0f, A.y); A.z = fminf(1.0f, A.z); A.w = fminf(1.0f, A.w); }   // :358
}

// ---------------------------------------------------------------------------------------------------------
// Per-ray backward: get_final_image.grad + raycast.grad (:460-461, :470-471) without a tape.
// The ray is re-marched from its last active sample K-1 down to 0.  The transmittance before sample s is
// reconstructed as T_{s-1} = T_s / (1 - o_s); the last active sample uses the saved Tprev (so an opaque last
// sample never divides by ~0), and 1 - o_s > 1 - ert for every earlier sample because sample s+1 was active.
// dL/dA_s.rgb is constant along the ray (= grad_out.rgb); only dL/dA_s.w evolves: g.w -= C_s . g.
// TfSink::add(lo, f, dc) accumulates the TF gradient ((1-f)*dc into bin lo, f*dc into bin min(lo+1, R-1));
// VolSink::open/close/direct the volume gradient; both may hold a partial sum in registers and are flushed at the end
// of the ray.
// ---------------------------------------------------------------------------------------------------------
template <typename VT, int LAYOUT, int TAPS, bool WANT_VOL, bool WANT_TF, bool SR1, bool SKIP, typename VolSink, typename TfSink>
DR_HD void march_backward(const DrDesc& d, const VolView<VT>& vol, const Layout& L, const TfTable& tf, F3 cam,
                          const Ray& r, F4 Afinal, int K, float Tprev, F4 g, VolSink& vsink, TfSink& tsink,
                          const unsigned char* skip_grid = nullptr)
{
    float Tafter = 1.0f - Afinal.w;                     // transmittance after sample K-1
    // SKIP (volume gradient only): a sample in a macro-cell that is exactly transparent under the TF has o = 0 and both TF bins
    // transparent, so it contributes nothing and changes neither g.w nor the transmittance -- runs of them are jumped over with
    // the forward's skip grid (the TF gradient, in contrast, receives d(alpha) from every transparent sample)
    int pbx = -1, pby = -1, pbz = -1;
    bool in_empty = false;
    for (int s = K - 1; s >= 0; --s) {
        float ts;
        const F3 pos = sample_pos(r, cam, s, ts);
        Centre c;
        if (SKIP && !WANT_TF && TAPS != TAPS_GENERIC && skip_grid) {
            locate_centre(d, L, pos, c);
            if ((((c.cx.b ^ pbx) | (c.cy.b ^ pby) | (c.cz.b ^ pbz)) & ~(kMacro - 1)) != 0) {
                pbx = c.cx.b; pby = c.cy.b; pbz = c.cz.b;
                const int cx = lo_of(c.cx), cy = lo_of(c.cy), cz = lo_of(c.cz);
                in_empty = load_u8(skip_grid + macro_index(d, cx, cy, cz)) != 0;
                if (in_empty) {
                    s -= skip_run_back(d, r, cam, s, cx, cy, cz) - 1;
                    continue;
                }
            } else if (in_empty) {
                continue;
            }
            typename AddrOf<VT, LAYOUT, false>::type ad;
            ad.init(d, vol.p, L, c);
            eval_centre(ad, c);
        } else {
            sample_centre<VT, LAYOUT, TAPS>(d, vol, L, pos, c);
        }
        TfHit h;
        apply_tf(d, tf, c.I, h);
        tf_colour(h, WANT_VOL);
        const float o = opacity<SR1>(d, h.c.w);
        // Volume-only gradient: an exactly transparent sample whose two TF bins are both transparent (h.d.w == 0) has
        // dc.rgb = k*o*dC = 0 and dI = tf_len * dc.w * h.d.w = 0, C = 0 (g.w unchanged) and T_{s-1} = T_s: nothing to do.
        if (!WANT_TF && o == 0.0f && h.d.w == 0.0f) continue;
        Taps t;
        sample_normals<VT, LAYOUT, TAPS, TAPS == TAPS_ONE ? (WANT_VOL ? 2 : 1) : 0>(d, vol, L, pos, c, t);
        Shade sh;
        shade(d, r.dir, ts, t.g, true, sh);
        if (o == 0.0f) {
            // Exactly transparent sample (78 % of the active samples under the tf1 preset): C = 0, so g.w and the
            // transmittance do not change (T_{s-1} = T_s; the forward saved Tprev = T for it too), dc.rgb = k*o*T*g = 0 and
            // dk = 0 (no normal-path adjoint).  What is left of sample_adjoint is d(alpha) = k*(c.rgb . dC.rgb) + dC.w --
            // it needs the Phong factor, which is why the normal was evaluated -- going to the alpha channel of the two TF
            // bins and, where the bins' alphas differ, to the centre tap.
            const float cd = Tafter * (h.c.x * g.x + h.c.y * g.y + h.c.z * g.z);
            const float dcw = (sh.k * cd + Tafter * g.w) * (SR1 ? 1.0f : d.inv_sr);      // pow(1 - 0, 1/sr - 1) = 1
            if (WANT_TF) tsink.add_alpha(h.lo, h.f, dcw);
            if (WANT_VOL) {
                SampleAdj a;
                a.dI = (h.x > 0.0f) ? dcw * h.d.w * d.tf_len : 0.0f;
                if (a.dI != 0.0f) {
                    a.dg.x = a.dg.y = a.dg.z = 0.0f; a.has_dg = false;
                    scatter_volume_grad<VolSink, TAPS == TAPS_GENERIC, TAPS == TAPS_TWO>(d, L, vsink, t, a);
                }
            }
            continue;
        }
        const float T = (s == K - 1) ? Tprev : Tafter * fast_rcp(1.0f - o);     // 1-o in (0.01, 1]: one MUFU.RCP, no fix-up code
        Tafter = T;
        SampleAdj a;
        const float Cg = sample_adjoint<SR1>(d, r.dir, ts, t.g, h, o, sh, T, g, WANT_VOL, a);
        g.w -= Cg;
        if (WANT_TF) tsink.add(h.lo, h.f, a.dc);
        if (WANT_VOL && (a.has_dg || a.dI != 0.0f))          // exactly-zero contributions (transparent samples) are not scattered
            scatter_volume_grad<VolSink, TAPS == TAPS_GENERIC, TAPS == TAPS_TWO>(d, L, vsink, t, a);
    }
    if (WANT_TF) tsink.flush();
    if (WANT_VOL) vsink.flush();
}

}  // namespace dr
