// dr_host.h -- what the translation units of libdiffrender.so share on the host side: error reporting, the argument
// packs of the two march launches and the per-TU entry points (C++ linkage, not exported: the public interface is the
// C ABI of include/diffrender.h only).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "diffrender.h"

#define DR_INTERNAL __attribute__((visibility("hidden")))

namespace dr {

constexpr int kTfSlots = 1024;                  // privatised TF-gradient copies (power of two)
constexpr size_t kMaxTfSmem = 200 * 1024;       // shared-memory budget of the staged TF table (32 bytes per bin)

// set the thread-local message behind dr_last_error() and return `code` / DR_ECUDA
DR_INTERNAL int fail(int code, const char* msg);
DR_INTERNAL int fail_cuda(cudaError_t e, const char* where);

struct FwdArgs {
    const DrDesc* d; const void* vol; const float* tf; const float* cam; const float* jitter;
    float* out; int32_t* K; float* T; cudaStream_t st; const float* target; float* loss_sum;
    const unsigned char* skip_grid;      // [1 or BS][nbz*nby*nbx] bytes from dr_build_skip_grid, or null
};
struct BwdArgs {
    const DrDesc* d; const void* vol; const float* tf; const float* cam; const float* jitter;
    const float* gout; const float* out; const int32_t* K; const float* T; float4* gvol; float4* slots;
    cudaStream_t st; float mse_scale;
    const unsigned char* skip_grid;      // from dr_build_skip_grid (the forward's), or null; used by the volume-only backward
    const float* scale_dev;              // optional device scalar multiplied into mse_scale (the upstream gradient of the fused loss)
};

DR_INTERNAL int launch_forward_f32(const FwdArgs& a);      // dr_fwd_f32.cu
DR_INTERNAL int launch_forward_f16(const FwdArgs& a);      // dr_fwd_f16.cu
DR_INTERNAL int launch_backward_f32(const BwdArgs& a);     // dr_bwd_f32.cu
DR_INTERNAL int launch_backward_f16(const BwdArgs& a);     // dr_bwd_f16.cu
DR_INTERNAL int launch_forward_u8(const FwdArgs& a);       // dr_fwd_u8.cu   (cell-major copy only)
DR_INTERNAL int launch_backward_u8(const BwdArgs& a);      // dr_bwd_u8.cu

}  // namespace dr
