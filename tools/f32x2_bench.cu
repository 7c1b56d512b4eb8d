#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float lo(u64 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return x; }
__device__ __forceinline__ float hi(u64 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); return y; }

// MODE 0: 8 scalar FFMA chains; 1: 4 FFMA2 chains; 2: 8 FFMA + 8 integer ops; 3: 4 FFMA2 + 8 integer ops
template <int MODE> __global__ void k(float* out, int iters, float s, int m)
{
    float a[8]; u64 p[4]; int q[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; q[i] = threadIdx.x + i; }
    for (int i = 0; i < 4; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
    const u64 ss = pk(s, s), cc = pk(0.5f, 0.25f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 0 || MODE == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], s, 0.5f);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) p[i] = fma2(p[i], ss, cc);
            }
            if (MODE >= 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) q[i] = (q[i] ^ m) + i;
            }
        }
    }
    float r = 0; int qi = 0;
    for (int i = 0; i < 8; ++i) { r += a[i]; qi += q[i]; }
    for (int i = 0; i < 4; ++i) r += lo(p[i]) + hi(p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + qi;
}
template <int MODE> float run(float* d, int iters)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, 10, 1.0001f, 3);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, iters, 1.0001f, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main()
{
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int iters = 20000;
    float t0 = run<0>(d, iters), t1 = run<1>(d, iters), t2 = run<2>(d, iters), t3 = run<3>(d, iters);
    double fl = 148.0 * 8 * 256 * iters * 4.0 * 8;   // fp32 FMAs
    printf("8xFFMA        %.3f ms  %.2f TFMA/s\n4xFFMA2       %.3f ms  %.2f TFMA/s\n8xFFMA+8xINT  %.3f ms\n4xFFMA2+8xINT %.3f ms\n", t0, fl / t0 / 1e9, t1, fl / t1 / 1e9, t2, t3);
    return 0;
}
