// tools/atomic_bench.cu -- micro-benchmark that decided the gradient-scatter design (DESIGN.md "Backward").
// Each thread is a ray of an 8x4 pixel tile walking through a 256^3 volume 0.29 voxel per step (the march's access
// pattern); per step it issues the atomics one sample would: (a) 8 scalar REDs to the cell corners in the bricked layout,
// (b) 2 RED.v4 to a cell-major [cell][8] layout, (c) TF histogram: 2 RED.v4 into a privatised 2 KiB table,
// (d) TF histogram in shared memory with fp32 atomicAdd (CAS loop), (e) with int atomicAdd (native).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomic_bench tools/atomic_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ int offx(int x) { return ((x >> 3) << 9) | (x & 7); }
__device__ __forceinline__ int offy(int y, int sY) { return (y >> 3) * sY + ((y & 7) << 3); }
__device__ __forceinline__ int offz(int z, int sZ) { return (z >> 3) * sZ + ((z & 7) << 6); }

template <int MODE>
__global__ void __launch_bounds__(128) k(float* gvol, float4* gcell, float4* tfslots, int steps, float pix)
{
    __shared__ float s_f[512];
    __shared__ int s_i[512];
    for (int e = threadIdx.x; e < 512; e += 128) { s_f[e] = 0.f; s_i[e] = 0; }
    __syncthreads();
    const int N = 256, sY = 32 * 512, sZ = 32 * 32 * 512;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int tile = blockIdx.x;
    float px = 8.f + ((tile % 28) * 16 + (w & 1) * 8 + (l & 7)) * pix * 0.5f;
    float py = 8.f + ((tile / 28 % 28) * 8 + (w >> 1) * 4 + (l >> 3)) * pix * 0.5f;
    float pz = 4.f + (tile % 7);
    const float dx = 0.05f, dy = 0.07f, dz = 0.277f;       // |d| = 0.29 voxel per step
    float4* slot = tfslots + (blockIdx.x & 1023) * 128;
    for (int s = 0; s < steps; ++s) {
        px += dx; py += dy; pz += dz;
        if (pz > N - 3) { pz -= (N - 8); }
        if (px > N - 3) px -= (N - 8);
        if (py > N - 3) py -= (N - 8);
        const int x = (int)px, y = (int)py, z = (int)pz;
        const float fx = px - x, fy = py - y, fz = pz - z;
        const float v = fx * fy + fz;
        if (MODE == 0) {
            const int x0 = offx(x), x1 = offx(x + 1), y0 = offy(y, sY), y1 = offy(y + 1, sY), z0 = offz(z, sZ), z1 = offz(z + 1, sZ);
            atomicAdd(gvol + x0 + y0 + z0, v); atomicAdd(gvol + x1 + y0 + z0, v + 1);
            atomicAdd(gvol + x0 + y1 + z0, v + 2); atomicAdd(gvol + x1 + y1 + z0, v + 3);
            atomicAdd(gvol + x0 + y0 + z1, v + 4); atomicAdd(gvol + x1 + y0 + z1, v + 5);
            atomicAdd(gvol + x0 + y1 + z1, v + 6); atomicAdd(gvol + x1 + y1 + z1, v + 7);
        } else if (MODE == 1) {
            const size_t c = ((size_t)(z * N + y) * N + x) * 2;
            atomicAdd(gcell + c, make_float4(v, v + 1, v + 2, v + 3));
            atomicAdd(gcell + c + 1, make_float4(v + 4, v + 5, v + 6, v + 7));
        } else if (MODE == 2) {
            const int b = ((int)(fz * 40.f + x)) & 126;
            atomicAdd(slot + b, make_float4(v, v, v, v));
            atomicAdd(slot + b + 1, make_float4(v, v, v, v));
        } else if (MODE == 3) {
            const int b = (((int)(fz * 40.f + x)) & 126) * 4;
#pragma unroll
            for (int q = 0; q < 8; ++q) atomicAdd(&s_f[b + q], v);
        } else if (MODE == 4) {
            const int b = (((int)(fz * 40.f + x)) & 126) * 4;
#pragma unroll
            for (int q = 0; q < 8; ++q) atomicAdd(&s_i[b + q], (int)(v * 1000.f));
        } else if (MODE == 5) {          // bricked layout, x-pairs as RED.v2 when 8-byte aligned (even x), else 2 scalars
            const int y0 = offy(y, sY), y1 = offy(y + 1, sY), z0 = offz(z, sZ), z1 = offz(z + 1, sZ);
            const int x0 = offx(x), x1 = offx(x + 1);
            const int o[4] = { y0 + z0, y1 + z0, y0 + z1, y1 + z1 };
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if ((x & 1) == 0) atomicAdd(reinterpret_cast<float2*>(gvol + x0 + o[q]), make_float2(v, v + q));
                else { atomicAdd(gvol + x0 + o[q], v); atomicAdd(gvol + x1 + o[q], v + q); }
            }
        }
    }
    if (MODE == 3 || MODE == 4) {
        __syncthreads();
        if (threadIdx.x == 0) gvol[blockIdx.x] = s_f[3] + s_i[5];
    }
}

template <int MODE> void run(const char* name, float* gvol, float4* gcell, float4* tf, float pix, int per_sample)
{
    const int grid = 148 * 8, steps = 2000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<grid, 128>>>(gvol, gcell, tf, 200, pix);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<MODE><<<grid, 128>>>(gvol, gcell, tf, steps, pix);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double samples = (double)grid * 128 * steps;
    printf("%-44s pix=%.2f  %8.3f ms  %7.2f Gsamples/s  (%d atomic instr/sample, %.1f G lane-atomics/s)  err=%s\n", name, pix, ms,
           samples / ms / 1e6, per_sample, samples * per_sample / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    float *gvol; float4 *gcell, *tf;
    cudaMalloc(&gvol, (size_t)256 * 256 * 256 * 4);
    cudaMalloc(&gcell, (size_t)256 * 256 * 256 * 32);
    cudaMalloc(&tf, 1024 * 128 * 16);
    cudaMemset(gvol, 0, (size_t)256 * 256 * 256 * 4); cudaMemset(gcell, 0, (size_t)256 * 256 * 256 * 32); cudaMemset(tf, 0, 1024 * 128 * 16);
    for (float pix : { 0.37f, 0.74f }) {
        run<0>("vol: 8 scalar RED.F32, bricked", gvol, gcell, tf, pix, 8);
        run<5>("vol: RED.F32x2 on aligned x-pairs, bricked", gvol, gcell, tf, pix, 6);
        run<1>("vol: 2 RED.F32x4, cell-major [cell][8]", gvol, gcell, tf, pix, 2);
        run<2>("tf: 2 RED.F32x4, 1024 privatised L2 tables", gvol, gcell, tf, pix, 2);
        run<3>("tf: 8 smem fp32 atomicAdd (CAS loop)", gvol, gcell, tf, pix, 8);
        run<4>("tf: 8 smem int atomicAdd (native)", gvol, gcell, tf, pix, 8);
    }
    return 0;
}
