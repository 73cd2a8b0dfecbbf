// Microbenchmark: issue rate of FFMA vs FFMA2 (packed fp32x2 FMA, sm_100) per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

template <bool PACKED>
__global__ void kern(float *out, int iters, long long *cyc) {
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.001f, -0.001f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (PACKED) a[i] = __ffma2_rn(a[i], m, c);
            else { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int threads : {128, 256, 512, 1024}) {
        for (int packed = 0; packed < 2; ++packed) {
            for (int rep = 0; rep < 2; ++rep) {
                if (packed) kern<true><<<148, threads>>>(out, iters, cyc); else kern<false><<<148, threads>>>(out, iters, cyc);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            const double fma_per_warp = 16.0 * iters;                 // scalar FMAs per thread
            const double warps_per_smsp = threads / 32 / 4.0;
            printf("threads/SM %4d  %s  cycles %lld  -> %.2f lane-FMA/cycle/SMSP (warp-instr/cycle/SMSP %.3f)\n", threads,
                   packed ? "FFMA2" : "FFMA ", h, fma_per_warp * warps_per_smsp * 32 / h,
                   (packed ? 8.0 : 16.0) * iters * warps_per_smsp / h);
        }
    }
    return 0;
}
