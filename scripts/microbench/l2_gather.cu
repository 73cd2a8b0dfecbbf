// Microbenchmark: L2 -> SM bandwidth for 128-byte row gathers (the access pattern of the quadtree / window kernels):
// every warp-wide LDG.128 fetches 4 random 128-byte rows of a buffer that fits in L2.  Also a plain streaming read of
// the same buffer for comparison.     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_gather l2_gather.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void gather_kernel(const float4 *__restrict__ buf, unsigned rows, int iters, float *out) {
    const unsigned lane = threadIdx.x & 31, g = lane >> 3, dq = lane & 7;
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) / 32 * 2654435761u + 12345u;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int it = 0; it < iters; ++it) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            s = s * 1664525u + 1013904223u;                       // same sequence in all lanes of the warp
            const unsigned row = ((s >> 8) + g * 7919u) % rows;   // 4 different rows per instruction
            v[u] = __ldg(buf + (size_t)row * 8 + dq);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[0] = 1.f;
}

__global__ void stream_kernel(const float4 *__restrict__ buf, size_t n4, int reps, float *out) {
    float4 acc = make_float4(0, 0, 0, 0);
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            const float4 v = __ldg(buf + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[0] = 1.f;
}

int main() {
    const size_t mb[] = {11, 22, 44, 88};
    float *out; cudaMalloc(&out, 4);
    for (size_t m : mb) {
        const size_t bytes = m << 20;
        float4 *buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
        const unsigned rows = bytes / 128;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int warps_per_sm : {16, 32, 64}) {
            const int blocks = 148 * warps_per_sm / 8, iters = 256;
            gather_kernel<<<blocks, 256>>>(buf, rows, iters, out);        // warm L2
            cudaEventRecord(e0);
            gather_kernel<<<blocks, 256>>>(buf, rows, iters, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double gb = (double)blocks * 8 * iters * 8 * 512 / 1e9;
            printf("buffer %3zu MB  gather  %2d warps/SM: %7.1f GB/s\n", m, warps_per_sm, gb / (ms / 1e3));
        }
        stream_kernel<<<148 * 8, 256>>>(buf, bytes / 16, 1, out);
        cudaEventRecord(e0);
        stream_kernel<<<148 * 8, 256>>>(buf, bytes / 16, 20, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("buffer %3zu MB  stream  : %7.1f GB/s\n", m, 20.0 * bytes / 1e9 / (ms / 1e3));
        cudaFree(buf);
    }
    return 0;
}
