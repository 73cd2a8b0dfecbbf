// Micro-benchmark: cost of SMALL tcgen05.mma (kind::tf32) instructions on B200 -- the regime of the quadtree / cascade kernels
// (M = 64 or 128 rows, N = 32..256, K = 8 per instruction, operands already in shared memory / TMEM).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_small umma_small.cu && ./umma_small
// For every configuration one CTA issues `n` MMAs from one thread and waits for tcgen05.commit; the cycles per MMA are the
// slope between n = 32 and n = 256.  Variants: accumulator policy (one accumulator / rotating over R accumulators), A operand
// from shared memory (SS) or from TMEM (TS), operand descriptors advancing through a tile like a real k-loop or fixed.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WD;\nbra WL;\nWD:\n}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(const void *smem_tile) {
    const uint64_t addr = (uint64_t)(smem_u32(smem_tile) >> 4) & 0x3fffull;
    return addr | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc),
                 "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

template <int M, int N, int ROT, int TS, int ADV>
__global__ void __launch_bounds__(128, 1) bench_kernel(int n, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t *sm = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) ((float *)sm)[i] = 1.0f + (i & 7) * 0.125f;     // 192 KB of finite operands
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tm = tmem_slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        constexpr int a_bytes = M * 128, b_bytes = N * 128, NB = N > 128 ? 2 : 4;
        uint8_t *a0 = sm, *b0 = sm + 64 * 1024;
        uint64_t ad[4], bd[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { ad[t] = umma_desc(a0 + (ADV ? t : 0) * a_bytes); bd[t] = umma_desc(b0 + (ADV ? t % NB : 0) * b_bytes); }
        long long t0 = clock64();
        for (int g = 0; g < n; g += 16) {           // 16 MMAs per iteration, straight-line: 4 tiles x 4 k-steps
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    constexpr int dummy = 0;
                    const int i = 4 * t + k;
                    const uint32_t d = tm + (uint32_t)((i % ROT) * N);
                    const uint32_t acc = (g | (i >= ROT)) != 0 ? 1u : 0u;
                    if (TS) umma_ts(d, tm + 448 + 8 * (ADV ? k : dummy), bd[t] + 2 * (ADV ? k : 0), idesc, acc);
                    else umma_ss(d, ad[t] + 2 * (ADV ? k : 0), bd[t] + 2 * (ADV ? k : 0), idesc, acc);
                }
        }
        long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tm) : "memory");
    }
}

template <int M, int N, int ROT, int TS, int ADV>
void run(long long *d_out) {
    if (ROT * N > 448) return;
    const size_t smem = 1024 + 192 * 1024;
    cudaFuncSetAttribute(bench_kernel<M, N, ROT, TS, ADV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long r[2][2], h[2];
    const int ns[2] = {32, 288};
    for (int j = 0; j < 2; ++j) {
        bench_kernel<M, N, ROT, TS, ADV><<<1, 128, smem>>>(ns[j], d_out);      // warm
        bench_kernel<M, N, ROT, TS, ADV><<<1, 128, smem>>>(ns[j], d_out);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
        cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
        r[j][0] = h[0]; r[j][1] = h[1];
    }
    printf("%-5d %-5d %-4d %-3d %-4d %12.1f %12.1f %10lld\n", M, N, ROT, TS, ADV, (r[1][1] - r[0][1]) / 256.0, (r[1][0] - r[0][0]) / 256.0, r[0][1]);
}

template <int M, int N>
void run_shape(long long *d_out) {
    run<M, N, 1, 0, 0>(d_out); run<M, N, 1, 0, 1>(d_out); run<M, N, 2, 0, 1>(d_out); run<M, N, 4, 0, 1>(d_out); run<M, N, 8, 0, 1>(d_out);
    run<M, N, 1, 1, 0>(d_out); run<M, N, 1, 1, 1>(d_out); run<M, N, 4, 1, 1>(d_out);
}

int main() {
    long long *d_out;
    cudaMalloc(&d_out, 16);
    printf("%-5s %-5s %-4s %-3s %-4s %12s %12s %10s\n", "M", "N", "rot", "ts", "adv", "cyc/MMA", "issue/MMA", "t(32)");
    run_shape<64, 32>(d_out); run_shape<64, 64>(d_out); run_shape<64, 128>(d_out); run_shape<64, 256>(d_out);
    run_shape<128, 32>(d_out); run_shape<128, 64>(d_out); run_shape<128, 128>(d_out); run_shape<128, 256>(d_out);
    return 0;
}
