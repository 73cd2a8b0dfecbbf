// Op-level drop-ins for the reference's three pybind extensions (forward).  These keep the
// reference's tensor contracts exactly (int64 indices, [.., K, H] score layout) so that its
// unmodified Python (functions/quadtree_attention.py, cascade_functions.py) can run on them;
// the fused kernels (qtatt_*.cu, cascade_match.cu) are the fast path.
#include "common.cuh"
#include "kernels.cuh"

namespace {

// ---- R1: score_computation_kernal.cu:22-62.  One CTA per (b, n1); thread per (k, h) pair computes the
// 4 sibling dots; results staged in smem so the [4][K][H] slab is written with coalesced stores.
__global__ void __launch_bounds__(128) score5d_kernel(const float *__restrict__ query, const float *__restrict__ key,
                                                       const int64_t *__restrict__ index, float *__restrict__ out,
                                                       int N1, int N2, int H, int D, int K) {
    extern __shared__ __align__(16) float sm[];
    float *qs = sm;                  // [4][H*D]
    float *os = sm + 4 * H * D;      // [4][K*H]
    const size_t bn = blockIdx.x;    // b*N1 + n1
    const size_t b = bn / N1;
    const int HD = H * D, KH = K * H;
    for (int i = threadIdx.x; i < 4 * HD; i += blockDim.x) qs[i] = __ldg(query + bn * 4 * HD + i);
    __syncthreads();
    for (int kh = threadIdx.x; kh < KH; kh += blockDim.x) {
        const int h = kh % H;
        long long idx = index[bn * KH + kh];
        idx = idx < 0 ? 0 : (idx >= N2 ? N2 - 1 : idx);
        const float *kr = key + ((b * N2 + (size_t)idx) * H + h) * D;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (int d = 0; d < D; ++d) {
            const float kv = __ldg(kr + d);
            s0 = fmaf(qs[0 * HD + h * D + d], kv, s0);
            s1 = fmaf(qs[1 * HD + h * D + d], kv, s1);
            s2 = fmaf(qs[2 * HD + h * D + d], kv, s2);
            s3 = fmaf(qs[3 * HD + h * D + d], kv, s3);
        }
        os[0 * KH + kh] = s0; os[1 * KH + kh] = s1; os[2 * KH + kh] = s2; os[3 * KH + kh] = s3;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * KH; i += blockDim.x) out[bn * 4 * KH + i] = os[i];
}

// ---- R2: value_aggregation_kernel.cu:21-42.  Thread per output element; adjacent threads = adjacent d,
// so the gathered value rows are read coalesced and score/index loads broadcast within a row.
__global__ void __launch_bounds__(256) value_agg_kernel(const float *__restrict__ score, const float *__restrict__ value,
                                                         const int64_t *__restrict__ index, float *__restrict__ out,
                                                         size_t total, int N, int K, int H, int M, int D) {
    for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
        const int d = (int)(o % D);
        const int h = (int)((o / D) % H);
        const size_t bn = o / ((size_t)D * H);
        const size_t b = bn / N;
        const size_t s0 = bn * K * H + h;
        float acc = 0.f;
        for (int k = 0; k < K; ++k) {
            long long idx = __ldg(index + s0 + (size_t)k * H);
            idx = idx < 0 ? 0 : (idx >= M ? M - 1 : idx);
            acc = fmaf(__ldg(score + s0 + (size_t)k * H), __ldg(value + ((b * M + (size_t)idx) * H + h) * D + d), acc);
        }
        out[o] = acc;
    }
}

// ---- R3: score_cuda/src/score_computation_kernel.cu:23-40.  Warp per query row: the key row of each
// candidate is read coalesced (float4 per lane), partial dots of 8 candidates are combined with a
// transposing butterfly (9 shuffles / 8 candidates).
__global__ void __launch_bounds__(256) score3d_kernel(const float *__restrict__ query, const float *__restrict__ key,
                                                       const int64_t *__restrict__ index, float *__restrict__ out,
                                                       size_t rows, int N1, int N2, int C, int K) {
    const int lane = threadIdx.x & 31;
    const size_t row = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const size_t b = row / N1;
    const float *q = query + row * C;
    const float *kb = key + b * (size_t)N2 * C;
    const int64_t *ix = index + row * K;
    const int c4 = C >> 2;
    for (int k0 = 0; k0 < K; k0 += 8) {
        float part[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            part[j] = 0.f;
            const int k = k0 + j;
            if (k < K) {
                long long idx = __ldg(ix + k);
                idx = idx < 0 ? 0 : (idx >= N2 ? N2 - 1 : idx);
                const float *kr = kb + (size_t)idx * C;
                for (int c = lane; c < c4; c += 32) {
                    const float4 kv = ldg4(kr + 4 * c), qv = ldg4(q + 4 * c);
                    part[j] = fmaf(qv.x, kv.x, fmaf(qv.y, kv.y, fmaf(qv.z, kv.z, fmaf(qv.w, kv.w, part[j]))));
                }
            }
        }
        float v4[4], v2[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool hi = lane & 16;
            const float recv = __shfl_xor_sync(FULL_MASK, hi ? part[i] : part[i + 4], 16);
            v4[i] = (hi ? part[i + 4] : part[i]) + recv;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const bool hi = lane & 8;
            const float recv = __shfl_xor_sync(FULL_MASK, hi ? v4[i] : v4[i + 2], 8);
            v2[i] = (hi ? v4[i + 2] : v4[i]) + recv;
        }
        const bool hi = lane & 4;
        const float recv = __shfl_xor_sync(FULL_MASK, hi ? v2[0] : v2[1], 4);
        float s = (hi ? v2[1] : v2[0]) + recv;
        s += __shfl_xor_sync(FULL_MASK, s, 2);
        s += __shfl_xor_sync(FULL_MASK, s, 1);
        const int k = k0 + (lane >> 2);      // candidate = 4*bit4 + 2*bit3 + bit2 of the lane id
        if ((lane & 3) == 0 && k < K) out[row * K + k] = s;
    }
}

}  // namespace

int launch_score5d(const float *q, const float *key, const int64_t *idx, float *out,
                   int B, int N1, int N2, int H, int D, int K, cudaStream_t stream) {
    const size_t smem = sizeof(float) * (4 * (size_t)H * D + 4 * (size_t)K * H);
    CASMTR_REQUIRE(smem <= 227 * 1024, CASMTR_E_UNSUPPORTED, "score5d: H*D=%d, K*H=%d exceed shared memory", H * D, K * H);
    if ((size_t)B * N1 == 0) return CASMTR_OK;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(score5d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { casmtr_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return CASMTR_E_CUDA; }
    }
    LaunchScope ls(CASMTR_K_OPS, stream);
    score5d_kernel<<<(unsigned)((size_t)B * N1), 128, smem, stream>>>(q, key, idx, out, N1, N2, H, D, K);
    CASMTR_CHECK_LAUNCH("score5d_kernel");
    return CASMTR_OK;
}

int launch_value_agg(const float *score, const float *value, const int64_t *idx, float *out,
                     int B, int N, int K, int H, int M, int D, cudaStream_t stream) {
    const size_t total = (size_t)B * N * H * D;
    if (total == 0) return CASMTR_OK;
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    LaunchScope ls(CASMTR_K_OPS, stream);
    value_agg_kernel<<<(unsigned)blocks, 256, 0, stream>>>(score, value, idx, out, total, N, K, H, M, D);
    CASMTR_CHECK_LAUNCH("value_agg_kernel");
    return CASMTR_OK;
}

int launch_score3d(const float *q, const float *key, const int64_t *idx, float *out,
                   int B, int N1, int N2, int C, int K, cudaStream_t stream) {
    const size_t rows = (size_t)B * N1;
    if (rows == 0) return CASMTR_OK;
    LaunchScope ls(CASMTR_K_OPS, stream);
    score3d_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(q, key, idx, out, rows, N1, N2, C, K);
    CASMTR_CHECK_LAUNCH("score3d_kernel");
    return CASMTR_OK;
}
