// Op-level drop-ins for the BACKWARD halves of the reference's three pybind extensions (SURVEY.md section 8f "next" #4):
//   R1b  ScoreDataBackward            cuda_imp/QuadTreeAttention/QuadtreeAttention/src/score_computation_kernal.cu:94-144
//   R2b  value_aggregation backward   cuda_imp/QuadTreeAttention/QuadtreeAttention/src/value_aggregation_kernel.cu:55-87
//   R3b  score backward               cuda_imp/score_cuda/src/score_computation_kernel.cu:67-118
// so that the reference's unmodified autograd Functions (functions/quadtree_attention.py:7-54, cascade_functions.py:8-22)
// train on libcasmtr_b200.  Same tensor contracts as the forward drop-ins (ops.cu).
//
// Every op has a gather half (gradient of the query / score: one owner per output, plain stores) and a scatter half
// (gradient of the gathered key / value rows).  The reference issues one scalar atomicAdd per (sibling, candidate, channel)
// for BOTH halves; here the gather half has no atomics at all, and the scatter half first sums what it can on chip (the 4
// sibling queries of a cell share their candidates) and then issues one 16-byte vector reduction (red.global.add.v4.f32,
// sm_90+) per (candidate, head, 4 channels): 16x fewer atomic operations for score5d, 4x for the other two.  The order in
// which different query rows reach a key row is not fixed, so the scattered gradients are reproducible to fp32 rounding of
// the sum, not bit for bit -- exactly like the reference's.
#include "common.cuh"
#include "kernels.cuh"

namespace {

__device__ __forceinline__ void red_add4(float *addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ long long clamp_idx(long long i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }

// ---- R1b.  CTA = one (b, n1) cell; thread = (candidate group kg, head h, 4-channel chunk d4).  The cell's grad slab
// [4][K][H], index list [K][H] and 4 query rows are staged in shared memory; each thread walks candidates kg, kg + G, ...:
//   grad_query[f][h][4 d4..] += grad[f][k][h] * key[idx[k][h]][h][4 d4..]      (register accumulators, reduced over kg in smem)
//   grad_key[idx[k][h]][h][4 d4..] += sum_f grad[f][k][h] * query[f][h][4 d4..]  (one vector reduction per (k, h, d4))
__global__ void __launch_bounds__(256) score5d_bwd_kernel(const float *__restrict__ grad, const float *__restrict__ query,
                                                           const float *__restrict__ key, const int64_t *__restrict__ index,
                                                           float *__restrict__ gq, float *__restrict__ gk,
                                                           int N1, int N2, int H, int D, int K, int G) {
    extern __shared__ __align__(16) float sm[];
    const int HD = H * D, KH = K * H, HD4 = HD >> 2, D4 = D >> 2;
    float *qs = sm;                              // [4][HD]
    float *gs = qs + 4 * HD;                     // [4][KH]
    int *is = (int *)(gs + 4 * KH);              // [KH]
    float *red = (float *)(is + ((KH + 3) & ~3));     // [G][4][HD], 16-byte aligned
    const size_t bn = blockIdx.x, b = bn / N1;
    for (int i = threadIdx.x; i < 4 * HD; i += blockDim.x) qs[i] = __ldg(query + bn * 4 * HD + i);
    for (int i = threadIdx.x; i < 4 * KH; i += blockDim.x) gs[i] = __ldg(grad + bn * 4 * KH + i);
    for (int i = threadIdx.x; i < KH; i += blockDim.x) is[i] = (int)clamp_idx(index[bn * KH + i], N2);
    __syncthreads();
    const int kg = threadIdx.x / HD4, hd4 = threadIdx.x - kg * HD4;
    const int h = hd4 / D4, c = 4 * hd4;                       // c = h * D + 4 * d4: channel offset inside a token row
    float4 acc[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) acc[f] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kg < G) {
        float4 qv[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) qv[f] = *reinterpret_cast<const float4 *>(qs + f * HD + c);
        for (int k = kg; k < K; k += G) {
            const size_t row = ((b * N2 + (size_t)is[k * H + h]) * HD) + c;
            const float4 kv = ldg4(key + row);
            float4 ck = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const float g = gs[f * KH + k * H + h];
                acc[f].x = fmaf(g, kv.x, acc[f].x); acc[f].y = fmaf(g, kv.y, acc[f].y);
                acc[f].z = fmaf(g, kv.z, acc[f].z); acc[f].w = fmaf(g, kv.w, acc[f].w);
                ck.x = fmaf(g, qv[f].x, ck.x); ck.y = fmaf(g, qv[f].y, ck.y);
                ck.z = fmaf(g, qv[f].z, ck.z); ck.w = fmaf(g, qv[f].w, ck.w);
            }
            red_add4(gk + row, ck);
        }
#pragma unroll
        for (int f = 0; f < 4; ++f) *reinterpret_cast<float4 *>(red + (kg * 4 + f) * HD + c) = acc[f];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * HD; i += blockDim.x) {
        float s = 0.f;
        for (int g = 0; g < G; ++g) s += red[g * 4 * HD + i];
        gq[bn * 4 * HD + i] = s;
    }
}

// ---- R2b.  CTA = one (b, n) query row; thread = (candidate group kg, head h, chunk d4):
//   grad_score[k][h] = <grad_out[h][:], value[idx[k][h]][h][:]>      (shuffle reduction over the D/4 chunk lanes)
//   grad_value[idx[k][h]][h][4 d4..] += score[k][h] * grad_out[h][4 d4..]
__global__ void __launch_bounds__(256) value_agg_bwd_kernel(const float *__restrict__ grad, const float *__restrict__ score,
                                                             const float *__restrict__ value, const int64_t *__restrict__ index,
                                                             float *__restrict__ gscore, float *__restrict__ gvalue,
                                                             int N, int K, int H, int M, int D, int G) {
    const int HD = H * D, HD4 = HD >> 2, D4 = D >> 2, KH = K * H;
    const size_t bn = blockIdx.x, b = bn / N;
    const int kg = threadIdx.x / HD4, hd4 = threadIdx.x - kg * HD4;
    const int h = hd4 / D4, c = 4 * hd4;
    const bool live = kg < G;                                   // dead threads still take part in the shuffles
    const float4 go = live ? ldg4(grad + bn * HD + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k0 = 0; k0 < K; k0 += G) {
        const int k = k0 + kg;
        const bool on = live && k < K;
        float dot = 0.f;
        if (on) {
            const size_t s = bn * KH + (size_t)k * H + h;
            const size_t row = (b * M + (size_t)clamp_idx(__ldg(index + s), M)) * HD + c;
            const float4 vv = ldg4(value + row);
            dot = fmaf(go.x, vv.x, fmaf(go.y, vv.y, fmaf(go.z, vv.z, go.w * vv.w)));
            const float a = __ldg(score + s);
            red_add4(gvalue + row, make_float4(a * go.x, a * go.y, a * go.z, a * go.w));
        }
        for (int o = D4 >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(FULL_MASK, dot, o);      // D/4 is a power of two <= 32
        if (on && (hd4 % D4) == 0) gscore[bn * KH + (size_t)k * H + h] = dot;
    }
}

// ---- R3b.  Warp = one query row; lane = 4-channel chunks c4 = lane, lane + 32, ...:
//   grad_query[:] = sum_k grad[k] * key[idx[k]][:]            grad_key[idx[k]][:] += grad[k] * query[:]
template <int NC>       // chunks per lane (C <= 128 * NC)
__global__ void __launch_bounds__(256) score3d_bwd_kernel(const float *__restrict__ grad, const float *__restrict__ query,
                                                           const float *__restrict__ key, const int64_t *__restrict__ index,
                                                           float *__restrict__ gq, float *__restrict__ gk,
                                                           size_t rows, int N1, int N2, int C, int K) {
    const int lane = threadIdx.x & 31;
    const size_t row = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const size_t b = row / N1;
    const int c4n = C >> 2;
    float4 qv[NC], acc[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        const int c4 = lane + 32 * j;
        qv[j] = c4 < c4n ? ldg4(query + row * C + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int k = 0; k < K; ++k) {
        const float g = __ldg(grad + row * K + k);
        const size_t kr = (b * N2 + (size_t)clamp_idx(__ldg(index + row * K + k), N2)) * C;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const int c4 = lane + 32 * j;
            if (c4 < c4n) {
                const float4 kv = ldg4(key + kr + 4 * c4);
                acc[j].x = fmaf(g, kv.x, acc[j].x); acc[j].y = fmaf(g, kv.y, acc[j].y);
                acc[j].z = fmaf(g, kv.z, acc[j].z); acc[j].w = fmaf(g, kv.w, acc[j].w);
                red_add4(gk + kr + 4 * c4, make_float4(g * qv[j].x, g * qv[j].y, g * qv[j].z, g * qv[j].w));
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        const int c4 = lane + 32 * j;
        if (c4 < c4n) *reinterpret_cast<float4 *>(gq + row * C + 4 * c4) = acc[j];
    }
}

bool pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

}  // namespace

int launch_score5d_bwd(const float *grad, const float *q, const float *key, const int64_t *idx, float *gq, float *gk,
                       int B, int N1, int N2, int H, int D, int K, cudaStream_t stream) {
    CASMTR_REQUIRE(D % 4 == 0 && H * D <= 1024, CASMTR_E_UNSUPPORTED, "score5d_bwd: D=%d must be a multiple of 4 and H*D=%d <= 1024", D, H * D);
    CASMTR_REQUIRE((((uintptr_t)q | (uintptr_t)key | (uintptr_t)gq | (uintptr_t)gk) & 15) == 0, CASMTR_E_INVALID, "score5d_bwd: tensors must be 16-byte aligned");
    if (cudaMemsetAsync(gk, 0, sizeof(float) * (size_t)B * N2 * H * D, stream) != cudaSuccess) { casmtr_set_error("score5d_bwd: cudaMemsetAsync failed"); return CASMTR_E_CUDA; }
    if ((size_t)B * N1 == 0) return CASMTR_OK;
    const int HD4 = H * D / 4;
    int G = 256 / HD4;
    G = G < 1 ? 1 : (G > K ? K : G);
    const size_t smem = sizeof(float) * (4 * (size_t)H * D + 4 * (size_t)K * H + (((size_t)K * H + 3) & ~(size_t)3) + (size_t)G * 4 * H * D);
    CASMTR_REQUIRE(smem <= 227 * 1024, CASMTR_E_UNSUPPORTED, "score5d_bwd: H*D=%d, K*H=%d exceed shared memory", H * D, K * H);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(score5d_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { casmtr_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return CASMTR_E_CUDA; }
    }
    LaunchScope ls(CASMTR_K_OPS, stream);
    score5d_bwd_kernel<<<(unsigned)((size_t)B * N1), G * HD4, smem, stream>>>(grad, q, key, idx, gq, gk, N1, N2, H, D, K, G);
    CASMTR_CHECK_LAUNCH("score5d_bwd_kernel");
    return CASMTR_OK;
}

int launch_value_agg_bwd(const float *grad, const float *score, const float *value, const int64_t *idx, float *gscore, float *gvalue,
                         int B, int N, int K, int H, int M, int D, cudaStream_t stream) {
    CASMTR_REQUIRE(D % 4 == 0 && pow2(D / 4) && D <= 128 && H * D <= 1024, CASMTR_E_UNSUPPORTED,
                   "value_agg_bwd: D=%d must be 4, 8, 16, 32, 64 or 128 and H*D=%d <= 1024", D, H * D);
    CASMTR_REQUIRE((((uintptr_t)grad | (uintptr_t)value | (uintptr_t)gvalue) & 15) == 0, CASMTR_E_INVALID, "value_agg_bwd: tensors must be 16-byte aligned");
    if (cudaMemsetAsync(gvalue, 0, sizeof(float) * (size_t)B * M * H * D, stream) != cudaSuccess) { casmtr_set_error("value_agg_bwd: cudaMemsetAsync failed"); return CASMTR_E_CUDA; }
    if ((size_t)B * N == 0) return CASMTR_OK;
    const int HD4 = H * D / 4;
    int G = 256 / HD4;
    G = G < 1 ? 1 : (G > K ? K : G);
    const int threads = (G * HD4 + 31) / 32 * 32;               // whole warps: the shuffle reduction needs every lane
    LaunchScope ls(CASMTR_K_OPS, stream);
    value_agg_bwd_kernel<<<(unsigned)((size_t)B * N), threads, 0, stream>>>(grad, score, value, idx, gscore, gvalue, N, K, H, M, D, G);
    CASMTR_CHECK_LAUNCH("value_agg_bwd_kernel");
    return CASMTR_OK;
}

int launch_score3d_bwd(const float *grad, const float *q, const float *key, const int64_t *idx, float *gq, float *gk,
                       int B, int N1, int N2, int C, int K, cudaStream_t stream) {
    CASMTR_REQUIRE(C % 4 == 0 && C <= 512, CASMTR_E_UNSUPPORTED, "score3d_bwd: C=%d must be a multiple of 4, at most 512", C);
    CASMTR_REQUIRE((((uintptr_t)q | (uintptr_t)key | (uintptr_t)gq | (uintptr_t)gk) & 15) == 0, CASMTR_E_INVALID, "score3d_bwd: tensors must be 16-byte aligned");
    if (cudaMemsetAsync(gk, 0, sizeof(float) * (size_t)B * N2 * C, stream) != cudaSuccess) { casmtr_set_error("score3d_bwd: cudaMemsetAsync failed"); return CASMTR_E_CUDA; }
    const size_t rows = (size_t)B * N1;
    if (rows == 0) return CASMTR_OK;
    LaunchScope ls(CASMTR_K_OPS, stream);
    const unsigned blocks = (unsigned)((rows + 7) / 8);
    if (C <= 128) score3d_bwd_kernel<1><<<blocks, 256, 0, stream>>>(grad, q, key, idx, gq, gk, rows, N1, N2, C, K);
    else if (C <= 256) score3d_bwd_kernel<2><<<blocks, 256, 0, stream>>>(grad, q, key, idx, gq, gk, rows, N1, N2, C, K);
    else score3d_bwd_kernel<4><<<blocks, 256, 0, stream>>>(grad, q, key, idx, gq, gk, rows, N1, N2, C, K);
    CASMTR_CHECK_LAUNCH("score3d_bwd_kernel");
    return CASMTR_OK;
}
