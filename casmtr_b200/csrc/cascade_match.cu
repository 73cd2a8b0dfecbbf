// Fused cascade matching: sparse correlation volume (both directions) + window mask + softmax over
// the K window candidates + max/argmax + index gather, one launch.
// Reference: CascadeMatching.forward, inference branch
//   src/model/functions/cascade_matching.py:87-149
// which runs fast_score_computation (cuda_imp/score_cuda/src/score_computation_kernel.cu:23-40) twice
// and ~10 elementwise / reduction torch ops over the [B,L,K] volume per direction.
//
// One warp = one query token of one direction.  A candidate's key row (C floats) is read
// coalesced by the whole warp; partial dots of 8 candidates are combined with a transposing
// butterfly, leaving candidate 8c + (lane>>2) of chunk c in every lane (replicated x4).
// HBM traffic: features once (re-reads of key rows hit L2), idx once, conf (optional) once.
#include "common.cuh"
#include "kernels.cuh"

namespace {

constexpr int MAX_CHUNKS = 16;   // K <= 128

template <int NQ>
__global__ void __launch_bounds__(256) cascade_match_kernel(MatchParams p) {
    const int lane = threadIdx.x & 31;
    size_t row = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t rows0 = (size_t)p.B * p.L0, rows1 = (size_t)p.B * p.L1;
    if (row >= rows0 + rows1) return;
    const bool rev = row >= rows0;               // direction 1 -> 0
    if (rev) row -= rows0;
    const int Lq = rev ? p.L1 : p.L0, Lk = rev ? p.L0 : p.L1;
    const size_t b = row / Lq;
    const float *q = (rev ? p.feat1 : p.feat0) + row * p.C;
    const float *kb = (rev ? p.feat0 : p.feat1) + b * (size_t)Lk * p.C;
    const int64_t *ix = (rev ? p.idx10 : p.idx01) + row * p.K;
    const uint8_t *mq = rev ? p.mask1 : p.mask0;
    const uint8_t *mk = rev ? p.mask0 : p.mask1;
    float *conf = rev ? p.conf10 : p.conf01;
    const int K = p.K, c4 = p.C >> 2;

    float4 qv[NQ];
#pragma unroll
    for (int c = 0; c < NQ; ++c) {
        const int cc = lane + 32 * c;
        qv[c] = cc < c4 ? ldg4(q + 4 * cc) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const bool q_ok = mq ? (mq[row] != 0) : true;
    const int mine = lane >> 2;                  // candidate within a chunk this lane ends up holding

    float sc[MAX_CHUNKS];
#pragma unroll
    for (int ch = 0; ch < MAX_CHUNKS; ++ch) {
        sc[ch] = -INFINITY;
        const int k0 = 8 * ch;
        if (k0 < K) {                            // warp-uniform
            float part[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                part[j] = 0.f;
                if (k0 + j < K) {
                    long long idx = __ldg(ix + k0 + j);
                    idx = idx < 0 ? 0 : (idx >= Lk ? Lk - 1 : idx);
                    const float *kr = kb + (size_t)idx * p.C;
#pragma unroll
                    for (int c = 0; c < NQ; ++c) {
                        const int cc = lane + 32 * c;
                        if (cc < c4) {
                            const float4 kv = ldg4(kr + 4 * cc);
                            part[j] = fmaf(qv[c].x, kv.x, fmaf(qv[c].y, kv.y, fmaf(qv[c].z, kv.z, fmaf(qv[c].w, kv.w, part[j]))));
                        }
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
            s *= p.inv_scale;
            const int k = k0 + mine;
            if (k < K) {
                if (mq) {                        // window mask (:108-112, :125)
                    long long idx = __ldg(ix + k);
                    idx = idx < 0 ? 0 : (idx >= Lk ? Lk - 1 : idx);
                    if (!(q_ok && mk[b * Lk + idx] != 0)) s = -1e9f;
                }
                sc[ch] = s;
            }
        }
    }
    // softmax over K; every value is replicated in 4 lanes, reductions run over lane bits 2..4
    float m = sc[0];
#pragma unroll
    for (int ch = 1; ch < MAX_CHUNKS; ++ch) m = fmaxf(m, sc[ch]);
    m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 4));
    m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 8));
    m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 16));
    float sum = 0.f;
    int arg = 0x7fffffff;                        // first index attaining the max (torch.max on CPU / first-max rule)
#pragma unroll
    for (int ch = 0; ch < MAX_CHUNKS; ++ch) {
        if (sc[ch] == m) arg = min(arg, 8 * ch + mine);
        sc[ch] = exp_neg(sc[ch] - m);
        sum += sc[ch];
    }
    sum += __shfl_xor_sync(FULL_MASK, sum, 4);
    sum += __shfl_xor_sync(FULL_MASK, sum, 8);
    sum += __shfl_xor_sync(FULL_MASK, sum, 16);
    arg = min(arg, __shfl_xor_sync(FULL_MASK, arg, 4));
    arg = min(arg, __shfl_xor_sync(FULL_MASK, arg, 8));
    arg = min(arg, __shfl_xor_sync(FULL_MASK, arg, 16));
    if (conf) {                                  // lane (mine, r) stores chunks r, r+4, ..: 32 consecutive k per store
        const int r = lane & 3;
#pragma unroll
        for (int mI = 0; mI < MAX_CHUNKS / 4; ++mI) {
            const float v = r == 0 ? sc[4 * mI] : r == 1 ? sc[4 * mI + 1] : r == 2 ? sc[4 * mI + 2] : sc[4 * mI + 3];
            const int k = 32 * mI + 8 * r + mine;
            if (k < K) conf[row * K + k] = v / sum;
        }
    }
    if (lane == 0) {
        (rev ? p.next_conf10 : p.next_conf01)[row] = 1.0f / sum;     // exp(0) / sum
        (rev ? p.next_idx10 : p.next_idx01)[row] = ix[arg];
    }
}

}  // namespace

int launch_cascade_match(const MatchParams &p, cudaStream_t stream) {
    CASMTR_REQUIRE(p.K >= 1 && p.K <= 8 * MAX_CHUNKS, CASMTR_E_UNSUPPORTED, "cascade_match: K=%d must be in [1,%d]", p.K, 8 * MAX_CHUNKS);
    CASMTR_REQUIRE(p.C % 4 == 0 && p.C >= 4 && p.C <= 512, CASMTR_E_UNSUPPORTED, "cascade_match: C=%d must be a multiple of 4, <= 512", p.C);
    const size_t rows = (size_t)p.B * p.L0 + (size_t)p.B * p.L1;
    if (rows == 0) return CASMTR_OK;
    const unsigned blocks = (unsigned)((rows + 7) / 8);
    LaunchScope ls(CASMTR_K_CASCADE_MATCH, stream);
    if (p.C <= 128) cascade_match_kernel<1><<<blocks, 256, 0, stream>>>(p);
    else if (p.C <= 256) cascade_match_kernel<2><<<blocks, 256, 0, stream>>>(p);
    else cascade_match_kernel<4><<<blocks, 256, 0, stream>>>(p);
    CASMTR_CHECK_LAUNCH("cascade_match_kernel");
    return CASMTR_OK;
}
