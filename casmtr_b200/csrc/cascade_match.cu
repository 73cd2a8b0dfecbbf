// Fused cascade matching: sparse correlation volume (both directions) + window mask + softmax over
// the K window candidates + max/argmax + index gather, one launch.
// Reference: CascadeMatching.forward, inference branch
//   src/model/functions/cascade_matching.py:87-149
// which runs fast_score_computation (cuda_imp/score_cuda/src/score_computation_kernel.cu:23-40) twice
// and ~10 elementwise / reduction torch ops over the [B,L,K] volume per direction.
//
// Two kernels:
//  * cell kernel (grid shapes known and even, C = 128 or 64): one warp = the 2x2 sibling queries of one
//    parent cell.  In the cascade the candidate lists come from CascadeQTAttB's upsampled_idx, which is the
//    parent's window repeated for its 4 children (reference quadtree_attention.py:450), so the 4 index rows
//    are identical: the warp verifies that (it has to read the rows anyway), then gathers every candidate
//    key row ONCE -- 4x less L2->SM traffic, which is what bounds this kernel -- with cp.async (all K row
//    gathers of the cell in flight at once, each a coalesced full row) into a chunk-swizzled smem slab, and
//    computes with lane = candidate: conflict-free row reads, q chunks broadcast and shared by the candidate
//    rounds (8 LDS.128 per 64 FMAs per lane), no shuffles until the softmax.  If the rows differ the warp
//    falls back to the row algorithm for its 4 queries -- same results either way.
//  * row kernel (no grid information, odd grids, other C): one warp = one query row, register gathers.
// HBM traffic: features once (re-reads of key rows hit L2), idx once, conf (optional) once.
#include "common.cuh"
#include "kernels.cuh"

namespace {

constexpr int MAX_CHUNKS = 16;   // K <= 128

struct Dir {                     // one matching direction as seen by a warp
    const float *qf, *kf;        // query / key features
    const int64_t *idx;
    const uint8_t *mq, *mk;
    float *conf, *next_conf;
    int64_t *next_idx;
    int Lq, Lk;
};

__device__ __forceinline__ Dir direction(const MatchParams &p, bool rev) {
    Dir d;
    d.qf = rev ? p.feat1 : p.feat0; d.kf = rev ? p.feat0 : p.feat1;
    d.idx = rev ? p.idx10 : p.idx01;
    d.mq = rev ? p.mask1 : p.mask0; d.mk = rev ? p.mask0 : p.mask1;
    d.conf = rev ? p.conf10 : p.conf01;
    d.next_conf = rev ? p.next_conf10 : p.next_conf01;
    d.next_idx = rev ? p.next_idx10 : p.next_idx01;
    d.Lq = rev ? p.L1 : p.L0; d.Lk = rev ? p.L0 : p.L1;
    return d;
}

__device__ __forceinline__ long long clamp_idx(long long i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }

__device__ __forceinline__ float dot4acc(const float4 a, const float4 b, float acc) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
}

// ---- one query row per warp.  Every score ends up replicated in 4 lanes (candidate = lane>>2).
template <int NQ>
__device__ __forceinline__ void match_row(const MatchParams &p, const Dir &d, size_t row, int lane) {
    const size_t b = row / d.Lq;
    const float *q = d.qf + row * p.C;
    const float *kb = d.kf + b * (size_t)d.Lk * p.C;
    const int64_t *ix = d.idx + row * p.K;
    const int K = p.K, c4 = p.C >> 2;

    float4 qv[NQ];
#pragma unroll
    for (int c = 0; c < NQ; ++c) {
        const int cc = lane + 32 * c;
        qv[c] = cc < c4 ? ldg4(q + 4 * cc) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const bool q_ok = d.mq ? (d.mq[row] != 0) : true;
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
                    const float *kr = kb + (size_t)clamp_idx(__ldg(ix + k0 + j), d.Lk) * p.C;
#pragma unroll
                    for (int c = 0; c < NQ; ++c) {
                        const int cc = lane + 32 * c;
                        if (cc < c4) part[j] = dot4acc(qv[c], ldg4(kr + 4 * cc), part[j]);
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
                if (d.mq) {                      // window mask (:108-112, :125)
                    const long long idx = clamp_idx(__ldg(ix + k), d.Lk);
                    if (!(q_ok && d.mk[b * d.Lk + idx] != 0)) s = -1e9f;
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
    if (arg == 0x7fffffff) arg = 0;              // NaN scores: nothing compares equal to the max; keep the gather in bounds (the confidence is NaN, as in the reference)
    if (d.conf) {                                // lane (mine, r) stores chunks r, r+4, ..: 32 consecutive k per store
        const int r = lane & 3;
#pragma unroll
        for (int mI = 0; mI < MAX_CHUNKS / 4; ++mI) {
            const float v = r == 0 ? sc[4 * mI] : r == 1 ? sc[4 * mI + 1] : r == 2 ? sc[4 * mI + 2] : sc[4 * mI + 3];
            const int k = 32 * mI + 8 * r + mine;
            if (k < K) d.conf[row * K + k] = v / sum;
        }
    }
    if (lane == 0) {
        d.next_conf[row] = 1.0f / sum;           // exp(0) / sum
        d.next_idx[row] = ix[arg];
    }
}

template <int NQ>
__global__ void __launch_bounds__(256) cascade_match_row_kernel(MatchParams p) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    size_t row = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t rows0 = (size_t)p.B * p.L0, rows1 = (size_t)p.B * p.L1;
    if (row >= rows0 + rows1) return;
    const bool rev = row >= rows0;               // direction 1 -> 0
    if (rev) row -= rows0;
    const Dir d = direction(p, rev);
    match_row<NQ>(p, d, row, lane);
}

// ---- the 2x2 sibling queries of one parent cell per warp, candidate rows staged in shared memory
__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}

// per-warp staging: CSTAGES buffers of a 32-channel (128-byte) slice of the K candidate rows + the 4 query rows.
// Measured on the fallback list of the 832^2 workload (2164 cells): 1 buffer (13 KB per warp, 16 warps per SM) 34 us,
// 2 buffers with prefetch (27 KB, 8 warps per SM) 48 us, whole rows (53 KB, 4 warps per SM) 79 us -- cells in flight win.
constexpr int SLICE = 32;
#ifndef MATCH_CELL_STAGES
#define MATCH_CELL_STAGES 1
#endif
constexpr int CSTAGES = MATCH_CELL_STAGES;
__host__ __device__ inline int cell_slab_floats(int K, int C) { (void)C; return CSTAGES * (K + 4) * SLICE; }

template <int CH>
__device__ __forceinline__ void match_cell(const MatchParams &p, float *slab, size_t cell, int lane);

// CH = 16-byte chunks per feature row (C / 4): 32 or 16.  lane = candidate in the compute phase.
template <int CH>
__global__ void __launch_bounds__(256) cascade_match_cell_kernel(MatchParams p, int warps_per_cta) {
    pdl_sync();
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *slab = smem + (size_t)warp * cell_slab_floats(p.K, 4 * CH);
    if (p.cell_list) {                            // persistent over the tile kernel's fallback list
        const int n = *p.cell_count;
        for (int i = blockIdx.x * warps_per_cta + warp; i < n; i += gridDim.x * warps_per_cta) {
            match_cell<CH>(p, slab, (size_t)p.cell_list[i], lane);
            __syncwarp();
        }
        return;
    }
    const size_t cell = blockIdx.x * (size_t)warps_per_cta + warp;
    if (cell >= (size_t)p.B * (p.L0 >> 2) + (size_t)p.B * (p.L1 >> 2)) return;
    match_cell<CH>(p, slab, cell, lane);
}

template <int CH>
__device__ __forceinline__ void match_cell(const MatchParams &p, float *slab, size_t cell, int lane) {
    constexpr int C = 4 * CH;
    constexpr int R = MAX_CHUNKS / 4;             // candidate rounds of 32
    const size_t cells0 = (size_t)p.B * (p.L0 >> 2);
    const bool rev = cell >= cells0;
    if (rev) cell -= cells0;
    const Dir d = direction(p, rev);
    const int wq = rev ? p.w1 : p.w0;            // query grid width
    const int wp = wq >> 1, cells_per = d.Lq >> 2;
    const size_t b = cell / cells_per;
    const int pc = (int)(cell - b * cells_per);
    const int py = pc / wp, px = pc - py * wp;
    const size_t row00 = b * d.Lq + (size_t)(2 * py) * wq + 2 * px;
#define ROWQ(f) (row00 + (size_t)((f) >> 1) * wq + ((f) & 1))
    const int K = p.K;

    // candidate lists: lane holds entries lane, lane+32, .. of sibling 0 (= its candidates in the compute phase);
    // siblings 1..3 must carry the same list
    int cand[R];
    bool same = true;
#pragma unroll
    for (int m = 0; m < R; ++m) {
        const int k = lane + 32 * m;
        cand[m] = 0;
        if (k < K) {
            const long long i0 = __ldg(d.idx + ROWQ(0) * K + k);
            same = same && i0 == __ldg(d.idx + ROWQ(1) * K + k) && i0 == __ldg(d.idx + ROWQ(2) * K + k) &&
                   i0 == __ldg(d.idx + ROWQ(3) * K + k);
            cand[m] = (int)clamp_idx(i0, d.Lk);
        }
    }
    if (!__all_sync(FULL_MASK, same)) {           // arbitrary lists: per-row algorithm, identical results
#pragma unroll 1
        for (int f = 0; f < 4; ++f) match_row<(CH + 31) / 32>(p, d, ROWQ(f), lane);
        return;
    }

    // The rows are streamed as NS slices of 32 channels (128 bytes of every candidate row), so a warp needs 13 KB instead
    // of 53 KB of shared memory (C = 128) -- that is what bounds the number of cells in flight per SM.
    constexpr int NS = C / SLICE;
    const int stage_floats = (K + 4) * SLICE;
    const float *kb = d.kf + b * (size_t)d.Lk * C;
    const int sub = lane >> 3, ch = lane & 7;     // a cp.async instruction moves the slice of 4 rows
    auto stage = [&](int s) {
        float *Ks = slab + (s % CSTAGES) * stage_floats;         // [K][32], 16-byte chunks XOR-swizzled by (row & 7)
        float *Qs = Ks + (size_t)K * SLICE;                      // [4][32]
        cp_async16(Qs + sub * SLICE + 4 * ch, d.qf + ROWQ(sub) * C + s * SLICE + 4 * ch);
#pragma unroll
        for (int m = 0; m < R; ++m) {
            if (32 * m < K) {
#pragma unroll
                for (int l = 0; l < 32; l += 4) {
                    const int ci = __shfl_sync(FULL_MASK, cand[m], l + sub);
                    const int k = 32 * m + l + sub;
                    if (k < K) cp_async16(Ks + (size_t)k * SLICE + 4 * (ch ^ (k & 7)), kb + (size_t)ci * C + s * SLICE + 4 * ch);
                }
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    stage(0);

    // correlation: lane = candidate 32r + lane; q chunks are broadcast reads shared by the rounds
    float sc[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int f = 0; f < 4; ++f) sc[r][f] = 0.f;
    int koff[R], sw[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int c = min(32 * r + lane, K - 1);  // rows past K re-read the last row; masked below
        koff[r] = c * SLICE;
        sw[r] = c & 7;
    }
#pragma unroll 1
    for (int s = 0; s < NS; ++s) {
        if (CSTAGES == 1) {
            if (s > 0) stage(s);
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        } else if (s + 1 < NS) {
            stage(s + 1);
            asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        }
        __syncwarp();
        const float *Ks = slab + (s % CSTAGES) * stage_floats, *Qs = Ks + (size_t)K * SLICE;
#pragma unroll
        for (int j = 0; j < SLICE / 4; ++j) {
            float4 qv[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) qv[f] = *reinterpret_cast<const float4 *>(Qs + f * SLICE + 4 * j);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (32 * r < K) {                 // warp-uniform
                    const float4 kv = *reinterpret_cast<const float4 *>(Ks + koff[r] + 4 * (j ^ sw[r]));
#pragma unroll
                    for (int f = 0; f < 4; ++f) sc[r][f] = dot4acc(qv[f], kv, sc[r][f]);
                }
            }
        }
        __syncwarp();                             // the stage is rewritten two slices later
    }
    // scale, window mask (:108-112, :125)
    bool q_ok[4] = {true, true, true, true};
    if (d.mq) {
#pragma unroll
        for (int f = 0; f < 4; ++f) q_ok[f] = d.mq[ROWQ(f)] != 0;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool valid = 32 * r + lane < K;
        const bool k_ok = (d.mq && valid) ? d.mk[b * d.Lk + cand[r]] != 0 : true;
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            float s = sc[r][f] * p.inv_scale;
            if (d.mq && !(q_ok[f] && k_ok)) s = -1e9f;
            sc[r][f] = valid ? s : -INFINITY;
        }
    }
    // softmax over the K candidates, max / first arg-max, per sibling
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        float m = sc[0][f];
#pragma unroll
        for (int r = 1; r < R; ++r) m = fmaxf(m, sc[r][f]);
        m = warp_max(m);
        float sum = 0.f;
        int arg = 0x7fffffff;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (sc[r][f] == m) arg = min(arg, 32 * r + lane);
            sc[r][f] = exp_neg(sc[r][f] - m);
            sum += sc[r][f];
        }
        sum = warp_sum(sum);
        arg = __reduce_min_sync(FULL_MASK, arg);
        if (arg == 0x7fffffff) arg = 0;          // NaN scores (see the row kernel)
        const size_t row = ROWQ(f);
        if (d.conf) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (32 * r + lane < K) d.conf[row * K + 32 * r + lane] = sc[r][f] / sum;
        }
        if (lane == 0) {
            d.next_conf[row] = 1.0f / sum;        // exp(0) / sum
            d.next_idx[row] = d.idx[row * K + arg];
        }
    }
#undef ROWQ
}

}  // namespace

int launch_cascade_match(const MatchParams &p, cudaStream_t stream) {
    CASMTR_REQUIRE(p.K >= 1 && p.K <= 8 * MAX_CHUNKS, CASMTR_E_UNSUPPORTED, "cascade_match: K=%d must be in [1,%d]", p.K, 8 * MAX_CHUNKS);
    CASMTR_REQUIRE(p.C % 4 == 0 && p.C >= 4 && p.C <= 512, CASMTR_E_UNSUPPORTED, "cascade_match: C=%d must be a multiple of 4, <= 512", p.C);
    const size_t rows = (size_t)p.B * p.L0 + (size_t)p.B * p.L1;
    if (rows == 0) return CASMTR_OK;
    const bool quad = p.w0 > 0 && p.w1 > 0 && p.w0 % 2 == 0 && p.w1 % 2 == 0 && p.L0 % p.w0 == 0 && p.L1 % p.w1 == 0 &&
                      (p.L0 / p.w0) % 2 == 0 && (p.L1 / p.w1) % 2 == 0 && (p.C == 128 || p.C == 64);
    MatchParams q = p;
    q.cell_list = nullptr; q.cell_count = nullptr;
    bool listed = false;
    if (quad && match_tile_applicable(p)) {       // window-structured lists: TMA-tiled kernel, cell kernel over its fallback list
        const int rc = launch_cascade_match_tile(p, stream);
        if (rc != CASMTR_OK) return rc;
        q.cell_list = p.fb_list; q.cell_count = p.fb_count;
        listed = true;
    }
    LaunchScope ls(listed ? CASMTR_K_CASCADE_FALLBACK : CASMTR_K_CASCADE_MATCH, stream);
    if (quad) {
        const size_t per_warp = sizeof(float) * cell_slab_floats(p.K, p.C);
        int wpc = (int)((110 * 1024) / per_warp);               // two CTAs per SM
        wpc = wpc < 1 ? 1 : (wpc > 8 ? 8 : wpc);
        const size_t smem = per_warp * wpc;
        const unsigned blocks = listed ? 2 * (unsigned)casmtr_sm_count() : (unsigned)((rows / 4 + wpc - 1) / wpc);
        static PerDeviceOnce once;
        const int dev = PerDeviceOnce::device();
        if (!once.done(dev)) {
            cudaError_t e = cudaFuncSetAttribute(cascade_match_cell_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(cascade_match_cell_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(cascade_match_cell_kernel<32>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(cascade_match_cell_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) { casmtr_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return CASMTR_E_CUDA; }
            once.mark(dev);
        }
        if (p.C == 128) launch_k(cascade_match_cell_kernel<32>, blocks, wpc * 32, smem, stream, q, wpc);
        else launch_k(cascade_match_cell_kernel<16>, blocks, wpc * 32, smem, stream, q, wpc);
    } else {
        const unsigned blocks = (unsigned)((rows + 7) / 8);
        if (p.C <= 128) launch_k(cascade_match_row_kernel<1>, blocks, 256, 0, stream, p);
        else if (p.C <= 256) launch_k(cascade_match_row_kernel<2>, blocks, 256, 0, stream, p);
        else launch_k(cascade_match_row_kernel<4>, blocks, 256, 0, stream, p);
    }
    CASMTR_CHECK_LAUNCH("cascade_match_kernel");
    return CASMTR_OK;
}
