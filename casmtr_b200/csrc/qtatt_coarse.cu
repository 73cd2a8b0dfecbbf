// Coarsest quadtree level: dense softmax(Q K^T / sqrt(D)) over all S keys, exact row top-k,
// A.V over ALL keys (type B) or over the non-selected keys (type A).
// Reference: QTAttB.process_coarse_level
//   cuda_imp/QuadTreeAttention/QuadtreeAttention/modules/quadtree_attention.py:161-178
// and QTAttA.process_coarse_level :25-44.  The reference materialises QK and A as [B,L,S,nh]
// tensors in HBM and runs torch.topk on a strided dim; here a CTA keeps a ROWS x S score slab
// in shared memory, so HBM sees only Q, K, V once (K/V re-reads hit L2) plus the outputs.
//
// One CTA = ROWS (32) query rows of one (batch, head); 256 threads.
//   phase 1  S = scale * Q K^T.  lane = query row (its 32-float q row lives in registers), the K tile
//            (64 tokens, cp.async double-buffered) is read with BROADCAST LDS.128: 8 LDS per 32 FMAs per
//            lane and no bank conflicts; warp w owns tokens 8w..8w+7 of every tile.  The slab row stride
//            is odd, so the per-lane score stores are conflict-free too.
//   phase 2  row softmax in smem, one warp per row.
//   phase 3  exact top-k per row, one warp per row, no serial selection loop:
//            lane-local top-2 -> T = k-th largest of those 64 keys by rank counting (64 shuffles)
//            -> compact the survivors {a >= T} (k <= n, typically n ~ 40) -> rank every survivor among
//            the survivors by the same counting scheme; rank < k writes itself to slot `rank`, which
//            yields the list sorted descending like torch.topk(sorted=True) (ties: lower key index first).
//   phase 4  O = A V, lane = query row again (32 accumulators in registers), V tile broadcast from smem,
//            warps split the tokens, partial sums reduced across warps through smem.
//
// Why the two contractions stay on FFMA2 (instruction count per 16-row CTA, Sk = 676, 11 tiles):
//   now      phase 1: per warp and tile 4 tokens x (8 LDS.128 + 16 FFMA2) + 4 stores ~ 110 instructions; phase 4 about the same
//            -> ~19 K of the CTA's ~82 K instructions (ncu smsp__inst_executed / 344 CTAs); the soft-max + exact top-k is ~37 K.
//   mma.sync m16n8k8 with a 3-term TF32 split (needed for fp32-exact top-k): per warp and tile 4 k-steps x (2 LDS + 6 split
//            ALU + 3 MMA) = 44 (phase 1) and ~60 (phase 4) -> ~9 K.  Saves ~10 K of 82 K = 12 % of this kernel, ~2.5 % of a step;
//            tcgen05 needs M >= 64 rows per CTA, i.e. 88 CTAs for 148 SMs while the soft-max / top-k part needs every SM.
#include <cuda_pipeline.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"

namespace {

constexpr int NW = 8;           // warps per CTA
constexpr int TILE = 64;        // key/value tokens per smem tile
constexpr int D = 32;
constexpr int LIST_CAP = 64;    // survivor list capacity per warp (2 per lane)
constexpr int RED_LD = 33;

__device__ __forceinline__ unsigned key_of(float v) { return v >= 0.f ? __float_as_uint(v) + 1u : 0u; }

// k rounds of warp arg-max over vals[0..n) (entries < 0 are dead).  Lane `it` ends up holding
// the it-th largest (value, position-in-list).  Destroys the selected entries (sets them to -1).
// Slow path, only used when a row has so many ties that more than LIST_CAP entries survive.
__device__ __forceinline__ void warp_select_k(float *vals, int n, int k, int lane, float &res_val, int &res_j) {
    res_val = -1.f;
    res_j = -1;
    for (int it = 0; it < k; ++it) {
        unsigned best = 0;
        int bj = -1;
        for (int j = lane; j < n; j += 32) {
            const unsigned kk = key_of(vals[j]);
            if (kk > best) { best = kk; bj = j; }
        }
        const unsigned g = __reduce_max_sync(FULL_MASK, best);
        const int owner = __ffs(__ballot_sync(FULL_MASK, best == g)) - 1;
        const int js = __shfl_sync(FULL_MASK, bj, owner);
        if (js < 0) break;       // fewer than k live entries (cannot happen when k <= n)
        const float v = vals[js];
        if (lane == it) { res_val = v; res_j = js; }
        __syncwarp();
        if (lane == owner) vals[js] = -1.f;
        __syncwarp();
    }
}

// token t of a tile -> XOR mask for its 16-byte chunk index.  With G lane groups per warp (G = 32 / ROWS) reading G
// different tokens t, t + TPL, .. in one instruction, the groups must land in different bank groups: the group index
// (t / TPL) % G is spread over the 3 chunk-index bits.
template <int G>
__device__ __forceinline__ int tok_swizzle(int t) {
    constexpr int TPL = (TILE / NW) / G;
    return G == 1 ? 0 : (((t / TPL) % G) * (8 / G));
}

// Per-thread constants of the tile loader: a thread always moves the same two 16-byte chunks of a tile (tokens t0 and
// t0 + 32, chunk c), so a tile costs it two cp.async and one pointer bump instead of index arithmetic per chunk.
template <int G>
struct TileLoader {
    int t0, soff0, doff0, doff1;
    size_t soff1;
    __device__ __forceinline__ TileLoader(int tid, int C) {
        t0 = tid >> 3;
        const int c = tid & 7;
        soff0 = t0 * C + 4 * c;
        soff1 = (size_t)soff0 + (size_t)32 * C;
        doff0 = t0 * D + 4 * (c ^ tok_swizzle<G>(t0));
        doff1 = (t0 + 32) * D + 4 * (c ^ tok_swizzle<G>(t0 + 32));
    }
    // stage tile `tok0 .. tok0 + TILE` of one head (rows past Sk are zero-filled) and commit the group
    __device__ __forceinline__ void load(float *dst, const float *src_head, int tok0, int Sk, int C) const {
        const float *src = src_head + (size_t)tok0 * C;
        if (tok0 + t0 < Sk) __pipeline_memcpy_async(dst + doff0, src + soff0, 16);
        else *reinterpret_cast<float4 *>(dst + doff0) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tok0 + t0 + 32 < Sk) __pipeline_memcpy_async(dst + doff1, src + soff1, 16);
        else *reinterpret_cast<float4 *>(dst + doff1) = make_float4(0.f, 0.f, 0.f, 0.f);
        __pipeline_commit();
    }
};

// Softmax + exact top-k of RP score rows at once, the row values held in registers (NV per lane, Sk <= 32 * NV).
// Processing RP rows together gives the scheduler independent dependency chains (the single-row version is bound by
// LDS / shuffle latency, not by issue slots); keeping the values in registers removes two of the three passes over smem.
//   srow[rr]: the row's scores in the slab (in: scaled logits, out: unnormalised exp, padding columns zeroed)
//   returns per row the exp sum; lane < k holds output slot `lane` (value rv, key index ridx)
template <int NV, int RP>
__device__ __forceinline__ void rows_softmax_topk(float *const (&srow)[RP], const bool (&live)[RP], int Sk, int s_pad, int k, int lane,
                                                   float *const (&lv)[RP], int *const (&lp)[RP], float (&sum)[RP], float (&rv)[RP],
                                                   int (&ridx)[RP]) {
    float e[RP][NV];
    float m[RP];
#pragma unroll
    for (int rr = 0; rr < RP; ++rr) {
        m[rr] = -INFINITY;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            e[rr][i] = (live[rr] && c < Sk) ? srow[rr][c] : -INFINITY;
            m[rr] = fmaxf(m[rr], e[rr][i]);
        }
    }
#pragma unroll
    for (int rr = 0; rr < RP; ++rr) m[rr] = warp_max(m[rr]);
    float m1[RP], m2[RP];
#pragma unroll
    for (int rr = 0; rr < RP; ++rr) {
        sum[rr] = 0.f; m1[rr] = -1.f; m2[rr] = -1.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            const float ex = c < Sk ? exp_neg(e[rr][i] - m[rr]) : 0.f;
            e[rr][i] = c < Sk ? ex : -1.f;
            sum[rr] += ex;
            m2[rr] = fmaxf(m2[rr], fminf(m1[rr], ex));       // running two largest
            m1[rr] = fmaxf(m1[rr], ex);
            if (live[rr] && c < s_pad) srow[rr][c] = ex;      // padding columns become 0 for the AV tiles
        }
    }
#pragma unroll
    for (int rr = 0; rr < RP; ++rr) sum[rr] = warp_sum(sum[rr]);
    // T = k-th largest of the 64 lane-top-2 keys (rank counting); at least k entries of the row are >= T
    unsigned T[RP];
    {
        unsigned k1[RP], k2[RP];
        int c1[RP], c2[RP];
#pragma unroll
        for (int rr = 0; rr < RP; ++rr) { k1[rr] = key_of(m1[rr]); k2[rr] = key_of(m2[rr]); c1[rr] = c2[rr] = 0; }
#pragma unroll 4
        for (int l = 0; l < 32; ++l) {
#pragma unroll
            for (int rr = 0; rr < RP; ++rr) {
                const unsigned a1 = __shfl_sync(FULL_MASK, k1[rr], l), a2 = __shfl_sync(FULL_MASK, k2[rr], l);
                c1[rr] += (a1 > k1[rr]) + (a2 > k1[rr]);
                c2[rr] += (a1 > k2[rr]) + (a2 > k2[rr]);
            }
        }
#pragma unroll
        for (int rr = 0; rr < RP; ++rr)
            T[rr] = __reduce_min_sync(FULL_MASK, c2[rr] < k ? k2[rr] : (c1[rr] < k ? k1[rr] : 0xffffffffu));
    }
    // compact the survivors {value >= T} in key order, straight from the registers
    int n[RP];
#pragma unroll
    for (int rr = 0; rr < RP; ++rr) {
        n[rr] = 0;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const bool pred = key_of(e[rr][i]) >= T[rr] && T[rr] > 0 && e[rr][i] >= 0.f;
            const unsigned bal = __ballot_sync(FULL_MASK, pred);
            const int pos = n[rr] + __popc(bal & ((1u << lane) - 1u));
            if (pred && pos < LIST_CAP) { lv[rr][pos] = e[rr][i]; lp[rr][pos] = lane + 32 * i; }
            n[rr] += __popc(bal);
        }
    }
    __syncwarp();
    // rank every survivor (2 per lane) among the survivors: descending value, ties by list position (= key index)
    float va[RP], vb[RP];
    int pa[RP], pb[RP], ra[RP], rb[RP];
    unsigned ka[RP], kb[RP];
#pragma unroll
    for (int rr = 0; rr < RP; ++rr) {
        va[rr] = lane < n[rr] ? lv[rr][lane] : -1.f;
        vb[rr] = lane + 32 < n[rr] ? lv[rr][lane + 32] : -1.f;
        pa[rr] = lane < n[rr] ? lp[rr][lane] : 0;
        pb[rr] = lane + 32 < n[rr] ? lp[rr][lane + 32] : 0;
        ka[rr] = key_of(va[rr]); kb[rr] = key_of(vb[rr]);
        ra[rr] = rb[rr] = 0;
    }
#pragma unroll 4
    for (int l = 0; l < 32; ++l) {
#pragma unroll
        for (int rr = 0; rr < RP; ++rr) {
            const unsigned oa = __shfl_sync(FULL_MASK, ka[rr], l), ob = __shfl_sync(FULL_MASK, kb[rr], l);
            ra[rr] += (oa > ka[rr] || (oa == ka[rr] && l < lane)) + (ob > ka[rr]);
            rb[rr] += (oa >= kb[rr]) + (ob > kb[rr] || (ob == kb[rr] && l < lane));
        }
    }
    __syncwarp();
#pragma unroll
    for (int rr = 0; rr < RP; ++rr) {
        rv[rr] = -1.f; ridx[rr] = 0;
        if (n[rr] <= LIST_CAP && n[rr] >= k) {
            if (lane < n[rr] && ra[rr] < k) { lv[rr][ra[rr]] = va[rr]; lp[rr][ra[rr]] = pa[rr]; }
            if (lane + 32 < n[rr] && rb[rr] < k) { lv[rr][rb[rr]] = vb[rr]; lp[rr][rb[rr]] = pb[rr]; }
        }
    }
    __syncwarp();
#pragma unroll
    for (int rr = 0; rr < RP; ++rr) {
        if (n[rr] <= LIST_CAP && n[rr] >= k) {
            if (lane < k) { rv[rr] = lv[rr][lane]; ridx[rr] = lp[rr][lane]; }
        } else if (live[rr]) {                       // massive ties: exact but slow path on the row itself
            int rj;
            warp_select_k(srow[rr], Sk, k, lane, rv[rr], rj);
            ridx[rr] = rj;
            __syncwarp();
            if (lane < k) srow[rr][ridx[rr]] = rv[rr];     // restore
        }
    }
    __syncwarp();
}

template <int ROWS, int NVF>       // NVF: values per lane of the register soft-max / top-k path, sized to the padded row (Sk <= 32 * NVF)
__global__ void __launch_bounds__(NW * 32, ROWS == 32 ? 2 : 3) qtatt_coarse_kernel(CoarseParams p, int s_ld) {
    pdl_sync();
    constexpr int HALVES = 32 / ROWS;              // lane groups sharing a row set (1 for 32 rows, 2 for 16, 4 for 8)
    constexpr int TOK_PER_WARP = TILE / NW;        // 8
    extern __shared__ __align__(16) float smem[];
    float *KVs = smem;                              // [2][TILE][D]
    float *Ss = KVs + 2 * TILE * D;                 // [ROWS][s_ld]       (aliased by the AV reduction buffer)
    const int slab_floats = ROWS * s_ld > NW * ROWS * RED_LD ? ROWS * s_ld : NW * ROWS * RED_LD;
    float *lval = Ss + slab_floats;                 // [NW][2][LIST_CAP]
    int *lpos = (int *)(lval + NW * 2 * LIST_CAP);  // [NW][2][LIST_CAP]
    float *rsum = (float *)(lpos + NW * 2 * LIST_CAP);  // [32] row sums of exp

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y / p.nh, h = blockIdx.y % p.nh;
    const int row0 = blockIdx.x * ROWS;
    const int C = p.nh * D;
    const float *qb = p.q + (size_t)b * p.Sq * C + h * D;
    const float *kb = p.k + (size_t)b * p.Sk * C + h * D;
    const float *vb = p.v + (size_t)b * p.Sk * C + h * D;
    const int n_tiles = (p.Sk + TILE - 1) / TILE;
    const int s_pad = n_tiles * TILE;
    const float scale = rsqrtf((float)D);           // 1/sqrt(32), same fp32 value as 1.0 / D ** 0.5
    const int myrow = lane % ROWS, half = lane / ROWS;
    const int grow = min(row0 + myrow, p.Sq - 1);   // clamped: out-of-range rows compute garbage that is never stored

    // ---- phase 1: scores
    const TileLoader<HALVES> loader(tid, C);
    loader.load(KVs, kb, 0, p.Sk, C);
    float4 q[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) q[c] = ldg4(qb + (size_t)grow * C + 4 * c);
    for (int kt = 0; kt < n_tiles; ++kt) {
        float *cur = KVs + (kt & 1) * TILE * D;
        if (kt + 1 < n_tiles) loader.load(KVs + ((kt + 1) & 1) * TILE * D, kb, (kt + 1) * TILE, p.Sk, C);
        else __pipeline_commit();
        __pipeline_wait_prior(1);
        __syncthreads();
        // warp w: tokens w*8 .. w*8+7 of the tile; with 16-row CTAs the two lane halves take 4 tokens each
        constexpr int TPL = TOK_PER_WARP / HALVES;  // tokens per lane group
        constexpr int TB = TPL < 4 ? TPL : 4;       // tokens per inner step (independent accumulation chains)
#pragma unroll
        for (int t0 = 0; t0 < TPL; t0 += TB) {
            const int tb = warp * TOK_PER_WARP + half * TPL + t0;
            float2 acc[TB];
#pragma unroll
            for (int j = 0; j < TB; ++j) acc[j] = make_float2(0.f, 0.f);
            const int sw = tok_swizzle<HALVES>(tb);     // the TB tokens of a step belong to one lane group: same swizzle
#pragma unroll
            for (int c = 0; c < 8; ++c) {
#pragma unroll
                for (int j = 0; j < TB; ++j) {
                    const float4 kv = *reinterpret_cast<const float4 *>(cur + (tb + j) * D + 4 * (c ^ sw));
                    acc[j] = __ffma2_rn(make_float2(q[c].x, q[c].y), make_float2(kv.x, kv.y), acc[j]);
                    acc[j] = __ffma2_rn(make_float2(q[c].z, q[c].w), make_float2(kv.z, kv.w), acc[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < TB; ++j) {
                const int tok = kt * TILE + tb + j;
                Ss[myrow * s_ld + tok] = tok < p.Sk ? (acc[j].x + acc[j].y) * scale : -INFINITY;
            }
        }
        __syncthreads();
    }

    // ---- phase 2 + 3: softmax and top-k, warp `warp` owns rows warp, warp+NW, ...
    constexpr int RPW = (ROWS + NW - 1) / NW;        // rows per warp
    float *mylv = lval + warp * 2 * LIST_CAP;        // two survivor lists per warp (the fast path handles 2 rows at once)
    int *mylp = lpos + warp * 2 * LIST_CAP;
    constexpr int RP = (RPW < 2 || NVF > 24) ? 1 : 2;   // rows processed together by the register fast path (one when a row alone fills the registers)
    // register fast path, sized to the padded row: NVF values per lane; 832^2 -> 676 keys -> 22, 640x480 -> 300 -> 16
    if (s_pad <= 32 * NVF) {
        for (int j0 = 0; j0 < RPW; j0 += RP) {
            float *srow[RP], *lvp[RP];
            int *lpp[RP];
            bool live[RP];
            float sum[RP], rv[RP];
            int ridx[RP], rws[RP];
#pragma unroll
            for (int rr = 0; rr < RP; ++rr) {
                rws[rr] = warp + (j0 + rr) * NW;
                live[rr] = j0 + rr < RPW && rws[rr] < ROWS && row0 + rws[rr] < p.Sq;
                srow[rr] = Ss + (live[rr] ? rws[rr] : 0) * s_ld;
                lvp[rr] = mylv + rr * LIST_CAP;
                lpp[rr] = mylp + rr * LIST_CAP;
            }
            rows_softmax_topk<NVF, RP>(srow, live, p.Sk, s_pad, p.topk, lane, lvp, lpp, sum, rv, ridx);
#pragma unroll
            for (int rr = 0; rr < RP; ++rr) {
                if (!live[rr]) continue;
                if (lane == 0) rsum[rws[rr]] = sum[rr];
                if (lane < p.topk) {
                    const size_t o = (((size_t)b * p.Sq + row0 + rws[rr]) * p.nh + h) * p.topk + lane;
                    p.topk_idx[o] = ridx[rr];
                    p.topk_score[o] = rv[rr] / sum[rr];
                    if (p.type_a) srow[rr][ridx[rr]] = 0.f;      // QTAttA: selected keys leave the message (:37-42)
                }
            }
        }
    }
    else
    for (int r = warp; r < ROWS; r += NW) {
        if (row0 + r >= p.Sq) continue;            // warp-uniform
        float *srow = Ss + r * s_ld;
        float m = -INFINITY;
        for (int e = lane; e < p.Sk; e += 32) m = fmaxf(m, srow[e]);
        m = warp_max(m);
        // the slab keeps the UNNORMALISED exp(s - max): the order is the same, only the k selected scores and the
        // A.V result are divided by the row sum (phase 4)
        float sum = 0.f;
        float m1 = -1.f, m2 = -1.f;                 // lane-local two largest
        for (int e = lane; e < s_pad; e += 32) {
            float ex = 0.f;                         // padding columns become 0 for the AV tiles
            if (e < p.Sk) {
                ex = exp_neg(srow[e] - m);
                sum += ex;
                if (ex > m1) { m2 = m1; m1 = ex; } else if (ex > m2) { m2 = ex; }
            }
            srow[e] = ex;
        }
        sum = warp_sum(sum);
        if (lane == 0) rsum[r] = sum;
        // T = k-th largest of the 64 lane-top-2 keys: a key's rank is the number of keys strictly above it;
        // the k-th largest is the smallest key with rank < k.  At least k entries of the row are >= T.
        const unsigned k1 = key_of(m1), k2 = key_of(m2);
        int c1 = 0, c2 = 0;
#pragma unroll 8
        for (int l = 0; l < 32; ++l) {
            const unsigned a1 = __shfl_sync(FULL_MASK, k1, l), a2 = __shfl_sync(FULL_MASK, k2, l);
            c1 += (a1 > k1) + (a2 > k1);
            c2 += (a1 > k2) + (a2 > k2);
        }
        const unsigned candT = c2 < p.topk ? k2 : (c1 < p.topk ? k1 : 0xffffffffu);
        const unsigned T = __reduce_min_sync(FULL_MASK, candT);
        __syncwarp();
        // compact survivors {a >= T} in key order
        int n = 0;
        for (int e0 = 0; e0 < p.Sk; e0 += 32) {
            const int e = e0 + lane;
            const float a = e < p.Sk ? srow[e] : -1.f;
            const bool pred = key_of(a) >= T && T > 0;
            const unsigned bal = __ballot_sync(FULL_MASK, pred);
            if (bal) {
                const int pos = n + __popc(bal & ((1u << lane) - 1u));
                if (pred && pos < LIST_CAP) { mylv[pos] = a; mylp[pos] = e; }
                n += __popc(bal);
            }
        }
        __syncwarp();
        float rv = -1.f;
        int ridx = 0;
        if (n <= LIST_CAP && n >= p.topk) {
            // rank each survivor (2 per lane) among all survivors: descending value, ties by list position
            // (= key index, the list is in key order); rank < k -> that lane owns output slot `rank`
            const float va = lane < n ? mylv[lane] : -1.f, vb2 = lane + 32 < n ? mylv[lane + 32] : -1.f;
            const unsigned ka = key_of(va), kb2 = key_of(vb2);
            int ra = 0, rb = 0;
#pragma unroll 8
            for (int l = 0; l < 32; ++l) {
                const unsigned oa = __shfl_sync(FULL_MASK, ka, l), ob = __shfl_sync(FULL_MASK, kb2, l);
                ra += (oa > ka || (oa == ka && l < lane)) + (ob > ka);
                rb += (oa >= kb2) + (ob > kb2 || (ob == kb2 && l < lane));
            }
            __syncwarp();
            // scatter (value, key index) to the slot given by the rank, then lane `slot` picks it up
            const int pa = lane < n ? mylp[lane] : 0, pb = lane + 32 < n ? mylp[lane + 32] : 0;
            __syncwarp();
            if (lane < n && ra < p.topk) { mylv[ra] = va; mylp[ra] = pa; }
            if (lane + 32 < n && rb < p.topk) { mylv[rb] = vb2; mylp[rb] = pb; }
            __syncwarp();
            if (lane < p.topk) { rv = mylv[lane]; ridx = mylp[lane]; }
        } else {                                    // massive ties: exact but slow path on the row itself
            int rj;
            warp_select_k(srow, p.Sk, p.topk, lane, rv, rj);
            ridx = rj;
            __syncwarp();
            if (lane < p.topk) srow[ridx] = rv;     // restore
        }
        __syncwarp();
        if (lane < p.topk) {
            const size_t o = (((size_t)b * p.Sq + row0 + r) * p.nh + h) * p.topk + lane;
            p.topk_idx[o] = ridx;
            p.topk_score[o] = rv / sum;
            if (p.type_a) srow[ridx] = 0.f;         // QTAttA: selected keys leave the message (:37-42)
        }
    }
    __syncthreads();

    // ---- phase 4: O = A V.  lane = row, 32 output dims in registers; warp w takes tokens w*8.. of each tile
    float2 o[D / 2];
#pragma unroll
    for (int d = 0; d < D / 2; ++d) o[d] = make_float2(0.f, 0.f);
    loader.load(KVs, vb, 0, p.Sk, C);
    for (int vt = 0; vt < n_tiles; ++vt) {
        float *cur = KVs + (vt & 1) * TILE * D;
        if (vt + 1 < n_tiles) loader.load(KVs + ((vt + 1) & 1) * TILE * D, vb, (vt + 1) * TILE, p.Sk, C);
        else __pipeline_commit();
        __pipeline_wait_prior(1);
        __syncthreads();
        constexpr int TPL = TOK_PER_WARP / HALVES;
#pragma unroll 2
        for (int t = 0; t < TPL; ++t) {
            const int tl = warp * TOK_PER_WARP + half * TPL + t;
            const float a = Ss[myrow * s_ld + vt * TILE + tl];
            const float2 aa = make_float2(a, a);
            const int sw = tok_swizzle<HALVES>(tl);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 vv = *reinterpret_cast<const float4 *>(cur + tl * D + 4 * (c ^ sw));
                o[2 * c + 0] = __ffma2_rn(aa, make_float2(vv.x, vv.y), o[2 * c + 0]);
                o[2 * c + 1] = __ffma2_rn(aa, make_float2(vv.z, vv.w), o[2 * c + 1]);
            }
        }
        __syncthreads();
    }
    // lane groups of a warp hold partial sums of the same rows: fold them with shuffles, then reduce across warps
    // through smem (aliased onto the score slab)
#pragma unroll
    for (int m = ROWS; m < 32; m <<= 1)
#pragma unroll
        for (int d = 0; d < D / 2; ++d) {
            o[d].x += __shfl_xor_sync(FULL_MASK, o[d].x, m);
            o[d].y += __shfl_xor_sync(FULL_MASK, o[d].y, m);
        }
    float *red = Ss;                                // [NW][ROWS][RED_LD]
    if (half == 0) {
#pragma unroll
        for (int d = 0; d < D / 2; ++d) {
            red[(warp * ROWS + myrow) * RED_LD + 2 * d] = o[d].x;
            red[(warp * ROWS + myrow) * RED_LD + 2 * d + 1] = o[d].y;
        }
    }
    __syncthreads();
    float w0 = 1.f;
    if (p.level_weight) {                           // softmax over the level weights (:264)
        float mx = -INFINITY, den = 0.f;
        for (int l = 0; l < p.n_weights; ++l) mx = fmaxf(mx, __ldg(p.level_weight + l));
        for (int l = 0; l < p.n_weights; ++l) den += expf(__ldg(p.level_weight + l) - mx);
        w0 = expf(__ldg(p.level_weight) - mx) / den;
        if (p.wsm && blockIdx.x == 0 && blockIdx.y == 0 && tid < p.levels)      // the finer levels read their weight from here
            p.wsm[tid] = expf(__ldg(p.level_weight + tid) - mx) / den;
    }
    for (int i = tid; i < ROWS * D; i += NW * 32) {
        const int r = i >> 5, d = i & 31;
        if (row0 + r < p.Sq) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) s += red[(w * ROWS + r) * RED_LD + d];
            p.acc[((size_t)b * p.Sq + row0 + r) * C + h * D + d] = (s / rsum[r]) * w0;
        }
    }
}

int coarse_s_ld(int Sk) { return ((Sk + TILE - 1) / TILE * TILE) | 1; }     // odd: conflict-free lane=row access

size_t smem_bytes(int Sk, int rows) {
    const size_t slab = (size_t)rows * coarse_s_ld(Sk);
    const size_t red = (size_t)NW * rows * RED_LD;                           // aliased onto the slab
    return sizeof(float) * (2 * TILE * D + (slab > red ? slab : red) + NW * 2 * LIST_CAP + 32) + sizeof(int) * NW * 2 * LIST_CAP;
}

}  // namespace

size_t coarse_smem_bytes(int Sk) { return smem_bytes(Sk, 8); }

int launch_qtatt_coarse(const CoarseParams &p, cudaStream_t stream) {
    if (p.tc_ws != nullptr && coarse_tc_applicable(p.Sq, p.Sk, p.topk)) return launch_qtatt_coarse_tc(p, p.tc_ws, stream);
    // 32-row CTAs are the efficient shape (every lane owns a row).  With too few of them to give every SM two, halve the
    // rows per CTA (lane groups split the tokens).  Measured at 832^2, B = 1 (676 rows x 8 heads): 32 rows 76 us, 16 rows
    // 58 us, 8 rows 64 us (every CTA streams all K and V tiles, so below 16 rows the tile traffic and barriers dominate).
    const long long col = (long long)p.B * p.nh * casmtr_concurrency();      // CTA columns in flight, counting the caller's concurrent calls
    int rows = 32;
    while (rows > 16 && ((p.Sq + rows - 1) / rows) * col < 2 * 148) rows >>= 1;
    static const int rows_env = [] { const char *e = getenv("CASMTR_COARSE_ROWS"); return e ? atoi(e) : 0; }();
    if (rows_env == 8 || rows_env == 16 || rows_env == 32) rows = rows_env;
    while (rows > 8 && smem_bytes(p.Sk, rows) > 227 * 1024) rows >>= 1;
    const size_t smem = smem_bytes(p.Sk, rows);
    CASMTR_REQUIRE(smem <= 227 * 1024, CASMTR_E_UNSUPPORTED,
                   "coarsest level has %d keys; the dense level supports at most ~6000 (shared memory)", p.Sk);
    CASMTR_REQUIRE(p.topk >= 1 && p.topk <= 32 && p.topk <= p.Sk, CASMTR_E_INVALID, "coarse top-k %d must be in [1, min(32, %d)]", p.topk, p.Sk);
    // kernel variant: rows per CTA x values per lane of the register soft-max / top-k path (1024^2 -> 1024 keys -> 32, 1152^2 -> 1296 -> 42;
    // beyond 1344 keys the looped path inside the <.., 42> variant)
    using Kern = void (*)(CoarseParams, int);
    static const Kern table[3][6] = {
        {qtatt_coarse_kernel<32, 8>, qtatt_coarse_kernel<32, 16>, qtatt_coarse_kernel<32, 22>, qtatt_coarse_kernel<32, 24>, qtatt_coarse_kernel<32, 32>, qtatt_coarse_kernel<32, 42>},
        {qtatt_coarse_kernel<16, 8>, qtatt_coarse_kernel<16, 16>, qtatt_coarse_kernel<16, 22>, qtatt_coarse_kernel<16, 24>, qtatt_coarse_kernel<16, 32>, qtatt_coarse_kernel<16, 42>},
        {qtatt_coarse_kernel<8, 8>, qtatt_coarse_kernel<8, 16>, qtatt_coarse_kernel<8, 22>, qtatt_coarse_kernel<8, 24>, qtatt_coarse_kernel<8, 32>, qtatt_coarse_kernel<8, 42>}};
    static PerDeviceOnce once;
    const int dev = PerDeviceOnce::device();
    if (!once.done(dev)) {
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 6; ++j) {
                if (e == cudaSuccess) e = cudaFuncSetAttribute(table[i][j], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                if (e == cudaSuccess) e = cudaFuncSetAttribute(table[i][j], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            }
        if (e != cudaSuccess) { casmtr_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return CASMTR_E_CUDA; }
        once.mark(dev);
    }
    const int s_pad = (p.Sk + TILE - 1) / TILE * TILE;
    const int nv = s_pad <= 32 * 8 ? 0 : (s_pad <= 32 * 16 ? 1 : (s_pad <= 32 * 22 ? 2 : (s_pad <= 32 * 24 ? 3 : (s_pad <= 32 * 32 ? 4 : 5))));   // 5: <= 1344 keys, else looped
    dim3 grid((p.Sq + rows - 1) / rows, p.B * p.nh);
    LaunchScope ls(CASMTR_K_QT_COARSE, stream);
    launch_k(table[rows == 32 ? 0 : (rows == 16 ? 1 : 2)][nv], grid, NW * 32, smem, stream, p, coarse_s_ld(p.Sk));
    CASMTR_CHECK_LAUNCH("qtatt_coarse_kernel");
    return CASMTR_OK;
}
