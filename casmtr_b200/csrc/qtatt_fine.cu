// Fused quadtree fine level / cascade window attention ("quad attention"), one work item per warp.
//
// One warp = one (batch, parent token, head): the 4 sibling queries of the parent attend to the
// KC = 4*kp candidate keys (the 4 children of each of the parent's kp selected coarser keys).  The
// kernel fuses what the reference does with ~40 torch ops and 2 custom kernels per level:
// candidate expansion, gathered Q.K^T, softmax, top-k for the next level, gathered A.V, the
// child-major -> raster reorder and the (weighted) merge with the coarser levels' message.
// No index / QK / A tensor ever reaches HBM.
// Reference: QTAttB.process_fine_level + merge
//   cuda_imp/QuadTreeAttention/QuadtreeAttention/modules/quadtree_attention.py:180-229, 262-284
// QTAttA.process_fine_level :46-99, merge :130-138;  CascadeQTAttB.forward :400-452;
// kernels replaced: src/score_computation_kernal.cu:22-62, src/value_aggregation_kernel.cu:21-42.
//
// Data movement.  The kernel is a gather: every candidate K and V row of a head is one 128-byte line
// of the token-major feature map, mostly served by L2 (the maps are 11-22 MB).  A warp first issues ALL
// of its gathers as 16-byte cp.async (LDGSTS) straight into its private shared-memory slab -- 8 lanes
// per row, 4 rows per instruction, no registers held -- and only then computes:
//   (Measured alternative, round 2: one cp.async.bulk -- TMA without a tensor map -- per 128-byte row, completed on a per-warp
//   mbarrier, rows unswizzled and read in a per-lane rotated chunk order.  It removes ~130 of the ~1100 instructions of a last-level
//   item and is 30 % SLOWER, 81 -> 107 us per launch: the copy engine's per-request cost dominates at 128 bytes per request.)
//   Q.K^T    lane = candidate (round r: candidate 32r + lane).  K rows are XOR-swizzled in 16-byte chunks
//            so the per-lane row reads are bank-conflict free; the 4 sibling q rows are broadcast reads,
//            loaded once per chunk and reused across the rounds.  FMAs are issued as packed FFMA2
//            (2 fp32 FMAs per instruction, sm_100): the kernel is issue-bound, not bandwidth-bound.
//   softmax  warp reductions (type B: over all candidates; type A: over the 4 children of a parent
//            candidate = 4 adjacent lanes, times the parent's score).
//   top-k    threshold + rank counting, no serial selection: T = k-th largest of the 32 lane maxima,
//            survivors {a >= T} are compacted (k <= n, typically n < 2k) and ranked against each other
//            with shuffles; rank < k owns output slot `rank` (descending order, ties: lower candidate).
//   A.V      lane = (child slot g, 16-byte chunk dq): 3 LDS.128 per 8 FFMA2, then a transposing
//            butterfly over g leaves query g's output chunk in the lane; merged and stored raster.
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"

namespace {

constexpr int D = 32;

__device__ __forceinline__ unsigned okey(float v) { return v >= 0.f ? __float_as_uint(v) + 1u : 0u; }   // order preserving, 0 = dead

__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// acc.x + acc.y accumulates dot(q, k) over a 4-float chunk with two packed FMAs
__device__ __forceinline__ float2 dot4p(const float4 q, const float4 k, float2 acc) {
    acc = __ffma2_rn(make_float2(q.x, q.y), make_float2(k.x, k.y), acc);
    return __ffma2_rn(make_float2(q.z, q.w), make_float2(k.z, k.w), acc);
}

template <int R>
__device__ __forceinline__ float pick(const float (&x)[R][4], int r, int f) {      // x[r][f] with a runtime f, no local memory
    return f == 0 ? x[r][0] : f == 1 ? x[r][1] : f == 2 ? x[r][2] : x[r][3];
}

// floats of shared memory per work item for kp parent candidates:
//   K slab [max(KC,32)][32] (later aliased by A2[KC][8], top-k staging and scratch), V slab [KC][32], Q [4][32]
__host__ __device__ inline int staged_slab_floats(int kp) {          // K + V + Q (CTA-per-item kernel)
    const int kc = 4 * kp;
    const int krows = kc < 32 ? 32 : kc;
    return krows * D + kc * D + 4 * D;
}
// warp-per-item kernel: only K and Q are staged; V rows are streamed straight into registers in the A.V loop (each lane
// consumes exactly what it loads, so staging V would only cost shared memory -- and shared memory is what caps the
// number of resident warps of this latency-bound kernel: 8.7 KB instead of 16.9 KB per warp at kp = 16)
__host__ __device__ inline int warp_slab_floats(int kp) {
    const int kc = 4 * kp;
    const int krows = kc < 32 ? 32 : kc;
    return krows * D + 4 * D;
}

// KP > 0: compile-time number of parent candidates (gather loops fully unrolled); KP == 0: runtime p.kp
template <int KP, int R, bool CASCADE, bool TYPE_A, bool DO_TOPK, bool PE = false>
__device__ __forceinline__ void quad_item(const FineParams &p, float *Ks, int b, int parent, int h, int lane) {
    const int g = lane >> 3, dq = lane & 7;
    const int wp = p.w0 >> 1;
    const int Np = (p.h0 >> 1) * wp;
    const int py = p.d_wp.div(parent), px = parent - py * wp;
    const int C = p.nh * D, L0 = p.h0 * p.w0, L1 = p.h1 * p.w1;
    const int kp = KP ? KP : p.kp, KC = 4 * kp;
    const float scale = rsqrtf((float)D);

    // Ks: this warp's slab, [max(KC,32)][32], chunk-swizzled
    float *Qs = Ks + (KC < 32 ? 32 : KC) * D;                  // [4][32]
    // after Q.K^T the K slab and Q are dead and get reused:
    float *A2 = Ks;                                            // [KC][8]: (a0,a0,a1,a1,a2,a2,a3,a3)
    float *stg_sc = Ks + 8 * KC;                               // [4][32] top-k scores        (8*KC + 256 <= max(KC,32)*32)
    unsigned *lkey = (unsigned *)(Ks + 8 * KC + 128);          // [64] survivor keys
    int *lpos = (int *)(Ks + 8 * KC + 192);                    // [64] survivor candidate positions
    int *stg_idx = (int *)Qs;                                  // [4][32] top-k token indices

    const int off_g = (g >> 1) * p.dil * p.w1 + (g & 1) * p.dil;
    const int qtok0 = 2 * py * p.w0 + 2 * px;      // sibling f of the parent is query token qtok0 + (f>>1)*w0 + (f&1)
#define QTOK(f) (qtok0 + ((f) >> 1) * p.w0 + ((f) & 1))
    // everything that does not depend on the candidate list goes out first: the q rows and the coarser levels' message
    cp_async16(Qs + g * D + 4 * dq, p.q + ((size_t)b * L0 + QTOK(g)) * C + h * D + 4 * dq);
    float4 ap = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.acc_prev) ap = ldg4(p.acc_prev + ((size_t)b * Np + parent) * C + h * D + 4 * dq);

    // ---- candidate bases: lane k (< kp) holds the top-left child of parent-candidate k
    int base = 0;
    float pscore = 0.f;
    if (lane < kp) {
        if (CASCADE) {
            if (p.next_idx != nullptr) {             // window derived from the parent's match on the previous grid
                const int hv = p.h1 >> 1, wv = p.w1 >> 1;
                const int idx = (int)__ldg(p.next_idx + (size_t)b * Np + parent);
                int ir, ic, lr, lc;
                p.d_wv.divmod(idx, ir, ic);
                p.d_win.divmod(lane, lr, lc);
                const int r0 = window_origin(ir, p.win, hv), c0 = window_origin(ic, p.win, wv);
                base = 2 * (r0 + lr) * p.w1 + 2 * (c0 + lc);
            } else {
                const int64_t *tp = p.topk_pos + (((size_t)b * Np + parent) * kp + lane) * 2;
                base = (int)(2 * tp[0] * p.w1 + 2 * tp[1]);
            }
        } else {
            const size_t o = (((size_t)b * Np + parent) * p.nh + h) * kp + lane;
            const int idx = p.prev_idx[o];
            const int r = p.d_wprev.div(idx);
            base = 2 * r * p.w1 + 2 * (idx - r * p.w_prev);
            if (TYPE_A) pscore = p.prev_score[o];
        }
    }

    // ---- issue every K gather of this item (one LDGSTS = the 4 children of a parent candidate)
    const float *vb = p.v + (size_t)b * L1 * C + h * D + 4 * dq;
    {
        const float *kb = p.k + (size_t)b * L1 * C + h * D + 4 * dq;
        // row 4u+g lands at chunk dq ^ ((4u+g) & 7) = dq ^ (4*(u&1) + g)
        float *kd0 = Ks + g * D + 4 * (dq ^ g), *kd1 = Ks + g * D + 4 * (dq ^ (4 + g));
#pragma unroll
        for (int u = 0; u < (KP ? KP : 32); ++u) {
            if (KP == 0 && u >= kp) break;
            int tok = __shfl_sync(FULL_MASK, base, u) + off_g;
            if (CASCADE) tok = min(max(tok, 0), L1 - 1);       // the reference's clamp (:428); QTAtt children are always in range
            cp_async16(((u & 1) ? kd1 : kd0) + u * 4 * D, kb + (size_t)tok * C);
        }
        cp_async_commit();
    }

    // ---- gathered Q.K^T, lane = candidate
    cp_async_wait<0>();
    __syncwarp();
    float sc[R][4];
    {
        float2 acc[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[r][f] = make_float2(0.f, 0.f);
        const float *krow[R];                       // rows past KC re-read the last valid row; their scores are masked below
        int ksw[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int c = min(32 * r + lane, KC - 1);
            krow[r] = Ks + c * D;
            ksw[r] = c & 7;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 qv[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) qv[f] = *reinterpret_cast<const float4 *>(Qs + f * D + 4 * j);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 kv = *reinterpret_cast<const float4 *>(krow[r] + 4 * (j ^ ksw[r]));
#pragma unroll
                for (int f = 0; f < 4; ++f) acc[r][f] = dot4p(qv[f], kv, acc[r][f]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int f = 0; f < 4; ++f) sc[r][f] = acc[r][f].x + acc[r][f].y;
    }
    __syncwarp();                                   // K slab and Q are dead from here on (A2 / staging alias them)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool valid = 32 * r + lane < KC;
#pragma unroll
        for (int f = 0; f < 4; ++f) sc[r][f] = valid ? sc[r][f] * scale : -INFINITY;
    }
    if (CASCADE && p.rel_pos != nullptr) {          // relative position bias read as a tensor (one test per item, not per score)
        const float *rp = p.rel_pos + ((size_t)b * p.nh + h) * L0 * KC;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int c = 32 * r + lane;
            if (c < KC) {
#pragma unroll
                for (int f = 0; f < 4; ++f) sc[r][f] += __ldg(rp + (size_t)QTOK(f) * KC + c);
            }
        }
    }
    if (CASCADE && PE) {                            // the same bias from its embedding tables (get_relative_pe, transformer.py:473-509)
        const int2 qt = relpe_query_term(p.pe, b, 2 * py, 2 * px);     // s is even: sibling f adds (f & 1, f >> 1)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int c = 32 * r + lane, cf = c & 3;
            // the reference derives (row, col) from the flat, un-clamped child index (:489-499)
            const int tok = __shfl_sync(FULL_MASK, base, (c >> 2) & 31) + (cf >> 1) * p.dil * p.w1 + (cf & 1) * p.dil;
            const int ky = tok / p.w1, kx = tok - ky * p.w1;
            if (c < KC) {
#pragma unroll
                for (int f = 0; f < 4; ++f) sc[r][f] += relpe_bias(p.pe, p.nh, h, qt, ky + (f >> 1), kx + (f & 1));
            }
        }
    }

    // ---- softmax -> a[r][f]; sc[r][f] becomes the selection score (>= 0, -1 = invalid)
    float a[R][4];
    // (Deferring the 1 / sum scaling to the output, as cascade_tile.cu does, was measured here and lost: the compiler then hoists all
    // streamed V loads above the soft-max, 72 -> 122 registers, last level 84 -> 85 us uncapped and 93 us capped with spills.)
    if (!TYPE_A) {      // over all candidates of sibling f
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            float m = sc[0][f];
#pragma unroll
            for (int r = 1; r < R; ++r) m = fmaxf(m, sc[r][f]);
            m = warp_max(m);
            float sum = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r) { a[r][f] = exp_neg(sc[r][f] - m); sum += a[r][f]; }
            sum = warp_sum(sum);
            const float inv = 1.0f / sum;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                a[r][f] = a[r][f] * inv;
                sc[r][f] = 32 * r + lane < KC ? a[r][f] : -1.f;
            }
        }
    } else {            // QTAttA: over the 4 children of each parent candidate (4 adjacent lanes), times the parent's score (:72-77)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool valid = 32 * r + lane < KC;
            const float ps = __shfl_sync(FULL_MASK, pscore, (8 * r + (lane >> 2)) & 31);
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                float m = sc[r][f];
                m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 1));
                m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 2));
                const float e = valid ? exp_neg(sc[r][f] - m) : 0.f;
                float sum = e;
                sum += __shfl_xor_sync(FULL_MASK, sum, 1);
                sum += __shfl_xor_sync(FULL_MASK, sum, 2);
                a[r][f] = valid ? (e / sum) * ps : 0.f;
                sc[r][f] = valid ? a[r][f] : -1.f;      // type A selects on the redistributed score
            }
        }
    }

    // ---- top-k for the next level
    if (DO_TOPK) {
        const int k = p.topk;
#pragma unroll 1
        for (int f = 0; f < 4; ++f) {
            unsigned key[R];
            unsigned km = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                key[r] = okey(pick<R>(sc, r, f));
                km = max(km, key[r]);
            }
            // T = k-th largest of the 32 lane maxima; at least k candidates are >= T
            int above = 0;
#pragma unroll 8
            for (int l = 0; l < 32; ++l) above += __shfl_sync(FULL_MASK, km, l) > km;
            const unsigned T = __reduce_min_sync(FULL_MASK, (above < k && km > 0) ? km : 0xffffffffu);
            // compact the survivors {key >= T} in candidate order
            int n = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool pred = key[r] >= T && key[r] > 0;
                const unsigned bal = __ballot_sync(FULL_MASK, pred);
                const int pos = n + __popc(bal & ((1u << lane) - 1u));
                if (pred && pos < 64) { lkey[pos] = key[r]; lpos[pos] = 32 * r + lane; }
                n += __popc(bal);
            }
            __syncwarp();
            int slot_c = 0;                          // candidate position owning output slot `lane`
            if (n <= 64) {
                // rank every survivor among the survivors: descending key, ties by list position (= candidate order)
                const unsigned ka = lane < n ? lkey[lane] : 0u, kb2 = lane + 32 < n ? lkey[lane + 32] : 0u;
                const int pa = lane < n ? lpos[lane] : 0, pb = lane + 32 < n ? lpos[lane + 32] : 0;
                int ra = 0, rb = 0;
                if (n <= 32) {
#pragma unroll 8
                    for (int l = 0; l < 32; ++l) {
                        const unsigned oa = __shfl_sync(FULL_MASK, ka, l);
                        ra += (oa > ka) || (oa == ka && l < lane);
                    }
                } else {
#pragma unroll 8
                    for (int l = 0; l < 32; ++l) {
                        const unsigned oa = __shfl_sync(FULL_MASK, ka, l), ob = __shfl_sync(FULL_MASK, kb2, l);
                        ra += ((oa > ka) || (oa == ka && l < lane)) + (ob > ka);
                        rb += (oa >= kb2) + ((ob > kb2) || (ob == kb2 && l < lane));
                    }
                }
                __syncwarp();                        // every lane holds its list entries in registers
                if (lane < n && ra < k) lpos[ra] = pa;
                if (lane + 32 < n && rb < k) lpos[rb] = pb;
                __syncwarp();
                if (lane < k) slot_c = lpos[lane];
            } else {
                // massive ties (> 64 survivors): serial warp arg-max, exact but slow
                for (int it = 0; it < k; ++it) {
                    unsigned best = 0;
                    int br = 0;
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (key[r] > best) { best = key[r]; br = r; }
                    const unsigned gmax = __reduce_max_sync(FULL_MASK, best);
                    const int owner = __ffs(__ballot_sync(FULL_MASK, best == gmax)) - 1;
                    const int c = __shfl_sync(FULL_MASK, 32 * br + lane, owner);
                    if (lane == owner) {
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            if (r == br) key[r] = 0;
                    }
                    if (lane == it) slot_c = c;
                }
            }
            __syncwarp();
            // slot -> (token index at this level, score); the owner lane of candidate c holds its score
            float av = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float t = __shfl_sync(FULL_MASK, pick<R>(a, r, f), slot_c & 31);
                if ((slot_c >> 5) == r) av = t;
            }
            const int cf = slot_c & 3;
            int cand = __shfl_sync(FULL_MASK, base, (slot_c >> 2) & 31) + (cf >> 1) * p.dil * p.w1 + (cf & 1) * p.dil;
            if (CASCADE) cand = min(max(cand, 0), L1 - 1);
            if (lane < k) { stg_idx[f * 32 + lane] = cand; stg_sc[f * 32 + lane] = av; }
            if (TYPE_A && !p.final_level) {          // selected keys leave the message (:81-84)
                for (int s = 0; s < k; ++s) {
                    const int cs = __shfl_sync(FULL_MASK, slot_c, s);
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (cs == 32 * r + lane) {
                            if (f == 0) a[r][0] = 0.f; else if (f == 1) a[r][1] = 0.f; else if (f == 2) a[r][2] = 0.f; else a[r][3] = 0.f;
                        }
                }
            }
        }
        __syncwarp();
        for (int i = lane; i < 4 * k; i += 32) {
            const int f = i / k, kk = i - f * k;
            const size_t o = (((size_t)b * L0 + QTOK(f)) * p.nh + h) * k + kk;
            p.topk_idx[o] = stg_idx[f * 32 + kk];
            p.topk_score[o] = stg_sc[f * 32 + kk];
        }
    }

    // ---- A -> smem as [candidate][sibling pairs]: the A.V loop reads broadcast pairs for FFMA2
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (32 * r + lane < KC) {
            float4 *dst = reinterpret_cast<float4 *>(A2 + (32 * r + lane) * 8);
            dst[0] = make_float4(a[r][0], a[r][0], a[r][1], a[r][1]);
            dst[1] = make_float4(a[r][2], a[r][2], a[r][3], a[r][3]);
        }
    __syncwarp();

    // ---- gathered A.V: lane accumulates all 4 siblings x its 4 dims over the candidates of child slot g
    float2 o[4][2];
#pragma unroll
    for (int f = 0; f < 4; ++f) o[f][0] = o[f][1] = make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < (KP ? KP : 32); ++u) {
        if (KP == 0 && u >= kp) break;
        int vtok = __shfl_sync(FULL_MASK, base, u) + off_g;
        if (CASCADE) vtok = min(max(vtok, 0), L1 - 1);
        const float4 vv = ldg4(vb + (size_t)vtok * C);          // streamed: issued ahead by the unrolled loop
        const float4 a01 = *reinterpret_cast<const float4 *>(A2 + (4 * u + g) * 8);
        const float4 a23 = *reinterpret_cast<const float4 *>(A2 + (4 * u + g) * 8 + 4);
        const float2 vlo = make_float2(vv.x, vv.y), vhi = make_float2(vv.z, vv.w);
        o[0][0] = __ffma2_rn(make_float2(a01.x, a01.y), vlo, o[0][0]); o[0][1] = __ffma2_rn(make_float2(a01.x, a01.y), vhi, o[0][1]);
        o[1][0] = __ffma2_rn(make_float2(a01.z, a01.w), vlo, o[1][0]); o[1][1] = __ffma2_rn(make_float2(a01.z, a01.w), vhi, o[1][1]);
        o[2][0] = __ffma2_rn(make_float2(a23.x, a23.y), vlo, o[2][0]); o[2][1] = __ffma2_rn(make_float2(a23.x, a23.y), vhi, o[2][1]);
        o[3][0] = __ffma2_rn(make_float2(a23.z, a23.w), vlo, o[3][0]); o[3][1] = __ffma2_rn(make_float2(a23.z, a23.w), vhi, o[3][1]);
    }
    // reduce over g (lane bits 3,4), transposing: lane ends with sibling f = g
    float ov[4][4];
#pragma unroll
    for (int f = 0; f < 4; ++f) { ov[f][0] = o[f][0].x; ov[f][1] = o[f][0].y; ov[f][2] = o[f][1].x; ov[f][3] = o[f][1].y; }
    float r2[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float recv = __shfl_xor_sync(FULL_MASK, (g & 2) ? ov[i][c] : ov[i + 2][c], 16);
            r2[i][c] = ((g & 2) ? ov[i + 2][c] : ov[i][c]) + recv;
        }
    float m4[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float recv = __shfl_xor_sync(FULL_MASK, (g & 1) ? r2[0][c] : r2[1][c], 8);
        m4[c] = ((g & 1) ? r2[1][c] : r2[0][c]) + recv;
    }

    // ---- merge with the coarser levels and write raster (:262-284)
    const float wl = p.wsm ? __ldg(p.wsm + p.level) : 1.f;      // softmax(weight)[level], normalised once by the coarse kernel
    float4 res = make_float4(m4[0] * wl, m4[1] * wl, m4[2] * wl, m4[3] * wl);
    res.x = ap.x + res.x; res.y = ap.y + res.y; res.z = ap.z + res.z; res.w = ap.w + res.w;
    *reinterpret_cast<float4 *>(p.out + ((size_t)b * L0 + QTOK(g)) * C + h * D + 4 * dq) = res;

    // ---- cascade: the window's key indices for the 4 children (the reference's upsampled_idx, :450)
    if (CASCADE && p.upsampled_idx != nullptr && h == 0) {
        for (int s0 = 0; s0 < KC; s0 += 32) {
            const int s = s0 + lane;
            const int cf = s & 3;
            const int cand = min(max(__shfl_sync(FULL_MASK, base, (s >> 2) & 31) + (cf >> 1) * p.dil * p.w1 + (cf & 1) * p.dil, 0), L1 - 1);
            if (s < KC) {
#pragma unroll
                for (int f = 0; f < 4; ++f) p.upsampled_idx[((size_t)b * L0 + QTOK(f)) * KC + s] = cand;
            }
        }
    }
#undef QTOK
}

// ------------------------------------------------------------------------------------------------------------------
// QTAtt fine levels, CTA-per-item variant: 4 warps = the 4 sibling queries of one (parent, head).  The K / V / Q slab is
// shared by the CTA, so shared memory is spent per ITEM while the warp count is 4x that of the warp-per-item kernel:
// 13 (kp = 16) or 6 (kp = 32) resident items per SM = 52 / 24 warps instead of 12 / 6, which is what these latency-bound
// gathers need.  Each warp: issues a quarter of the item's cp.async gathers, then for ITS sibling lane=candidate Q.K^T,
// softmax, top-k and A.V.  Top-k needs no shared memory and no sort: T = k-th largest of the 32 lane maxima (rank
// counting), survivors {score >= T} (n >= k, typically n - k < 8) are trimmed by removing the minimum n - k times, and
// the k remaining candidates are emitted in candidate order (the reference's torch.topk order is by score; only the SET
// is consumed downstream, and every comparison in tests/ is on sets).
// SV ("stream V"): V is not staged.  After the soft-max every warp publishes its sibling's weights as column f of A4[KC][4];
// warp f then takes the parent candidates f, f+4, .. of ALL four siblings: its lanes = (child slot g, chunk dq) load each V row
// chunk once, straight from L2 into registers, and feed 8 FFMA2 from one LDS.128 of weights; the per-warp partial sums are
// transposed over g with a butterfly and added across the warps through a 2 KB buffer aliased onto the (dead) K slab.  Shared
// memory per item drops from 35 KB to 19 KB (kp = 32): 9 resident items (register-limited) instead of 6, and the A.V loop issues
// 8 LDS.128 per warp instead of 64 LDS.  Measured at 832^2 (B200): 44.4 -> 40.4 us per launch; with the two directions of a layer
// co-running in the step's CUDA graph 2.179 -> 2.119 ms per step (the smaller footprint also lets the other direction's CTAs in).
// SV is the default; CASMTR_MID_STREAMV=0 in the environment selects the staged variant (A/B and fallback).
template <int KP, int R, bool TYPE_A, bool DO_TOPK, bool SV>
__global__ void __launch_bounds__(128, SV ? 9 : 1) quad_cta_kernel(FineParams p) {
    pdl_sync();
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, f = threadIdx.x >> 5;       // warp = sibling f
    const int g = lane >> 3, dq = lane & 7;
    const int wp = p.w0 >> 1;
    const int Np = (p.h0 >> 1) * wp;
    const int b = blockIdx.y;
    const int parent = p.d_nh.div((int)blockIdx.x), h = blockIdx.x - parent * p.nh;
    const int py = p.d_wp.div(parent), px = parent - py * wp;
    const int C = p.nh * D, L0 = p.h0 * p.w0, L1 = p.h1 * p.w1;
    const int kp = KP ? KP : p.kp, KC = 4 * kp;
    const float scale = rsqrtf((float)D);

    float *Ks = smem;                                          // [max(KC,32)][32], chunk-swizzled
    float *Vs = Ks + (KC < 32 ? 32 : KC) * D;                  // [KC][32]                         (not SV)
    float *Qs = SV ? Vs : Vs + KC * D;                         // [4][32]
    float *As = Qs + 4 * D;                                    // [4][KC] attention weights, one row per sibling; SV: A4 [KC][4]

    const int off_g = (g >> 1) * p.w1 + (g & 1);
    const int qtok = (2 * py + (f >> 1)) * p.w0 + 2 * px + (f & 1);     // this warp's query token
    // independent of the candidate list: the q rows (warp 0) and the coarser levels' message go out first
    if (f == 0) {
        const int qt = (2 * py + (g >> 1)) * p.w0 + 2 * px + (g & 1);
        cp_async16(Qs + g * D + 4 * dq, p.q + ((size_t)b * L0 + qt) * C + h * D + 4 * dq);
    }
    float4 ap = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.acc_prev && g == 0) ap = ldg4(p.acc_prev + ((size_t)b * Np + parent) * C + h * D + 4 * dq);

    // ---- candidate bases: lane k (< kp) holds the top-left child of parent-candidate k
    int base = 0;
    float pscore = 0.f;
    if (lane < kp) {
        const size_t o = (((size_t)b * Np + parent) * p.nh + h) * kp + lane;
        const int idx = p.prev_idx[o];
        const int r = p.d_wprev.div(idx);
        base = 2 * r * p.w1 + 2 * (idx - r * p.w_prev);
        if (TYPE_A) pscore = p.prev_score[o];
    }

    // ---- gathers: warp f takes parent candidates u = f, f+4, ..
    {
        const float *kb = p.k + (size_t)b * L1 * C + h * D + 4 * dq;
        const float *vb = p.v + (size_t)b * L1 * C + h * D + 4 * dq;
        float *kd0 = Ks + g * D + 4 * (dq ^ g), *kd1 = Ks + g * D + 4 * (dq ^ (4 + g));
        float *vd = Vs + g * D + 4 * dq;
#pragma unroll
        for (int i = 0; i < (KP ? (KP + 3) / 4 : 8); ++i) {
            const int u = 4 * i + f;
            const int tok = __shfl_sync(FULL_MASK, base, u & 31) + off_g;
            if (u < kp) {
                const size_t off = (size_t)tok * C;
                cp_async16(((u & 1) ? kd1 : kd0) + u * 4 * D, kb + off);
                if (!SV) cp_async16(vd + u * 4 * D, vb + off);
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
    }
    __syncthreads();

    // ---- Q.K^T for sibling f, lane = candidate
    float sc[R];
    {
        float2 acc[R];
        const float *krow[R];
        int ksw[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            acc[r] = make_float2(0.f, 0.f);
            const int c = min(32 * r + lane, KC - 1);          // rows past KC re-read the last valid row; masked below
            krow[r] = Ks + c * D;
            ksw[r] = c & 7;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 qv = *reinterpret_cast<const float4 *>(Qs + f * D + 4 * j);
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = dot4p(qv, *reinterpret_cast<const float4 *>(krow[r] + 4 * (j ^ ksw[r])), acc[r]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) sc[r] = 32 * r + lane < KC ? (acc[r].x + acc[r].y) * scale : -INFINITY;
    }

    // ---- softmax -> a[r]; sel[r] = selection score (>= 0) or -1
    float a[R], sel[R];
    if (!TYPE_A) {
        float m = sc[0];
#pragma unroll
        for (int r = 1; r < R; ++r) m = fmaxf(m, sc[r]);
        m = warp_max(m);
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) { a[r] = exp_neg(sc[r] - m); sum += a[r]; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
#pragma unroll
        for (int r = 0; r < R; ++r) { a[r] *= inv; sel[r] = 32 * r + lane < KC ? a[r] : -1.f; }
    } else {            // QTAttA: over the 4 children of each parent candidate (4 adjacent lanes), times the parent's score (:72-77)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool valid = 32 * r + lane < KC;
            const float ps = __shfl_sync(FULL_MASK, pscore, (8 * r + (lane >> 2)) & 31);
            float m = sc[r];
            m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 1));
            m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 2));
            const float e = valid ? exp_neg(sc[r] - m) : 0.f;
            float sum = e;
            sum += __shfl_xor_sync(FULL_MASK, sum, 1);
            sum += __shfl_xor_sync(FULL_MASK, sum, 2);
            a[r] = valid ? (e / sum) * ps : 0.f;
            sel[r] = valid ? a[r] : -1.f;
        }
    }

    // ---- top-k of sibling f for the next level
    if (DO_TOPK) {
        const int k = p.topk;
        unsigned key[R];
        float lmax = -1.f;
#pragma unroll
        for (int r = 0; r < R; ++r) { key[r] = okey(sel[r]); lmax = fmaxf(lmax, sel[r]); }
        // T = k-th largest lane maximum (at least k candidates are >= T): a 15-step bitonic sort of the lane maxima instead of 32
        // rounds of rank counting (16 % of this kernel's instructions, scripts/ncu_lines.py)
        const unsigned T = okey(__shfl_sync(FULL_MASK, warp_sort_desc(lmax, lane), (k - 1) & 31));
        int n = 0;
        bool sv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            sv[r] = key[r] >= T && key[r] > 0;
            n += __popc(__ballot_sync(FULL_MASK, sv[r]));
        }
        while (n > k) {                              // drop the smallest survivor (ties: lowest lane, lowest round)
            unsigned lm = 0xffffffffu;
#pragma unroll
            for (int r = 0; r < R; ++r) lm = min(lm, sv[r] ? key[r] : 0xffffffffu);
            const unsigned gm = __reduce_min_sync(FULL_MASK, lm);
            const int owner = __ffs(__ballot_sync(FULL_MASK, lm == gm)) - 1;
            if (lane == owner) {
                bool done = false;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (!done && sv[r] && key[r] == gm) { sv[r] = false; done = true; }
            }
            --n;
        }
        // emit in candidate order
        int slot0 = 0;
        const size_t o0 = (((size_t)b * L0 + qtok) * p.nh + h) * k;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const unsigned bal = __ballot_sync(FULL_MASK, sv[r]);
            const int c = 32 * r + lane;
            const int tok = __shfl_sync(FULL_MASK, base, (c >> 2) & 31) + ((c & 3) >> 1) * p.w1 + (c & 1);
            if (sv[r]) {
                const int slot = slot0 + __popc(bal & ((1u << lane) - 1u));
                p.topk_idx[o0 + slot] = tok;
                p.topk_score[o0 + slot] = a[r];
                if (TYPE_A && !p.final_level) a[r] = 0.f;      // selected keys leave the message (:81-84)
            }
            slot0 += __popc(bal);
        }
    }

    if (SV) {
        // ---- A.V, V streamed: warp f = parent candidates f, f+4, .. for all four siblings
        constexpr int NU = KP ? (KP + 3) / 4 : 8;
        const float *vb = p.v + (size_t)b * L1 * C + h * D + 4 * dq;
        float4 vv[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) {                         // issued before the barrier: the L2 latency hides behind it
            const int u = 4 * i + f;
            const int tok = __shfl_sync(FULL_MASK, base, u & 31) + off_g;
            vv[i] = u < kp ? ldg4(vb + (size_t)tok * C) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (32 * r + lane < KC) As[(32 * r + lane) * 4 + f] = a[r];
        __syncthreads();                                       // A4 complete; every warp is done with the K slab and Q
        float2 o[4][2];
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) o[s4][0] = o[s4][1] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            const int u = 4 * i + f;
            if (u < kp) {
                const float4 aw = *reinterpret_cast<const float4 *>(As + (4 * u + g) * 4);
                const float2 vlo = make_float2(vv[i].x, vv[i].y), vhi = make_float2(vv[i].z, vv[i].w);
                o[0][0] = __ffma2_rn(make_float2(aw.x, aw.x), vlo, o[0][0]); o[0][1] = __ffma2_rn(make_float2(aw.x, aw.x), vhi, o[0][1]);
                o[1][0] = __ffma2_rn(make_float2(aw.y, aw.y), vlo, o[1][0]); o[1][1] = __ffma2_rn(make_float2(aw.y, aw.y), vhi, o[1][1]);
                o[2][0] = __ffma2_rn(make_float2(aw.z, aw.z), vlo, o[2][0]); o[2][1] = __ffma2_rn(make_float2(aw.z, aw.z), vhi, o[2][1]);
                o[3][0] = __ffma2_rn(make_float2(aw.w, aw.w), vlo, o[3][0]); o[3][1] = __ffma2_rn(make_float2(aw.w, aw.w), vhi, o[3][1]);
            }
        }
        // reduce over g (lane bits 3,4), transposing: the lane ends with sibling g's chunk dq (this warp's candidates only)
        float ov[4][4];
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) { ov[s4][0] = o[s4][0].x; ov[s4][1] = o[s4][0].y; ov[s4][2] = o[s4][1].x; ov[s4][3] = o[s4][1].y; }
        float r2[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float recv = __shfl_xor_sync(FULL_MASK, (g & 2) ? ov[i][c] : ov[i + 2][c], 16);
                r2[i][c] = ((g & 2) ? ov[i + 2][c] : ov[i][c]) + recv;
            }
        float4 m4;
        {
            float t[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float recv = __shfl_xor_sync(FULL_MASK, (g & 1) ? r2[0][c] : r2[1][c], 8);
                t[c] = ((g & 1) ? r2[1][c] : r2[0][c]) + recv;
            }
            m4 = make_float4(t[0], t[1], t[2], t[3]);
        }
        float *red = Ks;                                       // [4 warps][4 siblings][32], the K slab is dead
        *reinterpret_cast<float4 *>(red + (f * 4 + g) * D + 4 * dq) = m4;
        __syncthreads();
        if (g == 0) {                                          // warp f finishes sibling f: sum of the four warps' partials, in warp order
            float4 res = *reinterpret_cast<const float4 *>(red + (0 * 4 + f) * D + 4 * dq);
#pragma unroll
            for (int w = 1; w < 4; ++w) {
                const float4 t = *reinterpret_cast<const float4 *>(red + (w * 4 + f) * D + 4 * dq);
                res.x += t.x; res.y += t.y; res.z += t.z; res.w += t.w;
            }
            const float wl = p.wsm ? __ldg(p.wsm + p.level) : 1.f;
            res.x *= wl; res.y *= wl; res.z *= wl; res.w *= wl;
            res.x += ap.x; res.y += ap.y; res.z += ap.z; res.w += ap.w;
            *reinterpret_cast<float4 *>(p.out + ((size_t)b * L0 + qtok) * C + h * D + 4 * dq) = res;
        }
        return;
    }
    // ---- A.V for sibling f: lane = (child slot g, chunk dq)
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (32 * r + lane < KC) As[f * KC + 32 * r + lane] = a[r];
    __syncwarp();
    float2 o0 = make_float2(0.f, 0.f), o1 = o0;
    {
        const float *arow = As + f * KC + g;
        const float *vrow = Vs + g * D + 4 * dq;
#pragma unroll
        for (int u = 0; u < (KP ? KP : 32); ++u) {
            if (KP == 0 && u >= kp) break;
            const float4 vv = *reinterpret_cast<const float4 *>(vrow + u * 4 * D);
            const float aw = arow[4 * u];
            o0 = __ffma2_rn(make_float2(aw, aw), make_float2(vv.x, vv.y), o0);
            o1 = __ffma2_rn(make_float2(aw, aw), make_float2(vv.z, vv.w), o1);
        }
    }
    float4 res = make_float4(o0.x, o0.y, o1.x, o1.y);
#pragma unroll
    for (int m = 8; m <= 16; m <<= 1) {
        res.x += __shfl_xor_sync(FULL_MASK, res.x, m); res.y += __shfl_xor_sync(FULL_MASK, res.y, m);
        res.z += __shfl_xor_sync(FULL_MASK, res.z, m); res.w += __shfl_xor_sync(FULL_MASK, res.w, m);
    }
    // ---- merge with the coarser levels and write raster (:262-284)
    if (g == 0) {
        const float wl = p.wsm ? __ldg(p.wsm + p.level) : 1.f;
        res.x *= wl; res.y *= wl; res.z *= wl; res.w *= wl;
        res.x += ap.x; res.y += ap.y; res.z += ap.z; res.w += ap.w;
        *reinterpret_cast<float4 *>(p.out + ((size_t)b * L0 + qtok) * C + h * D + 4 * dq) = res;
    }
}

__host__ __device__ inline int cta_slab_floats(int kp) { return staged_slab_floats(kp) + 16 * kp; }
__host__ __device__ inline int cta_sv_slab_floats(int kp) { return warp_slab_floats(kp) + 16 * kp; }     // K + Q + A4

template <int KP, int R, bool TYPE_A, bool DO_TOPK, bool SV>
int launch_cta_t(const FineParams &p, cudaStream_t stream) {
    const long long items = (long long)(p.h0 / 2) * (p.w0 / 2) * p.nh;       // per batch element
    if (items == 0 || p.B == 0) return CASMTR_OK;
    CASMTR_REQUIRE(items <= 0x7fffffffLL && p.B <= 65535, CASMTR_E_UNSUPPORTED, "quad attention grid too large");
    const size_t smem = sizeof(float) * (SV ? cta_sv_slab_floats(p.kp) : cta_slab_floats(p.kp));
    auto kern = quad_cta_kernel<KP, R, TYPE_A, DO_TOPK, SV>;
    static PerDeviceOnce once;                                                // one flag per template instance and device
    const int dev = PerDeviceOnce::device();
    if (!once.done(dev)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { casmtr_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return CASMTR_E_CUDA; }
        once.mark(dev);
    }
    LaunchScope ls(DO_TOPK ? CASMTR_K_QT_FINE_MID : CASMTR_K_QT_FINE_LAST, stream);
    launch_k(kern, dim3((unsigned)items, p.B), 128, smem, stream, p);
    CASMTR_CHECK_LAUNCH("quad_cta_kernel");
    return CASMTR_OK;
}

// grid.x covers the (parent, head) items of one batch element, grid.y = batch
// (QTAtt levels: at most 85 registers = the 24 warps per SM that the 8.5 KB slabs allow at kp = 16; 84.4 -> 83.4 us per launch)
template <int KP, int R, bool CASCADE, bool TYPE_A, bool DO_TOPK, bool PE = false>
__global__ void __launch_bounds__(256, (CASCADE || R > 2) ? 1 : 3) quad_attention_kernel(FineParams p, int warps_per_cta) {
    pdl_sync();
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Np = (p.h0 >> 1) * (p.w0 >> 1);
    const unsigned item = blockIdx.x * (unsigned)warps_per_cta + warp;
    if (item >= (unsigned)(Np * p.nh)) return;
    const int parent = p.d_nh.div((int)item), h = item - parent * p.nh;
    quad_item<KP, R, CASCADE, TYPE_A, DO_TOPK, PE>(p, smem + (size_t)warp * warp_slab_floats(KP ? KP : p.kp), blockIdx.y, parent, h, lane);
}

// persistent variant over a device-side list of cells (b * Np + parent), all heads of each: the cascade fallback
// (two 8-warp CTAs per SM: 128 registers; one more register halves the resident warps of this latency-bound kernel, 41 -> 49 us)
template <int KP, int R, bool PE = false>
__global__ void __launch_bounds__(256, 2) quad_attention_list_kernel(FineParams p, int warps_per_cta) {
    pdl_sync();
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Np = (p.h0 >> 1) * (p.w0 >> 1);
    const int n = *p.item_count * p.nh;
    float *slab = smem + (size_t)warp * warp_slab_floats(KP ? KP : p.kp);
    for (int i = blockIdx.x * warps_per_cta + warp; i < n; i += gridDim.x * warps_per_cta) {
        int ci, h, cb, cp;
        p.d_nh.divmod(i, ci, h);
        p.d_np.divmod(p.item_list[ci], cb, cp);
        quad_item<KP, R, true, false, false, PE>(p, slab, cb, cp, h, lane);
        __syncwarp();
    }
}

template <int KP, int R, bool CASCADE, bool TYPE_A, bool DO_TOPK, bool PE = false>
int launch_t(const FineParams &p, cudaStream_t stream) {
    const long long items = (long long)(p.h0 / 2) * (p.w0 / 2) * p.nh;       // per batch element
    if (items == 0 || p.B == 0) return CASMTR_OK;
    const size_t per_warp = sizeof(float) * warp_slab_floats(p.kp);
    int wpc = (int)((113 * 1024) / per_warp);
    // warps (= items) per CTA: small CTAs.  Same resident warps per SM (shared memory and registers are per warp here), but a
    // 17 KB CTA fits into the gaps other kernels leave and the tail is finer.  Measured at 832^2 (CASMTR_WPC overrides):
    // 8 -> 47.9 us, 4 -> 45.5 us, 2 -> 45.4 us per launch; whole-step graph 2.120 -> 2.092 -> 2.084 ms.
    static const int wpc_max = [] { const char *e = getenv("CASMTR_WPC"); const int v = e ? atoi(e) : 2; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
    wpc = wpc < 1 ? 1 : (wpc > wpc_max ? wpc_max : wpc);
    const size_t smem = per_warp * wpc;
    const long long blocks = (items + wpc - 1) / wpc;
    CASMTR_REQUIRE(blocks <= 0x7fffffffLL && p.B <= 65535, CASMTR_E_UNSUPPORTED, "quad attention grid too large");
    auto kern = quad_attention_kernel<KP, R, CASCADE, TYPE_A, DO_TOPK, PE>;
    static PerDeviceOnce once;                                                // one flag per template instance and device
    const int dev = PerDeviceOnce::device();
    if (!once.done(dev)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { casmtr_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return CASMTR_E_CUDA; }
        once.mark(dev);
    }
    LaunchScope ls(CASCADE ? CASMTR_K_CASCADE_ATT : (DO_TOPK ? CASMTR_K_QT_FINE_MID : CASMTR_K_QT_FINE_LAST), stream);
    launch_k(kern, dim3((unsigned)blocks, p.B), wpc * 32, smem, stream, p, wpc);
    CASMTR_CHECK_LAUNCH("quad_attention_kernel");
    return CASMTR_OK;
}

template <int KP, int R, bool PE>
int launch_list_t(const FineParams &p, cudaStream_t stream) {
    const size_t per_warp = sizeof(float) * warp_slab_floats(p.kp);
    int wpc = (int)((113 * 1024) / per_warp);
    wpc = wpc < 1 ? 1 : (wpc > 8 ? 8 : wpc);
    const size_t smem = per_warp * wpc;
    auto kern = quad_attention_list_kernel<KP, R, PE>;
    static PerDeviceOnce once;                                                // one flag per template instance and device
    const int dev = PerDeviceOnce::device();
    if (!once.done(dev)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { casmtr_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return CASMTR_E_CUDA; }
        once.mark(dev);
    }
    LaunchScope ls(CASMTR_K_CASCADE_FALLBACK, stream);
    launch_k(kern, 2 * 148, wpc * 32, smem, stream, p, wpc);             // persistent: the list length is only known on the device
    CASMTR_CHECK_LAUNCH("quad_attention_list_kernel");
    return CASMTR_OK;
}

template <int KP, int R>
int launch_by_flags(const FineParams &p, cudaStream_t stream) {
    if (p.topk_pos || p.next_idx)                    // cascade (gather path): warp per item
        return p.pe.w_tab ? launch_t<KP, R, true, false, false, true>(p, stream) : launch_t<KP, R, true, false, false>(p, stream);
    // QTAtt fine levels.  Intermediate levels (top-k, few items, long per-item chain): CTA per item, 4 warps = 4 siblings.
    // Last level (4x the items, no top-k): warp per item -- the CTA variant re-reads the K slab once per sibling warp and
    // becomes shared-memory-pipe bound there (ncu: l1tex 85 %), measured 69 us vs 64 us at 832^2.
    const bool topk = p.topk_idx != nullptr;
    static const bool sv = [] { const char *e = getenv("CASMTR_MID_STREAMV"); return !(e && e[0] == '0'); }();
    if (p.type_a) return topk ? (sv ? launch_cta_t<KP, R, true, true, true>(p, stream) : launch_cta_t<KP, R, true, true, false>(p, stream))
                              : launch_t<KP, R, false, true, false>(p, stream);
    return topk ? (sv ? launch_cta_t<KP, R, false, true, true>(p, stream) : launch_cta_t<KP, R, false, true, false>(p, stream))
                : launch_t<KP, R, false, false, false>(p, stream);
}

}  // namespace

int launch_quad_attention(const FineParams &p_in, cudaStream_t stream) {
    CASMTR_REQUIRE(p_in.kp >= 1 && p_in.kp <= 32, CASMTR_E_UNSUPPORTED, "parent candidate count %d must be in [1,32]", p_in.kp);
    CASMTR_REQUIRE((p_in.h0 % 2) == 0 && (p_in.w0 % 2) == 0, CASMTR_E_INVALID, "query grid %dx%d must be even", p_in.h0, p_in.w0);
    CASMTR_REQUIRE(p_in.nh >= 1 && p_in.h0 >= 2 && p_in.w0 >= 2 && p_in.h1 >= 2 && p_in.w1 >= 2, CASMTR_E_INVALID, "quad attention: empty grid");
    FineParams p = p_in;                             // + the multiply-shift forms of the divisors the kernels use
    p.d_nh = make_fastdiv(p.nh);
    p.d_wp = make_fastdiv(p.w0 / 2);
    p.d_np = make_fastdiv((p.h0 / 2) * (p.w0 / 2));
    p.d_wprev = make_fastdiv(p.w_prev > 0 ? p.w_prev : 1);
    p.d_wv = make_fastdiv(p.w1 / 2);
    p.d_win = make_fastdiv(p.win > 0 ? p.win : 1);
    if (p.item_list) {                               // cascade fallback cells of the tile kernel
        CASMTR_REQUIRE((p.topk_pos || p.next_idx) && p.kp == 25, CASMTR_E_INVALID, "item lists are a cascade (k = 25) feature");
        return p.pe.w_tab ? launch_list_t<25, 4, true>(p, stream) : launch_list_t<25, 4, false>(p, stream);
    }
    if (p.topk_idx) CASMTR_REQUIRE(p.topk >= 1 && p.topk <= 32 && p.topk <= 4 * p.kp, CASMTR_E_INVALID, "top-k %d must be in [1, min(32, %d)]", p.topk, 4 * p.kp);
    // the shipped configurations get fully unrolled gather loops; everything else runs the runtime-kp variant
    if (p.kp == 32) return launch_by_flags<32, 4>(p, stream);
    if (p.kp == 16) return launch_by_flags<16, 2>(p, stream);
    if (p.kp == 25) return launch_by_flags<25, 4>(p, stream);
    if (p.kp <= 8) return launch_by_flags<0, 1>(p, stream);
    if (p.kp <= 16) return launch_by_flags<0, 2>(p, stream);
    if (p.kp <= 24) return launch_by_flags<0, 3>(p, stream);
    return launch_by_flags<0, 4>(p, stream);
}
