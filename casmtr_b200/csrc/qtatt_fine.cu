// Fused quadtree fine level / cascade window attention ("quad attention").
//
// One warp = one (batch, parent token, head): the 4 sibling queries of the parent attend to the
// 4*kp candidate keys (the 4 children of each of the parent's kp selected coarser keys).  The
// kernel fuses what the reference does with ~40 torch ops and 2 custom kernels per level:
// candidate expansion, gathered Q.K^T, softmax, top-k for the next level, gathered A.V, the
// child-major -> raster reorder and the (weighted) merge with the coarser levels' message.
// No index / QK / A tensor ever reaches HBM.
// Reference: QTAttB.process_fine_level + merge
//   cuda_imp/QuadTreeAttention/QuadtreeAttention/modules/quadtree_attention.py:180-229, 262-284
// QTAttA.process_fine_level :46-99, merge :130-138;  CascadeQTAttB.forward :400-452;
// kernels replaced: src/score_computation_kernal.cu:22-62, src/value_aggregation_kernel.cu:21-42.
//
// Data movement: a (token, head) K or V row is one 128-byte line; 8 lanes read one row with
// LDG.128, so a warp-wide load touches exactly 4 lines (no sector waste).  Every loaded K/V
// element feeds 4 FMAs (one per sibling query) straight from registers; partial dot products are
// combined with a transposing butterfly (7 shuffles per 8 candidates), after which each lane
// owns one (query, candidate) logit per step -- the layout softmax and top-k work in.
//
// lane = g*8 + dq:  g in 0..3 (child slot of the candidate being loaded / query row in the
// epilogue), dq in 0..7 (which float4 of the 32-dim head row).  After the butterfly lane holds
// logit(query fq = dq&3, parent-candidate 2t + (dq>>2), child g) for step t.
#include "common.cuh"
#include "kernels.cuh"

namespace {

constexpr int D = 32;
constexpr int WARPS = 8;

__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

template <int T, bool CASCADE, bool TYPE_A, bool DO_TOPK>
__global__ void __launch_bounds__(WARPS * 32, 2) quad_attention_kernel(FineParams p) {
    __shared__ __align__(16) float Asm[WARPS][8 * T * 4];
    __shared__ int stg_idx[DO_TOPK ? WARPS : 1][4 * 32];
    __shared__ float stg_sc[DO_TOPK ? WARPS : 1][4 * 32];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, dq = lane & 7, fq = dq & 3, jj = dq >> 2;
    const int wp = p.w0 >> 1;
    const int Np = (p.h0 >> 1) * wp;
    const long long item = (long long)blockIdx.x * WARPS + warp;
    if (item >= (long long)p.B * Np * p.nh) return;
    const int h = (int)(item % p.nh);
    const int parent = (int)((item / p.nh) % Np);
    const int b = (int)(item / ((long long)p.nh * Np));
    const int py = parent / wp, px = parent - py * wp;
    const int C = p.nh * D, L0 = p.h0 * p.w0, L1 = p.h1 * p.w1;
    const int kp = p.kp, KC = 4 * kp;
    const float scale = rsqrtf((float)D);

    // ---- candidate bases: lane k (< kp) holds the top-left child of parent-candidate k
    int base = 0;
    float pscore = 0.f;
    if (lane < kp) {
        if (CASCADE) {
            const int64_t *tp = p.topk_pos + (((size_t)b * Np + parent) * kp + lane) * 2;
            base = (int)(2 * tp[0] * p.w1 + 2 * tp[1]);
        } else {
            const size_t o = (((size_t)b * Np + parent) * p.nh + h) * kp + lane;
            const int idx = p.prev_idx[o];
            const int r = idx / p.w_prev;
            base = 2 * r * p.w1 + 2 * (idx - r * p.w_prev);
            if (TYPE_A) pscore = p.prev_score[o];
        }
    }
    const int off_g = (g >> 1) * p.dil * p.w1 + (g & 1) * p.dil;
    int qtok[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) qtok[f] = (2 * py + (f >> 1)) * p.w0 + 2 * px + (f & 1);

    float4 q[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) q[f] = ldg4(p.q + ((size_t)b * L0 + qtok[f]) * C + h * D + 4 * dq);

    const float *kb = p.k + (size_t)b * L1 * C + h * D + 4 * dq;
    const float *vb = p.v + (size_t)b * L1 * C + h * D + 4 * dq;

    // ---- gathered Q.K^T
    float sc[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int ca = min(max(__shfl_sync(FULL_MASK, base, (2 * t) & 31) + off_g, 0), L1 - 1);
        const int cb = min(max(__shfl_sync(FULL_MASK, base, (2 * t + 1) & 31) + off_g, 0), L1 - 1);
        float4 ka = make_float4(0.f, 0.f, 0.f, 0.f), kbv = ka;
        if (2 * t < kp) ka = ldg4(kb + (size_t)ca * C);
        if (2 * t + 1 < kp) kbv = ldg4(kb + (size_t)cb * C);
        float v[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            const float pa = dot4(q[f], ka), pb = dot4(q[f], kbv);
            const float recv = __shfl_xor_sync(FULL_MASK, jj ? pa : pb, 4);
            v[f] = (jj ? pb : pa) + recv;
        }
        float w[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float recv = __shfl_xor_sync(FULL_MASK, (dq & 2) ? v[i] : v[i + 2], 2);
            w[i] = ((dq & 2) ? v[i + 2] : v[i]) + recv;
        }
        const float recv = __shfl_xor_sync(FULL_MASK, (dq & 1) ? w[0] : w[1], 1);
        float s = (((dq & 1) ? w[1] : w[0]) + recv) * scale;
        const bool valid = 2 * t + jj < kp;
        if (CASCADE && p.rel_pos != nullptr && valid)
            s += __ldg(p.rel_pos + (((size_t)b * p.nh + h) * L0 + qtok[fq]) * KC + 4 * (2 * t + jj) + g);
        sc[t] = valid ? s : -INFINITY;
    }

    // ---- softmax
    float a[T];
    if (!TYPE_A) {      // over all 4*kp candidates of query fq (lanes differing in bits 2,3,4)
        float m = sc[0];
#pragma unroll
        for (int t = 1; t < T; ++t) m = fmaxf(m, sc[t]);
        m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 4));
        m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 8));
        m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 16));
        float sum = 0.f;
#pragma unroll
        for (int t = 0; t < T; ++t) { a[t] = exp_neg(sc[t] - m); sum += a[t]; }
        sum += __shfl_xor_sync(FULL_MASK, sum, 4);
        sum += __shfl_xor_sync(FULL_MASK, sum, 8);
        sum += __shfl_xor_sync(FULL_MASK, sum, 16);
#pragma unroll
        for (int t = 0; t < T; ++t) a[t] = a[t] / sum;
    } else {            // QTAttA: over the 4 children of each parent candidate, times the parent's score (:72-77)
#pragma unroll
        for (int t = 0; t < T; ++t) {
            float m = sc[t];
            m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 8));
            m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, 16));
            const bool valid = 2 * t + jj < kp;
            float e = valid ? exp_neg(sc[t] - m) : 0.f;
            float sum = e;
            sum += __shfl_xor_sync(FULL_MASK, sum, 8);
            sum += __shfl_xor_sync(FULL_MASK, sum, 16);
            const float ps = __shfl_sync(FULL_MASK, pscore, (2 * t + jj) & 31);
            a[t] = valid ? (e / sum) * ps : 0.f;
            sc[t] = valid ? a[t] : -INFINITY;       // type A selects on the redistributed score
        }
    }

    // ---- top-k for the next level: the 4 queries run in parallel in their own lane groups
    if (DO_TOPK) {
        const unsigned gmask = 0x11111111u << fq;
        for (int it = 0; it < p.topk; ++it) {
            float lm = sc[0];
#pragma unroll
            for (int t = 1; t < T; ++t) lm = fmaxf(lm, sc[t]);
            float gm = fmaxf(lm, __shfl_xor_sync(FULL_MASK, lm, 4));
            gm = fmaxf(gm, __shfl_xor_sync(FULL_MASK, gm, 8));
            gm = fmaxf(gm, __shfl_xor_sync(FULL_MASK, gm, 16));
            const unsigned bal = __ballot_sync(FULL_MASK, lm == gm) & gmask;
            const int owner = __ffs(bal) - 1;
            int ts = -1;
            float av = 0.f;
#pragma unroll
            for (int t = 0; t < T; ++t)
                if (ts < 0 && sc[t] == gm) { ts = t; av = a[t]; }
            if (lane == owner) {
#pragma unroll
                for (int t = 0; t < T; ++t)
                    if (t == ts) {
                        sc[t] = -INFINITY;
                        if (TYPE_A && !p.final_level) a[t] = 0.f;   // selected keys leave the message (:81-84)
                    }
            }
            const int slot = __shfl_sync(FULL_MASK, 4 * (2 * ts + jj) + g, owner & 31);
            const float aval = __shfl_sync(FULL_MASK, av, owner & 31);
            const int cf = slot & 3;
            const int cand = min(max(__shfl_sync(FULL_MASK, base, (slot >> 2) & 31) + (cf >> 1) * p.dil * p.w1 + (cf & 1) * p.dil, 0), L1 - 1);
            if (lane == fq) { stg_idx[warp][fq * 32 + it] = cand; stg_sc[warp][fq * 32 + it] = aval; }
        }
        __syncwarp();
        for (int i = lane; i < 4 * p.topk; i += 32) {
            const int f = i / p.topk, kk = i - f * p.topk;
            const size_t o = (((size_t)b * L0 + qtok[f]) * p.nh + h) * p.topk + kk;
            p.topk_idx[o] = stg_idx[warp][f * 32 + kk];
            p.topk_score[o] = stg_sc[warp][f * 32 + kk];
        }
    }

    // ---- A -> smem as [slot][query] so the A.V loop reads the 4 sibling weights with one LDS.128
#pragma unroll
    for (int t = 0; t < T; ++t) Asm[warp][(4 * (2 * t + jj) + g) * 4 + fq] = a[t];
    __syncwarp();

    // ---- gathered A.V: lane accumulates all 4 queries x its 4 dims over the candidates of child slot g
    float o[4][4];
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int c = 0; c < 4; ++c) o[f][c] = 0.f;
#pragma unroll
    for (int u = 0; u < 2 * T; ++u) {
        const int cand = min(max(__shfl_sync(FULL_MASK, base, u & 31) + off_g, 0), L1 - 1);
        if (u < kp) {
            const float4 vv = ldg4(vb + (size_t)cand * C);
            const float4 aw = *reinterpret_cast<const float4 *>(&Asm[warp][(4 * u + g) * 4]);
            o[0][0] = fmaf(aw.x, vv.x, o[0][0]); o[0][1] = fmaf(aw.x, vv.y, o[0][1]); o[0][2] = fmaf(aw.x, vv.z, o[0][2]); o[0][3] = fmaf(aw.x, vv.w, o[0][3]);
            o[1][0] = fmaf(aw.y, vv.x, o[1][0]); o[1][1] = fmaf(aw.y, vv.y, o[1][1]); o[1][2] = fmaf(aw.y, vv.z, o[1][2]); o[1][3] = fmaf(aw.y, vv.w, o[1][3]);
            o[2][0] = fmaf(aw.z, vv.x, o[2][0]); o[2][1] = fmaf(aw.z, vv.y, o[2][1]); o[2][2] = fmaf(aw.z, vv.z, o[2][2]); o[2][3] = fmaf(aw.z, vv.w, o[2][3]);
            o[3][0] = fmaf(aw.w, vv.x, o[3][0]); o[3][1] = fmaf(aw.w, vv.y, o[3][1]); o[3][2] = fmaf(aw.w, vv.z, o[3][2]); o[3][3] = fmaf(aw.w, vv.w, o[3][3]);
        }
    }
    // reduce over g (lane bits 3,4), transposing: lane ends with query f = g
    float r2[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float recv = __shfl_xor_sync(FULL_MASK, (g & 2) ? o[i][c] : o[i + 2][c], 16);
            r2[i][c] = ((g & 2) ? o[i + 2][c] : o[i][c]) + recv;
        }
    float m4[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float recv = __shfl_xor_sync(FULL_MASK, (g & 1) ? r2[0][c] : r2[1][c], 8);
        m4[c] = ((g & 1) ? r2[1][c] : r2[0][c]) + recv;
    }

    // ---- merge with the coarser levels and write raster (:262-284)
    float wl = 1.f;
    if (p.level_weight) {
        float mx = -INFINITY, den = 0.f;
        for (int l = 0; l < p.levels; ++l) mx = fmaxf(mx, __ldg(p.level_weight + l));
        for (int l = 0; l < p.levels; ++l) den += expf(__ldg(p.level_weight + l) - mx);
        wl = expf(__ldg(p.level_weight + p.level) - mx) / den;
    }
    float4 res = make_float4(m4[0] * wl, m4[1] * wl, m4[2] * wl, m4[3] * wl);
    if (p.acc_prev) {
        const float4 ap = ldg4(p.acc_prev + ((size_t)b * Np + parent) * C + h * D + 4 * dq);
        res.x = ap.x + res.x; res.y = ap.y + res.y; res.z = ap.z + res.z; res.w = ap.w + res.w;
    }
    *reinterpret_cast<float4 *>(p.out + ((size_t)b * L0 + qtok[g]) * C + h * D + 4 * dq) = res;

    // ---- cascade: the window's key indices for the 4 children (the reference's upsampled_idx, :450)
    if (CASCADE && p.upsampled_idx != nullptr && h == 0) {
        for (int s0 = 0; s0 < KC; s0 += 32) {
            const int s = s0 + lane;
            const int cf = s & 3;
            const int cand = min(max(__shfl_sync(FULL_MASK, base, (s >> 2) & 31) + (cf >> 1) * p.dil * p.w1 + (cf & 1) * p.dil, 0), L1 - 1);
            if (s < KC) {
#pragma unroll
                for (int f = 0; f < 4; ++f) p.upsampled_idx[((size_t)b * L0 + qtok[f]) * KC + s] = cand;
            }
        }
    }
}

template <int T, bool CASCADE, bool TYPE_A, bool DO_TOPK>
int launch_t(const FineParams &p, cudaStream_t stream) {
    const long long items = (long long)p.B * (p.h0 / 2) * (p.w0 / 2) * p.nh;
    const long long blocks = (items + WARPS - 1) / WARPS;
    CASMTR_REQUIRE(blocks <= 0x7fffffffLL, CASMTR_E_UNSUPPORTED, "quad attention grid too large");
    if (blocks == 0) return CASMTR_OK;
    LaunchScope ls(CASCADE ? CASMTR_K_CASCADE_ATT : (DO_TOPK ? CASMTR_K_QT_FINE_MID : CASMTR_K_QT_FINE_LAST), stream);
    quad_attention_kernel<T, CASCADE, TYPE_A, DO_TOPK><<<(unsigned)blocks, WARPS * 32, 0, stream>>>(p);
    CASMTR_CHECK_LAUNCH("quad_attention_kernel");
    return CASMTR_OK;
}

template <int T>
int launch_by_flags(const FineParams &p, cudaStream_t stream) {
    if (p.topk_pos) return launch_t<T, true, false, false>(p, stream);
    const bool topk = p.topk_idx != nullptr;
    if (p.type_a) return topk ? launch_t<T, false, true, true>(p, stream) : launch_t<T, false, true, false>(p, stream);
    return topk ? launch_t<T, false, false, true>(p, stream) : launch_t<T, false, false, false>(p, stream);
}

}  // namespace

int launch_quad_attention(const FineParams &p, cudaStream_t stream) {
    CASMTR_REQUIRE(p.kp >= 1 && p.kp <= 32, CASMTR_E_UNSUPPORTED, "parent candidate count %d must be in [1,32]", p.kp);
    CASMTR_REQUIRE((p.h0 % 2) == 0 && (p.w0 % 2) == 0, CASMTR_E_INVALID, "query grid %dx%d must be even", p.h0, p.w0);
    if (p.topk_idx) CASMTR_REQUIRE(p.topk >= 1 && p.topk <= 32 && p.topk <= 4 * p.kp, CASMTR_E_INVALID, "top-k %d must be in [1, min(32, %d)]", p.topk, 4 * p.kp);
    if (p.kp <= 8) return launch_by_flags<4>(p, stream);
    if (p.kp <= 16) return launch_by_flags<8>(p, stream);
    if (p.kp <= 26) return launch_by_flags<13>(p, stream);
    return launch_by_flags<16>(p, stream);
}
