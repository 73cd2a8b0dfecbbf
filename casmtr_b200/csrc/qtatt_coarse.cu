// Coarsest quadtree level: dense softmax(Q K^T / sqrt(D)) over all S keys, exact row top-k,
// A.V over ALL keys (type B) or over the non-selected keys (type A).
// Reference: QTAttB.process_coarse_level
//   cuda_imp/QuadTreeAttention/QuadtreeAttention/modules/quadtree_attention.py:161-178
// and QTAttA.process_coarse_level :25-44.  The reference materialises QK and A as [B,L,S,nh]
// tensors in HBM and runs torch.topk on a strided dim; here a CTA keeps a 16 x S score slab
// in shared memory, so HBM sees only Q, K, V once (K/V re-reads hit L2) plus the outputs.
//
// One CTA = 16 query rows of one (batch, head); 128 threads.
//   phase 1  S = scale * Q K^T   (4x4 register tiles, K streamed through a 128-token smem tile)
//   phase 2  row softmax in smem
//   phase 3  exact top-k per row (one warp per row): lane-local top-2 -> threshold by bitwise
//            bisection with ballots -> compact survivors -> k rounds of warp arg-max
//   phase 4  O = A V (V streamed through the same tile buffer), split over the 4 warps, reduced
#include "common.cuh"
#include "kernels.cuh"

namespace {

constexpr int ROWS = 16;        // query rows per CTA
constexpr int TILE = 128;       // key/value tokens per smem tile
constexpr int KV_LD = 36;       // padded row length of the tile (floats): conflict-free LDS.128
constexpr int Q_LD = 36;
constexpr int LIST_CAP = 256;   // survivor list capacity per warp
constexpr int D = 32;

__device__ __forceinline__ unsigned key_of(float v) { return v >= 0.f ? __float_as_uint(v) + 1u : 0u; }

// k rounds of warp arg-max over vals[0..n) (entries < 0 are dead).  Lane `it` ends up holding
// the it-th largest (value, position-in-list).  Destroys the selected entries (sets them to -1).
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

__global__ void __launch_bounds__(128) qtatt_coarse_kernel(CoarseParams p, int s_ld) {
    extern __shared__ __align__(16) float smem[];
    float *Ss = smem;                               // [ROWS][s_ld]
    float *Qs = Ss + ROWS * s_ld;                   // [ROWS][Q_LD]
    float *KVs = Qs + ROWS * Q_LD;                  // [TILE][KV_LD]   (aliased by the AV reduction)
    float *lval = KVs + TILE * KV_LD;               // [4][LIST_CAP]
    int *lpos = (int *)(lval + 4 * LIST_CAP);       // [4][LIST_CAP]

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

    // ---- Q rows -> smem
    for (int i = tid; i < ROWS * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < p.Sq) v = ldg4(qb + (size_t)(row0 + r) * C + 4 * c);
        *reinterpret_cast<float4 *>(Qs + r * Q_LD + 4 * c) = v;
    }

    // ---- phase 1: scores
    const int tx = lane, ty = warp;                 // tokens tx+32j, rows 4ty+i
    for (int kt = 0; kt < n_tiles; ++kt) {
        __syncthreads();
        for (int i = tid; i < TILE * 8; i += 128) {
            const int t = i >> 3, c = i & 7;
            const int tok = kt * TILE + t;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tok < p.Sk) v = ldg4(kb + (size_t)tok * C + 4 * c);
            *reinterpret_cast<float4 *>(KVs + t * KV_LD + 4 * c) = v;
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
        for (int dq = 0; dq < 8; ++dq) {
            float4 qv[4], kv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4 *>(Qs + (4 * ty + i) * Q_LD + 4 * dq);
#pragma unroll
            for (int j = 0; j < 4; ++j) kv[j] = *reinterpret_cast<const float4 *>(KVs + (tx + 32 * j) * KV_LD + 4 * dq);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = fmaf(qv[i].x, kv[j].x, acc[i][j]);
                    acc[i][j] = fmaf(qv[i].y, kv[j].y, acc[i][j]);
                    acc[i][j] = fmaf(qv[i].z, kv[j].z, acc[i][j]);
                    acc[i][j] = fmaf(qv[i].w, kv[j].w, acc[i][j]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int tok = kt * TILE + tx + 32 * j;
                Ss[(4 * ty + i) * s_ld + tok] = tok < p.Sk ? acc[i][j] * scale : -INFINITY;
            }
    }
    __syncthreads();

    // ---- phase 2 + 3: softmax and top-k, warp `warp` owns rows warp, warp+4, warp+8, warp+12
    float *mylv = lval + warp * LIST_CAP;
    int *mylp = lpos + warp * LIST_CAP;
    for (int rr = 0; rr < 4; ++rr) {
        const int r = warp + 4 * rr;
        if (row0 + r >= p.Sq) continue;            // warp-uniform
        float *srow = Ss + r * s_ld;
        float m = -INFINITY;
        for (int e = lane; e < p.Sk; e += 32) m = fmaxf(m, srow[e]);
        m = warp_max(m);
        float sum = 0.f;
        for (int e = lane; e < p.Sk; e += 32) {
            const float ex = exp_neg(srow[e] - m);
            srow[e] = ex;
            sum += ex;
        }
        sum = warp_sum(sum);
        float m1 = -1.f, m2 = -1.f;                 // lane-local two largest
        for (int e = lane; e < s_pad; e += 32) {
            float a = 0.f;
            if (e < p.Sk) {
                a = srow[e] / sum;
                if (a > m1) { m2 = m1; m1 = a; } else if (a > m2) { m2 = a; }
            }
            srow[e] = a;                            // padding columns become 0 for the AV tiles
        }
        // threshold: largest T such that at least k of the 64 lane-top-2 keys are >= T
        const unsigned k1 = key_of(m1), k2 = key_of(m2);
        unsigned T = 0;
        for (int bit = 31; bit >= 0; --bit) {
            const unsigned cand = T | (1u << bit);
            const int c = __popc(__ballot_sync(FULL_MASK, k1 >= cand)) + __popc(__ballot_sync(FULL_MASK, k2 >= cand));
            if (c >= p.topk) {
                T = cand;
                if (c == p.topk) break;
            }
        }
        __syncwarp();
        // compact survivors
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
        float rv;
        int rj, ridx = 0;
        if (n <= LIST_CAP && n >= p.topk) {
            warp_select_k(mylv, n, p.topk, lane, rv, rj);
            if (lane < p.topk) ridx = mylp[rj];
        } else {                                    // massive ties: exact but slow path on the row itself
            warp_select_k(srow, p.Sk, p.topk, lane, rv, rj);
            ridx = rj;
            __syncwarp();
            if (lane < p.topk) srow[ridx] = rv;        // restore
        }
        if (lane < p.topk) {
            const size_t o = (((size_t)b * p.Sq + row0 + r) * p.nh + h) * p.topk + lane;
            p.topk_idx[o] = ridx;
            p.topk_score[o] = rv;
            if (p.type_a) srow[ridx] = 0.f;         // QTAttA: selected keys leave the message (:37-42)
        }
    }
    __syncthreads();

    // ---- phase 4: O = A V.  lane = (rg, dq): rows rg+4i, dims 4dq..4dq+3; warp w takes tokens 32w.. of each tile
    const int rg = lane >> 3, dq = lane & 7;
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) o[i][c] = 0.f;
    for (int vt = 0; vt < n_tiles; ++vt) {
        __syncthreads();
        for (int i = tid; i < TILE * 8; i += 128) {
            const int t = i >> 3, c = i & 7;
            const int tok = vt * TILE + t;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tok < p.Sk) v = ldg4(vb + (size_t)tok * C + 4 * c);
            *reinterpret_cast<float4 *>(KVs + t * KV_LD + 4 * c) = v;
        }
        __syncthreads();
#pragma unroll 2
        for (int tt = 0; tt < 8; ++tt) {
            const int t = 32 * warp + 4 * tt;
            float4 a[4], vv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4 *>(Ss + (rg + 4 * i) * s_ld + vt * TILE + t);
#pragma unroll
            for (int j = 0; j < 4; ++j) vv[j] = *reinterpret_cast<const float4 *>(KVs + (t + j) * KV_LD + 4 * dq);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                o[i][0] = fmaf(a[i].x, vv[0].x, fmaf(a[i].y, vv[1].x, fmaf(a[i].z, vv[2].x, fmaf(a[i].w, vv[3].x, o[i][0]))));
                o[i][1] = fmaf(a[i].x, vv[0].y, fmaf(a[i].y, vv[1].y, fmaf(a[i].z, vv[2].y, fmaf(a[i].w, vv[3].y, o[i][1]))));
                o[i][2] = fmaf(a[i].x, vv[0].z, fmaf(a[i].y, vv[1].z, fmaf(a[i].z, vv[2].z, fmaf(a[i].w, vv[3].z, o[i][2]))));
                o[i][3] = fmaf(a[i].x, vv[0].w, fmaf(a[i].y, vv[1].w, fmaf(a[i].z, vv[2].w, fmaf(a[i].w, vv[3].w, o[i][3]))));
            }
        }
    }
    __syncthreads();
    float *red = KVs;                               // [4][ROWS][D]
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4 *>(red + (warp * ROWS + rg + 4 * i) * D + 4 * dq) = make_float4(o[i][0], o[i][1], o[i][2], o[i][3]);
    __syncthreads();
    float w0 = 1.f;
    if (p.level_weight) {                           // softmax over the level weights (:264)
        float mx = -INFINITY, den = 0.f;
        for (int l = 0; l < p.levels; ++l) mx = fmaxf(mx, __ldg(p.level_weight + l));
        for (int l = 0; l < p.levels; ++l) den += expf(__ldg(p.level_weight + l) - mx);
        w0 = expf(__ldg(p.level_weight) - mx) / den;
    }
    {
        const int r = tid >> 3, c = tid & 7;        // 16 rows x 8 float4
        if (row0 + r < p.Sq) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const float4 t = *reinterpret_cast<const float4 *>(red + (w * ROWS + r) * D + 4 * c);
                s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
            }
            s.x *= w0; s.y *= w0; s.z *= w0; s.w *= w0;
            *reinterpret_cast<float4 *>(p.acc + ((size_t)b * p.Sq + row0 + r) * C + h * D + 4 * c) = s;
        }
    }
}

}  // namespace

static int coarse_s_ld(int Sk) { return (Sk + TILE - 1) / TILE * TILE + 4; }

size_t coarse_smem_bytes(int Sk) {
    return sizeof(float) * ((size_t)ROWS * coarse_s_ld(Sk) + ROWS * Q_LD + TILE * KV_LD + 4 * LIST_CAP) + sizeof(int) * 4 * LIST_CAP;
}

int launch_qtatt_coarse(const CoarseParams &p, cudaStream_t stream) {
    const size_t smem = coarse_smem_bytes(p.Sk);
    CASMTR_REQUIRE(smem <= 227 * 1024, CASMTR_E_UNSUPPORTED,
                   "coarsest level has %d keys; the dense level supports at most ~3300 (shared memory)", p.Sk);
    CASMTR_REQUIRE(p.topk >= 1 && p.topk <= 32 && p.topk <= p.Sk, CASMTR_E_INVALID, "coarse top-k %d must be in [1, min(32, %d)]", p.topk, p.Sk);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(qtatt_coarse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { casmtr_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return CASMTR_E_CUDA; }
        attr_set = true;
    }
    dim3 grid((p.Sq + ROWS - 1) / ROWS, p.B * p.nh);
    LaunchScope ls(CASMTR_K_QT_COARSE, stream);
    qtatt_coarse_kernel<<<grid, 128, smem, stream>>>(p, coarse_s_ld(p.Sk));
    CASMTR_CHECK_LAUNCH("qtatt_coarse_kernel");
    return CASMTR_OK;
}
