// CascadeQTAttB with TMA-staged key/value window tiles.
// Reference: CascadeQTAttB.forward
//   cuda_imp/QuadTreeAttention/QuadtreeAttention/modules/quadtree_attention.py:400-452
// window construction: CascadeFeatureTransformer.get_window_warp_idx  src/model/modules/transformer.py:416-440
//
// In the cascade every query cell (the 2x2 siblings of a parent token) attends to a 5x5-parent = 10x10-token
// window of the other image around the previous stage's match, the same window for all heads.  Neighbouring
// cells have neighbouring windows wherever the match field is coherent, so gathering 100 rows per cell and head
// (quad_attention_kernel: 1.1 GB of L2->SM traffic per call at 1/4 of 832^2) reads the same lines over and over.
// Here one CTA owns a 4x4 block of parent cells of one head and loads ONE 18x20-token K tile and V tile with two
// TMA tensor copies (cp.async.bulk.tensor.4d, 128-byte swizzle, mbarrier completion):
//   * tile origin = median window position of the block (robust to outlier matches), 1 parent of slack;
//   * a cell whose window lies inside the tile and is a regular 5x5 window is computed from shared memory
//     (lane = candidate Q.K^T with conflict-free swizzled reads, warp softmax, A.V, raster store, upsampled_idx);
//   * any other cell (outlier match, irregular / dilated window) is appended to a fallback list that the per-item
//     gather kernel (qtatt_fine.cu) processes afterwards -- same arithmetic, identical results.
// L2->SM traffic drops from 25.6 KB to 5.8 KB per (cell, head) and the 50 LDGSTS + address computations per
// item disappear; the kernel is bound by instruction issue (packed FFMA2 for all dot products).
#include <cuda.h>

#include "common.cuh"
#include "kernels.cuh"

namespace {

constexpr int D = 32;
constexpr int TP = 4;                   // parent cells per tile edge
constexpr int TH = 18, TW = 20;         // tile extent in key tokens (9 x 10 parents); TW % 8 == 4 keeps lane=candidate reads conflict-free
constexpr int TILE_BYTES = TH * TW * D * 4;
constexpr int NWARP = 8;
constexpr int KC = 100;

__device__ __forceinline__ float2 dot4p(const float4 q, const float4 k, float2 acc) {
    acc = __ffma2_rn(make_float2(q.x, q.y), make_float2(k.x, k.y), acc);
    return __ffma2_rn(make_float2(q.z, q.w), make_float2(k.z, k.w), acc);
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(smem_u32(dst)),
        "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}

struct TileParams {
    const float *q;             // token-major [B, h0*w0, C]
    const int64_t *topk_pos;    // [B, Np, 25, 2]
    const float *rel_pos;       // [B, nh, h0*w0, 100] or NULL
    float *out;                 // [B, h0*w0, C]
    int64_t *upsampled_idx;     // [B, h0*w0, 100] or NULL
    int *fb_list, *fb_count;    // fallback parents (b * Np + parent)
    int B, nh, h0, w0, h1, w1;
};

__global__ void __launch_bounds__(NWARP * 32, 2)
cascade_att_tile_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, TileParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);       // 128B-swizzled TMA tiles want 1024-byte alignment
    float *Kt = (float *)sm;                                   // [TH*TW][32], 16-byte chunk c of row r at chunk c ^ (r & 7)
    float *Vt = (float *)(sm + TILE_BYTES);
    float *Aw = (float *)(sm + 2 * TILE_BYTES);                // [NWARP][100][4]
    float *Qw = Aw + NWARP * KC * 4;                           // [NWARP][4][32]
    uint64_t *bar = (uint64_t *)(Qw + NWARP * 4 * D);
    int *s_org = (int *)(bar + 1);                             // tile origin (row, col) in parent units

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 3, dq = lane & 7;
    const int hp = p.h0 >> 1, wp = p.w0 >> 1, Np = hp * wp;
    const int tiles_x = (wp + TP - 1) / TP;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int h = blockIdx.y, b = blockIdx.z;
    const int C = p.nh * D, L0 = p.h0 * p.w0, L1 = p.h1 * p.w1;

    if (tid == 0) mbar_init(bar, 1);
    // ---- tile origin: median over the block's cells of (window row - local row, window col - local col)
    if (warp == 0) {
        const int ly = (lane >> 2) & 3, lx = lane & 3;
        const int py = ty * TP + ly, px = tx * TP + lx;
        const bool have = lane < TP * TP && py < hp && px < wp;
        int vr = 0x3fffffff, vc = 0x3fffffff;                  // absent cells sort last
        if (have) {
            const int64_t *tp = p.topk_pos + ((size_t)b * Np + (size_t)py * wp + px) * 50;
            vr = (int)tp[0] - ly;
            vc = (int)tp[1] - lx;
        }
        const int n = __popc(__ballot_sync(FULL_MASK, have));
        int rr = 0, rc = 0;
#pragma unroll
        for (int l = 0; l < TP * TP; ++l) {
            const int orr = __shfl_sync(FULL_MASK, vr, l), oc = __shfl_sync(FULL_MASK, vc, l);
            rr += (orr < vr) || (orr == vr && l < lane);
            rc += (oc < vc) || (oc == vc && l < lane);
        }
        const int mid = (n - 1) >> 1;
        const int src_r = __ffs(__ballot_sync(FULL_MASK, lane < TP * TP && rr == mid)) - 1;
        const int src_c = __ffs(__ballot_sync(FULL_MASK, lane < TP * TP && rc == mid)) - 1;
        const int org_r = __shfl_sync(FULL_MASK, vr, src_r);           // rows: cells at local row ly sit at org_r + ly (+1 tolerated)
        const int org_c = __shfl_sync(FULL_MASK, vc, src_c) - 1;       // cols: one parent of slack on both sides
        if (lane == 0) {
            s_org[0] = org_r;
            s_org[1] = org_c;
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            mbar_expect_tx(bar, 2 * TILE_BYTES);
            tma_load_4d(Kt, &tmK, h * D, 2 * org_c, 2 * org_r, b, bar);
            tma_load_4d(Vt, &tmV, h * D, 2 * org_c, 2 * org_r, b, bar);
        }
    }
    __syncthreads();                                           // barrier initialised, origin visible
    const int org_r = s_org[0], org_c = s_org[1];
    bool tile_ready = false;

    float *As = Aw + warp * KC * 4;
    float *Qs = Qw + warp * 4 * D;
    const float scale = rsqrtf((float)D);

#pragma unroll 1
    for (int it = 0; it < (TP * TP) / NWARP; ++it) {
        const int cell = warp + it * NWARP;
        const int ly = cell >> 2, lx = cell & 3;
        const int py = ty * TP + ly, px = tx * TP + lx;
        if (py >= hp || px >= wp) continue;                    // warp-uniform
        const int parent = py * wp + px;
        // the cell's window: 25 (row, col) pairs; regular = row-major 5x5 block starting at entry 0
        const int64_t *tp = p.topk_pos + ((size_t)b * Np + parent) * 50;
        const int r0 = (int)__ldg(tp), c0 = (int)__ldg(tp + 1);
        bool ok = true;
        if (lane < 25) {
            const int wr = (int)__ldg(tp + 2 * lane), wc = (int)__ldg(tp + 2 * lane + 1);
            ok = wr == r0 + lane / 5 && wc == c0 + lane % 5;
        }
        const int wy0 = r0 - org_r, wx0 = c0 - org_c;          // window offset inside the tile, parent units
        ok = __all_sync(FULL_MASK, ok) && wy0 >= 0 && wy0 <= TH / 2 - 5 && wx0 >= 0 && wx0 <= TW / 2 - 5;
        if (!ok) {                                             // outlier / irregular window: the gather kernel takes all heads of this cell
            if (h == 0 && lane == 0) p.fb_list[atomicAdd(p.fb_count, 1)] = b * Np + parent;
            continue;
        }
        const int qtok0 = 2 * py * p.w0 + 2 * px;
#define QTOK(f) (qtok0 + ((f) >> 1) * p.w0 + ((f) & 1))
        // sibling q rows -> smem (broadcast operand of Q.K^T)
        __syncwarp();
        *reinterpret_cast<float4 *>(Qs + g * D + 4 * dq) = ldg4(p.q + ((size_t)b * L0 + QTOK(g)) * C + h * D + 4 * dq);
        if (!tile_ready) { mbar_wait(bar, 0); tile_ready = true; }
        __syncwarp();

        // ---- Q.K^T, lane = candidate c = 32r + lane = 4*(5*wy + wx) + f
        float sc[4][4];
        {
            float2 acc[4][4];
            const float *krow[4];
            int ksw[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int f = 0; f < 4; ++f) acc[r][f] = make_float2(0.f, 0.f);
                const int c = min(32 * r + lane, KC - 1);
                const int kk = c >> 2, f = c & 3;
                const int trow = (2 * (wy0 + kk / 5) + (f >> 1)) * TW + 2 * (wx0 + kk % 5) + (f & 1);
                krow[r] = Kt + trow * D;
                ksw[r] = trow & 7;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 qv[4];
#pragma unroll
                for (int f = 0; f < 4; ++f) qv[f] = *reinterpret_cast<const float4 *>(Qs + f * D + 4 * j);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 kv = *reinterpret_cast<const float4 *>(krow[r] + 4 * (j ^ ksw[r]));
#pragma unroll
                    for (int f = 0; f < 4; ++f) acc[r][f] = dot4p(qv[f], kv, acc[r][f]);
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int f = 0; f < 4; ++f) sc[r][f] = acc[r][f].x + acc[r][f].y;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int c = 32 * r + lane;
            const bool valid = c < KC;
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                float s = sc[r][f] * scale;
                if (p.rel_pos != nullptr && valid) s += __ldg(p.rel_pos + (((size_t)b * p.nh + h) * L0 + QTOK(f)) * KC + c);
                sc[r][f] = valid ? s : -INFINITY;
            }
        }
        // ---- softmax over the 100 candidates per sibling
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            float m = fmaxf(fmaxf(sc[0][f], sc[1][f]), fmaxf(sc[2][f], sc[3][f]));
            m = warp_max(m);
            float sum = 0.f;
#pragma unroll
            for (int r = 0; r < 4; ++r) { sc[r][f] = exp_neg(sc[r][f] - m); sum += sc[r][f]; }
            sum = warp_sum(sum);
            const float inv = 1.0f / sum;
#pragma unroll
            for (int r = 0; r < 4; ++r) sc[r][f] *= inv;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (32 * r + lane < KC) *reinterpret_cast<float4 *>(As + (32 * r + lane) * 4) = make_float4(sc[r][0], sc[r][1], sc[r][2], sc[r][3]);
        __syncwarp();

        // ---- A.V: lane = (child slot g, chunk dq) accumulates the 4 siblings over the 25 parent candidates
        float2 o[4][2];
#pragma unroll
        for (int f = 0; f < 4; ++f) o[f][0] = o[f][1] = make_float2(0.f, 0.f);
        const int vbase = (2 * wy0 + (g >> 1)) * TW + 2 * wx0 + (g & 1);
#pragma unroll
        for (int u = 0; u < 25; ++u) {
            const int trow = vbase + 2 * (u / 5) * TW + 2 * (u % 5);
            const float4 vv = *reinterpret_cast<const float4 *>(Vt + trow * D + 4 * (dq ^ (trow & 7)));
            const float4 aw = *reinterpret_cast<const float4 *>(As + (4 * u + g) * 4);
            const float2 vlo = make_float2(vv.x, vv.y), vhi = make_float2(vv.z, vv.w);
            o[0][0] = __ffma2_rn(make_float2(aw.x, aw.x), vlo, o[0][0]); o[0][1] = __ffma2_rn(make_float2(aw.x, aw.x), vhi, o[0][1]);
            o[1][0] = __ffma2_rn(make_float2(aw.y, aw.y), vlo, o[1][0]); o[1][1] = __ffma2_rn(make_float2(aw.y, aw.y), vhi, o[1][1]);
            o[2][0] = __ffma2_rn(make_float2(aw.z, aw.z), vlo, o[2][0]); o[2][1] = __ffma2_rn(make_float2(aw.z, aw.z), vhi, o[2][1]);
            o[3][0] = __ffma2_rn(make_float2(aw.w, aw.w), vlo, o[3][0]); o[3][1] = __ffma2_rn(make_float2(aw.w, aw.w), vhi, o[3][1]);
        }
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
        float4 res;
        {
            float m4[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float recv = __shfl_xor_sync(FULL_MASK, (g & 1) ? r2[0][c] : r2[1][c], 8);
                m4[c] = ((g & 1) ? r2[1][c] : r2[0][c]) + recv;
            }
            res = make_float4(m4[0], m4[1], m4[2], m4[3]);
        }
        *reinterpret_cast<float4 *>(p.out + ((size_t)b * L0 + QTOK(g)) * C + h * D + 4 * dq) = res;

        // ---- upsampled_idx (:450): the window's key token indices, identical for the 4 siblings
        if (p.upsampled_idx != nullptr && h == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int c = 32 * r + lane;
                if (c < KC) {
                    const int kk = c >> 2, f = c & 3;
                    int tok = (2 * (r0 + kk / 5) + (f >> 1)) * p.w1 + 2 * (c0 + kk % 5) + (f & 1);
                    tok = min(max(tok, 0), L1 - 1);
#pragma unroll
                    for (int q = 0; q < 4; ++q) p.upsampled_idx[((size_t)b * L0 + QTOK(q)) * KC + c] = tok;
                }
            }
        }
#undef QTOK
    }
    if (!tile_ready) mbar_wait(bar, 0);                        // never leave with a TMA copy into this CTA's smem in flight
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// token-major [B][h1][w1][C] fp32 -> 4-D tensor map, box = (one head's 32 channels) x TW x TH x 1, 128-byte swizzle
int make_tile_map(CUtensorMap *tm, const float *base, int B, int h1, int w1, int C) {
    EncodeTiledFn enc = encode_tiled();
    CASMTR_REQUIRE(enc != nullptr, CASMTR_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)w1, (cuuint64_t)h1, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)w1 * C * 4, (cuuint64_t)h1 * w1 * C * 4};
    const cuuint32_t box[4] = {D, TW, TH, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CASMTR_REQUIRE(r == CUDA_SUCCESS, CASMTR_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CASMTR_OK;
}

}  // namespace

size_t cascade_tile_smem_bytes() { return 1024 + 2 * TILE_BYTES + sizeof(float) * (NWARP * KC * 4 + NWARP * 4 * D) + 64; }

// Tile path of CascadeQTAttB for k == 25, dilated == 1.  q/k/v are the token-major copies.  Cells that cannot use their
// block's tile are appended to fb_list (count in *fb_count, zeroed here); the caller runs the gather kernel over that list.
int launch_cascade_att_tile(const float *q, const float *k, const float *v, const int64_t *topk_pos, const float *rel_pos,
                            float *out, int64_t *upsampled_idx, int *fb_list, int *fb_count,
                            int B, int nh, int h0, int w0, int h1, int w1, cudaStream_t stream) {
    const int C = nh * D;
    CUtensorMap tmK, tmV;
    int rc = make_tile_map(&tmK, k, B, h1, w1, C);
    if (rc != CASMTR_OK) return rc;
    rc = make_tile_map(&tmV, v, B, h1, w1, C);
    if (rc != CASMTR_OK) return rc;
    TileParams p;
    p.q = q; p.topk_pos = topk_pos; p.rel_pos = rel_pos; p.out = out; p.upsampled_idx = upsampled_idx;
    p.fb_list = fb_list; p.fb_count = fb_count;
    p.B = B; p.nh = nh; p.h0 = h0; p.w0 = w0; p.h1 = h1; p.w1 = w1;
    const size_t smem = cascade_tile_smem_bytes();
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(cascade_att_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(cascade_att_tile_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { casmtr_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return CASMTR_E_CUDA; }
        attr_set = true;
    }
    if (cudaMemsetAsync(fb_count, 0, sizeof(int), stream) != cudaSuccess) { casmtr_set_error("cudaMemsetAsync failed"); return CASMTR_E_CUDA; }
    const int hp = h0 / 2, wp = w0 / 2;
    const unsigned tiles = (unsigned)(((hp + TP - 1) / TP) * ((wp + TP - 1) / TP));
    CASMTR_REQUIRE(nh <= 65535 && B <= 65535, CASMTR_E_UNSUPPORTED, "cascade tile grid too large");
    LaunchScope ls(CASMTR_K_CASCADE_ATT, stream);
    cascade_att_tile_kernel<<<dim3(tiles, nh, B), NWARP * 32, smem, stream>>>(tmK, tmV, p);
    CASMTR_CHECK_LAUNCH("cascade_att_tile_kernel");
    return CASMTR_OK;
}
