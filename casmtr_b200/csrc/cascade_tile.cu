// CascadeQTAttB with TMA-staged key/value window tiles.
// Reference: CascadeQTAttB.forward
//   cuda_imp/QuadTreeAttention/QuadtreeAttention/modules/quadtree_attention.py:400-452
// window construction: CascadeFeatureTransformer.get_window_warp_idx  src/model/modules/transformer.py:416-440
//
// In the cascade every query cell (the 2x2 siblings of a parent token) attends to a 5x5-parent = 10x10-token
// window of the other image around the previous stage's match, the same window for all heads.  Neighbouring
// cells have neighbouring windows wherever the match field is coherent, so gathering 100 rows per cell and head
// (quad_attention_kernel: 1.1 GB of L2->SM traffic per call at 1/4 of 832^2) reads the same lines over and over.
// Here a 4x4 block of parent cells of one head shares ONE 18x20-token K tile and V tile (plus the block's 8x8 query
// tokens), brought in by TMA tensor copies (cp.async.bulk.tensor.4d, 128-byte swizzle).  The kernel is persistent
// (one CTA per SM) and warp-specialised: a producer warp works out the block's window geometry and issues the TMA
// copies into a 2-stage shared-memory ring (full/empty mbarriers), 16 consumer warps compute one cell each:
//   * tile origin = median window position of the block (robust to outlier matches), 1 parent of slack;
//   * a cell whose window lies inside the tile and is a regular 5x5 window is computed from shared memory
//     (lane = candidate Q.K^T with conflict-free swizzled reads, warp softmax, A.V, raster store, upsampled_idx);
//   * any other cell (outlier match, irregular / dilated window) is appended to a fallback list that the per-item
//     gather kernel (qtatt_fine.cu) processes afterwards -- same arithmetic, identical results.
// L2->SM traffic drops from 25.6 KB to 5.8 KB per (cell, head) and the 50 LDGSTS + address computations per
// item disappear; the kernel is bound by instruction issue (packed FFMA2 for all dot products).
#include "common.cuh"
#include "kernels.cuh"
#include "tma.cuh"

namespace {

using namespace tma;
constexpr int D = 32;
constexpr int TP = 4;                   // parent cells per tile edge
constexpr int TH = 18, TW = 20;         // tile extent in key tokens (9 x 10 parents); TW % 8 == 4 keeps lane=candidate reads conflict-free
constexpr int TILE_BYTES = TH * TW * D * 4;
constexpr int KC = 100;

struct TileParams {
    const int64_t *topk_pos;    // [B, Np, 25, 2]
    const int64_t *next_idx;    // [B, Np] instead of topk_pos: the 5x5 window is derived from the parent's match
    const float *rel_pos;       // [B, nh, h0*w0, 100] or NULL
    RelPE pe;                   // or the bias computed from its embedding tables (pe.w_tab != NULL)
    float *out;                 // [B, h0*w0, C]
    int64_t *upsampled_idx;     // [B, h0*w0, 100] or NULL
    int *fb_list, *fb_count;    // fallback cells (b * Np + parent)
    int B, nh, h0, w0, h1, w1;
    int tiles_x, tiles_y;
};

// shared memory of the (single, persistent) CTA of an SM: 2 pipeline stages of {K tile, V tile, Q tile, cell metadata},
// per-consumer-warp attention weights, and the full/empty mbarriers of the ring
struct CellMeta { int flag, wy0, wx0, r0, c0; };   // flag: 1 = compute from the tile, 0 = absent / handed to the fallback list
constexpr int NCONS = TP * TP;                       // consumer warps = cells per tile
constexpr int Q_BYTES = (2 * TP) * (2 * TP) * D * 4; // the block's 8x8 query tokens of one head
constexpr int STAGE_BYTES = 2 * TILE_BYTES + Q_BYTES;
struct WorkMeta { int h, b, ty, tx; };              // the work item of a stage, decoded once by the producer
constexpr int SM_META = 2 * STAGE_BYTES;                              // CellMeta[2][NCONS]
constexpr int SM_WORK = SM_META + 2 * NCONS * (int)sizeof(CellMeta);  // WorkMeta[2] (inside the 64 bytes of slack below)
constexpr int SM_A = SM_META + 2 * NCONS * (int)sizeof(CellMeta) + 64; // float[NCONS][100][4], 16-byte aligned below
constexpr int SM_BAR = ((SM_A + 15) / 16) * 16 + NCONS * KC * 4 * 4;  // uint64 full[2], empty[2]
constexpr int SM_TOTAL = SM_BAR + 4 * 8;

template <bool PE>                                         // PE: relative position bias from its embedding tables (p.pe)
__global__ void __launch_bounds__((NCONS + 1) * 32, 1)     // 17 warps: one SM sub-partition holds 5 of them, which caps the kernel at 96 registers
cascade_att_tile_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                        const __grid_constant__ CUtensorMap tmQ, TileParams p) {
    pdl_sync();
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // 128B-swizzled TMA tiles want 1024-byte alignment; an offset
                                                                                 // on the __shared__ pointer (not an integer round trip) keeps LDS/STS
    uint64_t *full = (uint64_t *)(sm + SM_BAR), *empty = full + 2;
    CellMeta *meta = (CellMeta *)(sm + SM_META);
    WorkMeta *work = (WorkMeta *)(sm + SM_WORK);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hp = p.h0 >> 1, wp = p.w0 >> 1, Np = hp * wp;
    const int C = p.nh * D, L0 = p.h0 * p.w0, L1 = p.h1 * p.w1;
    const int n_tiles = p.tiles_x * p.tiles_y;
    const int n_work = n_tiles * p.nh * p.B;                   // work item = (b, tile, head), head fastest

    if (tid == 0) {
        mbar_init(full + 0, 1); mbar_init(full + 1, 1);
        mbar_init(empty + 0, NCONS); mbar_init(empty + 1, NCONS);
    }
    __syncthreads();

    if (warp == NCONS) {
        // ================= producer warp: window geometry of the block, TMA of its K / V / Q tiles =================
        int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
            const int s = it & 1;
            const int h = w % p.nh, tile = (w / p.nh) % n_tiles, b = w / (p.nh * n_tiles);
            const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
            // cell (ly, lx) of the block <-> lane (lanes 16..31 mirror 0..15)
            const int ly = (lane >> 2) & 3, lx = lane & 3;
            const int py = ty * TP + ly, px = tx * TP + lx;
            const bool have = py < hp && px < wp;
            const int64_t *tp = p.topk_pos + ((size_t)b * Np + (size_t)(have ? py * wp + px : 0)) * 50;
            int r0 = 0, c0 = 0;
            bool regular = have;
            if (have && p.next_idx != nullptr) {     // regular by construction: 8 bytes per cell instead of 400
                const int hv = p.h1 >> 1, wv = p.w1 >> 1;
                const int idx = (int)__ldg(p.next_idx + (size_t)b * Np + py * wp + px);
                r0 = window_origin(idx / wv, 5, hv);
                c0 = window_origin(idx % wv, 5, wv);
            } else if (have) {
                const longlong2 w0v = __ldg(reinterpret_cast<const longlong2 *>(tp));
                r0 = (int)w0v.x;
                c0 = (int)w0v.y;
                // lanes < 16 verify window entries 1..12, their mirrors entries 13..24: a regular window is the row-major 5x5 block
                const int k0 = lane < 16 ? 1 : 13;
                longlong2 wv[12];                       // all 12 (row, col) pairs in flight together (16-byte loads)
#pragma unroll
                for (int k = 0; k < 12; ++k) wv[k] = __ldg(reinterpret_cast<const longlong2 *>(tp) + k0 + k);
                int bad = 0;
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    const int kk = k0 + k;
                    bad |= ((int)wv[k].x ^ (r0 + kk / 5)) | ((int)wv[k].y ^ (c0 + kk % 5));
                }
                // ... that lies inside the key grid: an out-of-grid window reads clamped / row-wrapped tokens in the reference
                // (torch.clamp of the flat index, quadtree_attention.py:428) but zero-filled rows from an out-of-bounds TMA tile
                regular = bad == 0 && r0 >= 0 && c0 >= 0 && 2 * (r0 + 5) <= p.h1 && 2 * (c0 + 5) <= p.w1;
            }
            const bool mirror_ok = __shfl_xor_sync(FULL_MASK, regular, 16);
            regular = regular && mirror_ok;
            // tile origin: median over the block's cells of (window row - local row, window col - local col)
            const int vr = have ? r0 - ly : 0x3fffffff, vc = have ? c0 - lx : 0x3fffffff;      // absent cells sort last
            const int n = __popc(__ballot_sync(FULL_MASK, have) & 0xffffu);
            int rr = 0, rc = 0;
#pragma unroll
            for (int l = 0; l < NCONS; ++l) {
                const int orr = __shfl_sync(FULL_MASK, vr, l), oc = __shfl_sync(FULL_MASK, vc, l);
                rr += (orr < vr) || (orr == vr && l < (lane & 15));
                rc += (oc < vc) || (oc == vc && l < (lane & 15));
            }
            const int mid = (n - 1) >> 1;
            const int src_r = __ffs(__ballot_sync(FULL_MASK, lane < NCONS && rr == mid)) - 1;
            const int src_c = __ffs(__ballot_sync(FULL_MASK, lane < NCONS && rc == mid)) - 1;
            const int org_r = __shfl_sync(FULL_MASK, vr, src_r);           // rows: cells of local row ly sit at org_r + ly (+1 tolerated)
            const int org_c = __shfl_sync(FULL_MASK, vc, src_c) - 1;       // cols: one parent of slack on both sides
            const int wy0 = r0 - org_r, wx0 = c0 - org_c;
            const bool inside = wy0 >= 0 && wy0 <= TH / 2 - 5 && wx0 >= 0 && wx0 <= TW / 2 - 5;
            const bool use_tile = have && regular && inside;
            if (have && !use_tile && h == 0 && lane < NCONS)              // outlier / irregular window: the gather kernel takes all heads of the cell
                p.fb_list[atomicAdd(p.fb_count, 1)] = b * Np + py * wp + px;
            mbar_wait(empty + s, ((it >> 1) & 1) ^ 1);                     // the consumers have drained this stage
            if (lane < NCONS) {
                CellMeta m;
                m.flag = use_tile; m.wy0 = wy0; m.wx0 = wx0; m.r0 = r0; m.c0 = c0;
                meta[s * NCONS + lane] = m;
            }
            if (lane == 0) work[s] = WorkMeta{h, b, ty, tx};
            __syncwarp();
            if (lane == 0) {
                uint8_t *st = sm + s * STAGE_BYTES;
                mbar_expect_tx(full + s, STAGE_BYTES);
                tma_load_4d(st, &tmK, h * D, 2 * org_c, 2 * org_r, b, full + s);
                tma_load_4d(st + TILE_BYTES, &tmV, h * D, 2 * org_c, 2 * org_r, b, full + s);
                tma_load_4d(st + 2 * TILE_BYTES, &tmQ, h * D, 2 * TP * tx, 2 * TP * ty, b, full + s);
            }
        }
        return;
    }

    // ================= consumer warps: one cell of the block each =================
    const int g = lane >> 3, dq = lane & 7;
    const int ly = warp >> 2, lx = warp & 3;
    float *As = (float *)(sm + ((SM_A + 15) / 16) * 16) + warp * KC * 4;
    const float scale = rsqrtf((float)D);
    int it = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
        const int s = it & 1;
        const float *Kt = (const float *)(sm + s * STAGE_BYTES);              // [TH*TW][32], 16-byte chunk c of row r at chunk c ^ (r & 7)
        const float *Vt = (const float *)(sm + s * STAGE_BYTES + TILE_BYTES);
        const float *Qt = (const float *)(sm + s * STAGE_BYTES + 2 * TILE_BYTES);  // [8][8][32] query tokens of the block, unswizzled
        mbar_wait(full + s, (it >> 1) & 1);
        const CellMeta m = meta[s * NCONS + warp];
        if (m.flag) {
            const WorkMeta wm = work[s];
            const int h = wm.h, b = wm.b;
            const int py = wm.ty * TP + ly, px = wm.tx * TP + lx;
            const int qtok0 = 2 * py * p.w0 + 2 * px;
#define QTOK(f) (qtok0 + ((f) >> 1) * p.w0 + ((f) & 1))
            const float *Qs = Qt + ((2 * ly) * (2 * TP) + 2 * lx) * D;        // sibling f at + ((f>>1) * 8 + (f&1)) * 32
            const int wy0 = m.wy0, wx0 = m.wx0;

            // ---- Q.K^T, lane = candidate c = 32r + lane = 4*(5*wy + wx) + f
            float sc[4][4];
            {
                float2 acc[4][4];
                unsigned kaddr[4];                      // shared address of the row's swizzled chunk 0: chunk j is at kaddr ^ (j << 4)
                const unsigned kt = smem_u32(Kt);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
#pragma unroll
                    for (int f = 0; f < 4; ++f) acc[r][f] = make_float2(0.f, 0.f);
                    const int c = min(32 * r + lane, KC - 1);
                    const int kk = c >> 2, f = c & 3;
                    const int trow = (2 * (wy0 + kk / 5) + (f >> 1)) * TW + 2 * (wx0 + kk % 5) + (f & 1);
                    kaddr[r] = kt + (unsigned)trow * (D * 4) + ((unsigned)(trow & 7) << 4);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 qv[4];
#pragma unroll
                    for (int f = 0; f < 4; ++f) qv[f] = *reinterpret_cast<const float4 *>(Qs + ((f >> 1) * 2 * TP + (f & 1)) * D + 4 * j);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float4 kv = lds128(kaddr[r] ^ (unsigned)(j << 4));
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
                for (int f = 0; f < 4; ++f) sc[r][f] = valid ? sc[r][f] * scale : -INFINITY;
            }
            if (p.rel_pos != nullptr) {                       // relative position bias (indoor configuration only)
                const float *rp = p.rel_pos + ((size_t)b * p.nh + h) * L0 * KC;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int c = 32 * r + lane;
                    if (c < KC) {
#pragma unroll
                        for (int f = 0; f < 4; ++f) sc[r][f] += __ldg(rp + (size_t)QTOK(f) * KC + c);
                    }
                }
            }
            if (PE) {                                         // the same bias from its embedding tables (get_relative_pe, transformer.py:473-509)
                const int2 qt = relpe_query_term(p.pe, b, 2 * py, 2 * px);     // s is even: sibling f adds (f & 1, f >> 1)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int c = 32 * r + lane;
                    if (c < KC) {
                        const int kk = c >> 2, cf = c & 3;
                        const int ky = 2 * (m.r0 + kk / 5) + (cf >> 1), kx = 2 * (m.c0 + kk % 5) + (cf & 1);
#pragma unroll
                        for (int f = 0; f < 4; ++f) sc[r][f] += relpe_bias(p.pe, p.nh, h, qt, ky + (f >> 1), kx + (f & 1));
                    }
                }
            }
            // ---- softmax over the 100 candidates per sibling
            // (the weights stay un-normalised: the message is linear in them, 1 / sum scales the output at the end, and the row sums'
            // shuffle chains travel under the A.V loop)
            float inv_sum[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                float mx = fmaxf(fmaxf(sc[0][f], sc[1][f]), fmaxf(sc[2][f], sc[3][f]));
                mx = warp_max(mx);
                float sum = 0.f;
#pragma unroll
                for (int r = 0; r < 4; ++r) { sc[r][f] = exp_neg(sc[r][f] - mx); sum += sc[r][f]; }
                inv_sum[f] = 1.0f / warp_sum(sum);
            }
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (32 * r + lane < KC) *reinterpret_cast<float4 *>(As + (32 * r + lane) * 4) = make_float4(sc[r][0], sc[r][1], sc[r][2], sc[r][3]);
            __syncwarp();

            // ---- A.V: lane = (child slot g, chunk dq) accumulates the 4 siblings over the 25 parent candidates
            float2 o[4][2];
#pragma unroll
            for (int f = 0; f < 4; ++f) o[f][0] = o[f][1] = make_float2(0.f, 0.f);
            // V row of candidate (wy, wx), child g: tile row vbase + 2*wy*TW + 2*wx.  2*wy*TW = 40*wy is a multiple of 8, so the
            // swizzle (row & 7) depends on wx only: 5 swizzled base pointers per lane, everything else is an immediate offset
            const int vbase = (2 * wy0 + (g >> 1)) * TW + 2 * wx0 + (g & 1);
            const float *vp[5];
#pragma unroll
            for (int wx = 0; wx < 5; ++wx) vp[wx] = Vt + (vbase + 2 * wx) * D + 4 * (dq ^ ((vbase + 2 * wx) & 7));
#pragma unroll
            for (int u = 0; u < 25; ++u) {
                const float4 vv = *reinterpret_cast<const float4 *>(vp[u % 5] + 2 * (u / 5) * TW * D);
                const float4 aw = *reinterpret_cast<const float4 *>(As + (4 * u + g) * 4);
                const float2 vlo = make_float2(vv.x, vv.y), vhi = make_float2(vv.z, vv.w);
                o[0][0] = __ffma2_rn(make_float2(aw.x, aw.x), vlo, o[0][0]); o[0][1] = __ffma2_rn(make_float2(aw.x, aw.x), vhi, o[0][1]);
                o[1][0] = __ffma2_rn(make_float2(aw.y, aw.y), vlo, o[1][0]); o[1][1] = __ffma2_rn(make_float2(aw.y, aw.y), vhi, o[1][1]);
                o[2][0] = __ffma2_rn(make_float2(aw.z, aw.z), vlo, o[2][0]); o[2][1] = __ffma2_rn(make_float2(aw.z, aw.z), vhi, o[2][1]);
                o[3][0] = __ffma2_rn(make_float2(aw.w, aw.w), vlo, o[3][0]); o[3][1] = __ffma2_rn(make_float2(aw.w, aw.w), vhi, o[3][1]);
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(empty + s)) : "memory");   // this warp is done with the stage
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
            const float inv = g == 0 ? inv_sum[0] : g == 1 ? inv_sum[1] : g == 2 ? inv_sum[2] : inv_sum[3];      // the lane holds sibling g's output
            *reinterpret_cast<float4 *>(p.out + ((size_t)b * L0 + QTOK(g)) * C + h * D + 4 * dq) = make_float4(m4[0] * inv, m4[1] * inv, m4[2] * inv, m4[3] * inv);

            // ---- upsampled_idx (:450): the window's key token indices, identical for the 4 siblings
            if (p.upsampled_idx != nullptr && h == 0) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int c = 32 * r + lane;
                    if (c < KC) {
                        const int kk = c >> 2, f = c & 3;
                        int tok = (2 * (m.r0 + kk / 5) + (f >> 1)) * p.w1 + 2 * (m.c0 + kk % 5) + (f & 1);
                        tok = min(max(tok, 0), L1 - 1);
#pragma unroll
                        for (int q = 0; q < 4; ++q) p.upsampled_idx[((size_t)b * L0 + QTOK(q)) * KC + c] = tok;
                    }
                }
            }
#undef QTOK
        } else {
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(empty + s)) : "memory");
        }
    }
}

}  // namespace

size_t cascade_tile_smem_bytes() { return 1024 + SM_TOTAL; }

// Tile path of CascadeQTAttB for k == 25, dilated == 1.  q/k/v are the token-major copies.  Cells that cannot use their
// block's tile are appended to fb_list (count in *fb_count, zeroed here); the caller runs the gather kernel over that list.
int launch_cascade_att_tile(const float *q, const float *k, const float *v, const int64_t *topk_pos, const int64_t *next_idx, const float *rel_pos,
                            const RelPE &pe, float *out, int64_t *upsampled_idx, int *fb_list, int *fb_count,
                            int B, int nh, int h0, int w0, int h1, int w1, cudaStream_t stream) {
    const int C = nh * D;
    CUtensorMap tmK, tmV, tmQ;
    int rc = make_tile_map(&tmK, k, B, h1, w1, C, TW, TH, true);
    if (rc == CASMTR_OK) rc = make_tile_map(&tmV, v, B, h1, w1, C, TW, TH, true);
    if (rc == CASMTR_OK) rc = make_tile_map(&tmQ, q, B, h0, w0, C, 2 * TP, 2 * TP, false);
    if (rc != CASMTR_OK) return rc;
    TileParams p;
    p.topk_pos = topk_pos; p.next_idx = next_idx; p.rel_pos = rel_pos; p.pe = pe; p.out = out; p.upsampled_idx = upsampled_idx;
    p.fb_list = fb_list; p.fb_count = fb_count;
    p.B = B; p.nh = nh; p.h0 = h0; p.w0 = w0; p.h1 = h1; p.w1 = w1;
    const int hp = h0 / 2, wp = w0 / 2;
    p.tiles_x = (wp + TP - 1) / TP; p.tiles_y = (hp + TP - 1) / TP;
    const size_t smem = cascade_tile_smem_bytes();
    static PerDeviceOnce once;
    const int dev = PerDeviceOnce::device();
    if (!once.done(dev)) {
        cudaError_t e = cudaSuccess;
        for (auto kern : {cascade_att_tile_kernel<false>, cascade_att_tile_kernel<true>}) {
            if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        }
        if (e != cudaSuccess) { casmtr_set_error("cascade tile kernel setup: %s", cudaGetErrorString(e)); return CASMTR_E_CUDA; }
        once.mark(dev);
    }
    const int n_sm = casmtr_sm_count();
    if (cudaMemsetAsync(fb_count, 0, sizeof(int), stream) != cudaSuccess) { casmtr_set_error("cudaMemsetAsync failed"); return CASMTR_E_CUDA; }
    const long long n_work = (long long)p.tiles_x * p.tiles_y * nh * B;
    CASMTR_REQUIRE(n_work < 0x7fffffffLL, CASMTR_E_UNSUPPORTED, "cascade tile grid too large");
    const unsigned grid = (unsigned)(n_work < n_sm ? n_work : n_sm);          // persistent: one CTA per SM
    LaunchScope ls(CASMTR_K_CASCADE_ATT, stream);
    if (pe.w_tab) launch_k(cascade_att_tile_kernel<true>, grid, (NCONS + 1) * 32, smem, stream, tmK, tmV, tmQ, p);
    else launch_k(cascade_att_tile_kernel<false>, grid, (NCONS + 1) * 32, smem, stream, tmK, tmV, tmQ, p);
    CASMTR_CHECK_LAUNCH("cascade_att_tile_kernel");
    return CASMTR_OK;
}
