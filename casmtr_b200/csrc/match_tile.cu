// CascadeMatching correlation + softmax + argmax with TMA-staged key window tiles (both directions, one launch).
// Reference: CascadeMatching.forward, inference branch   src/model/functions/cascade_matching.py:87-149
// (fast_score_computation: cuda_imp/score_cuda/src/score_computation_kernel.cu:23-40).
//
// The candidate lists handed to CascadeMatching are CascadeQTAttB's upsampled_idx: for each 2x2 query cell the
// 10x10-token window around the previous stage's match, repeated for the 4 siblings.  The API still carries them as
// arbitrary int64 lists, so the kernel VERIFIES that structure per cell (it has to read the lists anyway: they are half
// of the compulsory HBM traffic) and only then uses it:
//   * a 4x4 block of query cells shares one 18x20-token key tile (median window origin, like cascade_tile.cu);
//     the C feature channels are streamed through a 4-stage shared-memory ring in 32-channel slices
//     (TMA cp.async.bulk.tensor.4d, 128-byte swizzle, full/empty mbarriers, warp-specialised producer);
//   * 16 consumer warps, one cell each: lane = candidate, partial dot products of the 4 siblings accumulate in
//     registers across the slices (packed FFMA2), then softmax over the K = 100 candidates, first arg-max,
//     confidence volume / next_conf / next_idx stores;
//   * cells whose lists are not such a window, or whose window falls outside the block's tile, go to a fallback list
//     processed by cascade_match_cell_kernel (cascade_match.cu) -- same results.
// L2->SM traffic per cell drops from 51 KB to 13.5 KB and no per-row address arithmetic is left in the loop.
#include "common.cuh"
#include "kernels.cuh"
#include "tma.cuh"

namespace {

using namespace tma;
constexpr int D = 32;                   // channels per slice
constexpr int TP = 4;                   // query cells per block edge
constexpr int TH = 18, TW = 20;         // key tile in tokens
constexpr int KC = 100;
constexpr int NCONS = TP * TP;
constexpr int NSTAGE = 4;
constexpr int KEY_BYTES = TH * TW * D * 4;
constexpr int Q_BYTES = (2 * TP) * (2 * TP) * D * 4;
constexpr int STAGE_BYTES = KEY_BYTES + Q_BYTES;          // 54272 = 53 * 1024
struct CellMeta { int flag, wy0, wx0, r0, c0; };
constexpr int SM_META = NSTAGE * STAGE_BYTES;
constexpr int SM_BAR = SM_META + NSTAGE * NCONS * (int)sizeof(CellMeta);
constexpr int SM_TOTAL = SM_BAR + 2 * NSTAGE * 8;

struct Maps { CUtensorMap key[2], qry[2]; };              // [direction]: key image / query image feature maps

__global__ void __launch_bounds__((NCONS + 1) * 32, 1) cascade_match_tile_kernel(const __grid_constant__ Maps maps, MatchParams p) {
    pdl_sync();
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // 128B-swizzled TMA tiles want 1024-byte alignment; an offset
                                                                                 // on the __shared__ pointer (not an integer round trip) keeps LDS/STS
    uint64_t *full = (uint64_t *)(sm + SM_BAR), *empty = full + NSTAGE;
    CellMeta *meta = (CellMeta *)(sm + SM_META);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NS = p.C / D;                                // channel slices per block
    const int h0 = p.L0 / p.w0, h1 = p.L1 / p.w1;
    const int tiles0 = ((h0 / 2 + TP - 1) / TP) * ((p.w0 / 2 + TP - 1) / TP);
    const int tiles1 = ((h1 / 2 + TP - 1) / TP) * ((p.w1 / 2 + TP - 1) / TP);
    const int n_work = p.B * (tiles0 + tiles1);            // direction 0 blocks of all batches, then direction 1

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NCONS); }
    }
    __syncthreads();

    int step = 0;                                          // ring position: one per (block, slice), identical in every warp
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const bool rev = w >= p.B * tiles0;
        const int wi = rev ? w - p.B * tiles0 : w;
        const int tiles = rev ? tiles1 : tiles0;
        const int b = wi / tiles, tile = wi - b * tiles;
        const int wq = rev ? p.w1 : p.w0, hq = rev ? h1 : h0;       // query grid
        const int wk = rev ? p.w0 : p.w1;                           // key grid width
        const int Lq = rev ? p.L1 : p.L0, Lk = rev ? p.L0 : p.L1;
        const int hp = hq >> 1, wp = wq >> 1;
        const int tiles_x = (wp + TP - 1) / TP;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int64_t *idx = rev ? p.idx10 : p.idx01;

        if (warp == NCONS) {
            // ================= producer: block geometry from the first candidate of every cell, then the slice copies =================
            const int ly = (lane >> 2) & 3, lx = lane & 3;
            const int py = ty * TP + ly, px = tx * TP + lx;
            const bool have = py < hp && px < wp;
            int r0 = 0, c0 = 0;
            bool even = false;
            if (have) {
                const long long t0 = __ldg(idx + ((size_t)b * Lq + (size_t)(2 * py) * wq + 2 * px) * KC);
                const int y = (int)(t0 / wk), x = (int)(t0 - (long long)y * wk);
                r0 = y >> 1; c0 = x >> 1;
                // the whole 5x5-parent window must lie inside the key grid, or linear indices would wrap rows
                even = t0 >= 0 && t0 < Lk && !(y & 1) && !(x & 1) && 2 * (r0 + 5) <= Lk / wk && 2 * (c0 + 5) <= wk;
            }
            const bool cand = have && even;
            const int vr = cand ? r0 - ly : 0x3fffffff, vc = cand ? c0 - lx : 0x3fffffff;
            const int n = __popc(__ballot_sync(FULL_MASK, cand) & 0xffffu);
            int rr = 0, rc = 0;
#pragma unroll
            for (int l = 0; l < NCONS; ++l) {
                const int orr = __shfl_sync(FULL_MASK, vr, l), oc = __shfl_sync(FULL_MASK, vc, l);
                rr += (orr < vr) || (orr == vr && l < (lane & 15));
                rc += (oc < vc) || (oc == vc && l < (lane & 15));
            }
            const int mid = n > 0 ? (n - 1) >> 1 : 0;
            const int src_r = __ffs(__ballot_sync(FULL_MASK, lane < NCONS && rr == mid)) - 1;
            const int src_c = __ffs(__ballot_sync(FULL_MASK, lane < NCONS && rc == mid)) - 1;
            int org_r = __shfl_sync(FULL_MASK, vr, src_r), org_c = __shfl_sync(FULL_MASK, vc, src_c) - 1;
            if (n == 0) { org_r = 0; org_c = 0; }
            const int wy0 = r0 - org_r, wx0 = c0 - org_c;
            const bool inside = cand && wy0 >= 0 && wy0 <= TH / 2 - 5 && wx0 >= 0 && wx0 <= TW / 2 - 5;
            for (int sl = 0; sl < NS; ++sl, ++step) {
                const int s = step % NSTAGE;
                mbar_wait(empty + s, ((step / NSTAGE) & 1) ^ 1);
                if (sl == 0 && lane < NCONS) {
                    CellMeta m;
                    m.flag = have ? (inside ? 1 : 2) : 0;          // 1 = tile candidate, 2 = present but not tileable, 0 = outside the grid
                    m.wy0 = wy0; m.wx0 = wx0; m.r0 = r0; m.c0 = c0;
                    meta[s * NCONS + lane] = m;
                }
                __syncwarp();
                if (lane == 0) {
                    uint8_t *st = sm + s * STAGE_BYTES;
                    mbar_expect_tx(full + s, STAGE_BYTES);
                    tma_load_4d(st, &maps.key[rev], sl * D, 2 * org_c, 2 * org_r, b, full + s);
                    tma_load_4d(st + KEY_BYTES, &maps.qry[rev], sl * D, 2 * TP * tx, 2 * TP * ty, b, full + s);
                }
            }
            continue;
        }

        // ================= consumers: one query cell each =================
        const int ly = warp >> 2, lx = warp & 3;
        const int py = ty * TP + ly, px = tx * TP + lx;
        const size_t row00 = (size_t)b * Lq + (size_t)(2 * py) * wq + 2 * px;
#define ROWQ(f) (row00 + (size_t)((f) >> 1) * wq + ((f) & 1))
        // candidate lists of the 4 siblings: lane j (< 25) loads entries 4j..4j+3 (two 16-byte loads) of every sibling
        // and checks them against the regular 10x10 window anchored at entry 0 (= token (2*r0, 2*c0)): entry
        // c = 4*(5*wy + wx) + f  ->  token (2*(r0+wy) + (f>>1)) * wk + 2*(c0+wx) + (f&1)
        const bool present = py < hp && px < wp;
        bool lists_ok = present;
        {
            longlong2 lv[4][2];
            if (present && lane < 25) {
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    const longlong2 *row = reinterpret_cast<const longlong2 *>(idx + ROWQ(f) * KC);
                    lv[f][0] = __ldg(row + 2 * lane);
                    lv[f][1] = __ldg(row + 2 * lane + 1);
                }
            }
            const long long t0 = __shfl_sync(FULL_MASK, lv[0][0].x, 0);
            if (present && lane < 25) {
                const int wy = lane / 5, wx = lane - 5 * wy;           // lane j holds the 4 children of parent candidate j
                const long long t00 = t0 + (long long)(2 * wy) * wk + 2 * wx;
#pragma unroll
                for (int f = 0; f < 4; ++f)
                    lists_ok = lists_ok && lv[f][0].x == t00 && lv[f][0].y == t00 + 1 && lv[f][1].x == t00 + wk && lv[f][1].y == t00 + wk + 1;
            }
            lists_ok = __all_sync(FULL_MASK, lists_ok);
        }
        float2 acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[r][f] = make_float2(0.f, 0.f);
        bool use = false;
        CellMeta m;
        unsigned koff[4];                               // byte offset of the row's swizzled chunk 0 inside a key tile
        for (int sl = 0; sl < NS; ++sl, ++step) {
            const int s = step % NSTAGE;
            mbar_wait(full + s, (step / NSTAGE) & 1);
            if (sl == 0) {
                m = meta[s * NCONS + warp];
                use = m.flag == 1 && lists_ok;         // the producer derived (r0, c0) from the same entry 0
                if (m.flag != 0 && !use && lane == 0)                     // present but not servable from the tile
                    p.fb_list[atomicAdd(p.fb_count, 1)] = (int)((rev ? (size_t)p.B * (p.L0 >> 2) : 0) + (size_t)b * (Lq >> 2) + (size_t)py * wp + px);
                if (use) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int c = min(32 * r + lane, KC - 1);
                        const int kk = c >> 2, f = c & 3;
                        const int trow = (2 * (m.wy0 + kk / 5) + (f >> 1)) * TW + 2 * (m.wx0 + kk % 5) + (f & 1);
                        koff[r] = (unsigned)trow * (D * 4) + ((unsigned)(trow & 7) << 4);
                    }
                }
            }
            if (use) {
                const unsigned kt = smem_u32(sm + s * STAGE_BYTES);
                unsigned kaddr[4];                      // chunk j of the row is at kaddr ^ (j << 4) (tma.cuh, lds128)
#pragma unroll
                for (int r = 0; r < 4; ++r) kaddr[r] = kt + koff[r];
                const float *Qs = (const float *)(sm + s * STAGE_BYTES + KEY_BYTES) + ((2 * ly) * (2 * TP) + 2 * lx) * D;
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
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
        if (!use) continue;

        // ---- softmax over the K candidates, max / first arg-max, per sibling (cascade_matching.py:120-129)
        float *conf = rev ? p.conf10 : p.conf01;
        float *next_conf = rev ? p.next_conf10 : p.next_conf01;
        int64_t *next_idx = rev ? p.next_idx10 : p.next_idx01;
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            float sc[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) sc[r] = 32 * r + lane < KC ? (acc[r][f].x + acc[r][f].y) * p.inv_scale : -INFINITY;
            float mx = warp_max(fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3])));
            float sum = 0.f;
            int arg = 0x7fffffff;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (sc[r] == mx) arg = min(arg, 32 * r + lane);
                sc[r] = exp_neg(sc[r] - mx);
                sum += sc[r];
            }
            sum = warp_sum(sum);
            arg = __reduce_min_sync(FULL_MASK, arg);
            if (arg == 0x7fffffff) arg = 0;         // NaN scores: keep next_idx a valid token of the window
            const size_t row = ROWQ(f);
            if (conf) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (32 * r + lane < KC) conf[row * KC + 32 * r + lane] = sc[r] / sum;
            }
            if (lane == 0) {
                const int kk = arg >> 2, cf = arg & 3;
                next_conf[row] = 1.0f / sum;            // exp(0) / sum
                next_idx[row] = (long long)(2 * (m.r0 + kk / 5) + (cf >> 1)) * wk + 2 * (m.c0 + kk % 5) + (cf & 1);
            }
        }
#undef ROWQ
    }
}

}  // namespace

size_t match_tile_smem_bytes() { return 1024 + SM_TOTAL; }

bool match_tile_applicable(const MatchParams &p) {
    return p.K == KC && p.C % D == 0 && p.C >= D && p.C <= 512 && !p.mask0 && !p.mask1 && p.w0 > 0 && p.w1 > 0 &&
           p.w0 % 2 == 0 && p.w1 % 2 == 0 && p.L0 % p.w0 == 0 && p.L1 % p.w1 == 0 && (p.L0 / p.w0) % 2 == 0 && (p.L1 / p.w1) % 2 == 0 &&
           p.fb_list != nullptr && (((uintptr_t)p.idx01 | (uintptr_t)p.idx10 | (uintptr_t)p.feat0 | (uintptr_t)p.feat1) & 15) == 0 &&
           (long long)p.B * ((p.L0 >> 2) + (p.L1 >> 2)) < 0x7fffffffLL;
}

int launch_cascade_match_tile(const MatchParams &p, cudaStream_t stream) {
    Maps maps;
    const int h0 = p.L0 / p.w0, h1 = p.L1 / p.w1;
    // direction 0: queries = image 0, keys = image 1; direction 1 the other way round
    int rc = make_tile_map(&maps.key[0], p.feat1, p.B, h1, p.w1, p.C, TW, TH, true);
    if (rc == CASMTR_OK) rc = make_tile_map(&maps.key[1], p.feat0, p.B, h0, p.w0, p.C, TW, TH, true);
    if (rc == CASMTR_OK) rc = make_tile_map(&maps.qry[0], p.feat0, p.B, h0, p.w0, p.C, 2 * TP, 2 * TP, false);
    if (rc == CASMTR_OK) rc = make_tile_map(&maps.qry[1], p.feat1, p.B, h1, p.w1, p.C, 2 * TP, 2 * TP, false);
    if (rc != CASMTR_OK) return rc;
    const size_t smem = match_tile_smem_bytes();
    static PerDeviceOnce once;
    const int dev = PerDeviceOnce::device();
    if (!once.done(dev)) {
        cudaError_t e = cudaFuncSetAttribute(cascade_match_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(cascade_match_tile_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { casmtr_set_error("cascade match tile kernel setup: %s", cudaGetErrorString(e)); return CASMTR_E_CUDA; }
        once.mark(dev);
    }
    const int n_sm = casmtr_sm_count();
    if (cudaMemsetAsync(p.fb_count, 0, sizeof(int), stream) != cudaSuccess) { casmtr_set_error("cudaMemsetAsync failed"); return CASMTR_E_CUDA; }
    const int tiles0 = ((h0 / 2 + TP - 1) / TP) * ((p.w0 / 2 + TP - 1) / TP), tiles1 = ((h1 / 2 + TP - 1) / TP) * ((p.w1 / 2 + TP - 1) / TP);
    const long long n_work = (long long)p.B * (tiles0 + tiles1);
    const unsigned grid = (unsigned)(n_work < n_sm ? n_work : n_sm);
    LaunchScope ls(CASMTR_K_CASCADE_MATCH, stream);
    launch_k(cascade_match_tile_kernel, grid, (NCONS + 1) * 32, smem, stream, maps, p);
    CASMTR_CHECK_LAUNCH("cascade_match_tile_kernel");
    return CASMTR_OK;
}
