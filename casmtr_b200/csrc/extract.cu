// NMS + thresholds + border + mutual check + ORDERED compaction of the match list, no host sync.
// Reference: CascadeMatching.get_coarse_match, inference branch
//   src/model/functions/cascade_matching.py:170-261, 316-331
//   PostProcess.apply 'maxpool_nms' / None   src/model/functions/post_processing.py:41-44, 111-121
//   mask_window_border[_with_padding]        src/model/functions/cascade_functions.py:120-172
// The reference runs ~30 small torch kernels with three host syncs (mask.sum()==0, torch.where,
// the per-sample .int() loop of the padded border).  Here: K1 computes the keep flag per source
// token and per-block counts, K2 scans the block counts (single CTA), K3 writes the matches in
// torch.where order (row-major over [B, L0]) -- bit-exact integer work.
#include "common.cuh"
#include "kernels.cuh"

namespace {

constexpr int BLK = 256;

// valid (un-padded) extent per sample: h = max_x sum_y m[y][x], w = max_y sum_x m[y][x]  (cascade_functions.py:156-157)
__global__ void valid_extent_kernel(const uint8_t *__restrict__ m, int h, int w, int *__restrict__ out_hw) {
    __shared__ int best[2];
    const uint8_t *mb = m + (size_t)blockIdx.x * h * w;
    if (threadIdx.x < 2) best[threadIdx.x] = 0;
    __syncthreads();
    for (int x = threadIdx.x; x < w; x += blockDim.x) {
        int s = 0;
        for (int y = 0; y < h; ++y) s += mb[(size_t)y * w + x] != 0;
        atomicMax(&best[0], s);
    }
    for (int y = threadIdx.x; y < h; y += blockDim.x) {
        int s = 0;
        for (int x = 0; x < w; ++x) s += mb[(size_t)y * w + x] != 0;
        atomicMax(&best[1], s);
    }
    __syncthreads();
    if (threadIdx.x < 2) out_hw[2 * blockIdx.x + threadIdx.x] = best[threadIdx.x];
}

struct ExtractArgs {
    casmtr_extract_desc d;
    const float *conf;
    const int64_t *idx01, *idx10;
    const int *valid0, *valid1;     // [B,2] (h,w) or NULL
    uint8_t *mask;                  // [B*L0] keep flags (workspace or caller's mask_out)
    int *block_count;               // [nblocks]
    int *block_off;                 // [nblocks]
    int *total;                     // [2]: total kept, emitted count
};

__device__ __forceinline__ bool keep_token(const ExtractArgs &a, int b, int i) {
    const casmtr_extract_desc &d = a.d;
    const int L0 = d.h0 * d.w0;
    const int y = i / d.w0, x = i - y * d.w0;
    const float *cb = a.conf + (size_t)b * L0;
    const float c = cb[i];
    // 1. detection: maxpool NMS (first maximum of the window in row-major scan order wins) or plain threshold
    if (d.nms_window > 0) {
        const int r = d.nms_window >> 1;
        for (int dy = -r; dy <= r; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= d.h0) continue;
            for (int dx = -r; dx <= r; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= d.w0 || (dy == 0 && dx == 0)) continue;
                const float nb = cb[yy * d.w0 + xx];
                const bool before = dy < 0 || (dy == 0 && dx < 0);
                if (before ? !(c > nb) : !(c >= nb)) return false;
            }
        }
    }
    if (!(c > d.test_thr)) return false;                       // mask[conf <= thr] = False
    // 2. previous-stage confidence gates (nearest-neighbour upsampling)
    for (int s = 0; s < d.n_pre; ++s) {
        const float sy = (float)d.pre_h[s] / (float)d.h0, sx = (float)d.pre_w[s] / (float)d.w0;
        const int py = min((int)floorf(y * sy), d.pre_h[s] - 1), px = min((int)floorf(x * sx), d.pre_w[s] - 1);
        if (d.pre_conf[s][(size_t)b * d.pre_h[s] * d.pre_w[s] + py * d.pre_w[s] + px] <= d.pre_thr[s]) return false;
    }
    // 3. border removal on source and on target coordinate
    const long long j = a.idx01[(size_t)b * L0 + i];
    if (d.border_rm > 0) {
        const int bd = d.border_rm;
        const int hs0 = a.valid0 ? a.valid0[2 * b] : d.h0, ws0 = a.valid0 ? a.valid0[2 * b + 1] : d.w0;
        const int hs1 = a.valid1 ? a.valid1[2 * b] : d.h1, ws1 = a.valid1 ? a.valid1[2 * b + 1] : d.w1;
        if (y < bd || x < bd || y >= hs0 - bd || x >= ws0 - bd) return false;
        const long long ty = j / d.w1, tx = j - ty * d.w1;
        if (d.coarse_mode ? (tx < bd || tx >= ws1 - bd || ty < bd || ty >= hs1 - bd)      // mask_border: symmetric (coarse_matching.py:116-119)
                          : (tx < bd || tx > ws1 - bd || ty < bd || ty > hs1 - bd)) return false;
    }
    // 4. mutual nearest neighbour
    if (d.double_check) {
        const long long L1 = (long long)d.h1 * d.w1;
        if (j < 0 || j >= L1) return false;
        if (a.idx10[(size_t)b * L1 + j] != i) return false;
    }
    return true;
}

__global__ void __launch_bounds__(BLK) extract_mask_kernel(ExtractArgs a) {
    pdl_sync();
    __shared__ int cnt;
    const int L0 = a.d.h0 * a.d.w0;
    const size_t n = (size_t)a.d.B * L0;
    const size_t o = blockIdx.x * (size_t)BLK + threadIdx.x;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    bool keep = false;
    if (o < n) {
        keep = keep_token(a, (int)(o / L0), (int)(o % L0));
        a.mask[o] = keep;
    }
    const unsigned bal = __ballot_sync(FULL_MASK, keep);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(&cnt, __popc(bal));
    __syncthreads();
    if (threadIdx.x == 0) a.block_count[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(1024) extract_scan_kernel(ExtractArgs a, int nblocks, int32_t *count_out) {
    pdl_sync();
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nblocks ? a.block_count[i] : 0;
        int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL_MASK, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) warp_tot[warp] = s;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL_MASK, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const int excl = carry + (warp ? warp_tot[warp - 1] : 0) + s - v;
        if (i < nblocks) a.block_off[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.total[0] = carry;
        const int emitted = (carry == 0 && !a.d.coarse_mode) ? a.d.B : carry;        // "if mask.sum() == 0: mask[:, 0] = True" (:254-255); the 1/8 stage has no such fallback
        a.total[1] = emitted;
        *count_out = emitted;
    }
}

struct EmitOut {
    uint8_t *mask_out;
    int64_t *b_ids, *i_ids, *j_ids;
    float *mconf, *mk0, *mk1;
    int capacity;
};

__device__ __forceinline__ void emit(const ExtractArgs &a, const EmitOut &e, int pos, int b, int i) {
    if (pos >= e.capacity) return;
    const casmtr_extract_desc &d = a.d;
    const size_t o = (size_t)b * d.h0 * d.w0 + i;
    const long long j = a.idx01[o];
    e.b_ids[pos] = b; e.i_ids[pos] = i; e.j_ids[pos] = j;
    e.mconf[pos] = a.conf[o];
    // (x, y) * scale [* scale0[b]]  -- int64 -> fp32, then fp32 multiplies as torch does (:317-321)
    float s0x = d.scale, s0y = d.scale, s1x = d.scale, s1y = d.scale;
    if (d.scale0) { s0x = d.scale * d.scale0[2 * b]; s0y = d.scale * d.scale0[2 * b + 1]; }
    if (d.scale1) { s1x = d.scale * d.scale1[2 * b]; s1y = d.scale * d.scale1[2 * b + 1]; }
    e.mk0[2 * pos] = (float)(i % d.w0) * s0x; e.mk0[2 * pos + 1] = (float)(i / d.w0) * s0y;
    e.mk1[2 * pos] = (float)(j % d.w1) * s1x; e.mk1[2 * pos + 1] = (float)(j / d.w1) * s1y;
}

__global__ void __launch_bounds__(BLK) extract_emit_kernel(ExtractArgs a, EmitOut e) {
    pdl_sync();
    __shared__ int warp_cnt[BLK / 32];
    const int L0 = a.d.h0 * a.d.w0;
    const size_t n = (size_t)a.d.B * L0;
    const size_t o = blockIdx.x * (size_t)BLK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (a.total[0] == 0 && a.d.coarse_mode) {                  // CoarseMatching: an empty list stays empty
        if (o < n && e.mask_out) e.mask_out[o] = 0;
        return;
    }
    if (a.total[0] == 0) {                                     // fallback: element 0 of every sample (:254-255)
        if (o < n && e.mask_out) e.mask_out[o] = 0;            // mask_out reports the flags BEFORE the fallback
        if (o < n && (o % L0) == 0) {
            const int b = (int)(o / L0);
            emit(a, e, b, b, 0);
        }
        return;
    }
    const bool keep = o < n ? a.mask[o] != 0 : false;
    if (o < n && e.mask_out) e.mask_out[o] = keep;
    const unsigned bal = __ballot_sync(FULL_MASK, keep);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int pre = a.block_off[blockIdx.x];
    for (int w = 0; w < warp; ++w) pre += warp_cnt[w];
    if (keep) emit(a, e, pre + __popc(bal & ((1u << lane) - 1u)), (int)(o / L0), (int)(o % L0));
}

}  // namespace

// ---- packed match records for the multi-GPU all-gather: row 0 = count, rows 1..M = 44-byte records
__global__ void pack_matches_kernel(const int64_t *__restrict__ b_ids, const int64_t *__restrict__ i_ids, const int64_t *__restrict__ j_ids,
                                    const float *__restrict__ mconf, const float *__restrict__ mk0, const float *__restrict__ mk1,
                                    int M, long long pair_offset, unsigned char *__restrict__ out, const int32_t *__restrict__ count_dev) {
    pdl_sync();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (count_dev != nullptr) M = min(M, (int)__ldg(count_dev));      // M is then the capacity, the count is read on the device
    if (r > M) return;
    unsigned *w = reinterpret_cast<unsigned *>(out + (size_t)r * 44);          // 44-byte rows are 4-byte aligned
    if (r == 0) {
        w[0] = (unsigned)M; w[1] = 0;
#pragma unroll
        for (int k = 2; k < 11; ++k) w[k] = 0;
        return;
    }
    const int m = r - 1;
    const long long b = b_ids[m] + pair_offset, i = i_ids[m], j = j_ids[m];
    w[0] = (unsigned)b; w[1] = (unsigned)(b >> 32);
    w[2] = (unsigned)i; w[3] = (unsigned)(i >> 32);
    w[4] = (unsigned)j; w[5] = (unsigned)(j >> 32);
    w[6] = __float_as_uint(mconf[m]);
    w[7] = __float_as_uint(mk0[2 * m]); w[8] = __float_as_uint(mk0[2 * m + 1]);
    w[9] = __float_as_uint(mk1[2 * m]); w[10] = __float_as_uint(mk1[2 * m + 1]);
}

int launch_pack_matches(const int64_t *b_ids, const int64_t *i_ids, const int64_t *j_ids, const float *mconf, const float *mk0,
                        const float *mk1, int M, long long pair_offset, unsigned char *out, const int32_t *count_dev, cudaStream_t stream) {
    LaunchScope ls(CASMTR_K_EXTRACT, stream);
    launch_k(pack_matches_kernel, (M + 1 + 255) / 256, 256, 0, stream, b_ids, i_ids, j_ids, mconf, mk0, mk1, M, pair_offset, out, count_dev);
    CASMTR_CHECK_LAUNCH("pack_matches_kernel");
    return CASMTR_OK;
}

static int extract_blocks(const casmtr_extract_desc &d) {
    return (int)(((size_t)d.B * d.h0 * d.w0 + BLK - 1) / BLK);
}

size_t match_extract_workspace(const casmtr_extract_desc &d) {
    Workspace ws(nullptr, 0);
    const int nb = extract_blocks(d);
    ws.take<uint8_t>((size_t)d.B * d.h0 * d.w0);
    ws.take<int>(nb);
    ws.take<int>(nb);
    ws.take<int>(2);
    ws.take<int>(2 * (size_t)d.B);
    ws.take<int>(2 * (size_t)d.B);
    return ws.off;
}

int launch_match_extract(const casmtr_extract_desc &d, const float *next_conf01, const int64_t *next_idx01,
                         const int64_t *next_idx10, uint8_t *mask_out, int64_t *b_ids, int64_t *i_ids,
                         int64_t *j_ids, float *mconf, float *mkpts0, float *mkpts1, int capacity,
                         int32_t *count_out, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
    Workspace ws(workspace, workspace_bytes);
    const int nb = extract_blocks(d);
    ExtractArgs a;
    a.d = d;
    a.conf = next_conf01; a.idx01 = next_idx01; a.idx10 = next_idx10;
    a.mask = ws.take<uint8_t>((size_t)d.B * d.h0 * d.w0);
    a.block_count = ws.take<int>(nb);
    a.block_off = ws.take<int>(nb);
    a.total = ws.take<int>(2);
    int *v0 = ws.take<int>(2 * (size_t)d.B);
    int *v1 = ws.take<int>(2 * (size_t)d.B);
    CASMTR_REQUIRE(ws.ok(), CASMTR_E_WORKSPACE, "match_extract: workspace %zu < %zu bytes", workspace_bytes, ws.off);
    a.valid0 = a.valid1 = nullptr;
    if (d.pad_mask0 && d.pad_mask1 && d.border_rm > 0) {
        { LaunchScope ls(CASMTR_K_EXTRACT, stream); valid_extent_kernel<<<d.B, 256, 0, stream>>>(d.pad_mask0, d.h0, d.w0, v0); }
        { LaunchScope ls(CASMTR_K_EXTRACT, stream); valid_extent_kernel<<<d.B, 256, 0, stream>>>(d.pad_mask1, d.h1, d.w1, v1); }
        CASMTR_CHECK_LAUNCH("valid_extent_kernel");
        a.valid0 = v0; a.valid1 = v1;
    }
    if (nb == 0) {
        cudaMemsetAsync(count_out, 0, sizeof(int32_t), stream);
        return CASMTR_OK;
    }
    { LaunchScope ls(CASMTR_K_EXTRACT, stream); launch_k(extract_mask_kernel, nb, BLK, 0, stream, a); }
    CASMTR_CHECK_LAUNCH("extract_mask_kernel");
    { LaunchScope ls(CASMTR_K_EXTRACT, stream); launch_k(extract_scan_kernel, 1, 1024, 0, stream, a, nb, count_out); }
    CASMTR_CHECK_LAUNCH("extract_scan_kernel");
    EmitOut e{mask_out, b_ids, i_ids, j_ids, mconf, mkpts0, mkpts1, capacity};
    { LaunchScope ls(CASMTR_K_EXTRACT, stream); launch_k(extract_emit_kernel, nb, BLK, 0, stream, a, e); }
    CASMTR_CHECK_LAUNCH("extract_emit_kernel");
    return CASMTR_OK;
}
