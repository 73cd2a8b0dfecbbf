// Layout kernels: NCHW -> token-major feature rows, internal top-k lists -> reference layout.
//
// The reference rearranges every pyramid level 'b c h w -> b (h w) c' and .contiguous()s it
// (cuda_imp/QuadTreeAttention/QuadtreeAttention/modules/quadtree_attention.py:166-168,185-189);
// here all (up to 12) maps of one attention call go through ONE batched tile-transpose launch.
// Token-major makes one (token, head) row exactly one 128-byte line, which is what the gather
// kernels want.  HBM-bound: 4 B read + 4 B written per element.
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

__global__ void __launch_bounds__(256) transpose_jobs_kernel(TransposeJobs jobs) {
    __shared__ float tile[32][33];
    int t = blockIdx.x;
    int j = 0;
#pragma unroll 1
    while (j + 1 < jobs.n && t >= jobs.job[j + 1].tile_begin) ++j;
    const TransposeJob jb = jobs.job[j];
    t -= jb.tile_begin;
    const int tiles_c = (jb.C + 31) >> 5;
    const int tc = t % tiles_c, tt = t / tiles_c;
    const int b = blockIdx.y;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float *src = jb.src + (size_t)b * jb.C * jb.HW;
    float *dst = jb.dst + (size_t)b * jb.C * jb.HW;
    const int tok_r = tt * 32 + tx;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int c = tc * 32 + ty + 8 * r;
        if (c < jb.C && tok_r < jb.HW) tile[ty + 8 * r][tx] = __ldg(src + (size_t)c * jb.HW + tok_r);
    }
    __syncthreads();
    const int c_w = tc * 32 + tx;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int tok = tt * 32 + ty + 8 * r;
        if (tok < jb.HW && c_w < jb.C) dst[(size_t)tok * jb.C + c_w] = tile[tx][ty + 8 * r];
    }
}

// Vectorised variant (every map has C % 32 == 0 and HW % 16 == 0): no shared memory at all.  One warp moves a
// 32-channel x 16-token block: lane = (channel group cg = lane & 7, token quad tq = lane >> 3) loads one float4 (4 tokens)
// from each of its 4 channels -- 64-byte contiguous pieces per channel -- transposes the 4x4 block in registers and
// stores 4 float4, one per token; for a fixed token the 8 channel groups of a warp write one full 128-byte line.
// 8 memory instructions per 16 elements per thread: the scalar tile kernel above is issue-bound (ncu: 77 % issue
// active at 3.6 TB/s), this one is HBM-bound.
struct VecJobs {
    TransposeJob job[12];
    int blk_begin[13];          // first 32ch x 128tok block of every job (a CTA = 8 warps = 128 tokens)
    int n;
};

__global__ void __launch_bounds__(256) transpose_vec_kernel(VecJobs jobs) {
    pdl_sync();
    int t = blockIdx.x;
    int j = 0;
#pragma unroll 1
    while (j + 1 < jobs.n && t >= jobs.blk_begin[j + 1]) ++j;
    const TransposeJob jb = jobs.job[j];
    t -= jobs.blk_begin[j];
    const int cblocks = jb.C >> 5;
    const int cb = t % cblocks, tb = t / cblocks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cg = lane & 7, tq = lane >> 3;
    const int tok = tb * 128 + warp * 16 + tq * 4;
    if (tok >= jb.HW) return;
    const int c0 = cb * 32 + cg * 4;
    const size_t b = blockIdx.y;
    const float *src = jb.src + b * (size_t)jb.C * jb.HW + (size_t)c0 * jb.HW + tok;
    float *dst = jb.dst + b * (size_t)jb.C * jb.HW + (size_t)tok * jb.C + c0;
    const float4 r0 = ldg4_stream(src), r1 = ldg4_stream(src + jb.HW), r2 = ldg4_stream(src + 2 * (size_t)jb.HW),
                 r3 = ldg4_stream(src + 3 * (size_t)jb.HW);
    *reinterpret_cast<float4 *>(dst) = make_float4(r0.x, r1.x, r2.x, r3.x);
    *reinterpret_cast<float4 *>(dst + jb.C) = make_float4(r0.y, r1.y, r2.y, r3.y);
    *reinterpret_cast<float4 *>(dst + 2 * (size_t)jb.C) = make_float4(r0.z, r1.z, r2.z, r3.z);
    *reinterpret_cast<float4 *>(dst + 3 * (size_t)jb.C) = make_float4(r0.w, r1.w, r2.w, r3.w);
}

int launch_transpose_jobs(TransposeJobs &jobs, int B, cudaStream_t stream) {
    bool vec = true;
    for (int i = 0; i < jobs.n; ++i)
        vec = vec && jobs.job[i].C % 32 == 0 && jobs.job[i].HW % 4 == 0 &&
              (((uintptr_t)jobs.job[i].src | (uintptr_t)jobs.job[i].dst) & 15) == 0;
    if (vec && jobs.n > 0 && B > 0) {
        VecJobs vj;
        vj.n = jobs.n;
        int total = 0;
        for (int i = 0; i < jobs.n; ++i) {
            vj.job[i] = jobs.job[i];
            vj.blk_begin[i] = total;
            total += (jobs.job[i].C / 32) * ((jobs.job[i].HW + 127) / 128);
        }
        vj.blk_begin[jobs.n] = total;
        if (total == 0) return CASMTR_OK;
        LaunchScope ls(CASMTR_K_LAYOUT, stream);
        launch_k(transpose_vec_kernel, dim3(total, B), 256, 0, stream, vj);
        CASMTR_CHECK_LAUNCH("transpose_vec_kernel");
        return CASMTR_OK;
    }

    int total = 0;
    for (int i = 0; i < jobs.n; ++i) {
        jobs.job[i].tile_begin = total;
        total += ((jobs.job[i].C + 31) / 32) * ((jobs.job[i].HW + 31) / 32);
    }
    if (total == 0 || B == 0) return CASMTR_OK;
    LaunchScope ls(CASMTR_K_LAYOUT, stream);
    transpose_jobs_kernel<<<dim3(total, B), 256, 0, stream>>>(jobs);
    CASMTR_CHECK_LAUNCH("transpose_jobs_kernel");
    return CASMTR_OK;
}

// 2x2 average pooling of token-major maps [B, h*w, C] -> [B, (h/2)*(w/2), C]: one pyramid level of
// QuadtreeAttention.forward (src/model/modules/quadtree_attention.py:86-89, F.avg_pool2d(kernel 2, stride 2)); the four
// taps are summed in avg_pool2d's window order.  Thread = 4 channels of one output token; up to 3 maps (q, k, v) per launch.
__global__ void __launch_bounds__(256) pool_tokens_kernel(const __grid_constant__ PoolJobs jobs, int C4) {
    pdl_sync();
    const int j = blockIdx.z, b = blockIdx.y;
    const PoolJob jb = jobs.job[j];
    __builtin_assume(__isGlobal(jb.src) && __isGlobal(jb.dst));
    const int ho = jb.h >> 1, wo = jb.w >> 1;
    const size_t n = (size_t)ho * wo * C4;
    const float4 *src = reinterpret_cast<const float4 *>(jb.src) + (size_t)b * jb.h * jb.w * C4;
    float4 *dst = reinterpret_cast<float4 *>(jb.dst) + (size_t)b * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        const int t = (int)(i / C4), y = t / wo, x = t - y * wo;
        const float4 *p = src + ((size_t)(2 * y) * jb.w + 2 * x) * C4 + c;
        const float4 a = __ldg(p), bb = __ldg(p + C4), cc = __ldg(p + (size_t)jb.w * C4), d = __ldg(p + (size_t)(jb.w + 1) * C4);
        float4 o;
        o.x = (((a.x + bb.x) + cc.x) + d.x) * 0.25f;
        o.y = (((a.y + bb.y) + cc.y) + d.y) * 0.25f;
        o.z = (((a.z + bb.z) + cc.z) + d.z) * 0.25f;
        o.w = (((a.w + bb.w) + cc.w) + d.w) * 0.25f;
        dst[i] = o;
    }
}

// Two pyramid levels in one pass (h, w multiples of 4): thread = 4 channels of one level-2 token; its 4x4 block of level-0
// tokens is read once (two rows = 8 loads in flight at a time), the four level-1 averages are written and averaged again (the reference pools
// the pooled map).  A warp covers 4 consecutive level-2 tokens x 8 channel quads (every request = four full 128-byte lines), so
// that the operands the tensor-core coarsest level wants next to the level-2 maps come out of the same pass
// (qtatt_coarse_tc.cu; coarse_prep_kernel is the stand-alone version): lo2 = x - trunc_tf32(x) of the level-2 map (Q, K), and
// the level-2 map transposed to channel-major fp16 pairs V * 2^8 = hi + lo (V) -- a 4 x 4 exchange between the four lanes
// that hold a channel quad of four neighbouring tokens leaves each lane with one channel x 4 tokens = one 8-byte store per half.
__device__ __forceinline__ float4 avg4(const float4 a, const float4 b, const float4 c, const float4 d) {
    float4 o;
    o.x = (((a.x + b.x) + c.x) + d.x) * 0.25f;
    o.y = (((a.y + b.y) + c.y) + d.y) * 0.25f;
    o.z = (((a.z + b.z) + c.z) + d.z) * 0.25f;
    o.w = (((a.w + b.w) + c.w) + d.w) * 0.25f;
    return o;
}
__device__ __forceinline__ float tf32_residual(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ uint32_t f16_pair(float x) {          // x * 2^8 = hi + lo: hi in bits 0..15, lo in bits 16..31
    const float xs = fminf(fmaxf(x * 256.f, -65504.f), 65504.f);
    const __half hi = __float2half_rn(xs);
    const __half lo = __float2half_rn(xs - __half2float(hi));
    return (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
}

__global__ void __launch_bounds__(256, 4) pool2_tokens_kernel(const __grid_constant__ PoolJobs jobs, int C4) {
    pdl_sync();
    const int j = blockIdx.z, b = blockIdx.y;
    const PoolJob jb = jobs.job[j];
    __builtin_assume(__isGlobal(jb.src) && __isGlobal(jb.dst) && __isGlobal(jb.dst2));
    const int h1 = jb.h >> 1, w1 = jb.w >> 1, h2 = jb.h >> 2, w2 = jb.w >> 2;
    const int n_tok = h2 * w2;
    const int n_tok_pad = jb.vt_hi ? jb.Sp : n_tok;                // the transposed map is zero-filled up to its padded row length
    const int per_unit = 4 * C4;                                   // threads per unit of 4 tokens (C4 is a multiple of 8)
    const size_t n = (size_t)((n_tok_pad + 3) / 4) * per_unit;
    const float4 *src = reinterpret_cast<const float4 *>(jb.src) + (size_t)b * jb.h * jb.w * C4;
    float4 *d1 = reinterpret_cast<float4 *>(jb.dst) + (size_t)b * h1 * w1 * C4;
    float4 *d2 = reinterpret_cast<float4 *>(jb.dst2) + (size_t)b * n_tok * C4;
    float4 *l2 = jb.lo2 ? reinterpret_cast<float4 *>(jb.lo2) + (size_t)b * n_tok * C4 : nullptr;
    const int lane = threadIdx.x & 31;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int u = (int)(i / per_unit), r = (int)(i - (size_t)u * per_unit);
        const int c = (r >> 5) * 8 + (lane & 7);                   // i - lane is a multiple of 32: r & 31 == lane
        const int t = 4 * u + (lane >> 3);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < n_tok) {
            const int y = t / w2, x = t - y * w2;
            const float4 *p = src + ((size_t)(4 * y) * jb.w + 4 * x) * C4 + c;
            float4 m[4];
#pragma unroll
            for (int half = 0; half < 2; ++half) {                 // two level-0 rows at a time: 8 loads in flight, 64 registers, 4 CTAs per SM
                float4 a[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) a[q] = __ldg(p + ((size_t)(2 * half + (q >> 2)) * jb.w + (q & 3)) * C4);
#pragma unroll
                for (int e = 0; e < 2; ++e) {                      // level-1 token (2y + half, 2x + e): taps in avg_pool2d's window order
                    m[2 * half + e] = avg4(a[2 * e], a[2 * e + 1], a[4 + 2 * e], a[4 + 2 * e + 1]);
                    d1[((size_t)(2 * y + half) * w1 + 2 * x + e) * C4 + c] = m[2 * half + e];
                }
            }
            o = avg4(m[0], m[1], m[2], m[3]);
            d2[(size_t)t * C4 + c] = o;
            if (l2) l2[(size_t)t * C4 + c] = make_float4(tf32_residual(o.x), tf32_residual(o.y), tf32_residual(o.z), tf32_residual(o.w));
        }
        if (jb.vt_hi) {                                            // warp-uniform
            // w[e] = channel 4c + e of token t as an fp16 (hi, lo) pair; 4 x 4 transpose over the lanes {lane & 7} + 8 k:
            // lane k ends with channel 4c + k of tokens 4u .. 4u + 3
            uint32_t w[4] = {f16_pair(o.x), f16_pair(o.y), f16_pair(o.z), f16_pair(o.w)};
            const int k = lane >> 3;
            {
                const bool up = k & 2;                             // exchange 2 x 2 blocks with lane ^ 16
                const uint32_t s0 = up ? w[0] : w[2], s1 = up ? w[1] : w[3];
                const uint32_t r0 = __shfl_xor_sync(FULL_MASK, s0, 16), r1 = __shfl_xor_sync(FULL_MASK, s1, 16);
                if (up) { w[0] = r0; w[1] = r1; } else { w[2] = r0; w[3] = r1; }
            }
            {
                const bool up = k & 1;                             // exchange single elements with lane ^ 8
                const uint32_t s0 = up ? w[0] : w[1], s1 = up ? w[2] : w[3];
                const uint32_t r0 = __shfl_xor_sync(FULL_MASK, s0, 8), r1 = __shfl_xor_sync(FULL_MASK, s1, 8);
                if (up) { w[0] = r0; w[2] = r1; } else { w[1] = r0; w[3] = r1; }
            }
            // after the two steps w[m] of lane k = (token 4u + m', channel ...): see the index algebra in tests (bit-exact vs coarse_prep_kernel)
            const size_t row = ((size_t)b * 4 * C4 + 4 * c + k) * jb.Sp + 4 * u;
            const uint2 hi = make_uint2((w[0] & 0xffffu) | (w[1] << 16), (w[2] & 0xffffu) | (w[3] << 16));
            const uint2 lo = make_uint2((w[0] >> 16) | (w[1] & 0xffff0000u), (w[2] >> 16) | (w[3] & 0xffff0000u));
            if (4 * u < jb.Sp) {
                *reinterpret_cast<uint2 *>(jb.vt_hi + row) = hi;
                *reinterpret_cast<uint2 *>(jb.vt_lo + row) = lo;
            }
        }
    }
}

int launch_pool2_tokens(const PoolJobs &jobs, int B, int C, cudaStream_t stream) {
    if (jobs.n == 0 || B == 0) return CASMTR_OK;
    CASMTR_REQUIRE(C % 32 == 0, CASMTR_E_UNSUPPORTED, "pool2_tokens: C=%d must be a multiple of 32", C);
    size_t nmax = 0;
    for (int i = 0; i < jobs.n; ++i) {
        const PoolJob &jb = jobs.job[i];
        CASMTR_REQUIRE((((uintptr_t)jb.src | (uintptr_t)jb.dst | (uintptr_t)jb.dst2 | (uintptr_t)jb.lo2 | (uintptr_t)jb.vt_hi | (uintptr_t)jb.vt_lo) & 15) == 0,
                       CASMTR_E_INVALID, "pool_tokens: unaligned map");
        CASMTR_REQUIRE(jb.h % 4 == 0 && jb.w % 4 == 0, CASMTR_E_INVALID, "pool2_tokens: grid not a multiple of 4");
        CASMTR_REQUIRE((jb.vt_hi == nullptr) == (jb.vt_lo == nullptr), CASMTR_E_INVALID, "pool2_tokens: both halves of the transposed map or none");
        const int n_tok = (jb.h / 4) * (jb.w / 4);
        CASMTR_REQUIRE(jb.vt_hi == nullptr || (jb.Sp % 4 == 0 && jb.Sp >= n_tok), CASMTR_E_INVALID, "pool2_tokens: padded row length %d", jb.Sp);
        nmax = std::max(nmax, (size_t)(((jb.vt_hi ? jb.Sp : n_tok) + 3) / 4) * C);
    }
    if (nmax == 0) return CASMTR_OK;
    LaunchScope ls(CASMTR_K_LAYOUT, stream);
    launch_k(pool2_tokens_kernel, dim3((unsigned)std::min<size_t>((nmax + 255) / 256, 148 * 8), B, jobs.n), 256, 0, stream, jobs, C / 4);
    CASMTR_CHECK_LAUNCH("pool2_tokens_kernel");
    return CASMTR_OK;
}

int launch_pool_tokens(const PoolJobs &jobs, int B, int C, cudaStream_t stream) {
    if (jobs.n == 0 || B == 0) return CASMTR_OK;
    CASMTR_REQUIRE(C % 4 == 0, CASMTR_E_UNSUPPORTED, "pool_tokens: C=%d must be a multiple of 4", C);
    size_t nmax = 0;
    for (int i = 0; i < jobs.n; ++i) {
        CASMTR_REQUIRE((((uintptr_t)jobs.job[i].src | (uintptr_t)jobs.job[i].dst) & 15) == 0, CASMTR_E_INVALID, "pool_tokens: unaligned map");
        nmax = std::max(nmax, (size_t)(jobs.job[i].h / 2) * (jobs.job[i].w / 2) * (C / 4));
    }
    if (nmax == 0) return CASMTR_OK;
    LaunchScope ls(CASMTR_K_LAYOUT, stream);
    launch_k(pool_tokens_kernel, dim3((unsigned)std::min<size_t>((nmax + 255) / 256, 148 * 8), B, jobs.n), 256, 0, stream, jobs, C / 4);
    CASMTR_CHECK_LAUNCH("pool_tokens_kernel");
    return CASMTR_OK;
}

// get_window_warp_idx (src/model/modules/transformer.py:416-440): next_idx [rows] on an H x W grid -> (row, col) of the
// win x win window around it, shifted rigidly inside the grid: pos [rows, win*win, 2] int64.
__global__ void window_idx_kernel(const int64_t *__restrict__ next_idx, int64_t *__restrict__ pos, size_t rows, int H, int W, int win) {
    pdl_sync();
    const int ww = win * win;
    const size_t total = rows * ww;
    for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
        const size_t r = o / ww;
        const int k = (int)(o - r * ww);
        const long long idx = next_idx[r];
        const int r0 = window_origin((int)(idx / W), win, H), c0 = window_origin((int)(idx % W), win, W);
        reinterpret_cast<longlong2 *>(pos)[o] = make_longlong2(r0 + k / win, c0 + k % win);
    }
}

int launch_window_idx(const int64_t *next_idx, int64_t *pos, size_t rows, int H, int W, int win, cudaStream_t stream) {
    if (rows == 0) return CASMTR_OK;
    LaunchScope ls(CASMTR_K_LAYOUT, stream);
    launch_k(window_idx_kernel, (unsigned)std::min<size_t>((rows * win * win + 255) / 256, 148 * 8), 256, 0, stream, next_idx, pos, rows, H, W, win);
    CASMTR_CHECK_LAUNCH("window_idx_kernel");
    return CASMTR_OK;
}

// internal [B,L,nh,k] int32 / fp32  ->  reference [B,L,k,nh] int64 / fp32
// Internal top-k lists ([n_tok, nh, k] int32 / fp32, in whatever order the level kernels emit them) -> the reference's API
// layout [n_tok, k, nh] int64 / fp32 SORTED by descending score like torch.topk(sorted=True) (ties: lower key index first).
// One warp per (token, head); rank counting over the k <= 32 entries.  Off the hot path: only callers that ask for the lists pay.
__global__ void topk_to_api_kernel(const int *__restrict__ idx, const float *__restrict__ score,
                                   int64_t *__restrict__ idx_out, float *__restrict__ score_out,
                                   size_t n_tok, int nh, int k) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const size_t n_items = n_tok * nh;
    for (size_t it = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5; it < n_items; it += ((size_t)gridDim.x * blockDim.x) >> 5) {
        const size_t tok = it / nh;
        const int h = (int)(it - tok * nh);
        const float s = lane < k ? score[it * k + lane] : -INFINITY;
        const int ix = lane < k ? idx[it * k + lane] : 0x7fffffff;
        int rank = 0;
        for (int l = 0; l < k; ++l) {
            const float so = __shfl_sync(FULL_MASK, s, l);
            const int io = __shfl_sync(FULL_MASK, ix, l);
            rank += (so > s) || (so == s && io < ix);
        }
        if (lane < k) {
            const size_t o = (tok * k + rank) * nh + h;
            if (idx_out) idx_out[o] = ix;
            if (score_out) score_out[o] = s;
        }
    }
}

// QTAttGuided: topk_pos [2, B, Np, K, nh] int64 (row, col on the hv x wv grid) -> the level kernels' candidate lists [B, Np, nh, K]
// int32 (flat cell index), and wsm[0..n) = softmax(level_weight[0..n))
__global__ void guided_prep_kernel(const int64_t *__restrict__ pos, int *__restrict__ idx, size_t n_cells, int K, int nh, int hv, int wv,
                                   const float *__restrict__ weight, int n_w, float *__restrict__ wsm) {
    pdl_sync();
    const size_t total = n_cells * K * nh;
    for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
        const int h = (int)(o % nh);
        const int k = (int)((o / nh) % K);
        const size_t cell = o / ((size_t)nh * K);
        const int r = (int)min(max(pos[o], (int64_t)0), (int64_t)hv - 1), c = (int)min(max(pos[total + o], (int64_t)0), (int64_t)wv - 1);
        idx[(cell * nh + h) * K + k] = r * wv + c;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && wsm) {
        float mx = -INFINITY, den = 0.f;
        for (int l = 0; l < n_w; ++l) mx = fmaxf(mx, weight[l]);
        for (int l = 0; l < n_w; ++l) den += expf(weight[l] - mx);
        for (int l = 0; l < n_w && l < CASMTR_MAX_LEVELS; ++l) wsm[l] = expf(weight[l] - mx) / den;
    }
}

int launch_guided_prep(const int64_t *pos, int *idx, size_t n_cells, int K, int nh, int hv, int wv, const float *weight, int n_w, float *wsm,
                       cudaStream_t stream) {
    const size_t total = n_cells * K * nh;
    if (total == 0) return CASMTR_OK;
    LaunchScope ls(CASMTR_K_LAYOUT, stream);
    launch_k(guided_prep_kernel, (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16), 256, 0, stream, pos, idx, n_cells, K, nh, hv, wv, weight, n_w, wsm);
    CASMTR_CHECK_LAUNCH("guided_prep_kernel");
    return CASMTR_OK;
}

int launch_topk_to_api(const int *idx, const float *score, int64_t *idx_out, float *score_out,
                       size_t n_tok, int nh, int k, cudaStream_t stream) {
    const size_t total = n_tok * nh;            // one warp each
    if (total == 0 || k <= 0) return CASMTR_OK;
    CASMTR_REQUIRE(k <= 32, CASMTR_E_UNSUPPORTED, "top-k export: k=%d > 32", k);
    size_t blocks = (total + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    LaunchScope ls(CASMTR_K_LAYOUT, stream);
    launch_k(topk_to_api_kernel, (unsigned)blocks, 256, 0, stream, idx, score, idx_out, score_out, n_tok, nh, k);
    CASMTR_CHECK_LAUNCH("topk_to_api_kernel");
    return CASMTR_OK;
}
