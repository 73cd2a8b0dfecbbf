// Dense dual-softmax coarse matching statistics on the 5th-generation tensor cores ("next" #1).
// Reference: CoarseMatching.forward  src/model/functions/coarse_matching.py:40-89
//   sim = <f0/sqrt(C), f1/sqrt(C)> / T                       [B, L, S]   (468 MB per tensor at 832^2, four of them)
//   next_conf_c01, next_idx_c01 = max_j softmax_j(sim);  next_conf_c10, next_idx_c10 = max_i softmax_i(sim)
// The cascade stages consume only those four vectors, so the L x S matrix is never written: one kernel computes, for both
// directions, the row-wise (max, sum-exp, arg-max) of sim with the GEMM on tcgen05 and the softmax statistics in the
// epilogue, a second tiny kernel merges the column splits.
//
// fp32 accuracy on TF32 tensor cores: kind::tf32 reads fp32 words from shared memory and uses their top 19 bits, so
// x_hi = x is implicit and x_lo = x - trunc_tf32(x) is precomputed once per feature map; three MMAs per k-step
// (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM) give ~2^-21 relative error -- the arg-max must match the fp32 reference.
//
// CTA = 128 rows x a range of 256-column tiles.  Warp 0: TMA producer (4 operand tiles per 32-channel k-block: A_hi, A_lo
// 128x32 and B_hi, B_lo 256x32 fp32, 128-byte swizzle, 2-stage ring, 96 KB per stage).  Warp 1: TMEM allocation + MMA issue
// (UMMA 128x256x8, 12 per k-block), accumulators double-buffered in TMEM (2 x 256 columns) so the epilogue of tile n
// overlaps the MMAs of tile n+1.  Warps 2-5: epilogue, one accumulator row per thread (tcgen05.ld 32x32b), online softmax.
#include "common.cuh"
#include "kernels.cuh"
#include "tma.cuh"

namespace {

using namespace tma;
constexpr int BM = 128, BN = 256, BK = 32;          // rows per CTA, columns per tile, fp32 channels per k-block (= 128 B)
constexpr int NSTG = 2;
constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;              // 96 KB
constexpr int SM_BAR = NSTG * STAGE_BYTES;
constexpr int SM_TOTAL = SM_BAR + 128;

struct CoarseMaps { CUtensorMap a_hi[2], a_lo[2], b_hi[2], b_lo[2]; };      // [direction]: rows = image d, cols = image 1-d

struct CoarseStatParams {
    float *pmax, *psum;         // [2][nsplit][B * Lmax] partial row statistics (log2 domain)
    int *parg;
    int B, L[2], C, nsplit, Lmax;
    float scale_log2;           // log2(e) / (C * temperature)
    const uint32_t *mbits[2];   // padding masks of image 0 / 1 packed 32 tokens per word ([B][mwords]), or NULL (reference :64-65)
    int mwords[2];
    // second pass (mutual nearest neighbours of conf = softmax_i * softmax_j, reference :68, :122): per direction the log-sum-exp
    // (base 2) of every COLUMN of that direction's similarity, [B][Lmax]; the kernel then reduces 2 * sim - lse_col over each row
    const float *cbias[2];
    float *lse;                 // [2][B * Lmax] out of the first merge: log2-sum-exp of every row of direction 0 / 1
};

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}

// K-major operand tile [rows][32 fp32] with 128-byte swizzle: 8-row groups are 1024 B apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(const void *smem_tile) {
    const uint64_t addr = (uint64_t)(smem_u32(smem_tile) >> 4) & 0x3fffull;
    return addr | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

template <bool BIAS>
__global__ void __launch_bounds__(192, 1) coarse_rowstats_kernel(const __grid_constant__ CoarseMaps maps, CoarseStatParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // 128B-swizzled TMA tiles want 1024-byte alignment; an offset
                                                                                 // on the __shared__ pointer (not an integer round trip) keeps LDS/STS
    uint64_t *full = (uint64_t *)(sm + SM_BAR), *empty = full + NSTG, *tfull = empty + NSTG, *tempty = tfull + 2;
    uint32_t *tmem_slot = (uint32_t *)(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.z / p.B, b = blockIdx.z % p.B;
    const int Lr = p.L[dir], Lc = p.L[1 - dir];
    const int row0 = blockIdx.x * BM;
    if (row0 >= Lr) return;                                     // the grid is sized for the longer image
    const int n_tiles = (Lc + BN - 1) / BN;
    const int t_begin = (int)((long long)n_tiles * blockIdx.y / p.nsplit), t_end = (int)((long long)n_tiles * (blockIdx.y + 1) / p.nsplit);
    const int KB = p.C / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTG; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, 4); }
    }
    if (warp == 1) {                                            // TMEM: all 512 columns (two 128 x 256 fp32 accumulators)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int step = 0;
            for (int t = t_begin; t < t_end; ++t)
                for (int kb = 0; kb < KB; ++kb, ++step) {
                    const int s = step % NSTG;
                    mbar_wait(empty + s, ((step / NSTG) & 1) ^ 1);
                    uint8_t *st = sm + s * STAGE_BYTES;
                    mbar_expect_tx(full + s, STAGE_BYTES);
                    tma_load_3d(st, &maps.a_hi[dir], kb * BK, row0, b, full + s);
                    tma_load_3d(st + A_BYTES, &maps.a_lo[dir], kb * BK, row0, b, full + s);
                    tma_load_3d(st + 2 * A_BYTES, &maps.b_hi[dir], kb * BK, t * BN, b, full + s);
                    tma_load_3d(st + 2 * A_BYTES + B_BYTES, &maps.b_lo[dir], kb * BK, t * BN, b, full + s);
                }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 256, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            int step = 0;
            for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
                const int a = it & 1;
                mbar_wait(tempty + a, ((it >> 1) & 1) ^ 1);     // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem_base + a * BN;
                for (int kb = 0; kb < KB; ++kb, ++step) {
                    const int s = step % NSTG;
                    mbar_wait(full + s, (step / NSTG) & 1);
                    tc_fence_after();
                    uint8_t *st = sm + s * STAGE_BYTES;
                    const uint64_t ah = umma_desc(st), al = umma_desc(st + A_BYTES), bh = umma_desc(st + 2 * A_BYTES),
                                   bl = umma_desc(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {          // UMMA_K = 8 tf32 = 32 bytes: +2 in descriptor address units
                        umma_tf32(d, ah + 2 * k, bh + 2 * k, idesc, (kb | k) != 0);
                        umma_tf32(d, ah + 2 * k, bl + 2 * k, idesc, 1);
                        umma_tf32(d, al + 2 * k, bh + 2 * k, idesc, 1);
                    }
                    umma_commit(empty + s);                      // smem stage reusable once these MMAs have read it
                }
                umma_commit(tfull + a);                          // accumulator complete
            }
        }
    } else {
        // ================= epilogue: one accumulator row per thread =================
        const int q = warp & 3;                                  // TMEM lane quarter this warp may access
        const int row = row0 + 32 * q + lane;
        float m = -INFINITY, l = 0.f;
        int arg = 0;
        const uint32_t *cmask = p.mbits[1 - dir] ? p.mbits[1 - dir] + (size_t)b * p.mwords[1 - dir] : nullptr;
        const float *cb = BIAS ? p.cbias[dir] + (size_t)b * p.Lmax : nullptr;
        const float sc = BIAS ? 2.f * p.scale_log2 : p.scale_log2;
        for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
            const int a = it & 1;
            mbar_wait(tfull + a, (it >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + a * BN + 32 * c, v);
                const int col0 = t * BN + 32 * c;
                // masked columns take no part (masked_fill_(-1e9) underflows to 0 in the soft-max, reference :64-65); col0 is a multiple of 32 = one mask word
                const uint32_t cbits = (cmask != nullptr && col0 < Lc) ? __ldg(cmask + (col0 >> 5)) : 0xffffffffu;
                float cm = -INFINITY;
                int ca = 0;
                if (BIAS) {                                       // log2 conf[i, j] + lse_row_i = 2 sim - lse_col_j: only max / arg-max are needed
                    float bias[32];
#pragma unroll
                    for (int i4 = 0; i4 < 8; ++i4) {
                        float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (col0 + 4 * i4 < Lc) t4 = ldg4(cb + col0 + 4 * i4);      // Lmax is padded to a multiple of 4
                        bias[4 * i4] = t4.x; bias[4 * i4 + 1] = t4.y; bias[4 * i4 + 2] = t4.z; bias[4 * i4 + 3] = t4.w;
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        v[i] = (col0 + i < Lc && ((cbits >> i) & 1u)) ? fmaf(v[i], sc, -bias[i]) : -INFINITY;
                        if (v[i] > cm) { cm = v[i]; ca = i; }
                    }
                    if (cm > m) { m = cm; arg = col0 + ca; }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        v[i] = (col0 + i < Lc && ((cbits >> i) & 1u)) ? v[i] * sc : -INFINITY;
                        if (v[i] > cm) { cm = v[i]; ca = i; }
                    }
                    if (cm > m) { l *= exp2f(m - cm); m = cm; arg = col0 + ca; }
                    if (m > -INFINITY) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) l += exp2f(v[i] - m);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + a);
        }
        if (row < Lr) {
            const size_t o = ((size_t)(dir * p.nsplit + blockIdx.y) * p.B + b) * p.Lmax + row;
            p.pmax[o] = m; p.psum[o] = l; p.parg[o] = arg;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base) : "memory");
    }
}

// lo = x - trunc_tf32(x): the part of x the tensor core drops when it reads x as tf32
__global__ void tf32_residual_kernel(const float4 *__restrict__ x, float4 *__restrict__ lo, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = x[i];
        float4 r;
        r.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
        r.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        r.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
        r.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
        lo[i] = r;
    }
}

// merge the column splits: conf = 1 / sum_j exp(s_j - max) (the soft-max value at the arg-max), idx = arg-max
__global__ void coarse_merge_kernel(CoarseStatParams p, float *conf01, int64_t *idx01, float *conf10, int64_t *idx10) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const int dir = blockIdx.y;
    const int Lr = p.L[dir];
    if (i >= (size_t)p.B * Lr) return;
    const int b = (int)(i / Lr), row = (int)(i % Lr);
    float m = -INFINITY;
    int arg = 0;
    for (int s = 0; s < p.nsplit; ++s) {
        const size_t o = ((size_t)(dir * p.nsplit + s) * p.B + b) * p.Lmax + row;
        if (p.pmax[o] > m) { m = p.pmax[o]; arg = p.parg[o]; }
    }
    float l = 0.f;
    for (int s = 0; s < p.nsplit; ++s) {
        const size_t o = ((size_t)(dir * p.nsplit + s) * p.B + b) * p.Lmax + row;
        if (p.pmax[o] > -INFINITY) l += p.psum[o] * exp2f(p.pmax[o] - m);
    }
    // the reference fills with -INF = -1e9 (coarse_matching.py:6), not -infinity: a padded row (or a row whose columns are all
    // padded) is a constant row, its soft-max is uniform and torch.max returns (1 / columns, 0); in every other row the
    // padded columns underflow to exactly 0
    const bool dead = m == -INFINITY || (p.mbits[dir] && !((p.mbits[dir][(size_t)b * p.mwords[dir] + (row >> 5)] >> (row & 31)) & 1u));
    (dir ? conf10 : conf01)[i] = dead ? 1.0f / (float)p.L[1 - dir] : 1.0f / l;
    (dir ? idx10 : idx01)[i] = dead ? 0 : arg;
    // log2-sum-exp of the row, for the second pass; +inf marks a dead row (its conf is 0: never a mutual match)
    if (p.lse) p.lse[((size_t)dir * p.B + b) * p.Lmax + row] = dead ? INFINITY : m + log2f(l);
}

// second merge: row maximum / arg-max of conf = softmax_i * softmax_j from the partial maxima of 2 sim - lse_col:
//   mconf_row[i] = max_j conf[i, j] = 2^(max_j(2 sim_ij - lse_col_j) - lse_row_i),  midx_row[i] = its arg-max; midx_col likewise over i
__global__ void coarse_merge2_kernel(CoarseStatParams p, float *mconf_row, int64_t *midx_row, int64_t *midx_col) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const int dir = blockIdx.y;
    const int Lr = p.L[dir];
    if (i >= (size_t)p.B * Lr) return;
    const int b = (int)(i / Lr), row = (int)(i % Lr);
    float m = -INFINITY;
    int arg = 0;
    for (int s = 0; s < p.nsplit; ++s) {
        const size_t o = ((size_t)(dir * p.nsplit + s) * p.B + b) * p.Lmax + row;
        if (p.pmax[o] > m) { m = p.pmax[o]; arg = p.parg[o]; }
    }
    const float lse = p.lse[((size_t)dir * p.B + b) * p.Lmax + row];
    const bool dead = m == -INFINITY || lse == INFINITY;
    if (dir == 0) {
        mconf_row[i] = dead ? 0.f : exp2f(m - lse);
        midx_row[i] = dead ? 0 : arg;
    } else {
        midx_col[i] = dead ? -1 : arg;              // -1: no row points back to a dead column
    }
}

// uint8 mask [B, L] -> one bit per token, 32 tokens per word: a warp per word
__global__ void mask_pack_kernel(const uint8_t *__restrict__ mask, uint32_t *__restrict__ bits, int B, int L, int words) {
    const size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= (size_t)B * words) return;
    const int b = (int)(w / words), tok = (int)(w % words) * 32 + lane;
    const unsigned bal = __ballot_sync(FULL_MASK, tok < L && mask[(size_t)b * L + tok] != 0);
    if (lane == 0) bits[w] = bal;
}

int make_map3(CUtensorMap *tm, const float *base, int B, int L, int C, int rows) {
    EncodeTiledFn enc = encode_tiled();
    CASMTR_REQUIRE(enc != nullptr, CASMTR_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)L * C * 4};
    const cuuint32_t box[3] = {BK, (cuuint32_t)rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CASMTR_REQUIRE(r == CUDA_SUCCESS, CASMTR_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CASMTR_OK;
}

// column splits per row block: enough CTAs for a few waves of the (one CTA per SM) grid, with the fullest last wave
int pick_nsplit(int B, int L0, int L1) {
    const int rb = ((L0 + BM - 1) / BM + (L1 + BM - 1) / BM) * B;        // row blocks of both directions
    const int nt = ((L0 < L1 ? L0 : L1) + BN - 1) / BN;
    int best = 1;
    double best_eff = 0.0;
    for (int ns = 1; ns <= nt && ns <= 16; ++ns) {
        const long long ctas = (long long)rb * ns, waves = (ctas + 147) / 148;
        if (waves > 12) break;
        const double eff = (double)ctas / (double)(waves * 148);
        if (ctas >= 2 * 148 && eff > best_eff + 0.02) { best_eff = eff; best = ns; }
        else if (ctas < 2 * 148) best = ns;                              // keep splitting until the machine is covered twice
    }
    return best;
}

}  // namespace

size_t coarse_match_workspace(int B, int L0, int L1, int C) {
    Workspace ws(nullptr, 0);
    const int Lmax = ((L0 > L1 ? L0 : L1) + 3) / 4 * 4, ns = pick_nsplit(B, L0, L1);
    ws.take<float>((size_t)2 * B * Lmax);
    ws.take<float>((size_t)B * L0 * C);
    ws.take<float>((size_t)B * L1 * C);
    ws.take<float>((size_t)2 * ns * B * Lmax);
    ws.take<float>((size_t)2 * ns * B * Lmax);
    ws.take<int>((size_t)2 * ns * B * Lmax);
    ws.take<uint32_t>((size_t)B * ((L0 + 31) / 32));
    ws.take<uint32_t>((size_t)B * ((L1 + 31) / 32));
    return ws.off;
}

int launch_coarse_match(const float *feat0, const float *feat1, const uint8_t *mask0, const uint8_t *mask1, float temperature,
                        float *conf01, int64_t *idx01, float *conf10,
                        int64_t *idx10, int B, int L0, int L1, int C, void *workspace, size_t workspace_bytes, cudaStream_t stream,
                        float *mconf_row, int64_t *midx_row, int64_t *midx_col) {
    CASMTR_REQUIRE(C % BK == 0 && C >= BK, CASMTR_E_UNSUPPORTED, "coarse_match: C=%d must be a multiple of %d", C, BK);
    CASMTR_REQUIRE((((uintptr_t)feat0 | (uintptr_t)feat1) & 15) == 0, CASMTR_E_INVALID, "coarse_match: features must be 16-byte aligned");
    Workspace ws(workspace, workspace_bytes);
    const int Lmax = ((L0 > L1 ? L0 : L1) + 3) / 4 * 4, ns = pick_nsplit(B, L0, L1);
    float *lse = ws.take<float>((size_t)2 * B * Lmax);
    float *lo0 = ws.take<float>((size_t)B * L0 * C), *lo1 = ws.take<float>((size_t)B * L1 * C);
    CoarseStatParams p;
    p.lse = mconf_row ? lse : nullptr;
    p.cbias[0] = p.cbias[1] = nullptr;
    p.pmax = ws.take<float>((size_t)2 * ns * B * Lmax);
    p.psum = ws.take<float>((size_t)2 * ns * B * Lmax);
    p.parg = ws.take<int>((size_t)2 * ns * B * Lmax);
    uint32_t *bits0 = ws.take<uint32_t>((size_t)B * ((L0 + 31) / 32)), *bits1 = ws.take<uint32_t>((size_t)B * ((L1 + 31) / 32));
    CASMTR_REQUIRE(ws.ok(), CASMTR_E_WORKSPACE, "coarse_match: workspace %zu < %zu bytes", workspace_bytes, ws.off);
    p.B = B; p.L[0] = L0; p.L[1] = L1; p.C = C; p.nsplit = ns; p.Lmax = Lmax;
    p.scale_log2 = LOG2E_F / ((float)C * temperature);
    p.mbits[0] = p.mbits[1] = nullptr;
    p.mwords[0] = (L0 + 31) / 32; p.mwords[1] = (L1 + 31) / 32;
    if (mask0 != nullptr) {
        LaunchScope ls(CASMTR_K_COARSE_MATCH, stream);
        mask_pack_kernel<<<(unsigned)(((size_t)B * p.mwords[0] * 32 + 255) / 256), 256, 0, stream>>>(mask0, bits0, B, L0, p.mwords[0]);
        mask_pack_kernel<<<(unsigned)(((size_t)B * p.mwords[1] * 32 + 255) / 256), 256, 0, stream>>>(mask1, bits1, B, L1, p.mwords[1]);
        CASMTR_CHECK_LAUNCH("mask_pack_kernel");
        p.mbits[0] = bits0; p.mbits[1] = bits1;
    }
    {
        LaunchScope ls(CASMTR_K_COARSE_MATCH, stream);
        tf32_residual_kernel<<<148 * 8, 256, 0, stream>>>((const float4 *)feat0, (float4 *)lo0, (size_t)B * L0 * C / 4);
    }
    {
        LaunchScope ls(CASMTR_K_COARSE_MATCH, stream);
        tf32_residual_kernel<<<148 * 8, 256, 0, stream>>>((const float4 *)feat1, (float4 *)lo1, (size_t)B * L1 * C / 4);
    }
    CASMTR_CHECK_LAUNCH("tf32_residual_kernel");
    CoarseMaps maps;
    const float *hi[2] = {feat0, feat1}, *lo[2] = {lo0, lo1};
    const int Ls[2] = {L0, L1};
    int rc = CASMTR_OK;
    for (int d = 0; d < 2 && rc == CASMTR_OK; ++d) {
        rc = make_map3(&maps.a_hi[d], hi[d], B, Ls[d], C, BM);
        if (rc == CASMTR_OK) rc = make_map3(&maps.a_lo[d], lo[d], B, Ls[d], C, BM);
        if (rc == CASMTR_OK) rc = make_map3(&maps.b_hi[d], hi[1 - d], B, Ls[1 - d], C, BN);
        if (rc == CASMTR_OK) rc = make_map3(&maps.b_lo[d], lo[1 - d], B, Ls[1 - d], C, BN);
    }
    if (rc != CASMTR_OK) return rc;
    const size_t smem = 1024 + SM_TOTAL;
    static PerDeviceOnce once;
    const int dev = PerDeviceOnce::device();
    if (!once.done(dev)) {
        cudaError_t e = cudaFuncSetAttribute(coarse_rowstats_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(coarse_rowstats_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { casmtr_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return CASMTR_E_CUDA; }
        once.mark(dev);
    }
    const int Lgrid = L0 > L1 ? L0 : L1;
    {
        LaunchScope ls(CASMTR_K_COARSE_MATCH, stream);
        coarse_rowstats_kernel<false><<<dim3((Lgrid + BM - 1) / BM, ns, 2 * B), 192, smem, stream>>>(maps, p);
        CASMTR_CHECK_LAUNCH("coarse_rowstats_kernel");
    }
    {
        LaunchScope ls(CASMTR_K_COARSE_MATCH, stream);
        coarse_merge_kernel<<<dim3(((size_t)B * Lgrid + 255) / 256, 2), 256, 0, stream>>>(p, conf01, idx01, conf10, idx10);
        CASMTR_CHECK_LAUNCH("coarse_merge_kernel");
    }
    if (mconf_row == nullptr) return CASMTR_OK;
    // second pass: the columns of direction d are the rows of direction 1 - d
    p.cbias[0] = lse + (size_t)B * Lmax;
    p.cbias[1] = lse;
    {
        LaunchScope ls(CASMTR_K_COARSE_MATCH, stream);
        coarse_rowstats_kernel<true><<<dim3((Lgrid + BM - 1) / BM, ns, 2 * B), 192, smem, stream>>>(maps, p);
        CASMTR_CHECK_LAUNCH("coarse_rowstats_kernel (mutual pass)");
    }
    {
        LaunchScope ls(CASMTR_K_COARSE_MATCH, stream);
        coarse_merge2_kernel<<<dim3(((size_t)B * Lgrid + 255) / 256, 2), 256, 0, stream>>>(p, mconf_row, midx_row, midx_col);
        CASMTR_CHECK_LAUNCH("coarse_merge2_kernel");
    }
    return CASMTR_OK;
}
