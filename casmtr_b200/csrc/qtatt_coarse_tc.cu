// Coarsest quadtree level on the 5th-generation tensor cores: S = scale * Q K^T and O = P V as tcgen05.mma (kind::tf32, fp32
// accuracy through error-compensated splits), exact row soft-max + top-k between them on the SIMT cores.
// Reference: QTAttB.process_coarse_level   cuda_imp/QuadTreeAttention/QuadtreeAttention/modules/quadtree_attention.py:161-178
//            QTAttA.process_coarse_level   same file :25-44
// The SIMT kernel of qtatt_coarse.cu spends 1.2 K of its 5.1 K warp-instructions per query row on the two contractions (FFMA2
// out of shared memory) and their tile traffic; here they cost the SM nothing but one elected thread issuing MMAs:
//
//   CTA = up to 64 query rows of one (batch, head), 512 threads, one CTA per SM (the score slab fills shared memory).  Row tiles: 64-row
//             tiles cut back to whole waves of SMs, the rows left over in every (batch, head) split evenly over one more wave of short
//             tiles (tc_row_tiles) -- one image pair is 9 x 64 + 9 x 12 rows per head instead of 176 CTAs on 148 SMs.
//   phase S   warp 0 = TMA producer: Q_hi / Q_lo and EVERY K_hi / K_lo tile at once -- the K operand of a 128-key chunk is loaded
//             into the very slab bytes that chunk's scores will occupy (32 KB either way), so nothing waits for a ring slot.
//             warp 1 = MMA issuer: per chunk 4 k-steps x 3 terms (hi*hi + hi*lo + lo*hi; x_lo = x - trunc_tf32(x) comes from the pooling
//             pass, layout.cu, or coarse_prep_kernel) of UMMA 64x128x8 (kind::tf32) into one of two TMEM accumulators.  warps 4-11 =
//             drain, two warps per TMEM lane quarter taking alternate k-blocks: tcgen05.ld (one accumulator row per thread), * D^-1/2
//             log2(e), 16-byte stores into the SLAB over the K tiles the finished MMAs no longer need: k-blocks of 32 keys, each a
//             [64 rows x 128 B] tile with the 128-byte swizzle.
//   phase T   all 16 warps, one warp per row, two rows at a time interleaved by hand, the row in registers (lane l holds keys 2l, 2l + 1
//             of every 64-key block): max, P = 2^(s - max), sum; exact top-k without a sort of the row: the k-th largest of the 64
//             lane-local top-2 values (two 32-lane bitonic sorts + one max = the upper half of their union) is a threshold T
//             below the row's k-th largest value, the survivors {P >= T} (k <= n, typically n ~ 40) are compacted through a
//             warp prefix sum with straight-line predicated stores and trimmed to exactly k by removing the minimum (an integer
//             REDUX: P >= 0) n - k times.  The k selected keys are emitted in list order; the order torch.topk would give them is
//             restored only where the lists leave the library (topk_to_api_kernel).  P goes back into the row's own slab bytes as an
//             fp16 PAIR: P * 2^14 = hi + lo, the two k-blocks (2 x 8 KB) that held the fp32 scores of 64 keys now hold a
//             [64 rows x 64 keys] fp16 tile of hi and, right behind it, one of lo.
//   phase PV  O = P V over all keys as kind::f16 MMAs (K = 16 per instruction) with BOTH operands stacked: [P_hi; P_lo] is one
//             M = 128 A operand, V^T * 2^8 = hi + lo (fp16, [32 dims x 64 keys] tiles, hi then lo, through a 6-slot TMA ring whose
//             first 4 slots are filled while phases S and T run) one N = 64 B operand: ONE 128x64x16 MMA per k-step yields all four
//             partial products, fp32 accumulation in TMEM.  (First version: P in fp32, TF32 MMAs in two passes over V: 35 K cycles per
//             CTA; three 64x32x16 MMAs per k-step: ~10 K; stacked: ~3 K.)
//   epilogue  warps 4-7: the four blocks of O from TMEM (the P_lo rows cross lane quarters through 8 KB of the idle staging area),
//             / (2^22 * row sum), * level weight, 128-byte row stores.
// Accuracy: every product carries ~2^-21 relative error (error-compensated splits, fp32 accumulation in TMEM), like
// coarse_match.cu; the top-k sets are those of the fp32 reference except at fp32 near-ties (tests/: sets + the tie rule).
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"
#include "tma.cuh"

namespace {

using namespace tma;
constexpr int D = 32;
constexpr int NWARP = 16;
constexpr int LIST_CAP = 64;
constexpr int LIST_KEY_OFF = NWARP * 2 * LIST_CAP * 4;      // bytes from a warp's survivor VALUE list to its KEY list (phase T: values of all warps, then keys)
constexpr int STAGE_AREA = 48 * 1024;           // Q_hi, Q_lo (8 KB each) in phase S, the selection lists (16 KB) in phase T, the V^T ring throughout T / PV

struct CoarseTcMaps { CUtensorMap q_hi, q_lo, k_hi, k_lo, vt_hi, vt_lo; };

__device__ __forceinline__ unsigned key_of(float v) { return v >= 0.f ? __float_as_uint(v) + 1u : 0u; }

__device__ __forceinline__ float ex2_approx(float x) {           // MUFU.EX2; 2^-inf = +0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float warp_min(float v) {
    float m;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
    return m;
}
// bitonic sort of one value per lane; DESC: lane 0 ends with the largest
template <bool DESC>
__device__ __forceinline__ float warp_sort(float v, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const float o = __shfl_xor_sync(FULL_MASK, v, stride);
            const bool up = ((lane & size) == 0) != DESC;        // this block of `size` lanes sorts ascending
            const bool lower = (lane & stride) == 0;
            v = (lower == up) ? fminf(v, o) : fmaxf(v, o);
        }
    }
    return v;
}
// descending sort of a bitonic sequence (one value per lane)
__device__ __forceinline__ float warp_merge_desc(float v, int lane) {
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) {
        const float o = __shfl_xor_sync(FULL_MASK, v, stride);
        v = (lane & stride) == 0 ? fmaxf(v, o) : fminf(v, o);
    }
    return v;
}

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
// K-major operand tile [rows][32 fp32] with 128-byte swizzle: 8-row groups are 1024 B apart (SBO), descriptor version 1 (sm_100)
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
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
// generic-proxy writes to shared memory (the slab) -> visible to the async proxy (the tensor core's operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

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

// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 64
__host__ __device__ constexpr uint32_t idesc_m64(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
}

// The slab, phase S -> T: k-block kb (keys 32 kb .. 32 kb + 31, fp32) is a [64 rows x 128 B] tile, rows in 8-row groups of 1024 B,
// the 16-byte chunk c of row r stored at chunk c ^ (r & 7) (the layout a 128-byte-swizzled K-major UMMA operand has, so that the K
// tiles TMA drops into the same bytes can be used as they are).  Phase T -> PV: the two k-blocks of a 64-key block hold the fp16
// tiles P_hi and P_lo, [64 rows x 64 keys] each, same row / chunk swizzle.  Byte offset of key j of row r inside the P_hi tile:
__device__ __forceinline__ int p16_off(int r, int j) {
    const int t = j & 63;
    return (j >> 6) * 16384 + r * 128 + ((((t >> 3)) ^ (r & 7)) << 4) + (t & 7) * 2;
}
__device__ __forceinline__ void p16_zero(uint8_t *slab, int r, int j) {
    const int o = p16_off(r, j);
    *reinterpret_cast<uint16_t *>(slab + o) = 0;
    *reinterpret_cast<uint16_t *>(slab + o + 8192) = 0;
}

constexpr float P_SCALE = 16384.f, V_SCALE = 256.f;     // fp16 splits: P * 2^14 <= 16384, |V| * 2^8 saturates at 65504 (|V| > 255)

__device__ __forceinline__ uint32_t pack_f16x2_rn(float lo_elem, float hi_elem) {       // {hi_elem, lo_elem} -> half2 bits, round to nearest
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t v) {
    float2 f;
    asm("{\n.reg .f16 l, h;\nmov.b32 {l, h}, %2;\ncvt.f32.f16 %0, l;\ncvt.f32.f16 %1, h;\n}" : "=f"(f.x), "=f"(f.y) : "r"(v));
    return f;
}

// Soft-max + exact top-k of RP slab rows at once, the rows held in registers.  NV values per lane: element i of lane l is key
// 64 (i / 2) + 2 l + (i & 1), i.e. a lane owns two neighbouring keys of every 64-key block.  In: base-2 logits, fp32, k-blocks of
// 32 keys (-inf in the padding columns).  Out: the same bytes hold P * 2^14 as fp16 hi / lo tiles of 64 keys (0 in the padding
// columns), sum[] the row sums of P; the k selected entries of row rr: lanes with keep_a / keep_b set own list entries `lane` /
// `lane + 32` (P value va / vb, key pa / pb), in no particular order.
// TIGHT: the key count fills all NV / 2 blocks of 64 (n_blk == NV / 2, e.g. 676 keys with NV = 22): only the last block can hold
// padding columns, so the per-block "is it there / is it full" tests of the unrolled loops fold away at compile time.
template <int NV, int RP, bool TIGHT>
__device__ __forceinline__ void rows_softmax_topk(uint8_t *slab, const int (&row)[RP], const bool (&live)[RP], int Sk, int n_blk, int k, int lane,
                                                   float *const (&lv)[RP], int *const (&lp)[RP], float (&sum)[RP],
                                                   float (&va)[RP], float (&vb)[RP], int (&pa)[RP], int (&pb)[RP], bool (&keep_a)[RP], bool (&keep_b)[RP],
                                                   long long *dbg = nullptr) {
    constexpr int NB = NV / 2;
#define T_STAMP(i) do { if (dbg && lane == 0) dbg[i] = clock64(); } while (0)
    T_STAMP(8);
    float e[RP][NV];
    float m[RP];
    uint8_t *pbase[RP];
    const int t = 2 * lane;                                       // first of the lane's two keys inside a 64-key block
#pragma unroll
    for (int rr = 0; rr < RP; ++rr) {
        const int r = row[rr], kk = t & 31;
        const float *sbase = reinterpret_cast<const float *>(slab) + (lane >> 4) * (64 * 32) + r * 32 + ((((kk >> 2)) ^ (r & 7)) << 2) + (kk & 3);
        pbase[rr] = slab + r * 128 + ((((t >> 3)) ^ (r & 7)) << 4) + (t & 7) * 2;
        m[rr] = -INFINITY;
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
            float2 v = make_float2(-INFINITY, -INFINITY);
            if (TIGHT || jb < n_blk) v = *reinterpret_cast<const float2 *>(sbase + jb * (2 * 64 * 32));
            e[rr][2 * jb] = v.x; e[rr][2 * jb + 1] = v.y;
            m[rr] = fmaxf(m[rr], fmaxf(v.x, v.y));
        }
    }
#pragma unroll
    for (int rr = 0; rr < RP; ++rr) m[rr] = warp_max(m[rr]);
    __syncwarp();                                                 // every lane holds its part of the rows: the bytes may be overwritten
    T_STAMP(9);
    float m1[RP], m2[RP];
    const int full_blk = Sk >> 6;                                 // 64-key blocks without a padding column
#pragma unroll
    for (int rr = 0; rr < RP; ++rr) {
        sum[rr] = 0.f; m1[rr] = -1.f; m2[rr] = -1.f;
#pragma unroll
        for (int jb = 0; jb < NB; ++jb) {
            float e0 = ex2_approx(e[rr][2 * jb] - m[rr]), e1 = ex2_approx(e[rr][2 * jb + 1] - m[rr]);
            sum[rr] += e0 + e1;
            if (TIGHT || jb < n_blk) {                            // P * 2^14 = hi + lo (fp16 each), packed pairs
                const uint32_t hi = pack_f16x2_rn(e0 * P_SCALE, e1 * P_SCALE);
                const float2 hf = unpack_f16x2(hi);
                const uint32_t lo = pack_f16x2_rn(fmaf(e0, P_SCALE, -hf.x), fmaf(e1, P_SCALE, -hf.y));
                *reinterpret_cast<uint32_t *>(pbase[rr] + jb * 16384) = hi;
                *reinterpret_cast<uint32_t *>(pbase[rr] + jb * 16384 + 8192) = lo;
            }
            if (TIGHT ? (jb == NB - 1 && full_blk < NB) : (jb >= full_blk)) {      // padding never takes part in the selection
                if (64 * jb + t >= Sk) e0 = -1.f;
                if (64 * jb + t + 1 >= Sk) e1 = -1.f;
            }
            e[rr][2 * jb] = e0; e[rr][2 * jb + 1] = e1;
            m2[rr] = fmaxf(m2[rr], fminf(m1[rr], e0));            // running two largest
            m1[rr] = fmaxf(m1[rr], e0);
            m2[rr] = fmaxf(m2[rr], fminf(m1[rr], e1));
            m1[rr] = fmaxf(m1[rr], e1);
        }
    }
#pragma unroll
    for (int rr = 0; rr < RP; ++rr) sum[rr] = warp_sum(sum[rr]);
    T_STAMP(10);
    // T = k-th largest of the 64 lane-top-2 values: max(sorted-descending m1, sorted-ascending m2) is the upper half of their union
    float T[RP];
    {
        float a[RP], b2[RP];
#pragma unroll
        for (int rr = 0; rr < RP; ++rr) { a[rr] = m1[rr]; b2[rr] = m2[rr]; }
#pragma unroll
        for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                const bool lower = (lane & stride) == 0, blk = (lane & size) == 0;
#pragma unroll
                for (int rr = 0; rr < RP; ++rr) {
                    const float oa = __shfl_xor_sync(FULL_MASK, a[rr], stride), ob = __shfl_xor_sync(FULL_MASK, b2[rr], stride);
                    a[rr] = (lower != blk) ? fminf(a[rr], oa) : fmaxf(a[rr], oa);        // descending
                    b2[rr] = (lower == blk) ? fminf(b2[rr], ob) : fmaxf(b2[rr], ob);     // ascending
                }
            }
        }
#pragma unroll
        for (int rr = 0; rr < RP; ++rr) {
            float tt = fmaxf(a[rr], b2[rr]);
            if (k == 32) T[rr] = warp_min(tt);
            else T[rr] = __shfl_sync(FULL_MASK, warp_merge_desc(tt, lane), k - 1);
        }
    }
    T_STAMP(11);
    // survivors {P >= T}: per-lane count, warp prefix sum, compaction into the warp's list
    // (straight-line predicated code: the branchy version of this loop -- if (e >= T) { if (pos < cap) store; ++pos; } -- cost 4.5 K of the
    // 13 K cycles a row pair takes, CASMTR_TC_DEBUG=1)
    // The rows of the pair are interleaved by hand below (counts, scans, stores, trim rounds): each is a dependent chain, and the
    // shuffles / volatile stores pin the order the compiler may issue them in.
    int n[RP], cnt[RP], inc[RP];
#pragma unroll
    for (int rr = 0; rr < RP; ++rr) {
        cnt[rr] = 0;
#pragma unroll
        for (int i = 0; i < NV; ++i) cnt[rr] += e[rr][i] >= T[rr];
        inc[rr] = cnt[rr];
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int rr = 0; rr < RP; ++rr) {
            const int u = __shfl_up_sync(FULL_MASK, inc[rr], o);
            if (lane >= o) inc[rr] += u;
        }
    }
    {
        // lane's survivors go to list entries [inc - cnt, inc): one running shared-memory address per row, the key list LIST_KEY_OFF
        // bytes after the value list.  Lists longer than LIST_CAP are never read (exact slow path below): their threshold becomes +inf.
        uint32_t addr[RP];
        float Ts[RP];
#pragma unroll
        for (int rr = 0; rr < RP; ++rr) {
            n[rr] = __shfl_sync(FULL_MASK, inc[rr], 31);
            addr[rr] = smem_u32(lv[rr]) + 4u * (uint32_t)(inc[rr] - cnt[rr]);
            Ts[rr] = n[rr] <= LIST_CAP ? T[rr] : INFINITY;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int key = 64 * (i >> 1) + t + (i & 1);
#pragma unroll
            for (int rr = 0; rr < RP; ++rr)
                asm volatile(
                    "{\n\t"
                    ".reg .pred q;\n\t"
                    "setp.ge.f32 q, %1, %2;\n\t"
                    "@q st.shared.f32 [%0], %1;\n\t"
                    "@q st.shared.s32 [%0 + %4], %3;\n\t"
                    "@q add.u32 %0, %0, 4;\n\t"
                    "}\n"
                    : "+r"(addr[rr])
                    : "f"(e[rr][i]), "f"(Ts[rr]), "r"(key), "n"(LIST_KEY_OFF)
                    : "memory");
        }
    }
    __syncwarp();
    T_STAMP(12);
    int rem[RP], rounds = 0;                                      // survivors to drop per row (fast path), and the longest of them
#pragma unroll
    for (int rr = 0; rr < RP; ++rr) {
        const bool fast = n[rr] <= LIST_CAP && n[rr] >= k;
        keep_a[rr] = fast && lane < n[rr]; keep_b[rr] = fast && lane + 32 < n[rr];
        va[rr] = keep_a[rr] ? lv[rr][lane] : 0.f; pa[rr] = keep_a[rr] ? lp[rr][lane] : 0;
        vb[rr] = keep_b[rr] ? lv[rr][lane + 32] : 0.f; pb[rr] = keep_b[rr] ? lp[rr][lane + 32] : 0;
        rem[rr] = fast ? n[rr] - k : 0;
        rounds = max(rounds, rem[rr]);
    }
    for (int it = 0; it < rounds; ++it) {                         // drop the smallest survivor (ties: lowest lane, first list half)
#pragma unroll
        for (int rr = 0; rr < RP; ++rr) {
            // P >= 0: the bit patterns order like the values, so the warp minimum is an integer REDUX (and +inf = 0x7f800000 stays on top)
            const unsigned ua = keep_a[rr] ? __float_as_uint(va[rr]) : 0x7f800000u, ub = keep_b[rr] ? __float_as_uint(vb[rr]) : 0x7f800000u;
            const unsigned lm = min(ua, ub);
            const unsigned gm = __reduce_min_sync(FULL_MASK, lm);
            const unsigned first = __ballot_sync(FULL_MASK, lm == gm);
            const bool mine = it < rem[rr] && (first & ((1u << lane) - 1u)) == 0 && lm == gm;      // lowest lane holding the minimum
            const bool drop_a = mine && ua == gm;
            keep_b[rr] = keep_b[rr] && !(mine && !drop_a);
            keep_a[rr] = keep_a[rr] && !drop_a;
        }
    }
#pragma unroll
    for (int rr = 0; rr < RP; ++rr) {
        if (!(n[rr] <= LIST_CAP && n[rr] >= k) && live[rr]) {
            // massive ties (more than LIST_CAP entries share the k-th value): k rounds of warp arg-max over the registers, exact but
            // slow; ties go to the lowest lane, then the lowest element
            for (int it = 0; it < k; ++it) {
                float best = -1.f;
                int bi = 0;
#pragma unroll
                for (int i = 0; i < NV; ++i)
                    if (e[rr][i] > best) { best = e[rr][i]; bi = i; }
                const float g = warp_max(best);
                const int owner = __ffs(__ballot_sync(FULL_MASK, best == g && best >= 0.f)) - 1;
                if (owner < 0) break;
                const int bkey = __shfl_sync(FULL_MASK, 64 * (bi >> 1) + t + (bi & 1), owner);
                if (lane == it) { keep_a[rr] = true; va[rr] = g; pa[rr] = bkey; }
                if (lane == owner) {
#pragma unroll
                    for (int i = 0; i < NV; ++i)
                        if (i == bi) e[rr][i] = -1.f;
                }
            }
        }
    }
    __syncwarp();
    T_STAMP(13);
    if (dbg && lane == 0) { dbg[14] = n[0]; dbg[15] = n[RP - 1]; }
#undef T_STAMP
}

struct TcParams {
    float *acc;                 // [B,Sq,C]
    int *topk_idx;              // [B,Sq,nh,k]
    float *topk_score;
    const float *level_weight;
    float *wsm;
    int levels, n_weights;
    int B, Sq, Sk, nh, topk;
    int n_kb;                   // k-blocks of 32 keys in the slab (even)
    int n_big, rows_small;      // row tiles of a (batch, head): blockIdx.y < n_big covers 64 rows, the tiles after them rows_small (even) each
    long long *dbg;             // CASMTR_TC_DEBUG=1: phase timestamps of tile dbg_tile of (batch 0, head 0) (development aid, NULL otherwise)
    int dbg_tile;
};

constexpr int ROWS = 64;
constexpr int KB_FLOATS = ROWS * 32;            // floats per k-block of the slab
constexpr int SCH = 128;                        // keys per S chunk (UMMA N); the last chunk may hold 64
constexpr int MAX_SCHUNK = 6;                   // n_kb <= 22 -> at most 5 chunks of 4 k-blocks + 1 of 2
constexpr int V_SLOT = 8192;                    // one 64-key block of V^T in fp16: hi 4 KB + lo 4 KB
constexpr int N_VSLOT = STAGE_AREA / V_SLOT;    // 6
constexpr int N_DRAIN = 8;                      // warps 4 .. 11 drain the S accumulators (phase S)
constexpr int V_LEAD = 2;                       // block j lives in slot (j + V_LEAD) % N_VSLOT: slots 0, 1 hold the selection lists during phase T

template <int NV, bool TYPE_A, bool TIGHT>
__global__ void __launch_bounds__(NWARP * 32, 1) qtatt_coarse_tc_kernel(const __grid_constant__ CoarseTcMaps maps, TcParams p) {
    pdl_sync();
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // swizzled tiles want 1024-byte alignment
    float *slab = (float *)sm;
    const int slab_bytes = p.n_kb * KB_FLOATS * 4;
    uint8_t *stage = sm + slab_bytes;                            // 48 KB, 1024-aligned (slab_bytes is a multiple of 8 KB)
    float *rsum = (float *)(stage + STAGE_AREA);                 // [64]
    uint64_t *bars = (uint64_t *)(rsum + 64);
    uint64_t *qfull = bars, *kfull = bars + 1, *tfull = kfull + MAX_SCHUNK, *tempty = tfull + 2;
    uint64_t *vfull = tempty + 2, *vempty = vfull + N_VSLOT, *pvdone = vempty + N_VSLOT;
    uint32_t *tmem_slot = (uint32_t *)(pvdone + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Row tiles: grid.x = (batch, head), grid.y = tile.  CTAs are dispatched x-fastest, so every 64-row tile is handed out before the
    // first short one: the launcher sizes the 64-row tiles to whole waves of SMs and splits the remaining rows of every (batch, head)
    // evenly over one more wave of short tiles (tc_row_tiles) instead of leaving a second wave that is mostly idle SMs.
    const int b = blockIdx.x / p.nh, h = blockIdx.x % p.nh;
    const int tile = blockIdx.y;
    const int row0 = tile < p.n_big ? tile * ROWS : p.n_big * ROWS + (tile - p.n_big) * p.rows_small;
    const int n_rows = min(tile < p.n_big ? ROWS : p.rows_small, p.Sq - row0);     // rows of the 64-row MMA tile this CTA owns
    if (n_rows <= 0) return;
    const int C = p.nh * D;
    const int n_chunks = (p.n_kb + 3) / 4;

    if (tid == 0) {
        mbar_init(qfull, 1);
        for (int s = 0; s < MAX_SCHUNK; ++s) mbar_init(kfull + s, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, N_DRAIN); }
        for (int s = 0; s < N_VSLOT; ++s) { mbar_init(vfull + s, 1); mbar_init(vempty + s, 1); }
        mbar_init(pvdone, 1);
    }
    if (warp == 2 && lane < 6) {                                // hide the descriptor fetch of the six tensor maps
        const CUtensorMap *tm = lane == 0 ? &maps.q_hi : lane == 1 ? &maps.q_lo : lane == 2 ? &maps.k_hi : lane == 3 ? &maps.k_lo : lane == 4 ? &maps.vt_hi : &maps.vt_lo;
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(tm) : "memory");
    }
    if (warp == 1) {                                            // TMEM: 2 S accumulators of 128 columns, O in the 64 columns after them
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_o = tmem_base + 2 * SCH;

    uint8_t *q_hi = stage, *q_lo = stage + 8192;
#define TC_STAMP(i) do { if (p.dbg && tid == 0 && blockIdx.x == 0 && blockIdx.y == p.dbg_tile) p.dbg[i] = clock64(); } while (0)
    TC_STAMP(0);
    // V^T block j (64 keys: hi 4 KB + lo 4 KB, fp16) travels through slot (j + V_LEAD) % N_VSLOT.  A slot's u-th use waits for parity
    // u & 1 (full) / (u & 1) ^ 1 (empty); slots 0 and 1 (the selection lists' bytes during phase T) are first used by blocks 4, 5.
    const int n_blk = p.n_kb / 2;
    auto v_slot = [](int j) { return (j + V_LEAD) % N_VSLOT; };
    auto v_use = [](int j) { const int x = j + V_LEAD; return x / N_VSLOT - (x % N_VSLOT < V_LEAD ? 1 : 0); };
    auto v_load = [&](int j) {
        const int s = v_slot(j);
        uint8_t *st = stage + s * V_SLOT;
        mbar_wait(vempty + s, (v_use(j) & 1) ^ 1);
        mbar_expect_tx(vfull + s, 8192);
        tma_load_3d(st, &maps.vt_hi, 64 * j, h * D, b, vfull + s);
        tma_load_3d(st + 4096, &maps.vt_lo, 64 * j, h * D, b, vfull + s);
    };
    const int n_pre = min(N_VSLOT - V_LEAD, n_blk);              // blocks whose slots are free from the start (not the lists' slots)
    // ================================================================ phase S: scores into the slab
    // The K operand of chunk c is loaded INTO the slab region chunk c's scores will occupy (4 k-blocks = 32 KB = K_hi + K_lo of
    // 128 keys): every chunk is in flight from the first cycle, nothing waits for a ring slot, and the drain overwrites a K tile
    // only after the MMAs that read it have completed (tfull).
    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(qfull, 2 * 8192);
            tma_load_3d(q_hi, &maps.q_hi, h * D, row0, b, qfull);
            tma_load_3d(q_lo, &maps.q_lo, h * D, row0, b, qfull);
            for (int c = 0; c < n_chunks; ++c) {
                const int nk = min(4, p.n_kb - 4 * c);           // 4 or 2 k-blocks
                uint8_t *kh = sm + (size_t)c * 4 * KB_FLOATS * 4, *kl = kh + nk * 4096;
                mbar_expect_tx(kfull + c, nk * 8192);
                for (int j = 0; j < nk / 2; ++j) {               // boxes of 64 keys; stacked they form one [nk * 32 rows x 128 B] tile
                    tma_load_3d(kh + j * 8192, &maps.k_hi, h * D, c * SCH + j * 64, b, kfull + c);
                    tma_load_3d(kl + j * 8192, &maps.k_lo, h * D, c * SCH + j * 64, b, kfull + c);
                }
            }
            for (int j = 0; j < n_pre; ++j) v_load(j);           // the first V^T blocks travel under phases S and T
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait(qfull, 0);
            tc_fence_after();
            const uint64_t ah = umma_desc(q_hi), al = umma_desc(q_lo);
            for (int c = 0; c < n_chunks; ++c) {
                const int a = c & 1;
                const int nk = min(4, p.n_kb - 4 * c);
                mbar_wait(tempty + a, ((c >> 1) & 1) ^ 1);       // the drain warps have emptied this accumulator
                mbar_wait(kfull + c, 0);
                tc_fence_after();
                uint8_t *kh = sm + (size_t)c * 4 * KB_FLOATS * 4;
                const uint64_t bh = umma_desc(kh), bl = umma_desc(kh + nk * 4096);
                const uint32_t d = tmem_base + a * SCH;
                const uint32_t idesc = nk == 4 ? idesc_m64(128) : idesc_m64(64);
#pragma unroll
                for (int k = 0; k < 4; ++k) {                    // UMMA_K = 8 tf32 = 32 bytes: +2 in descriptor address units
                    umma_tf32(d, ah + 2 * k, bh + 2 * k, idesc, k != 0);
                    umma_tf32(d, ah + 2 * k, bl + 2 * k, idesc, 1);
                    umma_tf32(d, al + 2 * k, bh + 2 * k, idesc, 1);
                }
                umma_commit(tfull + a);
            }
        }
    } else if (warp >= 4 && warp < 4 + N_DRAIN) {
        // drain: TMEM lane quarter q holds rows 16 q .. 16 q + 15 of the 64-row accumulator in its first 16 lanes (M = 64 layout).
        // A warp reaches the quarter warp % 4 only, so two warps share each quarter and take alternate k-blocks of a chunk: with
        // one warp per quarter the drain (7 K cycles per tile) was longer than the MMAs it follows (4.6 K).
        const int q = warp & 3, part = (warp - 4) >> 2;
        const int r = 16 * q + lane;                             // accumulator row of this thread (lanes >= 16: none)
        const bool mine = lane < 16;
        const float scale = rsqrtf((float)D) * LOG2E_F;          // logits in base 2: P = 2^(s - max) needs no multiply per element
        for (int c = 0; c < n_chunks; ++c) {
            const int a = c & 1;
            const int nk = min(4, p.n_kb - 4 * c);
            mbar_wait(tfull + a, (c >> 1) & 1);
            tc_fence_after();
            for (int kb = part; kb < nk; kb += N_DRAIN / 4) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + a * SCH + 32 * kb, v);
                const int col0 = (4 * c + kb) * 32;
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= scale;
                if (col0 + 32 > p.Sk) {                          // the last k-blocks: padding columns never win and weigh 0
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (col0 + i >= p.Sk) v[i] = -INFINITY;
                }
                if (mine) {
                    float *dst = slab + (4 * c + kb) * KB_FLOATS + r * 32;
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch)
                        *reinterpret_cast<float4 *>(dst + ((ch ^ (r & 7)) << 2)) = make_float4(v[4 * ch], v[4 * ch + 1], v[4 * ch + 2], v[4 * ch + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + a);
        }
    }
    __syncthreads();                                             // the slab is complete; the staging area is free
    TC_STAMP(1);

    // ================================================================ phase T: soft-max + top-k, one warp per row
    {
        constexpr int RP = 2;
        float *lval = (float *)stage + warp * 2 * LIST_CAP;      // two survivor lists per warp (16 KB in all = V slots 0, 1)
        int *lpos = (int *)(stage + LIST_KEY_OFF) + warp * 2 * LIST_CAP;
        // pass j: warp w takes rows 32 j + 2 w, + 1 -- a short tile keeps as many warps busy as it has row pairs
        for (int j0 = RP * warp; j0 < n_rows; j0 += RP * NWARP) {
            int row[RP];
            bool live[RP];
            float *lvp[RP];
            int *lpp[RP];
            float sum[RP], va[RP], vb[RP];
            int pa[RP], pb[RP];
            bool keep_a[RP], keep_b[RP];
#pragma unroll
            for (int rr = 0; rr < RP; ++rr) {
                row[rr] = j0 + rr;
                live[rr] = row[rr] < n_rows;
                lvp[rr] = lval + rr * LIST_CAP;
                lpp[rr] = lpos + rr * LIST_CAP;
            }
            long long *tdbg = (p.dbg && warp == 0 && j0 == 0 && blockIdx.x == 0 && blockIdx.y == p.dbg_tile) ? p.dbg : nullptr;
            rows_softmax_topk<NV, RP, TIGHT>(sm, row, live, p.Sk, p.n_kb / 2, p.topk, lane, lvp, lpp, sum, va, vb, pa, pb, keep_a, keep_b, tdbg);
#pragma unroll
            for (int rr = 0; rr < RP; ++rr) {
                if (!live[rr]) continue;                     // warp-uniform
                if (lane == 0) rsum[row[rr]] = sum[rr];
                const unsigned ba = __ballot_sync(FULL_MASK, keep_a[rr]), bb = __ballot_sync(FULL_MASK, keep_b[rr]);
                const unsigned lt = (1u << lane) - 1u;
                const size_t o = (((size_t)b * p.Sq + row0 + row[rr]) * p.nh + h) * p.topk;
                const float inv = 1.0f / sum[rr];
                if (keep_a[rr]) {
                    const int slot = __popc(ba & lt);
                    p.topk_idx[o + slot] = pa[rr];
                    p.topk_score[o + slot] = va[rr] * inv;
                    if (TYPE_A) p16_zero(sm, row[rr], pa[rr]);                     // QTAttA: selected keys leave the message (:37-42)
                }
                if (keep_b[rr]) {
                    const int slot = __popc(ba) + __popc(bb & lt);
                    p.topk_idx[o + slot] = pb[rr];
                    p.topk_score[o + slot] = vb[rr] * inv;
                    if (TYPE_A) p16_zero(sm, row[rr], pb[rr]);
                }
            }
            if (tdbg && lane == 0) tdbg[16] = clock64();
        }
    }
    fence_async_smem();                                          // P in the slab -> visible to the tensor core's operand reads
    __syncthreads();
    TC_STAMP(2);

    // ================================================================ phase PV: O = P V (fp16 hi / lo splits of both operands)
    if (warp == 0) {
        if (lane == 0)
            for (int j = n_pre; j < n_blk; ++j) v_load(j);
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D = F32, A = B = F16, both K-major, K = 16 per instruction = 32 bytes.  Both operands are stacked: a V
            // slot holds V_hi^T (32 rows of 128 B) with V_lo^T right behind it = ONE K-major B operand of N = 64 rows, and the P_hi tile of
            // a 64-key block (64 rows of 128 B) has its P_lo tile right behind it = ONE A operand of M = 128 rows.  A single 128 x 64 x 16
            // MMA per k-step therefore yields P_hi V_hi | P_hi V_lo in accumulator rows 0..63 and P_lo V_hi | P_lo V_lo in rows 64..127
            // (48 cycles against 3 x 44 for three 64 x 32 MMAs: small MMAs cost their issue floor); the epilogue adds the four blocks.
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(2 * D >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int j = 0; j < n_blk; ++j) {
                const int s = v_slot(j);
                mbar_wait(vfull + s, v_use(j) & 1);
                tc_fence_after();
                uint8_t *st = stage + s * V_SLOT;
                const uint64_t ph = umma_desc(sm + j * 16384);
                const uint64_t vh = umma_desc(st);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(tmem_o, ph + 2 * k, vh + 2 * k, idesc, (j | k) != 0);
                umma_commit(vempty + s);
            }
            umma_commit(pvdone);
        }
    }

    // ================================================================ epilogue: O / (scales * row sum) * level weight -> acc
    if (warp >= 4 && warp < 8) {
        // M = 128 accumulator: row i in TMEM lane i.  Lane quarters 0, 1 (warps 4, 5) hold the P_hi rows 0..63, quarters 2, 3 (warps 6, 7)
        // the P_lo rows of the same queries: those go through 8 KB of the (now idle) staging area to the warps that own the rows.
        const int q = warp & 3;
        float w0 = 1.f;
        if (p.level_weight) {                           // softmax over the level weights (:264)
            float mx = -INFINITY, den = 0.f;
            for (int l = 0; l < p.n_weights; ++l) mx = fmaxf(mx, __ldg(p.level_weight + l));
            for (int l = 0; l < p.n_weights; ++l) den += expf(__ldg(p.level_weight + l) - mx);
            w0 = expf(__ldg(p.level_weight) - mx) / den;
            if (p.wsm && blockIdx.x == 0 && blockIdx.y == 0 && warp == 4 && lane < p.levels)      // the finer levels read their weight from here (tile 0 always exists)
                p.wsm[lane] = expf(__ldg(p.level_weight + lane) - mx) / den;
        }
        mbar_wait(pvdone, 0);
        tc_fence_after();
        if (warp == 4 && lane == 0 && p.dbg && blockIdx.x == 0 && blockIdx.y == p.dbg_tile) p.dbg[3] = clock64();
        float v[32], v2[32];
        tmem_ld32(tmem_o + ((uint32_t)(32 * q) << 16), v);             // x V_hi
        tmem_ld32(tmem_o + ((uint32_t)(32 * q) << 16) + D, v2);        // x V_lo
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += v2[i];
        const int r = 32 * (q & 1) + lane;                             // query row of this lane (both halves)
        float *xch = reinterpret_cast<float *>(stage) + r * 32;        // [64 rows][32], 16-byte chunks swizzled by the row
        if (q >= 2) {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
                *reinterpret_cast<float4 *>(xch + ((ch ^ (r & 7)) << 2)) = make_float4(v[4 * ch], v[4 * ch + 1], v[4 * ch + 2], v[4 * ch + 3]);
        }
        asm volatile("bar.sync 1, 128;\n" ::: "memory");              // warps 4..7 only
        if (q < 2 && r < n_rows) {
            const float f = w0 / (rsum[r] * (P_SCALE * V_SCALE));
            float *dst = p.acc + ((size_t)b * p.Sq + row0 + r) * C + h * D;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                const float4 t = *reinterpret_cast<const float4 *>(xch + ((ch ^ (r & 7)) << 2));
                *reinterpret_cast<float4 *>(dst + 4 * ch) = make_float4((v[4 * ch] + t.x) * f, (v[4 * ch + 1] + t.y) * f, (v[4 * ch + 2] + t.z) * f, (v[4 * ch + 3] + t.w) * f);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    TC_STAMP(5);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base) : "memory");
    }
}

// ---- operand preparation: x_lo = x - trunc_tf32(x) for Q and K (token-major, like the inputs), and V transposed per batch
// element to channel-major [C][Sp] as an fp16 pair, V * 2^8 = hi + lo: the K-major layout of the second GEMM's B operand.
__global__ void __launch_bounds__(256) coarse_prep_kernel(const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v,
                                                           float *__restrict__ q_lo, float *__restrict__ k_lo, __half *__restrict__ vt_hi,
                                                           __half *__restrict__ vt_lo, int Sq, int Sk, int Sp, int C, int nq4, int nk4) {
    pdl_sync();
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    if (blockIdx.y == 0) {
        // grid-stride residuals of Q and K of this batch element
        const float4 *q4 = reinterpret_cast<const float4 *>(q + (size_t)b * Sq * C);
        const float4 *k4 = reinterpret_cast<const float4 *>(k + (size_t)b * Sk * C);
        float4 *ql4 = reinterpret_cast<float4 *>(q_lo + (size_t)b * Sq * C), *kl4 = reinterpret_cast<float4 *>(k_lo + (size_t)b * Sk * C);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nq4 + nk4; i += gridDim.x * blockDim.x) {
            const bool isq = i < nq4;
            float4 x = isq ? __ldg(q4 + i) : __ldg(k4 + (i - nq4));
            x.x -= __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
            x.y -= __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
            x.z -= __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
            x.w -= __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
            if (isq) ql4[i] = x; else kl4[i - nq4] = x;
        }
        return;
    }
    // transposes: 32 x 32 tiles of V [Sk][C] -> [C][Sp]
    const int tiles_c = C / 32, tiles_s = (Sp + 31) / 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;                     // 32 x 8
    for (int t = blockIdx.x; t < tiles_c * tiles_s; t += gridDim.x) {
        const int ts = t / tiles_c, tc = t - ts * tiles_c;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int s = ts * 32 + ty + 8 * j;
            tile[ty + 8 * j][tx] = s < Sk ? __ldg(v + ((size_t)b * Sk + s) * C + tc * 32 + tx) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = tc * 32 + ty + 8 * j, s = ts * 32 + tx;
            if (s < Sp) {
                const float x = fminf(fmaxf(tile[tx][ty + 8 * j] * V_SCALE, -65504.f), 65504.f);
                const __half hi = __float2half_rn(x);
                const size_t o = ((size_t)b * C + c) * Sp + s;
                vt_hi[o] = hi;
                vt_lo[o] = __float2half_rn(x - __half2float(hi));
            }
        }
    }
}

int make_map3(CUtensorMap *tm, const void *base, int elem_bytes, int d0, int d1, int d2, size_t stride1, size_t stride2, int box0, int box1) {
    EncodeTiledFn enc = encode_tiled();
    CASMTR_REQUIRE(enc != nullptr, CASMTR_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    const cuuint64_t strides[2] = {(cuuint64_t)stride1 * elem_bytes, (cuuint64_t)stride2 * elem_bytes};
    const cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(tm, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CASMTR_REQUIRE(r == CUDA_SUCCESS, CASMTR_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CASMTR_OK;
}

size_t tc_smem_bytes(int n_kb) { return 1024 + (size_t)n_kb * ROWS * 128 + STAGE_AREA + 64 * 4 + 32 * 8 + 16; }

// values per lane of the register-resident row for a key count, or 0 when the tensor-core kernel does not cover it (the slab of a
// 64-row tile must fit shared memory: at most 22 k-blocks = 704 keys; larger levels run the SIMT kernel)
void tc_variant(int Sk, int &nv, int &n_kb) {
    n_kb = ((Sk + 63) / 64) * 2;
    nv = n_kb <= 8 ? 8 : (n_kb <= 16 ? 16 : (n_kb <= 22 ? 22 : 0));
}

// Row tiles of one (batch, head) for a grid of `bh` of them on n_sm SMs (one CTA per SM).  Plain tiling = floor(Sq / 64) tiles of 64
// rows + the remainder.  When that makes 1 - 3 full waves plus a mostly idle one (one or two image pairs per launch), the 64-row
// tiles are cut back to whole waves and the rows left over in every (batch, head) are split evenly over one more wave of short tiles:
// 832^2, one pair: 9 x 64 + 9 x 12 rows per head (a wave of long CTAs, then a wave of CTAs whose row phase is a fifth as long) instead
// of 10 x 64 + 36 = 176 CTAs on 148 SMs.
void tc_row_tiles(int Sq, int bh, int n_sm, int &n_big, int &n_small, int &rows_small) {
    const int n_full = Sq / ROWS, plain = (Sq + ROWS - 1) / ROWS;
    n_big = n_full; n_small = plain - n_full; rows_small = ((Sq - n_full * ROWS) + 1) & ~1;
    static const bool balance = [] { const char *e = getenv("CASMTR_TC_BALANCE"); return !(e && e[0] == '0'); }();
    const long long waves = (long long)bh * plain / n_sm;
    if (!balance || waves < 1 || waves > 3 || (long long)bh * plain % n_sm == 0) return;
    const int nb = (int)std::min<long long>(n_full, waves * n_sm / bh);
    if (nb < 1) return;
    const int left = Sq - nb * ROWS;
    if (left <= 0) { n_big = nb; n_small = 0; rows_small = 0; return; }
    int ns = std::max(1, n_sm / bh);
    int rs = ((left + ns - 1) / ns + 1) & ~1;
    if (rs > ROWS) { ns = (left + ROWS - 1) / ROWS; rs = ((left + ns - 1) / ns + 1) & ~1; }
    n_big = nb; n_small = ns; rows_small = rs;
}

template <int NV>
int launch_variant(const CoarseTcMaps &maps, const TcParams &tp, bool type_a, int grid_x, int grid_y, size_t smem, cudaStream_t stream) {
    static PerDeviceOnce once;
    const int dev = PerDeviceOnce::device();
    if (!once.done(dev)) {
        cudaError_t e = cudaSuccess;
        for (auto kern : {qtatt_coarse_tc_kernel<NV, false, false>, qtatt_coarse_tc_kernel<NV, true, false>, qtatt_coarse_tc_kernel<NV, false, true>,
                          qtatt_coarse_tc_kernel<NV, true, true>})
            if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { casmtr_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return CASMTR_E_CUDA; }
        once.mark(dev);
    }
    LaunchScope ls(CASMTR_K_QT_COARSE, stream);
    const bool tight = tp.n_kb == NV;
    auto kern = type_a ? (tight ? qtatt_coarse_tc_kernel<NV, true, true> : qtatt_coarse_tc_kernel<NV, true, false>)
                       : (tight ? qtatt_coarse_tc_kernel<NV, false, true> : qtatt_coarse_tc_kernel<NV, false, false>);
    launch_k(kern, dim3(grid_x, grid_y), NWARP * 32, smem, stream, maps, tp);
    CASMTR_CHECK_LAUNCH("qtatt_coarse_tc_kernel");
    return CASMTR_OK;
}

}  // namespace

void coarse_tc_row_tiles(int Sq, int bh, int n_sm, int &n_big, int &n_small, int &rows_small) { tc_row_tiles(Sq, bh, n_sm, n_big, n_small, rows_small); }

bool coarse_tc_applicable(int Sq, int Sk, int topk) {
    static const bool env_on = [] { const char *e = getenv("CASMTR_COARSE_TC"); return !(e && e[0] == '0'); }();
    int nv, n_kb;
    tc_variant(Sk, nv, n_kb);
    return env_on && nv != 0 && topk >= 1 && topk <= 32 && Sk >= 64 && Sq >= 64;      // smaller levels: nothing to gain, SIMT kernel
}

// floats of workspace the tensor-core path needs next to the token-major q / k / v of the coarsest level: Q_lo, K_lo (fp32) and the
// two fp16 halves of V^T (rows padded to 8 keys = 16 bytes, a TMA stride requirement)
size_t coarse_tc_workspace_floats(int B, int Sq, int Sk, int C) {
    const size_t Sp = (size_t)(Sk + 7) / 8 * 8;
    return align_up((size_t)B * Sq * C, 64) + align_up((size_t)B * Sk * C, 64) + 2 * align_up((size_t)B * C * Sp / 2, 64);
}

CoarseTcOperands coarse_tc_operands(float *ws, int B, int Sq, int Sk, int C) {
    CoarseTcOperands o;
    o.Sp = (Sk + 7) / 8 * 8;
    o.q_lo = ws;
    o.k_lo = o.q_lo + align_up((size_t)B * Sq * C, 64);
    float *vh = o.k_lo + align_up((size_t)B * Sk * C, 64);
    o.vt_hi = reinterpret_cast<unsigned short *>(vh);
    o.vt_lo = reinterpret_cast<unsigned short *>(vh + align_up((size_t)B * C * o.Sp / 2, 64));
    return o;
}

int launch_qtatt_coarse_tc(const CoarseParams &p, float *ws, cudaStream_t stream) {
    int nv, n_kb;
    tc_variant(p.Sk, nv, n_kb);
    const int rows = ROWS;
    CASMTR_REQUIRE(nv != 0, CASMTR_E_UNSUPPORTED, "coarse level with %d keys is outside the tensor-core kernel", p.Sk);
    CASMTR_REQUIRE((((uintptr_t)p.q | (uintptr_t)p.k | (uintptr_t)p.v | (uintptr_t)ws) & 15) == 0, CASMTR_E_INVALID, "coarse level operands must be 16-byte aligned");
    const int C = p.nh * D;
    const CoarseTcOperands op = coarse_tc_operands(ws, p.B, p.Sq, p.Sk, C);
    const int Sp = op.Sp;
    float *q_lo = op.q_lo, *k_lo = op.k_lo;
    __half *vt_hi = reinterpret_cast<__half *>(op.vt_hi), *vt_lo = reinterpret_cast<__half *>(op.vt_lo);
    if (!p.tc_prepped) {                                         // callers that pooled the pyramid here already wrote them (pool2_tokens_kernel)
        LaunchScope ls(CASMTR_K_LAYOUT, stream);
        const int nq4 = p.Sq * C / 4, nk4 = p.Sk * C / 4;
        const int tiles = (C / 32) * ((Sp + 31) / 32);           // one transposed tile per CTA, the residuals in at most two sweeps
        launch_k(coarse_prep_kernel, dim3((unsigned)std::max(tiles, 48), 2, p.B), 256, 0, stream, p.q, p.k, p.v, q_lo, k_lo, vt_hi, vt_lo, p.Sq, p.Sk, Sp, C, nq4, nk4);
        CASMTR_CHECK_LAUNCH("coarse_prep_kernel");
    }
    CoarseTcMaps maps;
    int rc = make_map3(&maps.q_hi, p.q, 4, C, p.Sq, p.B, C, (size_t)p.Sq * C, 32, 64);
    if (rc == CASMTR_OK) rc = make_map3(&maps.q_lo, q_lo, 4, C, p.Sq, p.B, C, (size_t)p.Sq * C, 32, 64);
    if (rc == CASMTR_OK) rc = make_map3(&maps.k_hi, p.k, 4, C, p.Sk, p.B, C, (size_t)p.Sk * C, 32, 64);
    if (rc == CASMTR_OK) rc = make_map3(&maps.k_lo, k_lo, 4, C, p.Sk, p.B, C, (size_t)p.Sk * C, 32, 64);
    if (rc == CASMTR_OK) rc = make_map3(&maps.vt_hi, vt_hi, 2, p.Sk, C, p.B, Sp, (size_t)C * Sp, 64, 32);       // [32 dims x 64 keys] fp16 = 128-byte rows
    if (rc == CASMTR_OK) rc = make_map3(&maps.vt_lo, vt_lo, 2, p.Sk, C, p.B, Sp, (size_t)C * Sp, 64, 32);
    if (rc != CASMTR_OK) return rc;
    TcParams tp;
    tp.acc = p.acc; tp.topk_idx = p.topk_idx; tp.topk_score = p.topk_score; tp.level_weight = p.level_weight; tp.wsm = p.wsm;
    tp.levels = p.levels; tp.n_weights = p.n_weights; tp.B = p.B; tp.Sq = p.Sq; tp.Sk = p.Sk; tp.nh = p.nh; tp.topk = p.topk; tp.n_kb = n_kb;
    static const bool dbg_on = [] { const char *e = getenv("CASMTR_TC_DEBUG"); return e && e[0] == '1'; }();
    static long long *dbg_buf = nullptr;
    if (dbg_on && !dbg_buf) cudaMalloc(&dbg_buf, 24 * sizeof(long long));
    tp.dbg = dbg_on ? dbg_buf : nullptr;
    const size_t smem = tc_smem_bytes(n_kb);
    CASMTR_REQUIRE(smem <= 227 * 1024, CASMTR_E_UNSUPPORTED, "coarse tensor-core kernel: %zu bytes of shared memory", smem);
    int n_big, n_small, rows_small;
    tc_row_tiles(p.Sq, p.B * p.nh, casmtr_sm_count(), n_big, n_small, rows_small);
    tp.n_big = n_big; tp.rows_small = rows_small;
    if (tp.dbg) { static const int t = [] { const char *e = getenv("CASMTR_TC_DEBUG_TILE"); return e ? atoi(e) : 0; }(); tp.dbg_tile = t; }
    const int gx = p.B * p.nh, gy = n_big + n_small;
    const bool a = p.type_a != 0;
    if (nv == 8) rc = launch_variant<8>(maps, tp, a, gx, gy, smem, stream);
    else if (nv == 16) rc = launch_variant<16>(maps, tp, a, gx, gy, smem, stream);
    else rc = launch_variant<22>(maps, tp, a, gx, gy, smem, stream);
    if (tp.dbg && rc == CASMTR_OK) {                 // development aid: synchronises, never on in the product path
        long long t[24];
        cudaStreamSynchronize(stream);
        cudaMemcpy(t, tp.dbg, sizeof(t), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[tc coarse %dx%d rows=%d] cycles: S %lld  T %lld  PV %lld  epilogue %lld  total %lld\n", p.Sq, p.Sk, rows,
                t[1] - t[0], t[2] - t[1], t[3] - t[2], t[5] - t[3], t[5] - t[0]);
        fprintf(stderr, "[tc coarse phase T, warp 0, first row pair] load+max %lld  exp+pack %lld  threshold %lld  compaction %lld  trim %lld  (survivors %lld, %lld)  [start +%lld after phase S, outputs until +%lld]\n",
                t[9] - t[8], t[10] - t[9], t[11] - t[10], t[12] - t[11], t[13] - t[12], t[14], t[15], t[8] - t[1], t[16] - t[13]);
    }
    return rc;
}
