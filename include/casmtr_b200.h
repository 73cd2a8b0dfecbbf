/*
 * casmtr_b200.h -- C ABI of libcasmtr_b200.so: the B200 (sm_100a) replacement for the three
 * CUDA extensions of ewrfcas/CasMTR and for the torch op sequences around them on the
 * coarse-to-fine matching hot path.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; all tensors are dense,
 *     row-major, fp32 / int64 exactly as the reference's Python modules hand them over;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); every
 *     launch goes to that stream (or, inside casmtr_qtatt_fwd, to a side stream forked from and
 *     joined back into it -- see casmtr_set_overlap), nothing synchronises;
 *   - return value 0 = success, <0 = error (see CASMTR_E_*); the message of the last error
 *     on the calling thread is returned by casmtr_last_error_string();
 *   - no allocation inside the library: fused entry points take a caller-owned workspace
 *     whose size the matching *_workspace_bytes() function reports.
 *
 * Reference interfaces replaced (paths relative to the reference tree):
 *   R1 cuda_imp/QuadTreeAttention/QuadtreeAttention/src/score_computation.cpp:11-20,35-38
 *      (pybind `score_computation_cuda.score_forward`)         -> casmtr_score5d_fwd
 *   R2 cuda_imp/QuadTreeAttention/QuadtreeAttention/src/value_aggregation.cpp:9-31,62-65
 *      (pybind `value_aggregation_cuda.value_aggregation_forward`) -> casmtr_value_agg_fwd
 *   R3 cuda_imp/score_cuda/src/score_computation.cpp:11-17,29-32
 *      (pybind `fast_score_computation.score_forward`)         -> casmtr_score3d_fwd
 *   R4 cuda_imp/QuadTreeAttention/QuadtreeAttention/modules/quadtree_attention.py:8-140 (QTAttA.forward),
 *      :144-286 (QTAttB.forward)                               -> casmtr_qtatt_fwd
 *   R5 same file :392-452 (CascadeQTAttB.forward)               -> casmtr_cascade_qtatt_fwd
 *   R6 src/model/functions/cascade_matching.py:87-149 (CascadeMatching.forward, inference)
 *                                                               -> casmtr_cascade_match_fwd
 *   R7 src/model/functions/cascade_matching.py:170-261,316-331 (get_coarse_match, inference) +
 *      src/model/functions/post_processing.py:41-44,111-121 + cascade_functions.py:120-172
 *                                                               -> casmtr_match_extract
 *   R8 src/model/functions/fine_matching.py:77-137, 201-261     -> casmtr_fine_match_fwd
 */
#ifndef CASMTR_B200_H
#define CASMTR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CASMTR_VERSION 200          /* 0.2.0: casmtr_qtatt_desc grew flags / concurrent_calls; casmtr_set_concurrency removed */
#define CASMTR_MAX_LEVELS 4

#define CASMTR_OK 0
#define CASMTR_E_INVALID (-1)       /* bad argument (null pointer, non-positive size, odd grid ...) */
#define CASMTR_E_UNSUPPORTED (-2)   /* configuration outside what the kernels implement */
#define CASMTR_E_CUDA (-3)          /* a CUDA runtime call or launch failed */
#define CASMTR_E_WORKSPACE (-4)     /* workspace too small */

typedef void *casmtr_stream_t;      /* cudaStream_t */

#if defined(__GNUC__)
#define CASMTR_API __attribute__((visibility("default")))
#else
#define CASMTR_API
#endif

CASMTR_API int casmtr_version(void);
CASMTR_API const char *casmtr_last_error_string(void);
/* SM count / L2 bytes of the current device (host-side helper for the bench harness). */
CASMTR_API int casmtr_device_info(int *sm_count, size_t *l2_bytes);

/* ---------------------------------------------------------------- launch accounting / kernel timing
 * Kernel kinds, for the per-kernel breakdown of the bench harness (the reference's analogue is the
 * named-region InferenceProfiler, src/utils/profiler.py:8-40). */
enum {
    CASMTR_K_LAYOUT = 0,        /* NCHW -> token-major transposes, top-k list export */
    CASMTR_K_QT_COARSE = 1,     /* dense coarsest quadtree level */
    CASMTR_K_QT_FINE_MID = 2,   /* intermediate quadtree levels (emit top-k) */
    CASMTR_K_QT_FINE_LAST = 3,  /* finest quadtree level (merged message only) */
    CASMTR_K_CASCADE_ATT = 4,   /* CascadeQTAttB window attention (TMA-tiled, or the gather kernel when not tileable) */
    CASMTR_K_CASCADE_MATCH = 5, /* fused correlation + softmax + argmax */
    CASMTR_K_EXTRACT = 6,       /* NMS / gates / scan / ordered emit */
    CASMTR_K_FINE_MATCH = 7,
    CASMTR_K_OPS = 8,           /* op-level drop-ins (score5d / value_agg / score3d) */
    CASMTR_K_CASCADE_FALLBACK = 9, /* gather kernel over the cells the TMA-tiled cascade kernels could not serve */
    CASMTR_K_COARSE_MATCH = 10, /* dense dual-softmax statistics on tcgen05 (SURVEY section 8f "next" #1) */
    CASMTR_K_COUNT = 11
};
/* Total number of kernels this library has launched in this process (all threads). */
CASMTR_API uint64_t casmtr_launch_count(void);
/* on != 0: bracket every subsequent kernel launch with a CUDA event pair on its launch stream. */
CASMTR_API int casmtr_profile_enable(int on);
/* Synchronises the recorded events, ADDS per-kind device milliseconds / launch counts into the two
 * HOST arrays (CASMTR_K_COUNT entries each; either may be NULL) and clears the record list. */
CASMTR_API int casmtr_profile_collect(double *ms_by_kind, uint64_t *launches_by_kind);
CASMTR_API const char *casmtr_kernel_kind_name(int kind);
/* Host-only introspection of two launch-time computations (no device needed; used by the CPU tests):
 *  casmtr_plan_dense_tiles: the row tiling the tensor-core coarsest level uses for `rows` query rows per (batch, head), `bh` of them, on
 *    `n_sm` SMs: out[0] tiles of 64 rows, then out[1] tiles of out[2] rows each (every row is covered exactly once, in order).
 *  casmtr_fastdiv: the multiply-high / shift pair the kernels divide by `d` with: n / d == umulhi(n, out[0]) >> out[1] for 0 <= n < 2^31
 *    (out[0] == 0: d == 1). */
CASMTR_API int casmtr_plan_dense_tiles(int rows, int bh, int n_sm, int out[3]);
CASMTR_API int casmtr_fastdiv(int d, unsigned out[2]);

/* ---------------------------------------------------------------- op-level drop-ins (R1-R3) */

/* out[b,n,f,k,h] = sum_d query[b,n,f,h,d] * key[b, index[b,n,k,h], h, d]
 * query [B,N1,4,H,D], key [B,N2,H,D], index [B,N1,K,H] int64, out [B,N1,4,K,H] (fully written).
 * Lifts the reference's limits H<=8 and N1<=65535; index values are clamped to [0,N2-1]. */
CASMTR_API int casmtr_score5d_fwd(const float *query, const float *key, const int64_t *index, float *out,
                       int B, int N1, int N2, int H, int D, int K, casmtr_stream_t stream);

/* out[b,n,h,d] = sum_k score[b,n,k,h] * value[b, index[b,n,k,h], h, d]
 * score/index [B,N,K,H], value [B,M,H,D], out [B,N,H,D] (overwritten, need not be zeroed). */
CASMTR_API int casmtr_value_agg_fwd(const float *score, const float *value, const int64_t *index, float *out,
                         int B, int N, int K, int H, int M, int D, casmtr_stream_t stream);

/* out[b,n,k] = sum_c query[b,n,c] * key[b, index[b,n,k], c]
 * query [B,N1,C], key [B,N2,C], index [B,N1,K] int64, out [B,N1,K].  C % 4 == 0. */
CASMTR_API int casmtr_score3d_fwd(const float *query, const float *key, const int64_t *index, float *out,
                       int B, int N1, int N2, int C, int K, casmtr_stream_t stream);

/* Backward halves of the three ops (SURVEY 8f "next" #4; reference score_computation.cpp:22-33 `score_backward`,
 * value_aggregation.cpp:33-60 `value_aggregation_backward`, score_cuda/src/score_computation.cpp:19-27 `score_backward`), so
 * that the reference's autograd Functions train on this library.  Tensor contracts as in the forward calls.  grad_query /
 * grad_score are fully written (no atomics); grad_key / grad_value are zeroed here and then accumulated with 16-byte vector
 * reductions -- reproducible to fp32 rounding of the sum, not bit for bit (the reference's scalar atomicAdd is the same).
 *   score5d:   grad_out [B,N1,4,K,H] -> grad_query [B,N1,4,H,D], grad_key [B,N2,H,D]        D % 4 == 0, H*D <= 1024
 *   value_agg: grad_out [B,N,H,D]    -> grad_score [B,N,K,H],   grad_value [B,M,H,D]        D in {4,..,128} power of two
 *   score3d:   grad_out [B,N1,K]     -> grad_query [B,N1,C],    grad_key [B,N2,C]           C % 4 == 0, C <= 512 */
CASMTR_API int casmtr_score5d_bwd(const float *grad_out, const float *query, const float *key, const int64_t *index,
                       float *grad_query, float *grad_key, int B, int N1, int N2, int H, int D, int K, casmtr_stream_t stream);
CASMTR_API int casmtr_value_agg_bwd(const float *grad_out, const float *score, const float *value, const int64_t *index,
                         float *grad_score, float *grad_value, int B, int N, int K, int H, int M, int D, casmtr_stream_t stream);
CASMTR_API int casmtr_score3d_bwd(const float *grad_out, const float *query, const float *key, const int64_t *index,
                       float *grad_query, float *grad_key, int B, int N1, int N2, int C, int K, casmtr_stream_t stream);

/* NCHW -> token-major: src [B,C,HW] -> dst [B,HW,C] (exported for tests / callers that want
 * to keep features token-major between layers). */
CASMTR_API int casmtr_nchw_to_tokens(const float *src, float *dst, int B, int C, int HW, casmtr_stream_t stream);

/* ---------------------------------------------------------------- fused QuadTree attention (R4) */

typedef struct {
    int B;                          /* batch */
    int nhead;                      /* heads; channels C = nhead * D */
    int D;                          /* head dim; the fused kernels implement D == 32 */
    int levels;                     /* pyramid depth, 1..CASMTR_MAX_LEVELS */
    int type;                       /* 0 = QTAttB, 1 = QTAttA */
    int qh[CASMTR_MAX_LEVELS];      /* query grid per level, [0] = FINEST (the reference's list order) */
    int qw[CASMTR_MAX_LEVELS];
    int kh[CASMTR_MAX_LEVELS];      /* key/value grid per level */
    int kw[CASMTR_MAX_LEVELS];
    int topks[CASMTR_MAX_LEVELS];   /* the reference's `topks`: [0] = coarsest level */
    int flags;                      /* CASMTR_QT_* bits, 0 = defaults */
    int weight_len;                 /* entries of level_weight (QTAttB.weight has `scale` of them, reference :159); the soft-max runs
                                     * over ALL of them (:264) even when the pyramid is shorter.  0 = `levels`. */
    int concurrent_calls;           /* launch-geometry hint: independent calls of this shape the caller keeps in flight at once
                                     * (other streams / graph branches); 0 or 1 = this call runs alone.  Results do not depend on it.
                                     * Callers that stack the two directions of a layer on the batch dimension (B = 2 x pairs) need
                                     * no hint: every kernel sizes its grid from B. */
} casmtr_qtatt_desc;

#define CASMTR_QT_NO_OVERLAP 1      /* casmtr_qtatt_fwd: keep the transposes of the finer levels on the caller's stream */
#define CASMTR_QT_SIMT_COARSE 2     /* dense coarsest level on the fp32 SIMT kernel instead of the tcgen05 one (A/B, parity tests) */

CASMTR_API size_t casmtr_qtatt_workspace_bytes(const casmtr_qtatt_desc *desc);

/* queries/keys/values: HOST arrays of `levels` device pointers, [l] = [B,C,h_l,w_l] NCHW fp32,
 * l = 0 finest (exactly the lists QTAttB.forward receives).
 * level_weight: device [weight_len], the raw `weight` parameter of QTAttB (the softmax over its entries is
 *   applied inside, reference :264; level i of the processing order uses entry i); ignored (may be NULL) for type A.
 * out: [B, qh[0]*qw[0], nhead, D].
 * topk_idx_out / topk_score_out: optional HOST arrays (or NULL) of `levels` device pointers in
 *   the reference's processing order ([0] = coarsest); entry i, if non-NULL, receives the
 *   level's top-k key indices [B, L_i, topks[i], nhead] int64 (raster order, as the reference
 *   holds them after :226-227) resp. their scores (fp32).  The last level's entry is ignored
 *   (the reference computes and discards it). */
CASMTR_API int casmtr_qtatt_fwd(const casmtr_qtatt_desc *desc,
                     const float *const *queries, const float *const *keys, const float *const *values,
                     const float *level_weight, float *out,
                     int64_t *const *topk_idx_out, float *const *topk_score_out,
                     void *workspace, size_t workspace_bytes, casmtr_stream_t stream);

/* Token-major entry (SURVEY 8f "next" #2; src/model/modules/quadtree_attention.py:68-99): q0 / k0 / v0 are the FINEST level
 * only, token-major [B, h*w, C] fp32 (what a 1x1 convolution is when applied as a linear layer to the [B,N,C] tokens
 * QuadtreeAttention.forward receives).  The avg_pool2d(2,2) pyramid (:86-89) is built inside, in the layout the kernels
 * read, so neither the NCHW maps, nor their transposes, nor the caller's pooling launches exist.  desc as above
 * (qh[l] = qh[l-1] / 2 ...); same workspace size; other arguments as casmtr_qtatt_fwd. */
CASMTR_API int casmtr_qtatt_tokens_fwd(const casmtr_qtatt_desc *desc, const float *q0, const float *k0, const float *v0,
                     const float *level_weight, float *out,
                     int64_t *const *topk_idx_out, float *const *topk_score_out,
                     void *workspace, size_t workspace_bytes, casmtr_stream_t stream);

/* One guided quadtree level (QTAttGuided.process_fine_level + merge, reference quadtree_attention.py:298-389; reachable through
 * SELF_ATTN_TYPE = 'topk'): the 4 children of every query cell attend to the 4 children of each of K externally supplied key cells.
 *   query [B,C,h0,w0], key / value [B,C,h1,w1] NCHW fp32; topk_pos [2,B,(h0/2)*(w0/2),K,nhead] int64 = (row, col) of the key cells on
 *   the (h1/2 x w1/2) grid, per head (clamped into the grid; the reference reads out of bounds otherwise);
 *   level_weight device [weight_len] = the module's raw `weight`: the message is scaled by softmax(weight)[0] (:375-380);
 *   out [B, h0*w0, nhead, D] in RASTER order of the query grid.
 * The reference's own merge (:385) reshapes with queries[-0] = the only level's height instead of half of it, which permutes the
 * tokens; the Python module applies that permutation on top of this entry point's raster output to stay a drop-in.
 * K <= 32, D == 32.  Workspace: casmtr_qtatt_guided_workspace_bytes. */
CASMTR_API size_t casmtr_qtatt_guided_workspace_bytes(int B, int C, int nhead, int h0, int w0, int h1, int w1, int K);
CASMTR_API int casmtr_qtatt_guided_fwd(const float *query, const float *key, const float *value, const int64_t *topk_pos,
                            const float *level_weight, int weight_len, float *out,
                            int B, int nhead, int D, int h0, int w0, int h1, int w1, int K,
                            void *workspace, size_t workspace_bytes, casmtr_stream_t stream);

/* Launch options.  Both are settings of the CALLING THREAD (thread-local, default on), not of the process: a second host thread
 * is unaffected, and a stream capture records what the capturing thread selected.  Per-call options live in the descriptors
 * (casmtr_qtatt_desc.flags / .concurrent_calls).
 *
 * Programmatic dependent launch (CASMTR_PDL=0 in the environment makes off the default): every hot-path kernel is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs are scheduled while the previous kernel of the stream drains
 * and wait (griddepcontrol.wait) for its completion before touching memory.  Helps back-to-back launches on ONE stream; turn it
 * off when independent calls already overlap on several streams.  Returns the previous setting of this thread. */
CASMTR_API int casmtr_set_pdl(int on);

/* Layout / compute overlap inside casmtr_qtatt_fwd (CASMTR_OVERLAP=0 in the environment makes off the default): the call forks the
 * NCHW -> token-major transposes of all but the coarsest pyramid level onto a library-owned side stream and joins them back
 * before the first fine level, so they run under the coarsest level's kernel.  The call stays fully ordered with respect to the
 * caller's stream (event fork after the caller's prior work, event join before the call's later kernels and anything the caller
 * enqueues afterwards) and can be stream-captured; concurrent callers on one device are serialised while they enqueue.  Turning
 * it on creates the side streams (2 per device) if they do not exist yet -- do that outside a capture.  CASMTR_QT_NO_OVERLAP in
 * a descriptor switches it off for that call.  Returns the previous setting of this thread. */
CASMTR_API int casmtr_set_overlap(int on);

/* ---------------------------------------------------------------- fused cascade window attention (R5) */

CASMTR_API size_t casmtr_cascade_qtatt_workspace_bytes(int B, int C, int h0, int w0, int h1, int w1);

/* query [B,C,h0,w0], key/value [B,C,h1,w1] NCHW; topk_pos [B,(h0/2)*(w0/2),k,2] int64 (row,col at
 * the previous level); rel_pos NULL or [B,nhead,h0*w0,4k]; message [B,h0*w0,C];
 * upsampled_idx NULL or [B,h0*w0,4k] int64.  k <= 32, D == 32. */
CASMTR_API int casmtr_cascade_qtatt_fwd(const float *query, const float *key, const float *value,
                             const int64_t *topk_pos, const float *rel_pos,
                             float *message, int64_t *upsampled_idx,
                             int B, int nhead, int D, int h0, int w0, int h1, int w1, int k, int dilated,
                             void *workspace, size_t workspace_bytes, casmtr_stream_t stream);

/* Same with token-major query [B,h0*w0,C] / key, value [B,h1*w1,C] (CascadeQuadtreeAttention.forward,
 * src/model/modules/quadtree_attention.py:152-171, with the 1x1 convolutions applied as linear layers): no transposes. */
CASMTR_API int casmtr_cascade_qtatt_tokens_fwd(const float *query, const float *key, const float *value,
                             const int64_t *topk_pos, const float *rel_pos,
                             float *message, int64_t *upsampled_idx,
                             int B, int nhead, int D, int h0, int w0, int h1, int w1, int k, int dilated,
                             void *workspace, size_t workspace_bytes, casmtr_stream_t stream);

/* Window-index plumbing between the stages (SURVEY 8f "next" #3; src/model/modules/transformer.py:416-440,
 * CascadeFeatureTransformer.get_window_warp_idx): next_idx [B,L] int64 on an H x W grid -> pos [B,L,window*window,2] int64
 * (row, col) of the window around each match, shifted rigidly inside the grid.  window odd, <= min(H, W). */
CASMTR_API int casmtr_window_idx_fwd(const int64_t *next_idx, int64_t *pos, int B, int L, int H, int W, int window,
                          casmtr_stream_t stream);

/* The fusion of the two: CascadeQTAttB fed with next_idx [B,(h0/2)*(w0/2)] int64 (each parent cell's match on the (h1/2 x w1/2)
 * grid) instead of the expanded topk_pos; the window is derived inside the kernels, so the [B,L/4,25,2] tensor never exists
 * and upsampled_idx (optional) is the only index tensor written.  query/key/value NCHW (token_major = 0) or token-major
 * [B,h*w,C] (token_major = 1).  window in {1,3,5}.  Workspace as casmtr_cascade_qtatt_workspace_bytes. */
CASMTR_API int casmtr_cascade_qtatt_window_fwd(const float *query, const float *key, const float *value,
                             const int64_t *next_idx, int window, const float *rel_pos,
                             float *message, int64_t *upsampled_idx,
                             int B, int nhead, int D, int h0, int w0, int h1, int w1, int token_major,
                             void *workspace, size_t workspace_bytes, casmtr_stream_t stream);

/* Relative position bias of the cascade cross attention (SURVEY 8f "next" #3; CascadeFeatureTransformer.get_relative_pe,
 * src/model/modules/transformer.py:473-509; tables and LB from its constructor :356-362; indoor config only).  For query
 * token (Y, X) of the current (h0 x w0) level and key token (ky, kx) of the other image's current level
 *     t  = tgt_idx[b, (Y / s) * w8 + X / s],  s = h0 / h8           (the query's 1/8 match on the other image)
 *     rx = X % s - ((t % w8_other) * s + s/2 - 1) + kx + LB,   ry likewise with Y, t / w8_other, ky
 *     bias[b, head, token, candidate] = w_table[rx, head] + h_table[ry, head].
 * The reference raises on a table index outside [0, n_emb); here it is clamped into the table. */
typedef struct casmtr_relpe_desc {
    const float *w_table;       /* w_pos_bias.weight [n_emb, nhead] */
    const float *h_table;       /* h_pos_bias.weight [n_emb, nhead] */
    const int64_t *tgt_idx;     /* [B, h8*w8] int64: data['stage_8c']['next_idx_c01' / 'next_idx_c10'] */
    int n_emb;                  /* table rows = 2 * LB + sr_ratio */
    int LB;
    int h8, w8;                 /* 1/8 grid of the query image (h0 % h8 == 0, w0 == w8 * (h0 / h8)) */
    int w8_other;               /* row length of the other image's 1/8 grid (w1 == w8_other * (h0 / h8)) */
} casmtr_relpe_desc;

/* Stand-alone drop-in for get_relative_pe: window_pos [B,(h0/2)*(w0/2),k,2] int64 (row, col on the previous level of the
 * other image, the first output of get_window_warp_idx) -> rel_pos [B,nhead,h0*w0,4k] fp32, candidate order as in
 * CascadeQTAttB (window entry major, then the 2x2 children row-major). */
CASMTR_API int casmtr_relative_pe_fwd(const casmtr_relpe_desc *pe, const int64_t *window_pos, float *rel_pos,
                           int B, int nhead, int h0, int w0, int k, casmtr_stream_t stream);

/* CascadeQTAttB with the bias computed inside the attention kernels from the two embedding tables: the
 * [B,nhead,h0*w0,4k] tensor (30.7 MB per pair and direction at 640x480) and the ~25 torch ops that build it never exist.
 * Give topk_pos [B,(h0/2)*(w0/2),k,2] (next_idx NULL) or next_idx [B,(h0/2)*(w0/2)] with window in {1,3,5}
 * (topk_pos NULL, k = window^2).  token_major as in casmtr_cascade_qtatt_window_fwd; dilated is 1 (the reference's
 * get_relative_pe ignores dilation, transformer.py:489-493).  Workspace as casmtr_cascade_qtatt_workspace_bytes. */
CASMTR_API int casmtr_cascade_qtatt_relpe_fwd(const float *query, const float *key, const float *value,
                             const int64_t *topk_pos, const int64_t *next_idx, int window, const casmtr_relpe_desc *pe,
                             float *message, int64_t *upsampled_idx,
                             int B, int nhead, int D, int h0, int w0, int h1, int w1, int k, int token_major,
                             void *workspace, size_t workspace_bytes, casmtr_stream_t stream);

/* ---------------------------------------------------------------- fused cascade matching (R6) */

/* feat0 [B,L0,C], feat1 [B,L1,C] (un-normalised; the 1/sqrt(C) of reference :88 is applied inside);
 * idx01 [B,L0,K], idx10 [B,L1,K] int64; mask0 [B,L0] / mask1 [B,L1] uint8 (both or neither NULL);
 * conf01 / conf10: NULL or [B,L,K] softmax over the K candidates; next_conf* [B,L] fp32 (row max);
 * next_idx* [B,L] int64 = idx[b,i,argmax].  K <= 128, C % 4 == 0.
 * w0 / w1: row length of the image-0 / image-1 token grids, or 0 if unknown.  With even grids the kernel
 * processes the 2x2 sibling tokens of a parent cell together and, when their candidate lists are identical
 * (they are when idx comes from CascadeQTAttB's upsampled_idx), reads every candidate row once for all four;
 * results do not depend on w0 / w1.
 * workspace: casmtr_cascade_match_workspace_bytes() bytes, or NULL.  With a workspace, K == 100, no masks and 16-byte
 * aligned inputs, cells whose lists are the regular 10x10 window of a CascadeQTAttB upsampled_idx are served by the
 * TMA-tiled kernel (match_tile.cu); all other cells, and every call without a workspace, by the gather kernels. */
/* Bytes of workspace casmtr_cascade_match_fwd wants (the tile kernel's fallback cell list). */
CASMTR_API size_t casmtr_cascade_match_workspace_bytes(int B, int L0, int L1);

CASMTR_API int casmtr_cascade_match_fwd(const float *feat0, const float *feat1,
                             const int64_t *idx01, const int64_t *idx10,
                             const uint8_t *mask0, const uint8_t *mask1, float temperature,
                             float *conf01, float *next_conf01, int64_t *next_idx01,
                             float *conf10, float *next_conf10, int64_t *next_idx10,
                             int B, int L0, int L1, int C, int K, int w0, int w1,
                             void *workspace, size_t workspace_bytes, casmtr_stream_t stream);

/* ---------------------------------------------------------------- dense coarse matching statistics
 * Reference: src/model/functions/coarse_matching.py:60-75.  feat0 [B,L0,C], feat1 [B,L1,C] fp32 (un-normalised; the
 * 1/sqrt(C) of :61 is applied inside), sim = <f0,f1> / (C * temperature).  Outputs, without ever forming the L0 x L1 matrix:
 *   next_conf01 [B,L0] = max_j softmax_j(sim)[i,:],  next_idx01 [B,L0] = argmax_j;  next_conf10 / next_idx10 [B,L1] likewise
 *   over i.  GEMM on tcgen05 (kind::tf32, 3-term split for fp32 accuracy), softmax statistics in the epilogue.  C % 32 == 0. */
CASMTR_API size_t casmtr_coarse_match_workspace_bytes(int B, int L0, int L1, int C);
CASMTR_API int casmtr_coarse_match_fwd(const float *feat0, const float *feat1, float temperature,
                            float *next_conf01, int64_t *next_idx01, float *next_conf10, int64_t *next_idx10,
                            int B, int L0, int L1, int C, void *workspace, size_t workspace_bytes, casmtr_stream_t stream);
/* With the padding masks of the two images (reference :64-65, sim.masked_fill_(~(mask0[:, :, None] * mask1[:, None]), -INF)):
 * mask0 [B,L0], mask1 [B,L1] uint8 (the reference's bool tensors; both or neither NULL).  Padded columns take no part in the
 * row soft-max / arg-max (exp(-1e9 - max) is exactly 0 in fp32); a padded row is constant (-INF is -1e9 there, :6), so its
 * soft-max is uniform: next_conf = 1 / columns, next_idx = 0, as torch.max returns.  Same workspace. */
CASMTR_API int casmtr_coarse_match_masked_fwd(const float *feat0, const float *feat1, const uint8_t *mask0, const uint8_t *mask1,
                            float temperature,
                            float *next_conf01, int64_t *next_idx01, float *next_conf10, int64_t *next_idx10,
                            int B, int L0, int L1, int C, void *workspace, size_t workspace_bytes, casmtr_stream_t stream);

/* The same plus what CoarseMatching.get_coarse_match needs (reference :91-153; the `temp_outputs` of a stage-1 model read its match
 * list, cascade_model_stage3.py:71-74,146-147): the row maxima / arg-maxima of conf = softmax_i(sim) * softmax_j(sim) (:68) and its
 * column arg-maxima, from a second tcgen05 pass over the similarity (2 sim - lse_col reduced over every row), again without the matrix:
 *   mconf_row [B,L0] = max_j conf[i,j],  midx_row [B,L0] = argmax_j conf[i,j],  midx_col [B,L1] = argmax_i conf[i,j]  (int64; -1 for a
 *   fully padded column).  (i, j) is a mutual nearest neighbour (:122) iff midx_row[i] == j and midx_col[j] == i; feed the three arrays to
 *   casmtr_match_extract with coarse_mode = 1, test_thr = thr, double_check = 1 for the ordered match list.  Same workspace. */
CASMTR_API int casmtr_coarse_match_mutual_fwd(const float *feat0, const float *feat1, const uint8_t *mask0, const uint8_t *mask1,
                            float temperature,
                            float *next_conf01, int64_t *next_idx01, float *next_conf10, int64_t *next_idx10,
                            float *mconf_row, int64_t *midx_row, int64_t *midx_col,
                            int B, int L0, int L1, int C, void *workspace, size_t workspace_bytes, casmtr_stream_t stream);

/* ---------------------------------------------------------------- NMS + match extraction (R7) */

typedef struct {
    int B, h0, w0, h1, w1;          /* source / target grids of this stage */
    int nms_window;                 /* 0 = plain threshold (post_processing.py:43-44); odd k = maxpool NMS */
    float test_thr;
    int border_rm;
    int double_check;
    int n_pre;                      /* previous-stage confidence gates, 0..2 (cascade_matching.py:199-206) */
    const float *pre_conf[2];       /* device [B, pre_h*pre_w] */
    int pre_h[2], pre_w[2];
    float pre_thr[2];
    const uint8_t *pad_mask0;       /* NULL, or device [B,h0,w0]: padded border variant (cascade_functions.py:142-172) */
    const uint8_t *pad_mask1;       /* NULL, or device [B,h1,w1] */
    float scale;                    /* hw0_i[0] / h0 (cascade_matching.py:317) */
    const float *scale0;            /* NULL or device [B,2] */
    const float *scale1;            /* NULL or device [B,2] */
    int coarse_mode;                /* 0: CascadeMatching.get_coarse_match.  1: CoarseMatching.get_coarse_match (coarse_matching.py:
                                     * 91-153): the border is removed symmetrically on the target grid as well (row / col >= size - b,
                                     * mask_border, cascade_functions.py:82-100, where the cascade helper tests > size - b) and an empty
                                     * result stays empty (no "keep element 0" fallback) */
} casmtr_extract_desc;

CASMTR_API size_t casmtr_match_extract_workspace_bytes(const casmtr_extract_desc *desc);

/* next_conf01 [B,L0] fp32, next_idx01 [B,L0] int64, next_idx10 [B,L1] int64.
 * Outputs (capacity entries each; capacity >= B, B*L0 always suffices): b_ids,i_ids,j_ids int64,
 * mconf fp32, mkpts0/mkpts1 [capacity,2] fp32 in torch.where order; mask_out NULL or [B,L0] uint8 (the
 * keep flags before the reference's "if mask.sum() == 0: mask[:, 0] = True" fallback, which only affects the list);
 * count_out: device int32, number of matches (entries beyond capacity are dropped, count is not clamped). */
CASMTR_API int casmtr_match_extract(const casmtr_extract_desc *desc,
                         const float *next_conf01, const int64_t *next_idx01, const int64_t *next_idx10,
                         uint8_t *mask_out, int64_t *b_ids, int64_t *i_ids, int64_t *j_ids,
                         float *mconf, float *mkpts0, float *mkpts1,
                         int capacity, int32_t *count_out,
                         void *workspace, size_t workspace_bytes, casmtr_stream_t stream);

/* Packs a match list for the multi-GPU exchange (the reference gathers pickled result dicts over gloo,
 * src/utils/comm.py:180-220): out [(capacity+1) x 44 bytes]; row 0 = count (int64 little endian), rows 1..M =
 * {b_id + pair_offset, i_id, j_id : int64; mconf, mkpts0[2], mkpts1[2] : fp32}.  Rows past M are left untouched. */
CASMTR_API int casmtr_pack_matches(const int64_t *b_ids, const int64_t *i_ids, const int64_t *j_ids, const float *mconf,
                        const float *mkpts0, const float *mkpts1, int M, int64_t pair_offset, int capacity,
                        unsigned char *out, casmtr_stream_t stream);

/* Same with the match count still on the device (count = the int32 casmtr_match_extract wrote; the arrays hold `capacity`
 * rows): min(*count, capacity) records are packed, no host synchronisation between extraction and the all-gather. */
CASMTR_API int casmtr_pack_matches_dev(const int64_t *b_ids, const int64_t *i_ids, const int64_t *j_ids, const float *mconf,
                        const float *mkpts0, const float *mkpts1, const int32_t *count, int64_t pair_offset, int capacity,
                        unsigned char *out, casmtr_stream_t stream);

/* Window gather of CascadeFinePreprocess (src/model/functions/fine_matching.py:47-55; SURVEY 8f "next" #4): for match m the
 * W x W window of the fine map feat [B,C,Hf,Wf] centred on coarse token ids[m] (grid width wc, fine = coarse * stride),
 * zero padded, written as out [M, W*W, C].  Equals F.unfold(feat, W, stride=stride, padding=W/2)[b_ids, ids] of the reference
 * without unfolding the whole map. */
CASMTR_API int casmtr_fine_window_gather(const float *feat, const int64_t *b_ids, const int64_t *ids, float *out,
                              int M, int C, int Hf, int Wf, int wc, int stride, int W, casmtr_stream_t stream);

/* ---------------------------------------------------------------- fine matching (R8) */

/* feat_f0/feat_f1 [M,WW,C]; mkpts1_c [M,2]; scale = hw0_i[0]/hw0_f[0]; scale1_b NULL or [B,2] with
 * b_ids [M] int64; expec_f [M,3] = (x, y, std); mkpts1_f [M,2].  WW <= 32 (W in {3,5}), C % 4 == 0. */
CASMTR_API int casmtr_fine_match_fwd(const float *feat_f0, const float *feat_f1, const float *mkpts1_c,
                          const float *scale1_b, const int64_t *b_ids, float scale,
                          float *expec_f, float *mkpts1_f, int M, int WW, int C, casmtr_stream_t stream);

/* Same with the match count on the device: rows [0, min(*count, capacity)) are computed, the rest of the outputs is left
 * untouched -- lets the whole path (extraction -> fine matching) run without the host reading the count in between. */
CASMTR_API int casmtr_fine_match_dev_fwd(const float *feat_f0, const float *feat_f1, const float *mkpts1_c,
                          const float *scale1_b, const int64_t *b_ids, float scale,
                          float *expec_f, float *mkpts1_f, const int32_t *count, int capacity, int WW, int C,
                          casmtr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CASMTR_B200_H */
