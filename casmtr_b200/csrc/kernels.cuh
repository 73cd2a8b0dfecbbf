// Internal launcher declarations shared between the .cu files and capi.cu.
#pragma once
#include "common.cuh"

// ---- layout.cu
struct TransposeJob {
    const float *src;   // [B, C, HW]
    float *dst;         // [B, HW, C]
    int C, HW;
    int tile_begin;     // filled by the launcher
};
struct TransposeJobs {
    TransposeJob job[12];
    int n;
};
int launch_transpose_jobs(TransposeJobs &jobs, int B, cudaStream_t stream);
struct PoolJob {
    const float *src;   // [B, h*w, C] token-major
    float *dst;         // [B, (h/2)*(w/2), C]
    int h, w;
    float *dst2;        // launch_pool2_tokens only: [B, (h/4)*(w/4), C]
    // launch_pool2_tokens only, optional operands of the tensor-core coarsest level derived from dst2 (qtatt_coarse_tc.cu):
    float *lo2;                     // x - trunc_tf32(x), same layout as dst2
    unsigned short *vt_hi, *vt_lo;  // dst2 * 2^8 = hi + lo as fp16, channel-major [B, C, Sp] (zero-filled up to Sp)
    int Sp;
};
struct PoolJobs {
    PoolJob job[3];
    int n;
};
int launch_pool_tokens(const PoolJobs &jobs, int B, int C, cudaStream_t stream);
int launch_pool2_tokens(const PoolJobs &jobs, int B, int C, cudaStream_t stream);
int launch_guided_prep(const int64_t *pos, int *idx, size_t n_cells, int K, int nh, int hv, int wv, const float *weight, int n_w, float *wsm,
                       cudaStream_t stream);
int launch_topk_to_api(const int *idx, const float *score, int64_t *idx_out, float *score_out,
                       size_t n_tok, int nh, int k, cudaStream_t stream);

// ---- qtatt_coarse.cu
struct CoarseParams {
    const float *q, *k, *v;     // token-major [B,Sq,C] / [B,Sk,C]
    float *acc;                 // [B,Sq,C]: level-0 contribution to the merged message
    int *topk_idx;              // [B,Sq,nh,k] key index at this level
    float *topk_score;          // [B,Sq,nh,k]
    const float *level_weight;  // raw QTAttB.weight (device) or NULL
    float *wsm;                 // [levels] out: softmax(level_weight), written once for the finer levels' kernels (NULL if no weights)
    int levels, n_weights;      // pyramid levels; entries of level_weight the soft-max runs over (>= levels)
    int B, Sq, Sk, nh, topk;
    int type_a;
    float *tc_ws;               // coarse_tc_workspace_floats() of scratch for the tensor-core kernel, or NULL: fp32 SIMT kernel
    bool tc_prepped;            // tc_ws already holds Q_lo / K_lo / V^T (written by launch_pool2_tokens, see coarse_tc_operands)
};
size_t coarse_smem_bytes(int Sk);
int launch_qtatt_coarse(const CoarseParams &p, cudaStream_t stream);
// qtatt_coarse_tc.cu: the same level with both contractions on tcgen05 (Sq, Sk >= 64, Sk <= 1344)
bool coarse_tc_applicable(int Sq, int Sk, int topk);
size_t coarse_tc_workspace_floats(int B, int Sq, int Sk, int C);
int launch_qtatt_coarse_tc(const CoarseParams &p, float *ws, cudaStream_t stream);
void coarse_tc_row_tiles(int Sq, int bh, int n_sm, int &n_big, int &n_small, int &rows_small);   // the kernel's row tiling (host logic)
struct CoarseTcOperands { float *q_lo, *k_lo; unsigned short *vt_hi, *vt_lo; int Sp; };
CoarseTcOperands coarse_tc_operands(float *ws, int B, int Sq, int Sk, int C);      // where those operands live inside tc_ws

// ---- relative position bias of the cascade cross attention, computed where it is consumed
// (CascadeFeatureTransformer.get_relative_pe, src/model/modules/transformer.py:473-509; casmtr_relpe_desc)
struct RelPE {
    const float *w_tab, *h_tab; // [n_emb, nh]; w_tab == NULL: no bias
    const int64_t *tgt_idx;     // [B, h8*w8]
    int n_emb, LB, s;           // s = h0 / h8
    int h8, w8, w8o;
};
#ifdef __CUDACC__
// query-side part of the two table indices of query token (Y, X): (LB + X % s - tgt_x, LB + Y % s - tgt_y)   (:474-485, :500-503)
__device__ __forceinline__ int2 relpe_query_term(const RelPE &pe, int b, int Y, int X) {
    const int t = (int)__ldg(pe.tgt_idx + (size_t)b * pe.h8 * pe.w8 + (Y / pe.s) * pe.w8 + X / pe.s);
    const int off = pe.s / 2 - 1;
    const int ty = t / pe.w8o;
    return make_int2(pe.LB + X % pe.s - ((t - ty * pe.w8o) * pe.s + off), pe.LB + Y % pe.s - (ty * pe.s + off));
}
// bias of key token (ky, kx) for head h: w_pos_bias[rx] + h_pos_bias[ry]   (:504-507)
__device__ __forceinline__ float relpe_bias(const RelPE &pe, int nh, int h, int2 qt, int ky, int kx) {
    const int rx = min(max(qt.x + kx, 0), pe.n_emb - 1), ry = min(max(qt.y + ky, 0), pe.n_emb - 1);
    return __ldg(pe.w_tab + rx * nh + h) + __ldg(pe.h_tab + ry * nh + h);
}
#endif
int launch_relative_pe(const RelPE &pe, const int64_t *window_pos, float *rel_pos, int B, int nh, int h0, int w0, int k, cudaStream_t stream);   // relpe.cu

// ---- qtatt_fine.cu
struct FineParams {
    const float *q;             // token-major raster [B, h0*w0, C]
    const float *k, *v;         // token-major raster [B, h1*w1, C]
    const int *prev_idx;        // [B, Np, nh, kp]   (QTAtt levels)  key index on the (h1/2 x w1/2) grid
    const float *prev_score;    // [B, Np, nh, kp]   (type A only)
    const int64_t *topk_pos;    // [B, Np, kp, 2]    (cascade)  row, col on the previous grid
    const int64_t *next_idx;    // [B, Np] (cascade, instead of topk_pos): match of the parent cell on the previous grid of the other
                                // image; the win x win window around it, shifted inside the grid, is derived in the kernel
    int win;                    // window side (kp == win * win) when next_idx is used
    const float *rel_pos;       // [B, nh, h0*w0, 4kp] or NULL (cascade)
    RelPE pe;                   // cascade: the same bias computed from its embedding tables (pe.w_tab != NULL; rel_pos NULL then)
    const float *acc_prev;      // [B, Np, C] merged message of the coarser levels, or NULL (cascade)
    float *out;                 // [B, h0*w0, C] raster: acc_prev[parent] + w * message
    int *topk_idx;              // [B, h0*w0, nh, k] or NULL (last level / cascade)
    float *topk_score;          // same shape, or NULL
    int64_t *upsampled_idx;     // [B, h0*w0, 4kp] or NULL (cascade)
    const float *wsm;           // softmax(QTAttB.weight) written by the coarse kernel, or NULL (type A, cascade: weight 1)
    int level;                  // this level's position in the weight vector
    int B, nh, h0, w0, h1, w1, w_prev;
    int kp, topk;               // candidates = 4*kp; topk selected for the next level
    int dil;                    // child offset dilation (cascade), 1 otherwise
    int type_a, final_level;
    const int *item_list;       // NULL, or device list of cells (b * Np + parent) to process for every head (cascade fallback)
    const int *item_count;      // device int: length of item_list
    // filled by launch_quad_attention: the divisors of the kernels' index arithmetic
    FastDiv d_nh, d_wp, d_np, d_wprev, d_wv, d_win;     // nh, w0 / 2, (h0 / 2) (w0 / 2), w_prev, w1 / 2, win
};
int launch_quad_attention(const FineParams &p, cudaStream_t stream);

// ---- cascade_tile.cu: TMA-tiled CascadeQTAttB (k = 25, dilated = 1); cells it cannot serve go to fb_list
size_t cascade_tile_smem_bytes();
// top-left corner of the win x win window centred on `centre`, shifted rigidly inside [0, n)
// (CascadeFeatureTransformer.get_window_warp_idx, src/model/modules/transformer.py:427-434)
__host__ __device__ inline int window_origin(int centre, int win, int n) {
    int o = centre - win / 2;
    if (o < 0) o = 0;
    if (o + win > n) o = n - win;
    return o;
}
int launch_window_idx(const int64_t *next_idx, int64_t *pos, size_t rows, int H, int W, int win, cudaStream_t stream);
int launch_cascade_att_tile(const float *q, const float *k, const float *v, const int64_t *topk_pos, const int64_t *next_idx, const float *rel_pos,
                            const RelPE &pe, float *out, int64_t *upsampled_idx, int *fb_list, int *fb_count,
                            int B, int nh, int h0, int w0, int h1, int w1, cudaStream_t stream);

// ---- ops.cu
int launch_score5d(const float *q, const float *key, const int64_t *idx, float *out,
                   int B, int N1, int N2, int H, int D, int K, cudaStream_t stream);
int launch_value_agg(const float *score, const float *value, const int64_t *idx, float *out,
                     int B, int N, int K, int H, int M, int D, cudaStream_t stream);
int launch_score3d(const float *q, const float *key, const int64_t *idx, float *out,
                   int B, int N1, int N2, int C, int K, cudaStream_t stream);

// ---- ops_bwd.cu: backward halves of the three op-level drop-ins (grad_key / grad_value are zeroed inside)
int launch_score5d_bwd(const float *grad, const float *q, const float *key, const int64_t *idx, float *gq, float *gk,
                       int B, int N1, int N2, int H, int D, int K, cudaStream_t stream);
int launch_value_agg_bwd(const float *grad, const float *score, const float *value, const int64_t *idx, float *gscore, float *gvalue,
                         int B, int N, int K, int H, int M, int D, cudaStream_t stream);
int launch_score3d_bwd(const float *grad, const float *q, const float *key, const int64_t *idx, float *gq, float *gk,
                       int B, int N1, int N2, int C, int K, cudaStream_t stream);

// ---- cascade_match.cu
struct MatchParams {
    const float *feat0, *feat1;
    const int64_t *idx01, *idx10;
    const uint8_t *mask0, *mask1;
    float inv_scale;            // 1 / (C * temperature)
    float *conf01, *conf10;
    float *next_conf01, *next_conf10;
    int64_t *next_idx01, *next_idx10;
    int B, L0, L1, C, K;
    int w0, w1;                 // grid widths of image 0 / 1 (0 = unknown: one warp per query row)
    int *fb_list, *fb_count;    // workspace of the TMA-tiled kernel: cells it could not serve (NULL = tile path off)
    const int *cell_list;       // cell kernel: NULL, or process only these cells (count in *cell_count)
    const int *cell_count;
};
int launch_cascade_match(const MatchParams &p, cudaStream_t stream);

// ---- match_tile.cu: TMA-tiled correlation for window-structured candidate lists; other cells go to p.fb_list
size_t match_tile_smem_bytes();
bool match_tile_applicable(const MatchParams &p);
int launch_cascade_match_tile(const MatchParams &p, cudaStream_t stream);

// ---- coarse_match.cu: dense dual-softmax row / column statistics on tcgen05
size_t coarse_match_workspace(int B, int L0, int L1, int C);
int launch_coarse_match(const float *feat0, const float *feat1, const uint8_t *mask0, const uint8_t *mask1, float temperature, float *conf01, int64_t *idx01, float *conf10,
                        int64_t *idx10, int B, int L0, int L1, int C, void *workspace, size_t workspace_bytes, cudaStream_t stream,
                        float *mconf_row = nullptr, int64_t *midx_row = nullptr, int64_t *midx_col = nullptr);

// ---- extract.cu
int launch_match_extract(const casmtr_extract_desc &d, const float *next_conf01, const int64_t *next_idx01,
                         const int64_t *next_idx10, uint8_t *mask_out, int64_t *b_ids, int64_t *i_ids,
                         int64_t *j_ids, float *mconf, float *mkpts0, float *mkpts1, int capacity,
                         int32_t *count_out, void *workspace, size_t workspace_bytes, cudaStream_t stream);
size_t match_extract_workspace(const casmtr_extract_desc &d);
int launch_pack_matches(const int64_t *b_ids, const int64_t *i_ids, const int64_t *j_ids, const float *mconf, const float *mk0,
                        const float *mk1, int M, long long pair_offset, unsigned char *out, const int32_t *count_dev, cudaStream_t stream);

// ---- fine_match.cu
int launch_fine_window_gather(const float *feat, const int64_t *b_ids, const int64_t *ids, float *out, int M, int C, int Hf, int Wf,
                              int wc, int stride, int W, cudaStream_t stream);
int launch_fine_match(const float *f0, const float *f1, const float *mkpts1_c, const float *scale1_b,
                      const int64_t *b_ids, float scale, float *expec_f, float *mkpts1_f,
                      int M, int WW, int C, const int32_t *count_dev, cudaStream_t stream);
