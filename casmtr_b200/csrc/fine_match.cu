// Sub-pixel fine matching: centre-vs-window correlation -> softmax -> DSNT expectation + std.
// Reference: CascadeFineMatching.forward / get_fine_match  src/model/functions/fine_matching.py:77-137
// (legacy FineMatching :201-261 is the same arithmetic).  kornia's spatial_expectation2d /
// create_meshgrid (normalised grid linspace(-1,1,W), x then y) are folded in.
// One warp per match: lane r (< WW) owns window position r of image 1.
#include "common.cuh"
#include "kernels.cuh"

namespace {

__global__ void __launch_bounds__(256) fine_match_kernel(const float *__restrict__ f0, const float *__restrict__ f1,
                                                          const float *__restrict__ mkpts1_c, const float *__restrict__ scale1_b,
                                                          const int64_t *__restrict__ b_ids, float scale,
                                                          float *__restrict__ expec_f, float *__restrict__ mkpts1_f,
                                                          int M, int WW, int W, int C, const int32_t *__restrict__ count_dev) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (count_dev != nullptr) M = min(M, (int)__ldg(count_dev));      // match count still on the device (no host sync yet)
    if (m >= M) return;
    const float *centre = f0 + ((size_t)m * WW + WW / 2) * C;          // feat_f0[:, WW//2, :]  (:105)
    float sim = -INFINITY;
    if (lane < WW) {
        const float *row = f1 + ((size_t)m * WW + lane) * C;
        float s = 0.f;
        for (int c = 0; c < C; c += 4) {
            const float4 a = ldg4(centre + c), bq = ldg4(row + c);
            s = fmaf(a.x, bq.x, fmaf(a.y, bq.y, fmaf(a.z, bq.z, fmaf(a.w, bq.w, s))));
        }
        sim = s * (1.0f / sqrtf((float)C));                            // softmax_temp * sim  (:107-108)
    }
    const float mx = warp_max(sim);
    const float e = lane < WW ? exp_neg(sim - mx) : 0.f;
    const float p = e / warp_sum(e);
    // normalised grid: (i / (W-1) - 0.5) * 2
    const float gx = lane < WW ? ((float)(lane % W) / (float)(W - 1) - 0.5f) * 2.f : 0.f;
    const float gy = lane < WW ? ((float)(lane / W) / (float)(W - 1) - 0.5f) * 2.f : 0.f;
    const float ex = warp_sum(gx * p), ey = warp_sum(gy * p);
    const float vx = warp_sum(gx * gx * p) - ex * ex, vy = warp_sum(gy * gy * p) - ey * ey;   // (:115)
    if (lane == 0) {
        const float sd = sqrtf(fmaxf(vx, 1e-10f)) + sqrtf(fmaxf(vy, 1e-10f));                 // (:116)
        expec_f[3 * (size_t)m] = ex; expec_f[3 * (size_t)m + 1] = ey; expec_f[3 * (size_t)m + 2] = sd;
        float sx = scale, sy = scale;
        if (scale1_b) {
            const long long b = b_ids[m];
            sx = scale * scale1_b[2 * b]; sy = scale * scale1_b[2 * b + 1];
        }
        const float half = (float)(W / 2);
        mkpts1_f[2 * (size_t)m] = mkpts1_c[2 * (size_t)m] + ex * half * sx;                    // (:131)
        mkpts1_f[2 * (size_t)m + 1] = mkpts1_c[2 * (size_t)m + 1] + ey * half * sy;
    }
}

// Window gather of CascadeFinePreprocess (src/model/functions/fine_matching.py:47-55): the reference unfolds the WHOLE fine
// map (F.unfold, 277 MB - 1.1 GB) and then selects M rows; this reads only the M windows.  CTA = one match: W*W pixels x C
// channels of the NCHW fine map -> out[m, W*W, C] (channel innermost), zero padding outside the map, via a smem transpose.
__global__ void __launch_bounds__(256) fine_window_gather_kernel(const float *__restrict__ feat, const int64_t *__restrict__ b_ids,
                                                                  const int64_t *__restrict__ ids, float *__restrict__ out,
                                                                  int C, int Hf, int Wf, int wc, int stride, int W) {
    extern __shared__ float tile[];                 // [W*W][C + 1]
    const int m = blockIdx.x, WW = W * W, ld = C + 1;
    const long long b = b_ids[m], id = ids[m];
    const int cy = (int)(id / wc) * stride - W / 2, cx = (int)(id % wc) * stride - W / 2;      // top-left fine pixel of the window
    const float *fb = feat + (size_t)b * C * Hf * Wf;
    for (int i = threadIdx.x; i < C * W; i += blockDim.x) {                                    // one window row of one channel per thread
        const int c = i / W, wy = i - c * W;
        const int y = cy + wy;
        const float *src = fb + ((size_t)c * Hf + y) * Wf;
        for (int wx = 0; wx < W; ++wx) {
            const int x = cx + wx;
            tile[(wy * W + wx) * ld + c] = (y >= 0 && y < Hf && x >= 0 && x < Wf) ? __ldg(src + x) : 0.f;
        }
    }
    __syncthreads();
    float *o = out + (size_t)m * WW * C;
    for (int i = threadIdx.x; i < WW * C; i += blockDim.x) o[i] = tile[(i / C) * ld + (i % C)];
}

}  // namespace

int launch_fine_window_gather(const float *feat, const int64_t *b_ids, const int64_t *ids, float *out, int M, int C, int Hf, int Wf,
                              int wc, int stride, int W, cudaStream_t stream) {
    if (M == 0) return CASMTR_OK;
    const size_t smem = sizeof(float) * (size_t)W * W * (C + 1);
    CASMTR_REQUIRE(smem <= 48 * 1024, CASMTR_E_UNSUPPORTED, "fine_window_gather: window %d x C=%d exceeds shared memory", W, C);
    LaunchScope ls(CASMTR_K_FINE_MATCH, stream);
    fine_window_gather_kernel<<<M, 256, smem, stream>>>(feat, b_ids, ids, out, C, Hf, Wf, wc, stride, W);
    CASMTR_CHECK_LAUNCH("fine_window_gather_kernel");
    return CASMTR_OK;
}

int launch_fine_match(const float *f0, const float *f1, const float *mkpts1_c, const float *scale1_b,
                      const int64_t *b_ids, float scale, float *expec_f, float *mkpts1_f,
                      int M, int WW, int C, const int32_t *count_dev, cudaStream_t stream) {
    int W = 1;
    while (W * W < WW) ++W;
    CASMTR_REQUIRE(W * W == WW && WW <= 32 && W >= 2, CASMTR_E_UNSUPPORTED, "fine_match: window %d must be a square <= 32 (W in 2..5)", WW);
    CASMTR_REQUIRE(C % 4 == 0 && C > 0, CASMTR_E_UNSUPPORTED, "fine_match: C=%d must be a positive multiple of 4", C);
    if (M == 0) return CASMTR_OK;
    LaunchScope ls(CASMTR_K_FINE_MATCH, stream);
    launch_k(fine_match_kernel, (M + 7) / 8, 256, 0, stream, f0, f1, mkpts1_c, scale1_b, b_ids, scale, expec_f, mkpts1_f, M, WW, W, C, count_dev);
    CASMTR_CHECK_LAUNCH("fine_match_kernel");
    return CASMTR_OK;
}
