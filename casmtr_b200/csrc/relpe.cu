// Relative position bias of the cascade cross attention as a tensor: the stand-alone drop-in for
// CascadeFeatureTransformer.get_relative_pe (src/model/modules/transformer.py:473-509), which the reference builds with
// ~25 torch ops over [B,HW,4ww,2] int64 intermediates.  One thread per (batch, query token, candidate): the two table
// indices are integer arithmetic on the query position, the query cell's 1/8 match and the candidate's window entry;
// the nhead values of the candidate are two table rows added (the tables are a few hundred bytes: L1 hits).
// The attention kernels compute the same bias in place (kernels.cuh: relpe_query_term / relpe_bias) -- this kernel
// exists for callers that want the tensor itself.
#include "common.cuh"
#include "kernels.cuh"

namespace {

__global__ void __launch_bounds__(256) relative_pe_kernel(RelPE pe, const int64_t *__restrict__ window_pos, float *__restrict__ rel_pos,
                                                          int B, int nh, int h0, int w0, int k) {
    pdl_sync();
    const int KC = 4 * k, L0 = h0 * w0, wp = w0 >> 1, Np = (h0 >> 1) * wp;
    const long long W1 = (long long)pe.w8o * pe.s;
    const size_t n = (size_t)B * L0 * KC;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % KC);
        const int q = (int)((i / KC) % L0);
        const int b = (int)(i / ((size_t)KC * L0));
        const int Y = q / w0, X = q - Y * w0;
        const int2 qt = relpe_query_term(pe, b, Y, X);
        const int64_t *wpos = window_pos + (((size_t)b * Np + (Y >> 1) * wp + (X >> 1)) * k + (c >> 2)) * 2;
        // child (x_, y_) of window entry (row, col): flat index on the other image's current level, then back to 2-D (:489-499;
        // `%` is Python's non-negative modulo, the division truncates)
        const long long idx = (2 * wpos[0] + ((c & 3) >> 1)) * W1 + 2 * wpos[1] + (c & 1);
        const int kx = (int)(((idx % W1) + W1) % W1), ky = (int)(idx / W1);
        float *o = rel_pos + ((size_t)b * nh * L0 + q) * KC + c;
        for (int h = 0; h < nh; ++h) o[(size_t)h * L0 * KC] = relpe_bias(pe, nh, h, qt, ky, kx);
    }
}

}  // namespace

int launch_relative_pe(const RelPE &pe, const int64_t *window_pos, float *rel_pos, int B, int nh, int h0, int w0, int k, cudaStream_t stream) {
    const size_t n = (size_t)B * h0 * w0 * 4 * k;
    if (n == 0) return CASMTR_OK;
    const size_t blocks = (n + 255) / 256;
    LaunchScope ls(CASMTR_K_LAYOUT, stream);
    launch_k(relative_pe_kernel, dim3((unsigned)(blocks < 148 * 64 ? blocks : 148 * 64)), 256, 0, stream, pe, window_pos, rel_pos, B, nh, h0, w0, k);
    CASMTR_CHECK_LAUNCH("relative_pe_kernel");
    return CASMTR_OK;
}
