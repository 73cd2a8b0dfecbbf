// Relative position bias of the cascade cross attention as a tensor: the stand-alone drop-in for
// CascadeFeatureTransformer.get_relative_pe (src/model/modules/transformer.py:473-509), which the reference builds with
// ~25 torch ops over [B,HW,4ww,2] int64 intermediates.  One thread per (batch, query token, window entry): the table
// indices are integer arithmetic on the query position, the query cell's 1/8 match and the entry's 2x2 children;
// every output is two table values added (the tables are a few hundred bytes: L1 hits).  Write-bound.
// The attention kernels compute the same bias in place (kernels.cuh: relpe_query_term / relpe_bias) -- this kernel
// exists for callers that want the tensor itself.
#include "common.cuh"
#include "kernels.cuh"

namespace {

// thread = one window entry (4 candidates = the entry's 2x2 children) of one query token, all heads: the index arithmetic is
// paid once per 4 * nhead outputs and every head's 4 values leave as one 16-byte store (consecutive lanes, consecutive 16 bytes)
__global__ void __launch_bounds__(256) relative_pe_kernel(RelPE pe, const int64_t *__restrict__ window_pos, float *__restrict__ rel_pos,
                                                          int B, int nh, int h0, int w0, int k) {
    pdl_sync();
    const int L0 = h0 * w0, wp = w0 >> 1, Np = (h0 >> 1) * wp;
    const int W1 = pe.w8o * pe.s;
    const size_t n = (size_t)B * L0 * k;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int kk = (int)(i % k);
        const size_t bq = i / k;
        const int q = (int)(bq % L0), b = (int)(bq / L0);
        const int Y = q / w0, X = q - Y * w0;
        const int2 qt = relpe_query_term(pe, b, Y, X);
        const longlong2 wpos = __ldg(reinterpret_cast<const longlong2 *>(window_pos) + ((size_t)b * Np + (Y >> 1) * wp + (X >> 1)) * k + kk);
        float4 *o = reinterpret_cast<float4 *>(rel_pos + ((size_t)b * nh * L0 + q) * 4 * k) + kk;
        // child (x_, y_) of window entry (row, col): flat index on the other image's current level, then back to 2-D (:489-499;
        // `%` is Python's non-negative modulo, the division truncates)
        int kx[4], ky[4];
#pragma unroll
        for (int cf = 0; cf < 4; ++cf) {
            const long long idx = (2 * wpos.x + (cf >> 1)) * W1 + 2 * wpos.y + (cf & 1);
            if (idx >= 0 && idx < 0x7fffffffLL) {               // 32-bit arithmetic for every real grid
                ky[cf] = (int)idx / W1;
                kx[cf] = (int)idx - ky[cf] * W1;
            } else {
                ky[cf] = (int)(idx / W1);
                kx[cf] = (int)(((idx % W1) + W1) % W1);
            }
        }
        for (int h = 0; h < nh; ++h)
            o[(size_t)h * L0 * k] = make_float4(relpe_bias(pe, nh, h, qt, ky[0], kx[0]), relpe_bias(pe, nh, h, qt, ky[1], kx[1]),
                                                relpe_bias(pe, nh, h, qt, ky[2], kx[2]), relpe_bias(pe, nh, h, qt, ky[3], kx[3]));
    }
}

}  // namespace

int launch_relative_pe(const RelPE &pe, const int64_t *window_pos, float *rel_pos, int B, int nh, int h0, int w0, int k, cudaStream_t stream) {
    const size_t n = (size_t)B * h0 * w0 * k;
    if (n == 0) return CASMTR_OK;
    CASMTR_REQUIRE((((uintptr_t)window_pos | (uintptr_t)rel_pos) & 15) == 0, CASMTR_E_INVALID, "relative_pe: window_pos / rel_pos must be 16-byte aligned");
    const size_t blocks = (n + 255) / 256;
    LaunchScope ls(CASMTR_K_LAYOUT, stream);
    launch_k(relative_pe_kernel, dim3((unsigned)(blocks < 148 * 64 ? blocks : 148 * 64)), 256, 0, stream, pe, window_pos, rel_pos, B, nh, h0, w0, k);
    CASMTR_CHECK_LAUNCH("relative_pe_kernel");
    return CASMTR_OK;
}
