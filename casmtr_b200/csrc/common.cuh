// Shared device/host helpers for libcasmtr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/casmtr_b200.h"

#define FULL_MASK 0xffffffffu
#define LOG2E_F 1.4426950408889634f

void casmtr_set_error(const char *fmt, ...);

#define CASMTR_REQUIRE(cond, code, ...)          \
    do {                                         \
        if (!(cond)) {                           \
            casmtr_set_error(__VA_ARGS__);       \
            return (code);                       \
        }                                        \
    } while (0)

#define CASMTR_CHECK_LAUNCH(what)                                                   \
    do {                                                                            \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) {                                                   \
            casmtr_set_error("%s: %s", what, cudaGetErrorString(e__));              \
            return CASMTR_E_CUDA;                                                   \
        }                                                                           \
    } while (0)

// ---- launch accounting / optional per-kernel timing (capi.cu).  Every kernel launch of the library sits in
// a LaunchScope: it counts the launch and, when casmtr_profile_enable(1) is active, brackets it with a CUDA
// event pair recorded on the launch stream so bench.py can attribute time per kernel kind inside its timed region.
void casmtr_prof_begin(int kind, cudaStream_t stream, int *slot);
void casmtr_prof_end(int slot, cudaStream_t stream);
struct LaunchScope {
    int slot;
    cudaStream_t stream;
    LaunchScope(int kind, cudaStream_t s) : slot(-1), stream(s) { casmtr_prof_begin(kind, s, &slot); }
    ~LaunchScope() { if (slot >= 0) casmtr_prof_end(slot, stream); }
};

// ---- programmatic dependent launch.  Every hot-path kernel starts with pdl_sync(): it lets the NEXT kernel of the stream be
// scheduled onto SMs as this grid's tail frees them (griddepcontrol.launch_dependents) and then waits until the PREVIOUS
// grid has completed and flushed (griddepcontrol.wait) before touching memory -- launch latency and CTA ramp-up overlap the
// predecessor's tail, the data dependency stays a full one.  launch_k() sets the matching launch attribute
// (CASMTR_PDL=0 in the environment turns it off; without the attribute both instructions are no-ops).
bool casmtr_pdl_enabled();
int casmtr_concurrency();       // casmtr_qtatt_desc::concurrent_calls of the running entry point (>= 1)
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = casmtr_pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

// ---- per-device one-time setup.  cudaFuncSetAttribute applies to the CURRENT device only, so the "already done" flag of a
// launcher is one bit per device ordinal, not one per process.  Setting an attribute twice is harmless, so two host threads
// racing through the first call both do the setup and both publish the bit (release / acquire on the atomic).
#include <atomic>
struct PerDeviceOnce {
    std::atomic<uint64_t> bits{0};
    // device ordinal of the calling thread (0 if it cannot be read; ordinals >= 64 never cache and redo the setup each call)
    static int device() { int d = 0; if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = 0; } return d; }
    bool done(int dev) const { return dev >= 0 && dev < 64 && ((bits.load(std::memory_order_acquire) >> dev) & 1ull); }
    void mark(int dev) { if (dev >= 0 && dev < 64) bits.fetch_or(1ull << dev, std::memory_order_release); }
};
int casmtr_sm_count();          // SM count of the current device (cached per device; capi.cu)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Division of a non-negative int32 by a launch-time constant as multiply-high + shift (the grid widths and head counts the
// kernels divide by are runtime values, and an integer division is ~20 instructions on the SIMT cores).  Exact for 0 <= n < 2^31.
struct FastDiv {
    unsigned mul, shr;          // mul == 0: divisor 1
    int d;
#ifdef __CUDACC__
    __device__ __forceinline__ int div(int n) const { return mul ? (int)(__umulhi((unsigned)n, mul) >> shr) : n; }
    __device__ __forceinline__ void divmod(int n, int &q, int &r) const { q = div(n); r = n - q * d; }
#endif
};
static inline FastDiv make_fastdiv(int div) {
    FastDiv f;
    f.mul = 0; f.shr = 0; f.d = div > 0 ? div : 1;
    if (div > 1) {
        int lg = 0;
        while ((1ll << lg) < div) ++lg;                           // ceil(log2 d)
        const unsigned long long p2 = 1ull << (31 + lg);
        f.mul = (unsigned)((p2 + (unsigned)div - 1) / (unsigned)div);
        f.shr = (unsigned)(lg - 1);
    }
    return f;
}

// Bump allocator over the caller's workspace.
struct Workspace {
    char *base;
    size_t cap, off;
    Workspace(void *p, size_t n) : base((char *)p), cap(n), off(0) {}
    template <typename T>
    T *take(size_t count) {
        size_t bytes = align_up(count * sizeof(T), 256);
        T *p = (T *)(base ? base + off : nullptr);
        off += bytes;
        return p;
    }
    bool ok() const { return off <= cap; }
};

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// streaming 128-bit load that does not allocate in L1 (data touched once)
__device__ __forceinline__ float4 ldg4_stream(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// warp-wide fp32 max in ONE instruction: redux.sync.max.f32 (CREDUX.MAX.F32, sm_100a) instead of a 5-step shuffle chain
__device__ __forceinline__ float warp_max(float v) {
    float m;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
    return m;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// Bitonic sort of one fp32 value per lane, descending (lane 0 ends with the largest): 15 shuffle + FMNMX steps (the compare
// direction is a predicate of the lane index, which FMNMX takes as its min / max selector).
__device__ __forceinline__ float warp_sort_desc(float v, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const float o = __shfl_xor_sync(FULL_MASK, v, stride);
            const bool up = (lane & size) != 0;                  // this block of `size` lanes sorts ascending
            const bool lower = (lane & stride) == 0;
            v = (lower == up) ? fminf(v, o) : fmaxf(v, o);
        }
    }
    return v;
}

// exp(x) for x <= 0 as used by every softmax here: one FMUL + EX2 (2 ulp).
__device__ __forceinline__ float exp_neg(float x) { return exp2f(x * LOG2E_F); }
