// extern "C" entry points of libcasmtr_b200.so (see include/casmtr_b200.h): argument validation,
// workspace carving and kernel sequencing.  No allocation, no synchronisation, caller's stream (plus the fork/join side lanes below).
#include <cstdlib>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "kernels.cuh"

static thread_local char g_err[512] = "";

void casmtr_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- launch accounting + optional event-pair timing (see LaunchScope in common.cuh)
#include <atomic>
#include <mutex>
#include <vector>
namespace {
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_prof_on{0};
struct ProfRec { int kind; cudaEvent_t start, stop; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof_recs;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_free;
const char *const g_kind_names[CASMTR_K_COUNT] = {"layout", "qt_coarse", "qt_fine_mid", "qt_fine_last", "cascade_att",
                                                  "cascade_match", "extract", "fine_match", "ops", "cascade_fallback", "coarse_match"};
}  // namespace

// Launch options are per calling THREAD (thread_local), never process-global: two host threads driving two devices / streams
// cannot change each other's launch geometry, and what a stream capture bakes in is what the capturing thread asked for.
static int env_default(const char *name) {
    const char *e = getenv(name);
    return (e && e[0] == '0') ? 0 : 1;
}
static thread_local int t_pdl = -1;
bool casmtr_pdl_enabled() {
    if (t_pdl < 0) t_pdl = env_default("CASMTR_PDL");
    return t_pdl != 0;
}
int casmtr_set_pdl(int on) {
    const int prev = casmtr_pdl_enabled() ? 1 : 0;
    t_pdl = on ? 1 : 0;
    return prev;
}

// concurrent calls of the running entry point (casmtr_qtatt_desc::concurrent_calls), visible to the launchers below it
static thread_local int t_concurrency = 1;
int casmtr_concurrency() { return t_concurrency; }
struct ConcurrencyScope {
    int prev;
    explicit ConcurrencyScope(int n) : prev(t_concurrency) { t_concurrency = n < 1 ? 1 : (n > 64 ? 64 : n); }
    ~ConcurrencyScope() { t_concurrency = prev; }
};

int casmtr_sm_count() {
    static std::atomic<int> cache[64];
    const int dev = PerDeviceOnce::device();
    int n = (dev >= 0 && dev < 64) ? cache[dev].load(std::memory_order_relaxed) : 0;
    if (n <= 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
        if (dev >= 0 && dev < 64) cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

// ---- fork / join onto a library-owned side stream (casmtr_set_overlap).  A fused call whose first kernel needs only part of
// the re-laid inputs forks the rest of the layout work onto a side stream and joins it back before the first consumer, so the
// HBM-bound transposes run under the issue-bound first kernel.  Everything stays ordered with respect to the caller's stream:
// the fork waits for the caller's stream, the caller's stream waits for the join.  Both are event record / wait pairs, legal
// inside a stream capture (the side stream joins the capture and leaves it at the join).  Lanes are created once per device,
// outside any capture where possible (casmtr_set_overlap(1) or the first eager call), and handed out round-robin.
namespace {
struct SideLane { cudaStream_t stream; cudaEvent_t fork, join; };
// Two lanes per device: enough for the two directions of a layer, and few enough that caller stream + lanes + a copy stream + NCCL's
// stream stay within the 8 hardware work queues of a default CUDA context (streams that share a queue pick up false dependencies:
// a lane queued behind an NCCL kernel that waits for a peer rank stalls the call that joins it).
constexpr int SIDE_LANES = 2, SIDE_MAX_DEV = 16;
std::mutex g_side_mu;                       // creation of the lanes
std::mutex g_side_use[SIDE_MAX_DEV];        // one caller at a time enqueues fork ... join on a device's lanes (see LaneGuard)
SideLane g_side[SIDE_MAX_DEV][SIDE_LANES];
bool g_side_ready[SIDE_MAX_DEV];
unsigned g_side_next[SIDE_MAX_DEV];
thread_local int t_overlap = -1;

bool overlap_enabled() {
    if (t_overlap < 0) t_overlap = env_default("CASMTR_OVERLAP");
    return t_overlap != 0;
}

// A lane of the current device, or nullptr (overlap off / lanes unavailable: the caller then stays on its own stream).
// The returned guard owns the device's lane mutex: the fork record + wait and the join record + wait of a call are enqueued
// without another host thread's record landing between them (a lane has ONE fork and ONE join event; an interleaved record
// would make this call's wait observe the other caller's stream position).
struct LaneGuard {
    SideLane *lane = nullptr;
    std::unique_lock<std::mutex> lock;
};
LaneGuard side_lane(bool want) {
    LaneGuard g;
    if (!want || !overlap_enabled()) return g;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= SIDE_MAX_DEV) { cudaGetLastError(); return g; }
    {
        std::lock_guard<std::mutex> lk(g_side_mu);
        if (!g_side_ready[dev]) {
            for (int i = 0; i < SIDE_LANES; ++i) {
                SideLane &l = g_side[dev][i];
                if (cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) != cudaSuccess ||
                    cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming) != cudaSuccess) {
                    cudaGetLastError();
                    return g;
                }
            }
            g_side_ready[dev] = true;
        }
    }
    g.lock = std::unique_lock<std::mutex>(g_side_use[dev]);
    g.lane = &g_side[dev][g_side_next[dev]++ % SIDE_LANES];
    return g;
}
}  // namespace

int casmtr_set_overlap(int on) {
    const int prev = overlap_enabled() ? 1 : 0;
    t_overlap = on ? 1 : 0;
    if (on) side_lane(true);                // create the lanes now (outside a stream capture)
    return prev;
}

void casmtr_prof_begin(int kind, cudaStream_t stream, int *slot) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    *slot = -1;
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r;
    r.kind = kind;
    if (!g_prof_free.empty()) {
        r.start = g_prof_free.back().first; r.stop = g_prof_free.back().second;
        g_prof_free.pop_back();
    } else if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    cudaEventRecord(r.start, stream);
    g_prof_recs.push_back(r);
    *slot = (int)g_prof_recs.size() - 1;
}

void casmtr_prof_end(int slot, cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (slot >= 0 && slot < (int)g_prof_recs.size()) cudaEventRecord(g_prof_recs[slot].stop, stream);
}

extern "C" {

uint64_t casmtr_launch_count(void) { return g_launches.load(); }

int casmtr_profile_enable(int on) {
    g_prof_on.store(on ? 1 : 0);
    return CASMTR_OK;
}

int casmtr_profile_collect(double *ms_by_kind, uint64_t *launches_by_kind) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (const ProfRec &r : g_prof_recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.stop) != cudaSuccess || cudaEventElapsedTime(&ms, r.start, r.stop) != cudaSuccess) {
            casmtr_set_error("casmtr_profile_collect: %s", cudaGetErrorString(cudaGetLastError()));
            g_prof_recs.clear();
            return CASMTR_E_CUDA;
        }
        if (ms_by_kind) ms_by_kind[r.kind] += ms;
        if (launches_by_kind) launches_by_kind[r.kind] += 1;
        g_prof_free.emplace_back(r.start, r.stop);
    }
    g_prof_recs.clear();
    return CASMTR_OK;
}

const char *casmtr_kernel_kind_name(int kind) { return kind >= 0 && kind < CASMTR_K_COUNT ? g_kind_names[kind] : "?"; }

int casmtr_version(void) { return CASMTR_VERSION; }

int casmtr_plan_dense_tiles(int rows, int bh, int n_sm, int out[3]) {
    CASMTR_REQUIRE(out != nullptr && rows >= 1 && bh >= 1 && n_sm >= 1, CASMTR_E_INVALID, "plan_dense_tiles: rows=%d bh=%d n_sm=%d", rows, bh, n_sm);
    coarse_tc_row_tiles(rows, bh, n_sm, out[0], out[1], out[2]);
    return CASMTR_OK;
}

int casmtr_fastdiv(int d, unsigned out[2]) {
    CASMTR_REQUIRE(out != nullptr && d >= 1, CASMTR_E_INVALID, "fastdiv: divisor %d", d);
    const FastDiv f = make_fastdiv(d);
    out[0] = f.mul; out[1] = f.shr;
    return CASMTR_OK;
}

const char *casmtr_last_error_string(void) { return g_err; }

int casmtr_device_info(int *sm_count, size_t *l2_bytes) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        casmtr_set_error("casmtr_device_info: no CUDA device");
        return CASMTR_E_CUDA;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (l2_bytes) *l2_bytes = (size_t)prop.l2CacheSize;
    return CASMTR_OK;
}

// ------------------------------------------------------------------------------------------------ ops
int casmtr_score5d_fwd(const float *query, const float *key, const int64_t *index, float *out,
                       int B, int N1, int N2, int H, int D, int K, casmtr_stream_t stream) {
    CASMTR_REQUIRE(B >= 0 && N1 >= 0 && N2 > 0 && H > 0 && D > 0 && K > 0, CASMTR_E_INVALID, "score5d: bad sizes");
    CASMTR_REQUIRE((query && key && index && out) || (size_t)B * N1 == 0, CASMTR_E_INVALID, "score5d: null pointer");
    return launch_score5d(query, key, index, out, B, N1, N2, H, D, K, (cudaStream_t)stream);
}

int casmtr_value_agg_fwd(const float *score, const float *value, const int64_t *index, float *out,
                         int B, int N, int K, int H, int M, int D, casmtr_stream_t stream) {
    CASMTR_REQUIRE(B >= 0 && N >= 0 && K > 0 && H > 0 && M > 0 && D > 0, CASMTR_E_INVALID, "value_agg: bad sizes");
    CASMTR_REQUIRE((score && value && index && out) || (size_t)B * N == 0, CASMTR_E_INVALID, "value_agg: null pointer");
    return launch_value_agg(score, value, index, out, B, N, K, H, M, D, (cudaStream_t)stream);
}

int casmtr_score3d_fwd(const float *query, const float *key, const int64_t *index, float *out,
                       int B, int N1, int N2, int C, int K, casmtr_stream_t stream) {
    CASMTR_REQUIRE(B >= 0 && N1 >= 0 && N2 > 0 && C > 0 && K > 0, CASMTR_E_INVALID, "score3d: bad sizes");
    CASMTR_REQUIRE(C % 4 == 0, CASMTR_E_UNSUPPORTED, "score3d: C=%d must be a multiple of 4", C);
    CASMTR_REQUIRE((query && key && index && out) || (size_t)B * N1 == 0, CASMTR_E_INVALID, "score3d: null pointer");
    return launch_score3d(query, key, index, out, B, N1, N2, C, K, (cudaStream_t)stream);
}

int casmtr_nchw_to_tokens(const float *src, float *dst, int B, int C, int HW, casmtr_stream_t stream) {
    CASMTR_REQUIRE(B >= 0 && C > 0 && HW > 0, CASMTR_E_INVALID, "nchw_to_tokens: bad sizes");
    CASMTR_REQUIRE((src && dst) || B == 0, CASMTR_E_INVALID, "nchw_to_tokens: null pointer");
    TransposeJobs jobs;
    jobs.n = 1;
    jobs.job[0] = TransposeJob{src, dst, C, HW, 0};
    return launch_transpose_jobs(jobs, B, (cudaStream_t)stream);
}

int casmtr_score5d_bwd(const float *grad_out, const float *query, const float *key, const int64_t *index,
                       float *grad_query, float *grad_key, int B, int N1, int N2, int H, int D, int K, casmtr_stream_t stream) {
    CASMTR_REQUIRE(B >= 0 && N1 >= 0 && N2 > 0 && H > 0 && D > 0 && K > 0, CASMTR_E_INVALID, "score5d_bwd: bad sizes");
    CASMTR_REQUIRE(grad_key != nullptr || B == 0, CASMTR_E_INVALID, "score5d_bwd: null grad_key");
    CASMTR_REQUIRE((grad_out && query && key && index && grad_query) || (size_t)B * N1 == 0, CASMTR_E_INVALID, "score5d_bwd: null pointer");
    return launch_score5d_bwd(grad_out, query, key, index, grad_query, grad_key, B, N1, N2, H, D, K, (cudaStream_t)stream);
}

int casmtr_value_agg_bwd(const float *grad_out, const float *score, const float *value, const int64_t *index,
                         float *grad_score, float *grad_value, int B, int N, int K, int H, int M, int D, casmtr_stream_t stream) {
    CASMTR_REQUIRE(B >= 0 && N >= 0 && K > 0 && H > 0 && M > 0 && D > 0, CASMTR_E_INVALID, "value_agg_bwd: bad sizes");
    CASMTR_REQUIRE(grad_value != nullptr || B == 0, CASMTR_E_INVALID, "value_agg_bwd: null grad_value");
    CASMTR_REQUIRE((grad_out && score && value && index && grad_score) || (size_t)B * N == 0, CASMTR_E_INVALID, "value_agg_bwd: null pointer");
    return launch_value_agg_bwd(grad_out, score, value, index, grad_score, grad_value, B, N, K, H, M, D, (cudaStream_t)stream);
}

int casmtr_score3d_bwd(const float *grad_out, const float *query, const float *key, const int64_t *index,
                       float *grad_query, float *grad_key, int B, int N1, int N2, int C, int K, casmtr_stream_t stream) {
    CASMTR_REQUIRE(B >= 0 && N1 >= 0 && N2 > 0 && C > 0 && K > 0, CASMTR_E_INVALID, "score3d_bwd: bad sizes");
    CASMTR_REQUIRE(grad_key != nullptr || B == 0, CASMTR_E_INVALID, "score3d_bwd: null grad_key");
    CASMTR_REQUIRE((grad_out && query && key && index && grad_query) || (size_t)B * N1 == 0, CASMTR_E_INVALID, "score3d_bwd: null pointer");
    return launch_score3d_bwd(grad_out, query, key, index, grad_query, grad_key, B, N1, N2, C, K, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ QTAtt
static int check_qtatt_desc(const casmtr_qtatt_desc *d) {
    CASMTR_REQUIRE(d != nullptr, CASMTR_E_INVALID, "qtatt: null descriptor");
    CASMTR_REQUIRE(d->levels >= 1 && d->levels <= CASMTR_MAX_LEVELS, CASMTR_E_INVALID, "qtatt: levels=%d must be in [1,%d]", d->levels, CASMTR_MAX_LEVELS);
    CASMTR_REQUIRE(d->D == 32, CASMTR_E_UNSUPPORTED, "qtatt: head dim %d unsupported, the fused kernels implement D == 32", d->D);
    CASMTR_REQUIRE(d->B >= 1 && d->nhead >= 1, CASMTR_E_INVALID, "qtatt: B=%d nhead=%d", d->B, d->nhead);
    CASMTR_REQUIRE(d->type == 0 || d->type == 1, CASMTR_E_INVALID, "qtatt: type=%d must be 0 (B) or 1 (A)", d->type);
    CASMTR_REQUIRE(d->weight_len == 0 || (d->weight_len >= d->levels && d->weight_len <= 64), CASMTR_E_INVALID,
                   "qtatt: weight_len=%d must be 0 or in [levels=%d, 64]", d->weight_len, d->levels);
    for (int l = 0; l < d->levels; ++l) {
        CASMTR_REQUIRE(d->qh[l] > 0 && d->qw[l] > 0 && d->kh[l] > 0 && d->kw[l] > 0, CASMTR_E_INVALID, "qtatt: empty grid at level %d", l);
        if (l + 1 < d->levels)
            CASMTR_REQUIRE(d->qh[l] == 2 * d->qh[l + 1] && d->qw[l] == 2 * d->qw[l + 1] && d->kh[l] == 2 * d->kh[l + 1] && d->kw[l] == 2 * d->kw[l + 1],
                           CASMTR_E_INVALID, "qtatt: level %d grid must be exactly 2x level %d", l, l + 1);
        CASMTR_REQUIRE(d->topks[l] >= 1 && d->topks[l] <= 32, CASMTR_E_UNSUPPORTED, "qtatt: topks[%d]=%d must be in [1,32]", l, d->topks[l]);
    }
    return CASMTR_OK;
}

struct QtattBuffers {
    float *q[CASMTR_MAX_LEVELS], *k[CASMTR_MAX_LEVELS], *v[CASMTR_MAX_LEVELS];   // token-major, list order
    float *acc[CASMTR_MAX_LEVELS];                                                 // processing order
    float *wsm;                                                                    // softmax of the level weights
    int *tk_idx[CASMTR_MAX_LEVELS];
    float *tk_sc[CASMTR_MAX_LEVELS];
    float *tc_ws;                                                                  // scratch of the tensor-core coarsest level (or NULL)
};

static void carve_qtatt(const casmtr_qtatt_desc *d, Workspace &ws, QtattBuffers &bf) {
    const size_t C = (size_t)d->nhead * d->D;
    bf.wsm = ws.take<float>(CASMTR_MAX_LEVELS);
    {
        const int lc = d->levels - 1;
        const int Sq = d->qh[lc] * d->qw[lc], Sk = d->kh[lc] * d->kw[lc];
        const bool tc = !(d->flags & CASMTR_QT_SIMT_COARSE) && d->D == 32 && coarse_tc_applicable(Sq, Sk, d->topks[0]);
        bf.tc_ws = tc ? ws.take<float>(coarse_tc_workspace_floats(d->B, Sq, Sk, (int)C)) : nullptr;
    }
    for (int l = 0; l < d->levels; ++l) {
        bf.q[l] = ws.take<float>((size_t)d->B * d->qh[l] * d->qw[l] * C);
        bf.k[l] = ws.take<float>((size_t)d->B * d->kh[l] * d->kw[l] * C);
        bf.v[l] = ws.take<float>((size_t)d->B * d->kh[l] * d->kw[l] * C);
    }
    for (int i = 0; i < d->levels; ++i) {
        const int l = d->levels - 1 - i;
        const size_t Lq = (size_t)d->qh[l] * d->qw[l];
        const bool last = i == d->levels - 1;
        bf.acc[i] = last ? nullptr : ws.take<float>((size_t)d->B * Lq * C);
        const bool need_topk = !last || i == 0;      // a single-level call still reports its top-k
        bf.tk_idx[i] = need_topk ? ws.take<int>((size_t)d->B * Lq * d->nhead * d->topks[i]) : nullptr;
        bf.tk_sc[i] = need_topk ? ws.take<float>((size_t)d->B * Lq * d->nhead * d->topks[i]) : nullptr;
    }
}

size_t casmtr_qtatt_workspace_bytes(const casmtr_qtatt_desc *desc) {
    if (check_qtatt_desc(desc) != CASMTR_OK) return 0;
    Workspace ws(nullptr, 0);
    QtattBuffers bf;
    carve_qtatt(desc, ws, bf);
    return ws.off;
}

// the levels themselves, on token-major pyramids bf.q/k/v
static int qtatt_levels(const casmtr_qtatt_desc *d, const QtattBuffers &bf, const float *level_weight, float *out,
                        int64_t *const *topk_idx_out, float *const *topk_score_out, cudaStream_t stream, cudaEvent_t join = nullptr,
                        bool tc_prepped = false);

int casmtr_qtatt_fwd(const casmtr_qtatt_desc *d,
                     const float *const *queries, const float *const *keys, const float *const *values,
                     const float *level_weight, float *out,
                     int64_t *const *topk_idx_out, float *const *topk_score_out,
                     void *workspace, size_t workspace_bytes, casmtr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_qtatt_desc(d);
    if (rc != CASMTR_OK) return rc;
    CASMTR_REQUIRE(queries && keys && values && out && workspace, CASMTR_E_INVALID, "qtatt: null pointer");
    CASMTR_REQUIRE(d->type == 1 || level_weight != nullptr, CASMTR_E_INVALID, "qtatt: QTAttB needs the level weight vector");
    Workspace ws(workspace, workspace_bytes);
    QtattBuffers bf;
    carve_qtatt(d, ws, bf);
    CASMTR_REQUIRE(ws.ok(), CASMTR_E_WORKSPACE, "qtatt: workspace %zu < %zu bytes", workspace_bytes, ws.off);
    const int C = d->nhead * d->D;

    for (int l = 0; l < d->levels; ++l) CASMTR_REQUIRE(queries[l] && keys[l] && values[l], CASMTR_E_INVALID, "qtatt: null level pointer %d", l);
    auto add_level = [&](TransposeJobs &jobs, int l) {
        jobs.job[jobs.n++] = TransposeJob{queries[l], bf.q[l], C, d->qh[l] * d->qw[l], 0};
        jobs.job[jobs.n++] = TransposeJob{keys[l], bf.k[l], C, d->kh[l] * d->kw[l], 0};
        jobs.job[jobs.n++] = TransposeJob{values[l], bf.v[l], C, d->kh[l] * d->kw[l], 0};
    };
    TransposeJobs jobs;
    jobs.n = 0;
    // The dense coarsest level needs only the coarsest maps (1/16 of the finest): with a side lane the finer levels' maps
    // (95 % of the bytes, HBM-bound) are re-laid under the coarsest level's kernel (issue-bound) and joined before the first
    // fine level.  Without a lane: one batched launch on the caller's stream.
    LaneGuard guard = side_lane(d->levels >= 2 && !(d->flags & CASMTR_QT_NO_OVERLAP));      // holds the device's lane mutex until the call returns
    SideLane *lane = guard.lane;
    ConcurrencyScope cc(d->concurrent_calls);
    if (lane && (cudaEventRecord(lane->fork, stream) != cudaSuccess || cudaStreamWaitEvent(lane->stream, lane->fork, 0) != cudaSuccess)) {
        cudaGetLastError();
        lane = nullptr;
    }
    if (lane) {
        for (int l = 0; l + 1 < d->levels; ++l) add_level(jobs, l);
        rc = launch_transpose_jobs(jobs, d->B, lane->stream);
        // the join is recorded and waited for even after a failed launch: a capture must never be left with a dangling fork
        const bool joined = cudaEventRecord(lane->join, lane->stream) == cudaSuccess;
        if (rc != CASMTR_OK || !joined) {
            if (joined) cudaStreamWaitEvent(stream, lane->join, 0);
            if (rc == CASMTR_OK) { casmtr_set_error("qtatt: cudaEventRecord on the side stream failed"); rc = CASMTR_E_CUDA; }
            return rc;
        }
        jobs.n = 0;
        add_level(jobs, d->levels - 1);
    } else {
        for (int l = 0; l < d->levels; ++l) add_level(jobs, l);
    }
    rc = launch_transpose_jobs(jobs, d->B, stream);
    if (rc != CASMTR_OK) {
        if (lane) cudaStreamWaitEvent(stream, lane->join, 0);
        return rc;
    }
    return qtatt_levels(d, bf, level_weight, out, topk_idx_out, topk_score_out, stream, lane ? lane->join : nullptr);
}

int casmtr_qtatt_tokens_fwd(const casmtr_qtatt_desc *d, const float *q0, const float *k0, const float *v0,
                            const float *level_weight, float *out,
                            int64_t *const *topk_idx_out, float *const *topk_score_out,
                            void *workspace, size_t workspace_bytes, casmtr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_qtatt_desc(d);
    if (rc != CASMTR_OK) return rc;
    CASMTR_REQUIRE(q0 && k0 && v0 && out && workspace, CASMTR_E_INVALID, "qtatt_tokens: null pointer");
    CASMTR_REQUIRE(d->type == 1 || level_weight != nullptr, CASMTR_E_INVALID, "qtatt_tokens: QTAttB needs the level weight vector");
    for (int l = 1; l < d->levels; ++l)
        CASMTR_REQUIRE(d->qh[l] == d->qh[l - 1] / 2 && d->qw[l] == d->qw[l - 1] / 2 && d->kh[l] == d->kh[l - 1] / 2 && d->kw[l] == d->kw[l - 1] / 2,
                       CASMTR_E_INVALID, "qtatt_tokens: level %d is not the 2x2 pooling of level %d", l, l - 1);
    Workspace ws(workspace, workspace_bytes);
    QtattBuffers bf;
    carve_qtatt(d, ws, bf);
    CASMTR_REQUIRE(ws.ok(), CASMTR_E_WORKSPACE, "qtatt_tokens: workspace %zu < %zu bytes", workspace_bytes, ws.off);
    const int C = d->nhead * d->D;
    bf.q[0] = const_cast<float *>(q0); bf.k[0] = const_cast<float *>(k0); bf.v[0] = const_cast<float *>(v0);   // read in place
    ConcurrencyScope cc(d->concurrent_calls);
    bool tc_prepped = false;
    for (int l = 1; l < d->levels; ++l) {
        PoolJobs pj;
        pj.n = 3;
        const bool two = l + 1 < d->levels && C % 32 == 0 && d->qh[l - 1] % 4 == 0 && d->qw[l - 1] % 4 == 0 && d->kh[l - 1] % 4 == 0 && d->kw[l - 1] % 4 == 0;
        pj.job[0] = PoolJob{bf.q[l - 1], bf.q[l], d->qh[l - 1], d->qw[l - 1], two ? bf.q[l + 1] : nullptr};
        pj.job[1] = PoolJob{bf.k[l - 1], bf.k[l], d->kh[l - 1], d->kw[l - 1], two ? bf.k[l + 1] : nullptr};
        pj.job[2] = PoolJob{bf.v[l - 1], bf.v[l], d->kh[l - 1], d->kw[l - 1], two ? bf.v[l + 1] : nullptr};
        if (two && l + 1 == d->levels - 1 && bf.tc_ws) {
            // the pass that produces the coarsest maps also leaves the tensor-core level's split operands next to them
            const int lc = d->levels - 1;
            const CoarseTcOperands op = coarse_tc_operands(bf.tc_ws, d->B, d->qh[lc] * d->qw[lc], d->kh[lc] * d->kw[lc], C);
            pj.job[0].lo2 = op.q_lo;
            pj.job[1].lo2 = op.k_lo;
            pj.job[2].vt_hi = op.vt_hi; pj.job[2].vt_lo = op.vt_lo; pj.job[2].Sp = op.Sp;
            tc_prepped = true;
        }
        rc = two ? launch_pool2_tokens(pj, d->B, C, stream) : launch_pool_tokens(pj, d->B, C, stream);
        if (rc != CASMTR_OK) return rc;
        if (two) ++l;
    }
    return qtatt_levels(d, bf, level_weight, out, topk_idx_out, topk_score_out, stream, nullptr, tc_prepped);
}

static int qtatt_levels_impl(const casmtr_qtatt_desc *d, const QtattBuffers &bf, const float *level_weight, float *out,
                             int64_t *const *topk_idx_out, float *const *topk_score_out, cudaStream_t stream, cudaEvent_t &join, bool tc_prepped);

static int qtatt_levels(const casmtr_qtatt_desc *d, const QtattBuffers &bf, const float *level_weight, float *out,
                        int64_t *const *topk_idx_out, float *const *topk_score_out, cudaStream_t stream, cudaEvent_t join, bool tc_prepped) {
    const int rc = qtatt_levels_impl(d, bf, level_weight, out, topk_idx_out, topk_score_out, stream, join, tc_prepped);
    if (join) cudaStreamWaitEvent(stream, join, 0);     // an early error return must not leave the side lane un-joined
    return rc;
}

static int qtatt_levels_impl(const casmtr_qtatt_desc *d, const QtattBuffers &bf, const float *level_weight, float *out,
                             int64_t *const *topk_idx_out, float *const *topk_score_out, cudaStream_t stream, cudaEvent_t &join, bool tc_prepped) {
    int rc = CASMTR_OK;
    const float *wts = d->type == 0 ? level_weight : nullptr;
    for (int i = 0; i < d->levels; ++i) {
        if (join && i == 1) {                           // the finer levels' maps come from the side lane (casmtr_qtatt_fwd)
            if (cudaStreamWaitEvent(stream, join, 0) != cudaSuccess) { casmtr_set_error("qtatt: cudaStreamWaitEvent(join) failed"); return CASMTR_E_CUDA; }
            join = nullptr;
        }
        const int l = d->levels - 1 - i;
        const bool last = i == d->levels - 1;
        float *dst = last ? out : bf.acc[i];
        if (i == 0) {
            CoarseParams cp;
            cp.q = bf.q[l]; cp.k = bf.k[l]; cp.v = bf.v[l];
            cp.acc = dst; cp.topk_idx = bf.tk_idx[0]; cp.topk_score = bf.tk_sc[0];
            cp.level_weight = wts; cp.levels = d->levels; cp.n_weights = d->weight_len > 0 ? d->weight_len : d->levels; cp.wsm = wts ? bf.wsm : nullptr;
            cp.B = d->B; cp.Sq = d->qh[l] * d->qw[l]; cp.Sk = d->kh[l] * d->kw[l];
            cp.nh = d->nhead; cp.topk = d->topks[0]; cp.type_a = d->type; cp.tc_ws = bf.tc_ws; cp.tc_prepped = tc_prepped;
            rc = launch_qtatt_coarse(cp, stream);
        } else {
            FineParams fp;
            memset(&fp, 0, sizeof(fp));
            fp.q = bf.q[l]; fp.k = bf.k[l]; fp.v = bf.v[l];
            fp.prev_idx = bf.tk_idx[i - 1]; fp.prev_score = bf.tk_sc[i - 1];
            fp.acc_prev = bf.acc[i - 1]; fp.out = dst;
            fp.topk_idx = last ? nullptr : bf.tk_idx[i];
            fp.topk_score = last ? nullptr : bf.tk_sc[i];
            fp.wsm = wts ? bf.wsm : nullptr; fp.level = i;
            fp.B = d->B; fp.nh = d->nhead;
            fp.h0 = d->qh[l]; fp.w0 = d->qw[l]; fp.h1 = d->kh[l]; fp.w1 = d->kw[l]; fp.w_prev = d->kw[l + 1];
            fp.kp = d->topks[i - 1]; fp.topk = d->topks[i]; fp.dil = 1;
            fp.type_a = d->type; fp.final_level = last;
            rc = launch_quad_attention(fp, stream);
        }
        if (rc != CASMTR_OK) return rc;
        if (bf.tk_idx[i] && ((topk_idx_out && topk_idx_out[i]) || (topk_score_out && topk_score_out[i]))) {
            rc = launch_topk_to_api(bf.tk_idx[i], bf.tk_sc[i], topk_idx_out ? topk_idx_out[i] : nullptr,
                                    topk_score_out ? topk_score_out[i] : nullptr,
                                    (size_t)d->B * d->qh[l] * d->qw[l], d->nhead, d->topks[i], stream);
            if (rc != CASMTR_OK) return rc;
        }
    }
    return CASMTR_OK;
}

// ------------------------------------------------------------------------------------------------ guided quadtree level
static void carve_guided(Workspace &ws, int B, int C, int nhead, int h0, int w0, int h1, int w1, int K, float **q, float **k, float **v, int **idx, float **wsm) {
    *q = ws.take<float>((size_t)B * h0 * w0 * C);
    *k = ws.take<float>((size_t)B * h1 * w1 * C);
    *v = ws.take<float>((size_t)B * h1 * w1 * C);
    *idx = ws.take<int>((size_t)B * (h0 / 2) * (w0 / 2) * nhead * K);
    *wsm = ws.take<float>(CASMTR_MAX_LEVELS);
}

size_t casmtr_qtatt_guided_workspace_bytes(int B, int C, int nhead, int h0, int w0, int h1, int w1, int K) {
    if (B <= 0 || C <= 0 || nhead <= 0 || h0 <= 0 || w0 <= 0 || h1 <= 0 || w1 <= 0 || K <= 0) return 0;
    Workspace ws(nullptr, 0);
    float *q, *k, *v, *wsm;
    int *idx;
    carve_guided(ws, B, C, nhead, h0, w0, h1, w1, K, &q, &k, &v, &idx, &wsm);
    return ws.off;
}

int casmtr_qtatt_guided_fwd(const float *query, const float *key, const float *value, const int64_t *topk_pos,
                            const float *level_weight, int weight_len, float *out,
                            int B, int nhead, int D, int h0, int w0, int h1, int w1, int K,
                            void *workspace, size_t workspace_bytes, casmtr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CASMTR_REQUIRE(D == 32, CASMTR_E_UNSUPPORTED, "qtatt_guided: head dim %d unsupported (D == 32)", D);
    CASMTR_REQUIRE(B >= 1 && nhead >= 1 && h0 > 0 && w0 > 0 && h1 > 0 && w1 > 0, CASMTR_E_INVALID, "qtatt_guided: bad sizes");
    CASMTR_REQUIRE(h0 % 2 == 0 && w0 % 2 == 0 && h1 % 2 == 0 && w1 % 2 == 0, CASMTR_E_INVALID, "qtatt_guided: grids %dx%d / %dx%d must be even", h0, w0, h1, w1);
    CASMTR_REQUIRE(K >= 1 && K <= 32, CASMTR_E_UNSUPPORTED, "qtatt_guided: K=%d must be in [1,32]", K);
    CASMTR_REQUIRE(weight_len >= 1 && weight_len <= CASMTR_MAX_LEVELS, CASMTR_E_INVALID, "qtatt_guided: weight_len=%d must be in [1,%d]", weight_len, CASMTR_MAX_LEVELS);
    CASMTR_REQUIRE(query && key && value && topk_pos && level_weight && out && workspace, CASMTR_E_INVALID, "qtatt_guided: null pointer");
    const int C = nhead * D;
    Workspace ws(workspace, workspace_bytes);
    float *qt, *kt, *vt, *wsm;
    int *idx;
    carve_guided(ws, B, C, nhead, h0, w0, h1, w1, K, &qt, &kt, &vt, &idx, &wsm);
    CASMTR_REQUIRE(ws.ok(), CASMTR_E_WORKSPACE, "qtatt_guided: workspace %zu < %zu bytes", workspace_bytes, ws.off);
    TransposeJobs jobs;
    jobs.n = 3;
    jobs.job[0] = TransposeJob{query, qt, C, h0 * w0, 0};
    jobs.job[1] = TransposeJob{key, kt, C, h1 * w1, 0};
    jobs.job[2] = TransposeJob{value, vt, C, h1 * w1, 0};
    int rc = launch_transpose_jobs(jobs, B, stream);
    if (rc != CASMTR_OK) return rc;
    rc = launch_guided_prep(topk_pos, idx, (size_t)B * (h0 / 2) * (w0 / 2), K, nhead, h1 / 2, w1 / 2, level_weight, weight_len, wsm, stream);
    if (rc != CASMTR_OK) return rc;
    FineParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.q = qt; fp.k = kt; fp.v = vt;
    fp.prev_idx = idx; fp.out = out;
    fp.wsm = wsm; fp.level = 0;
    fp.B = B; fp.nh = nhead; fp.h0 = h0; fp.w0 = w0; fp.h1 = h1; fp.w1 = w1; fp.w_prev = w1 / 2;
    fp.kp = K; fp.topk = 0; fp.dil = 1; fp.final_level = 1;
    return launch_quad_attention(fp, stream);
}

// ------------------------------------------------------------------------------------------------ cascade attention
size_t casmtr_cascade_qtatt_workspace_bytes(int B, int C, int h0, int w0, int h1, int w1) {
    if (B <= 0 || C <= 0 || h0 <= 0 || w0 <= 0 || h1 <= 0 || w1 <= 0) return 0;
    Workspace ws(nullptr, 0);
    ws.take<float>((size_t)B * h0 * w0 * C);
    ws.take<float>((size_t)B * h1 * w1 * C);
    ws.take<float>((size_t)B * h1 * w1 * C);
    ws.take<int>((size_t)B * (h0 / 2) * (w0 / 2) + 1);       // fallback cell list + count of the tile kernel
    return ws.off;
}

static int cascade_qtatt_impl(bool token_major, const float *query, const float *key, const float *value,
                              const int64_t *topk_pos, const float *rel_pos, float *message, int64_t *upsampled_idx,
                              int B, int nhead, int D, int h0, int w0, int h1, int w1, int k, int dilated,
                              void *workspace, size_t workspace_bytes, cudaStream_t stream, const int64_t *next_idx = nullptr, int win = 0,
                              const RelPE *pe = nullptr);

int casmtr_window_idx_fwd(const int64_t *next_idx, int64_t *pos, int B, int L, int H, int W, int window, casmtr_stream_t stream) {
    CASMTR_REQUIRE(B >= 0 && L >= 0 && window >= 1 && (window & 1) && H >= window && W >= window, CASMTR_E_INVALID,
                   "window_idx: window %d must be odd and fit the %dx%d grid", window, H, W);
    CASMTR_REQUIRE((size_t)B * L == 0 || (next_idx && pos), CASMTR_E_INVALID, "window_idx: null pointer");
    return launch_window_idx(next_idx, pos, (size_t)B * L, H, W, window, (cudaStream_t)stream);
}

int casmtr_cascade_qtatt_window_fwd(const float *query, const float *key, const float *value,
                                    const int64_t *next_idx, int window, const float *rel_pos,
                                    float *message, int64_t *upsampled_idx,
                                    int B, int nhead, int D, int h0, int w0, int h1, int w1, int token_major,
                                    void *workspace, size_t workspace_bytes, casmtr_stream_t stream) {
    CASMTR_REQUIRE(window >= 1 && (window & 1) && window * window <= 32, CASMTR_E_UNSUPPORTED, "cascade_qtatt_window: window %d must be 1, 3 or 5", window);
    CASMTR_REQUIRE(h1 % 2 == 0 && w1 % 2 == 0 && h1 / 2 >= window && w1 / 2 >= window, CASMTR_E_INVALID,
                   "cascade_qtatt_window: the %dx%d key grid must be even and its parent grid must hold a %d-window", h1, w1, window);
    CASMTR_REQUIRE(next_idx != nullptr, CASMTR_E_INVALID, "cascade_qtatt_window: null next_idx");
    return cascade_qtatt_impl(token_major != 0, query, key, value, nullptr, rel_pos, message, upsampled_idx, B, nhead, D, h0, w0, h1, w1,
                              window * window, 1, workspace, workspace_bytes, (cudaStream_t)stream, next_idx, window);
}

// casmtr_relpe_desc -> RelPE for an (h0 x w0) query level whose keys live on rows of w1 tokens
static int relpe_from_desc(const casmtr_relpe_desc *d, int B, int nhead, int h0, int w0, int w1, RelPE *pe) {
    CASMTR_REQUIRE(d != nullptr && d->w_table && d->h_table && d->tgt_idx, CASMTR_E_INVALID, "relative_pe: null descriptor / table / tgt_idx");
    CASMTR_REQUIRE(B >= 1 && nhead >= 1 && d->h8 >= 1 && d->w8 >= 1 && d->w8_other >= 1 && d->n_emb >= 1 && d->LB >= 0, CASMTR_E_INVALID, "relative_pe: bad sizes");
    CASMTR_REQUIRE(h0 % d->h8 == 0 && w0 == d->w8 * (h0 / d->h8), CASMTR_E_INVALID,
                   "relative_pe: the %dx%d query level is not a multiple of the %dx%d 1/8 grid", h0, w0, d->h8, d->w8);
    CASMTR_REQUIRE((h0 / d->h8) % 2 == 0, CASMTR_E_UNSUPPORTED, "relative_pe: level / 1/8-grid ratio %d must be even (2 at 1/4, 4 at 1/2)", h0 / d->h8);
    CASMTR_REQUIRE(w1 == 0 || w1 == d->w8_other * (h0 / d->h8), CASMTR_E_INVALID,
                   "relative_pe: key rows of %d tokens do not match w8_other * s = %d", w1, d->w8_other * (h0 / d->h8));
    pe->w_tab = d->w_table; pe->h_tab = d->h_table; pe->tgt_idx = d->tgt_idx;
    pe->n_emb = d->n_emb; pe->LB = d->LB; pe->s = h0 / d->h8; pe->h8 = d->h8; pe->w8 = d->w8; pe->w8o = d->w8_other;
    return CASMTR_OK;
}

int casmtr_relative_pe_fwd(const casmtr_relpe_desc *d, const int64_t *window_pos, float *rel_pos,
                           int B, int nhead, int h0, int w0, int k, casmtr_stream_t stream) {
    CASMTR_REQUIRE(h0 > 0 && w0 > 0 && h0 % 2 == 0 && w0 % 2 == 0 && k >= 1, CASMTR_E_INVALID, "relative_pe: the %dx%d level must be even, k >= 1", h0, w0);
    CASMTR_REQUIRE(window_pos && rel_pos, CASMTR_E_INVALID, "relative_pe: null pointer");
    RelPE pe;
    const int rc = relpe_from_desc(d, B, nhead, h0, w0, 0, &pe);
    if (rc != CASMTR_OK) return rc;
    return launch_relative_pe(pe, window_pos, rel_pos, B, nhead, h0, w0, k, (cudaStream_t)stream);
}

int casmtr_cascade_qtatt_relpe_fwd(const float *query, const float *key, const float *value,
                                   const int64_t *topk_pos, const int64_t *next_idx, int window, const casmtr_relpe_desc *d,
                                   float *message, int64_t *upsampled_idx,
                                   int B, int nhead, int D, int h0, int w0, int h1, int w1, int k, int token_major,
                                   void *workspace, size_t workspace_bytes, casmtr_stream_t stream) {
    CASMTR_REQUIRE((topk_pos != nullptr) != (next_idx != nullptr), CASMTR_E_INVALID, "cascade_qtatt_relpe: give topk_pos or next_idx, not both");
    if (next_idx) {
        CASMTR_REQUIRE(window >= 1 && (window & 1) && window * window <= 32, CASMTR_E_UNSUPPORTED, "cascade_qtatt_relpe: window %d must be 1, 3 or 5", window);
        CASMTR_REQUIRE(h1 % 2 == 0 && w1 % 2 == 0 && h1 / 2 >= window && w1 / 2 >= window, CASMTR_E_INVALID,
                       "cascade_qtatt_relpe: the %dx%d key grid must be even and its parent grid must hold a %d-window", h1, w1, window);
        k = window * window;
    }
    CASMTR_REQUIRE(h0 > 0 && w0 > 0 && w1 > 0, CASMTR_E_INVALID, "cascade_qtatt_relpe: bad sizes");
    RelPE pe;
    const int rc = relpe_from_desc(d, B, nhead, h0, w0, w1, &pe);
    if (rc != CASMTR_OK) return rc;
    return cascade_qtatt_impl(token_major != 0, query, key, value, topk_pos, nullptr, message, upsampled_idx, B, nhead, D, h0, w0, h1, w1,
                              k, 1, workspace, workspace_bytes, (cudaStream_t)stream, next_idx, next_idx ? window : 0, &pe);
}

int casmtr_cascade_qtatt_fwd(const float *query, const float *key, const float *value,
                             const int64_t *topk_pos, const float *rel_pos,
                             float *message, int64_t *upsampled_idx,
                             int B, int nhead, int D, int h0, int w0, int h1, int w1, int k, int dilated,
                             void *workspace, size_t workspace_bytes, casmtr_stream_t stream) {
    return cascade_qtatt_impl(false, query, key, value, topk_pos, rel_pos, message, upsampled_idx, B, nhead, D, h0, w0, h1, w1, k, dilated,
                              workspace, workspace_bytes, (cudaStream_t)stream);
}

int casmtr_cascade_qtatt_tokens_fwd(const float *query, const float *key, const float *value,
                                    const int64_t *topk_pos, const float *rel_pos,
                                    float *message, int64_t *upsampled_idx,
                                    int B, int nhead, int D, int h0, int w0, int h1, int w1, int k, int dilated,
                                    void *workspace, size_t workspace_bytes, casmtr_stream_t stream) {
    return cascade_qtatt_impl(true, query, key, value, topk_pos, rel_pos, message, upsampled_idx, B, nhead, D, h0, w0, h1, w1, k, dilated,
                              workspace, workspace_bytes, (cudaStream_t)stream);
}

static int cascade_qtatt_impl(bool token_major, const float *query, const float *key, const float *value,
                              const int64_t *topk_pos, const float *rel_pos, float *message, int64_t *upsampled_idx,
                              int B, int nhead, int D, int h0, int w0, int h1, int w1, int k, int dilated,
                              void *workspace, size_t workspace_bytes, cudaStream_t stream, const int64_t *next_idx, int win, const RelPE *pe) {
    CASMTR_REQUIRE(D == 32, CASMTR_E_UNSUPPORTED, "cascade_qtatt: head dim %d unsupported (D == 32)", D);
    CASMTR_REQUIRE(B >= 1 && nhead >= 1 && h0 > 0 && w0 > 0 && h1 > 0 && w1 > 0, CASMTR_E_INVALID, "cascade_qtatt: bad sizes");
    CASMTR_REQUIRE(h0 % 2 == 0 && w0 % 2 == 0, CASMTR_E_INVALID, "cascade_qtatt: query grid %dx%d must be even", h0, w0);
    CASMTR_REQUIRE(k >= 1 && k <= 32, CASMTR_E_UNSUPPORTED, "cascade_qtatt: window size k=%d must be in [1,32]", k);
    CASMTR_REQUIRE(dilated >= 1, CASMTR_E_INVALID, "cascade_qtatt: dilated=%d", dilated);
    CASMTR_REQUIRE(query && key && value && (topk_pos || next_idx) && message && workspace, CASMTR_E_INVALID, "cascade_qtatt: null pointer");
    const int C = nhead * D;
    Workspace ws(workspace, workspace_bytes);
    const float *qt = query, *kt = key, *vt = value;
    if (!token_major) {
        float *q_ = ws.take<float>((size_t)B * h0 * w0 * C);
        float *k_ = ws.take<float>((size_t)B * h1 * w1 * C);
        float *v_ = ws.take<float>((size_t)B * h1 * w1 * C);
        qt = q_; kt = k_; vt = v_;
    }
    int *fb = ws.take<int>((size_t)B * (h0 / 2) * (w0 / 2) + 1);
    CASMTR_REQUIRE(ws.ok(), CASMTR_E_WORKSPACE, "cascade_qtatt: workspace %zu < %zu bytes", workspace_bytes, ws.off);
    int rc = CASMTR_OK;
    if (!token_major) {
        TransposeJobs jobs;
        jobs.n = 3;
        jobs.job[0] = TransposeJob{query, const_cast<float *>(qt), C, h0 * w0, 0};
        jobs.job[1] = TransposeJob{key, const_cast<float *>(kt), C, h1 * w1, 0};
        jobs.job[2] = TransposeJob{value, const_cast<float *>(vt), C, h1 * w1, 0};
        rc = launch_transpose_jobs(jobs, B, stream);
        if (rc != CASMTR_OK) return rc;
    }
    FineParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.q = qt; fp.k = kt; fp.v = vt;
    fp.topk_pos = topk_pos; fp.next_idx = next_idx; fp.win = win; fp.rel_pos = rel_pos;
    if (pe) fp.pe = *pe;
    fp.out = message; fp.upsampled_idx = upsampled_idx;
    fp.B = B; fp.nh = nhead; fp.h0 = h0; fp.w0 = w0; fp.h1 = h1; fp.w1 = w1;
    fp.kp = k; fp.dil = dilated;
    if (k == 25 && dilated == 1 && ((uintptr_t)topk_pos & 15) == 0) {
        // regular 5x5 windows: TMA-tiled kernel for the coherent cells, gather kernel for the listed outliers
        int *fb_count = fb + (size_t)B * (h0 / 2) * (w0 / 2);
        rc = launch_cascade_att_tile(qt, kt, vt, topk_pos, next_idx, rel_pos, fp.pe, message, upsampled_idx, fb, fb_count, B, nhead, h0, w0, h1, w1, stream);
        if (rc != CASMTR_OK) return rc;
        fp.item_list = fb; fp.item_count = fb_count;
    }
    return launch_quad_attention(fp, stream);
}

// ------------------------------------------------------------------------------------------------ cascade matching
size_t casmtr_cascade_match_workspace_bytes(int B, int L0, int L1) {
    if (B <= 0 || L0 <= 0 || L1 <= 0) return 0;
    Workspace ws(nullptr, 0);
    ws.take<int>((size_t)B * (L0 / 4 + L1 / 4) + 2);
    return ws.off;
}

int casmtr_cascade_match_fwd(const float *feat0, const float *feat1,
                             const int64_t *idx01, const int64_t *idx10,
                             const uint8_t *mask0, const uint8_t *mask1, float temperature,
                             float *conf01, float *next_conf01, int64_t *next_idx01,
                             float *conf10, float *next_conf10, int64_t *next_idx10,
                             int B, int L0, int L1, int C, int K, int w0, int w1,
                             void *workspace, size_t workspace_bytes, casmtr_stream_t stream) {
    CASMTR_REQUIRE(w0 >= 0 && w1 >= 0 && (w0 == 0 || L0 % w0 == 0) && (w1 == 0 || L1 % w1 == 0), CASMTR_E_INVALID,
                   "cascade_match: grid widths %d / %d do not divide L0=%d / L1=%d", w0, w1, L0, L1);
    CASMTR_REQUIRE(B >= 1 && L0 > 0 && L1 > 0 && C > 0 && K > 0, CASMTR_E_INVALID, "cascade_match: bad sizes");
    CASMTR_REQUIRE(feat0 && feat1 && idx01 && idx10 && next_conf01 && next_idx01 && next_conf10 && next_idx10,
                   CASMTR_E_INVALID, "cascade_match: null pointer");
    CASMTR_REQUIRE((mask0 == nullptr) == (mask1 == nullptr), CASMTR_E_INVALID, "cascade_match: give both masks or neither");
    CASMTR_REQUIRE(temperature > 0.f, CASMTR_E_INVALID, "cascade_match: temperature must be positive");
    MatchParams p;
    p.feat0 = feat0; p.feat1 = feat1; p.idx01 = idx01; p.idx10 = idx10; p.mask0 = mask0; p.mask1 = mask1;
    p.inv_scale = 1.0f / ((float)C * temperature);
    p.conf01 = conf01; p.conf10 = conf10; p.next_conf01 = next_conf01; p.next_conf10 = next_conf10;
    p.next_idx01 = next_idx01; p.next_idx10 = next_idx10;
    p.B = B; p.L0 = L0; p.L1 = L1; p.C = C; p.K = K; p.w0 = w0; p.w1 = w1;
    p.fb_list = p.fb_count = nullptr; p.cell_list = p.cell_count = nullptr;
    if (workspace != nullptr && workspace_bytes >= casmtr_cascade_match_workspace_bytes(B, L0, L1)) {
        p.fb_list = (int *)workspace;
        p.fb_count = p.fb_list + (size_t)B * (L0 / 4 + L1 / 4) + 1;
    }
    return launch_cascade_match(p, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ dense coarse matching
size_t casmtr_coarse_match_workspace_bytes(int B, int L0, int L1, int C) {
    if (B <= 0 || L0 <= 0 || L1 <= 0 || C <= 0) return 0;
    return coarse_match_workspace(B, L0, L1, C);
}

int casmtr_coarse_match_fwd(const float *feat0, const float *feat1, float temperature,
                            float *next_conf01, int64_t *next_idx01, float *next_conf10, int64_t *next_idx10,
                            int B, int L0, int L1, int C, void *workspace, size_t workspace_bytes, casmtr_stream_t stream) {
    return casmtr_coarse_match_masked_fwd(feat0, feat1, nullptr, nullptr, temperature, next_conf01, next_idx01, next_conf10, next_idx10,
                                          B, L0, L1, C, workspace, workspace_bytes, stream);
}

int casmtr_coarse_match_masked_fwd(const float *feat0, const float *feat1, const uint8_t *mask0, const uint8_t *mask1, float temperature,
                                   float *next_conf01, int64_t *next_idx01, float *next_conf10, int64_t *next_idx10,
                                   int B, int L0, int L1, int C, void *workspace, size_t workspace_bytes, casmtr_stream_t stream) {
    CASMTR_REQUIRE(B >= 1 && L0 > 0 && L1 > 0 && C > 0, CASMTR_E_INVALID, "coarse_match: bad sizes");
    CASMTR_REQUIRE((mask0 == nullptr) == (mask1 == nullptr), CASMTR_E_INVALID, "coarse_match: give both masks or neither");
    CASMTR_REQUIRE(feat0 && feat1 && next_conf01 && next_idx01 && next_conf10 && next_idx10 && workspace, CASMTR_E_INVALID,
                   "coarse_match: null pointer");
    CASMTR_REQUIRE(temperature > 0.f, CASMTR_E_INVALID, "coarse_match: temperature must be positive");
    CASMTR_REQUIRE(2 * B <= 65535, CASMTR_E_UNSUPPORTED, "coarse_match: batch too large");
    return launch_coarse_match(feat0, feat1, mask0, mask1, temperature, next_conf01, next_idx01, next_conf10, next_idx10, B, L0, L1, C,
                               workspace, workspace_bytes, (cudaStream_t)stream);
}

int casmtr_coarse_match_mutual_fwd(const float *feat0, const float *feat1, const uint8_t *mask0, const uint8_t *mask1, float temperature,
                                   float *next_conf01, int64_t *next_idx01, float *next_conf10, int64_t *next_idx10,
                                   float *mconf_row, int64_t *midx_row, int64_t *midx_col,
                                   int B, int L0, int L1, int C, void *workspace, size_t workspace_bytes, casmtr_stream_t stream) {
    CASMTR_REQUIRE(B >= 1 && L0 > 0 && L1 > 0 && C > 0, CASMTR_E_INVALID, "coarse_match_mutual: bad sizes");
    CASMTR_REQUIRE((mask0 == nullptr) == (mask1 == nullptr), CASMTR_E_INVALID, "coarse_match_mutual: give both masks or neither");
    CASMTR_REQUIRE(feat0 && feat1 && next_conf01 && next_idx01 && next_conf10 && next_idx10 && mconf_row && midx_row && midx_col && workspace,
                   CASMTR_E_INVALID, "coarse_match_mutual: null pointer");
    CASMTR_REQUIRE(temperature > 0.f, CASMTR_E_INVALID, "coarse_match_mutual: temperature must be positive");
    CASMTR_REQUIRE(2 * B <= 65535, CASMTR_E_UNSUPPORTED, "coarse_match_mutual: batch too large");
    return launch_coarse_match(feat0, feat1, mask0, mask1, temperature, next_conf01, next_idx01, next_conf10, next_idx10, B, L0, L1, C,
                               workspace, workspace_bytes, (cudaStream_t)stream, mconf_row, midx_row, midx_col);
}

// ------------------------------------------------------------------------------------------------ extraction
static int check_extract_desc(const casmtr_extract_desc *d) {
    CASMTR_REQUIRE(d != nullptr, CASMTR_E_INVALID, "match_extract: null descriptor");
    CASMTR_REQUIRE(d->B >= 1 && d->h0 > 0 && d->w0 > 0 && d->h1 > 0 && d->w1 > 0, CASMTR_E_INVALID, "match_extract: bad sizes");
    CASMTR_REQUIRE(d->nms_window == 0 || (d->nms_window % 2 == 1 && d->nms_window <= 15), CASMTR_E_UNSUPPORTED,
                   "match_extract: nms_window=%d must be 0 or odd <= 15", d->nms_window);
    CASMTR_REQUIRE(d->n_pre >= 0 && d->n_pre <= 2, CASMTR_E_INVALID, "match_extract: n_pre=%d", d->n_pre);
    for (int s = 0; s < d->n_pre; ++s)
        CASMTR_REQUIRE(d->pre_conf[s] && d->pre_h[s] > 0 && d->pre_w[s] > 0, CASMTR_E_INVALID, "match_extract: bad previous-stage gate %d", s);
    CASMTR_REQUIRE((d->pad_mask0 == nullptr) == (d->pad_mask1 == nullptr), CASMTR_E_INVALID, "match_extract: give both pad masks or neither");
    CASMTR_REQUIRE((size_t)d->B * d->h0 * d->w0 < 0x7fffffffu, CASMTR_E_UNSUPPORTED, "match_extract: B*L0 too large");
    return CASMTR_OK;
}

size_t casmtr_match_extract_workspace_bytes(const casmtr_extract_desc *desc) {
    if (check_extract_desc(desc) != CASMTR_OK) return 0;
    return match_extract_workspace(*desc);
}

int casmtr_match_extract(const casmtr_extract_desc *desc,
                         const float *next_conf01, const int64_t *next_idx01, const int64_t *next_idx10,
                         uint8_t *mask_out, int64_t *b_ids, int64_t *i_ids, int64_t *j_ids,
                         float *mconf, float *mkpts0, float *mkpts1,
                         int capacity, int32_t *count_out,
                         void *workspace, size_t workspace_bytes, casmtr_stream_t stream) {
    int rc = check_extract_desc(desc);
    if (rc != CASMTR_OK) return rc;
    CASMTR_REQUIRE(next_conf01 && next_idx01 && next_idx10 && b_ids && i_ids && j_ids && mconf && mkpts0 && mkpts1 && count_out && workspace,
                   CASMTR_E_INVALID, "match_extract: null pointer");
    CASMTR_REQUIRE(capacity >= desc->B, CASMTR_E_INVALID, "match_extract: capacity %d < B=%d", capacity, desc->B);
    return launch_match_extract(*desc, next_conf01, next_idx01, next_idx10, mask_out, b_ids, i_ids, j_ids, mconf,
                                mkpts0, mkpts1, capacity, count_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

int casmtr_pack_matches(const int64_t *b_ids, const int64_t *i_ids, const int64_t *j_ids, const float *mconf,
                        const float *mkpts0, const float *mkpts1, int M, int64_t pair_offset, int capacity,
                        unsigned char *out, casmtr_stream_t stream) {
    CASMTR_REQUIRE(M >= 0 && capacity >= M && out != nullptr, CASMTR_E_INVALID, "pack_matches: M=%d capacity=%d", M, capacity);
    CASMTR_REQUIRE(M == 0 || (b_ids && i_ids && j_ids && mconf && mkpts0 && mkpts1), CASMTR_E_INVALID, "pack_matches: null pointer");
    return launch_pack_matches(b_ids, i_ids, j_ids, mconf, mkpts0, mkpts1, M, pair_offset, out, nullptr, (cudaStream_t)stream);
}

int casmtr_pack_matches_dev(const int64_t *b_ids, const int64_t *i_ids, const int64_t *j_ids, const float *mconf,
                            const float *mkpts0, const float *mkpts1, const int32_t *count, int64_t pair_offset, int capacity,
                            unsigned char *out, casmtr_stream_t stream) {
    CASMTR_REQUIRE(capacity >= 0 && out && count && b_ids && i_ids && j_ids && mconf && mkpts0 && mkpts1, CASMTR_E_INVALID,
                   "pack_matches_dev: null pointer or negative capacity");
    return launch_pack_matches(b_ids, i_ids, j_ids, mconf, mkpts0, mkpts1, capacity, pair_offset, out, count, (cudaStream_t)stream);
}

int casmtr_fine_window_gather(const float *feat, const int64_t *b_ids, const int64_t *ids, float *out, int M, int C, int Hf, int Wf,
                              int wc, int stride, int W, casmtr_stream_t stream) {
    CASMTR_REQUIRE(M >= 0 && C > 0 && Hf > 0 && Wf > 0 && wc > 0 && stride > 0 && W > 0 && (W & 1), CASMTR_E_INVALID,
                   "fine_window_gather: bad sizes");
    CASMTR_REQUIRE(M == 0 || (feat && b_ids && ids && out), CASMTR_E_INVALID, "fine_window_gather: null pointer");
    return launch_fine_window_gather(feat, b_ids, ids, out, M, C, Hf, Wf, wc, stride, W, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ fine matching
int casmtr_fine_match_fwd(const float *feat_f0, const float *feat_f1, const float *mkpts1_c,
                          const float *scale1_b, const int64_t *b_ids, float scale,
                          float *expec_f, float *mkpts1_f, int M, int WW, int C, casmtr_stream_t stream) {
    CASMTR_REQUIRE(M >= 0 && WW > 0 && C > 0, CASMTR_E_INVALID, "fine_match: bad sizes");
    CASMTR_REQUIRE(M == 0 || (feat_f0 && feat_f1 && mkpts1_c && expec_f && mkpts1_f), CASMTR_E_INVALID, "fine_match: null pointer");
    CASMTR_REQUIRE(scale1_b == nullptr || b_ids != nullptr, CASMTR_E_INVALID, "fine_match: scale1_b needs b_ids");
    return launch_fine_match(feat_f0, feat_f1, mkpts1_c, scale1_b, b_ids, scale, expec_f, mkpts1_f, M, WW, C, nullptr, (cudaStream_t)stream);
}

int casmtr_fine_match_dev_fwd(const float *feat_f0, const float *feat_f1, const float *mkpts1_c,
                              const float *scale1_b, const int64_t *b_ids, float scale,
                              float *expec_f, float *mkpts1_f, const int32_t *count, int capacity, int WW, int C, casmtr_stream_t stream) {
    CASMTR_REQUIRE(capacity >= 0 && WW > 0 && C > 0 && count, CASMTR_E_INVALID, "fine_match_dev: bad sizes");
    CASMTR_REQUIRE(capacity == 0 || (feat_f0 && feat_f1 && mkpts1_c && expec_f && mkpts1_f), CASMTR_E_INVALID, "fine_match_dev: null pointer");
    CASMTR_REQUIRE(scale1_b == nullptr || b_ids != nullptr, CASMTR_E_INVALID, "fine_match_dev: scale1_b needs b_ids");
    return launch_fine_match(feat_f0, feat_f1, mkpts1_c, scale1_b, b_ids, scale, expec_f, mkpts1_f, capacity, WW, C, count, (cudaStream_t)stream);
}

}  // extern "C"
