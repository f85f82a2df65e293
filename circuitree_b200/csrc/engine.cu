// C ABI of the B200 rollout engine (include/circuitree_b200.h).  Host side only: validation, job
// tables, stream-ordered scratch, kernel launches.  No CPU fallback: without a CUDA device every
// compute entry point fails with CT_ERR_NO_DEVICE / CT_ERR_CUDA.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <nvtx3/nvToolsExt.h>      // header-only; ranges cost nothing unless a profiler is attached

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <array>
#include <future>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/circuitree_b200.h"
#include "ssa_pairwise.cuh"
#include "ssa_dimers.cuh"
#include "ssa_dimers_coop.cuh"
#include "ssa_pairwise_coop.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CT_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(CT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

struct KernelInfo {
    int blocks_per_sm = 0;
    size_t smem = 0;
};

// NVTX range for the duration of a C-ABI call (shows up on the timeline of nsys / ncu --nvtx)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// Stream-ordered scratch of one call: everything handed out is released (cudaFreeAsync on the same
// stream, i.e. after the work enqueued so far) when the object goes out of scope -- on every error
// return as well as on success.  keep() takes a buffer out (it then belongs to a Ticket).
struct StreamScratch {
    cudaStream_t stream;
    std::vector<void*> ptrs;
    explicit StreamScratch(cudaStream_t s) : stream(s) {}
    StreamScratch(const StreamScratch&) = delete;
    StreamScratch& operator=(const StreamScratch&) = delete;
    template <typename T>
    cudaError_t alloc(T** out, size_t bytes)
    {
        void* p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 1, stream);
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = static_cast<T*>(p);
        return e;
    }
    void keep(void* p)
    {
        for (auto& q : ptrs) if (q == p) q = nullptr;
    }
    ~StreamScratch()
    {
        for (void* p : ptrs) if (p) cudaFreeAsync(p, stream);
    }
};

// Process-wide, append-only table of dimer topologies (SPEC.md section 9); a k = 3 handle is
// kDimerTag | n_tf << 56 | index into this table.
constexpr uint64_t kDimerTag = 1ull << 63;
std::mutex g_dimer_mutex;
std::vector<ct::DimerTopo> g_dimer_topos;
std::map<std::array<uint8_t, sizeof(ct::DimerTopo)>, uint32_t> g_dimer_index;

inline bool is_dimer_handle(uint64_t h) { return (h & kDimerTag) != 0; }
inline int params_per_trajectory(const uint64_t* handles, int n_jobs)
{
    return (n_jobs > 0 && is_dimer_handle(handles[0])) ? CT_N_PARAMS_DIMER : CT_N_PARAMS;
}

}  // namespace

namespace {
// automatic schedule: dimer launches of up to this many times the resident one-lane-per-trajectory lanes
// (75,776 on B200) run on the cooperative kernel.  Measured (profiles/r02_dimer_schedules_*.txt): one topology,
// 2^16 trajectories: cooperative 2.5 s, one topology per warp 6.4 s; 2^20: one topology per warp 7 s (it
// refills its lanes from the same job and saturates at 4e10 steps/s, the cooperative kernel at ~1e10);
// batches of random five-TF topologies x 32 trajectories: 1024 leaves 36 s vs 112 s per-lane programs,
// 4096 leaves 31 s vs 86 s -- such batches last as long as their longest trajectory.
constexpr int kCoopAutoFactorOne = 2;
constexpr int kCoopAutoFactorMany = 8;

struct Ticket {                      // one asynchronous reward batch in flight
    cudaStream_t stream = nullptr;
    int32_t n_jobs = 0;
    uint64_t total = 0;
    float *d_params = nullptr, *d_reward = nullptr;
    int32_t* d_lag = nullptr;
    uint32_t* d_events = nullptr;
    double* d_mean = nullptr;
    uint64_t* d_jev = nullptr;
};
void release_ticket(Ticket& t);
}  // namespace

struct ct_engine {
    int device = 0;
    std::map<int64_t, Ticket> tickets;
    int64_t next_ticket = 1;
    int n_sm = 0;
    cudaStream_t stream = nullptr;   // used by the host-buffer entry points
    KernelInfo ki[3][2][2];         // [M-3][mixed][all jobs have M species]
    KernelInfo ki_fast[3];          // [M-3]: every gene has at most one regulator (ssa_step kFast)
    KernelInfo ki_dimers;
    int64_t launches = 0;
    unsigned long long* d_warp_iters = nullptr;   // debug counter (ct_engine_warp_iterations)
};

namespace {

template <int MM, bool kMixed, bool kAll, bool kFast = false>
int prepare_kernel(KernelInfo& ki)
{
    // per warp: reward scratch + the staging copy of one finished record (the records themselves live in global scratch)
    ki.smem = (size_t)(ct::kThreads / 32) * (ct::kSerStride * sizeof(float) + (size_t)MM * ct::kSerStride);
    CT_CUDA(cudaFuncSetAttribute(ct::ssa_pairwise_kernel<MM, kMixed, kAll, kFast>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ki.smem));
    CT_CUDA(cudaFuncSetAttribute(ct::ssa_pairwise_kernel<MM, kMixed, kAll, kFast>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
    CT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ki.blocks_per_sm, ct::ssa_pairwise_kernel<MM, kMixed, kAll, kFast>, ct::kThreads,
                                                          ki.smem));
    if (ki.blocks_per_sm < 1) return fail(CT_ERR_CUDA, "SSA kernel (M=%d) does not fit on an SM", MM);
    return CT_OK;
}

// Explicit parameter vectors must be finite and non-negative: with a NaN or infinite rate the clock of a trajectory
// never reaches the next grid point and the kernel would never end.  Checked where the buffer is on the host
// (ct_rewards_host, ct_rewards_submit); a device buffer handed to ct_launch_rewards is the caller's responsibility.
// Big buffers are scanned by eight threads (168 MB for 2^22 trajectories: ~10 ms against a 2.6 s launch).
int check_host_params(const float* p, size_t n_traj, size_t np)
{
    const size_t n = n_traj * np;
    auto scan = [](const float* q, size_t m) -> bool {
        bool bad = false;
        for (size_t i = 0; i < m; ++i) bad |= !(q[i] >= 0.0f && q[i] <= 3.0e38f);
        return bad;
    };
    bool bad = false;
    constexpr size_t kSerial = (size_t)1 << 20;
    if (n <= kSerial) {
        bad = scan(p, n);
    } else {
        constexpr int kThreadsScan = 8;
        std::future<bool> parts[kThreadsScan];
        const size_t len = (n + kThreadsScan - 1) / kThreadsScan;
        for (int x = 0; x < kThreadsScan; ++x) {
            const size_t a = (size_t)x * len, b = a + len < n ? a + len : n;
            parts[x] = std::async(std::launch::async, scan, p + (a < n ? a : n), b > a ? b - a : 0);
        }
        for (auto& f : parts) bad = f.get() || bad;
    }
    if (!bad) return CT_OK;
    size_t i = 0;
    while (i < n && p[i] >= 0.0f && p[i] <= 3.0e38f) ++i;
    return fail(CT_ERR_INVALID, "explicit parameter %zu of trajectory %zu is %g: rates must be finite and non-negative",
                i % np, i / np, (double)p[i]);
}

int validate_cfg(const ct_sim_config* cfg)
{
    if (!cfg) return fail(CT_ERR_INVALID, "cfg is NULL");
    if (cfg->nt < 1 || cfg->decim < 1 || cfg->nt < cfg->decim) return fail(CT_ERR_INVALID, "bad nt/decim %d/%d", cfg->nt, cfg->decim);
    if (cfg->nt / cfg->decim > CT_MAX_NDEC) return fail(CT_ERR_INVALID, "nt/decim = %d exceeds CT_MAX_NDEC", cfg->nt / cfg->decim);
    if (!(cfg->dt > 0.0f)) return fail(CT_ERR_INVALID, "dt must be positive");
    if (!(cfg->init_mean >= 0.0f) || cfg->init_mean > 80.0f) return fail(CT_ERR_INVALID, "init_mean out of range [0, 80]");
    if (cfg->schedule < 0 || cfg->schedule > 3)
        return fail(CT_ERR_INVALID, "schedule must be 0 (auto), 1 (uniform), 2 (mixed) or 3 (cooperative)");
    if (cfg->lpt < 0 || cfg->lpt > 1) return fail(CT_ERR_INVALID, "lpt must be 0 or 1");
    if (cfg->grid_limit < 0) return fail(CT_ERR_INVALID, "grid_limit must be >= 0");
    if (cfg->fast_path < 0 || cfg->fast_path > 1) return fail(CT_ERR_INVALID, "fast_path must be 0 or 1");
    return CT_OK;
}

}  // namespace

extern "C" {

const char* ct_last_error(void) { return g_err; }

int ct_version(void) { return 100; }

int ct_default_config(ct_sim_config* cfg)
{
    if (!cfg) return fail(CT_ERR_INVALID, "cfg is NULL");
    static const float r[CT_MAX_PARAMS][3] = {
        {-2.3f, -1.7f, 1.f}, {-1.0f, -0.4f, 1.f}, {-2.7f, -2.1f, 1.f}, {-0.3f, 0.3f, 1.f}, {0.0f, 0.6f, 1.f},
        {-3.5f, -2.5f, 1.f}, {-1.3f, -0.7f, 1.f}, {-1.9f, -1.3f, 1.f}, {-2.0f, -1.4f, 1.f}, {-2.7f, -2.1f, 1.f},
        {-5.3f, -4.7f, 1.f}, {-2.8f, -2.2f, 1.f}};
    cfg->nt = 2000;
    cfg->decim = 16;
    cfg->dt = 20.0f;
    cfg->init_mean = 10.0f;
    cfg->schedule = 0;
    cfg->lpt = 1;
    cfg->grid_limit = 0;
    cfg->fast_path = 1;
    memcpy(cfg->ranges, r, sizeof r);
    return CT_OK;
}

int ct_engine_create(int device, ct_engine** out)
{
    if (!out) return fail(CT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        (void)cudaGetLastError();
        return fail(CT_ERR_NO_DEVICE, "no CUDA device available (%s); this engine has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n_dev) return fail(CT_ERR_INVALID, "device %d out of range [0,%d)", device, n_dev);
    CT_CUDA(cudaSetDevice(device));
    // (owned by the guard until the last step has succeeded: every early return below destroys stream and engine)
    struct EngineGuard {
        ct_engine* e;
        ~EngineGuard() { if (e) ct_engine_destroy(e); }
    } guard{new ct_engine()};
    ct_engine* eng = guard.e;
    eng->device = device;
    cudaDeviceProp prop;
    CT_CUDA(cudaGetDeviceProperties(&prop, device));
    eng->n_sm = prop.multiProcessorCount;
    CT_CUDA(cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking));
    {
        // Every call allocates its scratch stream-ordered (cudaMallocAsync) and synchronises at the end; with the
        // default release threshold (0) the pool hands its memory back to the OS at every synchronisation and the
        // next call pays for mapping hundreds of MB again (measured: +37 ms on a 2^18 launch, -9 % end to end at
        // 2^22).  Keep what the pool has grown to.
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            unsigned long long threshold = ~0ull;
            (void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
        }
        (void)cudaGetLastError();
    }
    int rc;
#define CT_PREP(MM, MIXED, ALL) \
    if ((rc = prepare_kernel<MM, MIXED, ALL>(eng->ki[MM - 3][MIXED][ALL])) != CT_OK) return rc;
    CT_PREP(3, false, false) CT_PREP(3, false, true) CT_PREP(3, true, false) CT_PREP(3, true, true)
    CT_PREP(4, false, false) CT_PREP(4, false, true) CT_PREP(4, true, false) CT_PREP(4, true, true)
    CT_PREP(5, false, false) CT_PREP(5, false, true) CT_PREP(5, true, false) CT_PREP(5, true, true)
#undef CT_PREP
    if ((rc = prepare_kernel<3, false, true, true>(eng->ki_fast[0])) != CT_OK) return rc;
    if ((rc = prepare_kernel<4, false, true, true>(eng->ki_fast[1])) != CT_OK) return rc;
    if ((rc = prepare_kernel<5, false, true, true>(eng->ki_fast[2])) != CT_OK) return rc;
    {
        // the dimer kernel sizes its shared memory per launch (DimerLayout); allow the maximum here
        KernelInfo& ki = eng->ki_dimers;
        ct::DimerLayout lay{ct::kMaxSpecies, ct::kMaxEntries + 1};
        ki.smem = lay.warp_bytes() * (ct::kDimThreads / 32);
        cudaError_t e1 = cudaFuncSetAttribute(ct::ssa_dimers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ki.smem);
        cudaError_t e2 = cudaFuncSetAttribute(ct::ssa_dimers_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                              cudaSharedmemCarveoutMaxShared);
        cudaError_t e3 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ki.blocks_per_sm, ct::ssa_dimers_kernel,
                                                                       ct::kDimThreads, ki.smem);
        ct::DimerMixShape big{ct::kMaxSpecies, ct::kMaxIxn, ct::kMaxSpecies * ct::kMixPerTf + 2 * ct::kMaxIxn, 6};
        cudaError_t e4 = cudaFuncSetAttribute(ct::ssa_dimers_mixed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)(big.warp_bytes() * (ct::kDimThreads / 32)));
        cudaError_t e5 = cudaFuncSetAttribute(ct::ssa_dimers_mixed_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                              cudaSharedmemCarveoutMaxShared);
        const int coop8 = (int)(ct::coop_warp_bytes(8, ct::kMaxSpecies) * (ct::kCoopThreads / 32));
        const int coop4 = (int)(ct::coop_warp_bytes(4, 3) * (ct::kCoopThreads / 32));
        if (cudaFuncSetAttribute(ct::ssa_dimers_coop_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, coop8) != cudaSuccess ||
            cudaFuncSetAttribute(ct::ssa_dimers_coop_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, coop4) != cudaSuccess ||
            cudaFuncSetAttribute(ct::ssa_dimers_coop_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess ||
            cudaFuncSetAttribute(ct::ssa_dimers_coop_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess)
            e1 = cudaErrorUnknown;
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess || e5 != cudaSuccess ||
            ki.blocks_per_sm < 1)
            return fail(CT_ERR_CUDA, "dimer SSA kernel does not fit on an SM");
    }
    guard.e = nullptr;
    *out = eng;
    return CT_OK;
}

int ct_engine_destroy(ct_engine* eng)
{
    if (!eng) return CT_OK;
    cudaSetDevice(eng->device);
    for (auto& kv : eng->tickets) release_ticket(kv.second);     // batches nobody waited for
    if (eng->d_warp_iters) cudaFree(eng->d_warp_iters);
    if (eng->stream) cudaStreamDestroy(eng->stream);
    delete eng;
    return CT_OK;
}

int ct_engine_info(ct_engine* eng, int32_t* n_sm, int32_t* resident_lanes)
{
    if (!eng) return fail(CT_ERR_INVALID, "engine is NULL");
    if (n_sm) *n_sm = eng->n_sm;
    if (resident_lanes) *resident_lanes = eng->n_sm * eng->ki[0][0][1].blocks_per_sm * ct::kThreads;
    return CT_OK;
}

int64_t ct_engine_launch_count(ct_engine* eng) { return eng ? eng->launches : -1; }

// Debug: total hot-loop iterations (two lock-step SSA steps of 32 lanes each) of all warps since the
// previous call; the first call switches the counter on.  lane utilisation = events / (64 * value).
int64_t ct_engine_warp_iterations(ct_engine* eng)
{
    if (!eng) return -1;
    cudaSetDevice(eng->device);
    if (!eng->d_warp_iters) {
        if (cudaMalloc((void**)&eng->d_warp_iters, sizeof(unsigned long long)) != cudaSuccess) return -1;
        cudaMemset(eng->d_warp_iters, 0, sizeof(unsigned long long));
        return 0;
    }
    unsigned long long v = 0;
    cudaDeviceSynchronize();
    cudaMemcpy(&v, eng->d_warp_iters, sizeof v, cudaMemcpyDeviceToHost);
    cudaMemset(eng->d_warp_iters, 0, sizeof v);
    return (int64_t)v;
}

int ct_compile_topology(int32_t n_species, const int64_t* act, int32_t n_act, const int64_t* inh, int32_t n_inh,
                        int32_t k, uint64_t* handle)
{
    if (!handle) return fail(CT_ERR_INVALID, "handle is NULL");
    if (k != 2 && k != 3) return fail(CT_ERR_INVALID, "k must be 2 or 3, got %d", k);
    if (n_species < 1) return fail(CT_ERR_INVALID, "n_species must be >= 1");
    if (n_species > CT_MAX_SPECIES) return fail(CT_ERR_UNSUPPORTED, "n_species %d > CT_MAX_SPECIES", n_species);
    if (n_act < 0 || n_inh < 0 || (n_act && !act) || (n_inh && !inh)) return fail(CT_ERR_INVALID, "bad edge arrays");
    if (k == 3) {
        // SPEC.md section 9: interactions numbered activations first; dimers in order of first appearance
        if (n_act + n_inh > CT_MAX_INTERACTIONS)
            return fail(CT_ERR_UNSUPPORTED, "%d interactions > CT_MAX_INTERACTIONS", n_act + n_inh);
        ct::DimerTopo T;
        memset(&T, 0, sizeof T);
        T.n_tf = (uint8_t)n_species;
        int per_promoter[CT_MAX_SPECIES] = {0, 0, 0, 0, 0};
        for (int pass = 0; pass < 2; ++pass) {
            const int64_t* e = pass ? inh : act;
            const int n = pass ? n_inh : n_act;
            for (int x = 0; x < n; ++x) {
                const int64_t m1 = e[3 * x], m2 = e[3 * x + 1], tg = e[3 * x + 2];
                if (m1 < 0 || m2 < m1 || m2 >= n_species || tg < 0 || tg >= n_species)
                    return fail(CT_ERR_INVALID, "interaction (%lld,%lld,%lld) invalid for %d TFs (need 0 <= m1 <= m2 < n, target < n)",
                                (long long)m1, (long long)m2, (long long)tg, n_species);
                int d = -1;
                for (int y = 0; y < T.n_dim; ++y)
                    if (T.dim_a[y] == m1 && T.dim_b[y] == m2) d = y;
                for (int y = 0; y < T.n_ixn; ++y)
                    if (d >= 0 && T.ix_dim[y] == d && T.ix_tgt[y] == tg)
                        return fail(CT_ERR_INVALID, "duplicate interaction (%lld,%lld,%lld)", (long long)m1, (long long)m2, (long long)tg);
                if (d < 0) { d = T.n_dim++; T.dim_a[d] = (uint8_t)m1; T.dim_b[d] = (uint8_t)m2; }
                T.ix_dim[T.n_ixn] = (uint8_t)d;
                T.ix_tgt[T.n_ixn] = (uint8_t)tg;
                T.ix_act[T.n_ixn] = pass ? 0 : 1;
                T.n_ixn++;
                if ((int)tg + 1 > T.n_comp) T.n_comp = (uint8_t)(tg + 1);
                per_promoter[tg]++;
            }
        }
        std::array<uint8_t, sizeof(ct::DimerTopo)> key;
        memcpy(key.data(), &T, sizeof T);
        std::lock_guard<std::mutex> lock(g_dimer_mutex);
        auto it = g_dimer_index.find(key);
        uint32_t idx;
        if (it != g_dimer_index.end()) {
            idx = it->second;
        } else {
            idx = (uint32_t)g_dimer_topos.size();
            g_dimer_topos.push_back(T);
            g_dimer_index[key] = idx;
        }
        *handle = kDimerTag | ((uint64_t)n_species << 56) | idx;
        return CT_OK;
    }
    uint64_t code = (uint64_t)n_species << 56;
    for (int pass = 0; pass < 2; ++pass) {
        const int64_t* e = pass ? inh : act;
        const int n = pass ? n_inh : n_act;
        for (int x = 0; x < n; ++x) {
            const int64_t i = e[2 * x], j = e[2 * x + 1];
            if (i < 0 || j < 0 || i >= n_species || j >= n_species)
                return fail(CT_ERR_INVALID, "edge (%lld,%lld) out of range for %d species", (long long)i, (long long)j, n_species);
            const int sh = 2 * ((int)i * ct::kTopoStride + (int)j);
            if ((code >> sh) & 3u) return fail(CT_ERR_INVALID, "duplicate edge (%lld,%lld)", (long long)i, (long long)j);
            code |= (uint64_t)(pass ? 2u : 1u) << sh;
        }
    }
    *handle = code;
    return CT_OK;
}

int ct_check_params(const float* h_params, int64_t n_trajectories, int32_t n_params)
{
    if (n_trajectories < 0 || n_params < 1 || n_params > CT_MAX_PARAMS || (n_trajectories && !h_params))
        return fail(CT_ERR_INVALID, "bad arguments");
    return check_host_params(h_params, (size_t)n_trajectories, (size_t)n_params);
}

int ct_topology_info(uint64_t handle, int32_t* n_species, int32_t* n_reactions)
{
    if (is_dimer_handle(handle)) {
        std::lock_guard<std::mutex> lock(g_dimer_mutex);
        const uint64_t idx = handle & 0xffffffffull;
        if (idx >= g_dimer_topos.size()) return fail(CT_ERR_INVALID, "unknown dimer topology handle");
        const ct::DimerTopo& T = g_dimer_topos[idx];
        if (n_species) *n_species = T.n_tf;
        if (n_reactions) *n_reactions = 4 * T.n_tf + 2 * T.n_ixn + 2 * T.n_dim;
        return CT_OK;
    }
    const uint32_t nsp = ct::topo_nspecies(handle);
    if (nsp < 1 || nsp > CT_MAX_SPECIES) return fail(CT_ERR_INVALID, "bad topology handle");
    int edges = 0;
    for (int c = 0; c < 25; ++c) edges += ((handle >> (2 * c)) & 3u) != 0;
    if (n_species) *n_species = (int32_t)nsp;
    if (n_reactions) *n_reactions = 4 * (int32_t)nsp + 2 * edges;
    return CT_OK;
}

int ct_launch_rewards(ct_engine* eng, const uint64_t* handles, const uint32_t* n_traj, const uint64_t* traj_base,
                      const float* job_cost, int32_t n_jobs, const ct_sim_config* cfg, uint64_t seed, void* stream_,
                      const float* d_params, float* d_reward, int32_t* d_lag, uint32_t* d_events, uint8_t* d_series,
                      int32_t series_stride)
{
    NvtxRange nvtx_range("ct_launch_rewards");
    if (!eng) return fail(CT_ERR_INVALID, "engine is NULL");
    if (n_jobs < 0 || (n_jobs && (!handles || !n_traj || !traj_base))) return fail(CT_ERR_INVALID, "bad job arrays");
    int rc = validate_cfg(cfg);
    if (rc != CT_OK) return rc;
    if (n_jobs == 0) return CT_OK;
    if (!d_reward) return fail(CT_ERR_INVALID, "d_reward is NULL");
    cudaStream_t stream = (cudaStream_t)stream_;
    CT_CUDA(cudaSetDevice(eng->device));

    const bool dimers = is_dimer_handle(handles[0]);
    std::vector<ct::DimerProg> progs;   // one reaction program per job (dimer model)
    std::vector<ct::DimerLaneProg> lane_progs;   // the same topologies as per-lane programs (many small jobs)
    bool lane_progs_ok = true;          // false: some promoter has more interactions than the per-lane kernel has slots
    ct::DimerLayout lay{1, 1};
    ct::DimerMixShape shape{1, 0, 0, 0};
    if (dimers) {
        progs.resize((size_t)n_jobs);
        lane_progs.resize((size_t)n_jobs);
        std::lock_guard<std::mutex> lock(g_dimer_mutex);
        for (int j = 0; j < n_jobs; ++j) {
            const uint64_t idx = handles[j] & 0xffffffffull;
            if (!is_dimer_handle(handles[j]) || idx >= g_dimer_topos.size())
                return fail(CT_ERR_INVALID, "job %d: not a dimer topology handle (a launch takes one model only)", j);
            progs[j] = ct::dimer_program(g_dimer_topos[idx]);
            if (n_traj[j] == 0) continue;
            if (progs[j].n_tf > lay.ser_species) lay.ser_species = progs[j].n_tf;
            if (progs[j].n_entries + 1 > lay.cum_slots) lay.cum_slots = progs[j].n_entries + 1;
            lane_progs_ok = ct::dimer_lane_program(g_dimer_topos[idx], lane_progs[j]) && lane_progs_ok;
            if (progs[j].n_tf > shape.n_tf) shape.n_tf = progs[j].n_tf;
            if (progs[j].n_dim > shape.n_dim) shape.n_dim = progs[j].n_dim;
        }
        shape.n_entries = shape.n_tf * ct::kMixPerTf + 2 * shape.n_dim;
        while ((1 << shape.search_iters) < shape.n_entries + 1) ++shape.search_iters;
    }
    std::vector<ct::Job> jobs((size_t)n_jobs);
    uint64_t total = 0, tiles = 0;
    uint32_t max_nsp = 0;
    for (int j = 0; j < n_jobs; ++j) {
        if (!dimers && is_dimer_handle(handles[j]))
            return fail(CT_ERR_INVALID, "job %d: dimer handle in a pairwise launch (a launch takes one model only)", j);
        const uint32_t nsp = ct::topo_nspecies(handles[j]);
        if (nsp < 1 || nsp > CT_MAX_SPECIES) return fail(CT_ERR_INVALID, "job %d: bad topology handle", j);
        if (nsp > max_nsp) max_nsp = nsp;
        jobs[j].topo = handles[j];
        jobs[j].traj_base = traj_base[j];
        jobs[j].n_traj = n_traj[j];
        jobs[j].out_offset = (uint32_t)total;
        jobs[j].tile_start = (uint32_t)tiles;
        jobs[j].pad = 0;
        total += n_traj[j];
        tiles += (n_traj[j] + ct::kTile - 1) / ct::kTile;
        if (total > 0xffffffffull) return fail(CT_ERR_INVALID, "more than 2^32-1 trajectories in one launch");
    }
    if (total == 0) return CT_OK;
    if (d_series && series_stride < (int)max_nsp * (cfg->nt / cfg->decim))
        return fail(CT_ERR_INVALID, "series_stride %d too small", series_stride);

    StreamScratch scratch(stream);      // released in stream order when this call returns, whatever the path
    ct::Job* d_jobs = nullptr;
    uint32_t* d_counter = nullptr;
    CT_CUDA(scratch.alloc(&d_jobs, sizeof(ct::Job) * jobs.size()));
    CT_CUDA(scratch.alloc(&d_counter, sizeof(uint32_t)));
    CT_CUDA(cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(ct::Job) * jobs.size(), cudaMemcpyHostToDevice, stream));
    CT_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(uint32_t), stream));

    ct::LaunchArgs A;
    memset(&A, 0, sizeof A);
    A.jobs = d_jobs;
    A.n_jobs = (uint32_t)n_jobs;
    A.n_tiles = (uint32_t)tiles;
    A.tile_counter = d_counter;
    A.params = d_params;
    A.reward = d_reward;
    A.lag = d_lag;
    A.events = d_events;
    A.series = d_series;
    A.series_stride = series_stride;
    A.log_mask = 0;
    A.n_params = dimers ? CT_N_PARAMS_DIMER : CT_N_PARAMS;
    for (int i = 0; i < CT_MAX_PARAMS; ++i) {
        A.rng_lo[i] = cfg->ranges[i][0];
        A.rng_span[i] = cfg->ranges[i][1] - cfg->ranges[i][0];
        if (cfg->ranges[i][2] != 0.0f) A.log_mask |= 1u << i;
    }
    A.seed_lo = (uint32_t)seed;
    A.seed_hi = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        A.rk[2 * r] = A.seed_lo + (uint32_t)r * 0x9E3779B9u;
        A.rk[2 * r + 1] = A.seed_hi + (uint32_t)r * 0xBB67AE85u;
    }
    A.total = (uint32_t)total;
    A.order = nullptr;
    A.warp_iters = eng->d_warp_iters;
    A.nt = cfg->nt;
    A.decim = cfg->decim;
    A.n_dec = cfg->nt / cfg->decim;
    A.dt = cfg->dt;
    A.init_mean = cfg->init_mean;

    // schedule: big jobs -> one topology per warp (uniform branches); many small jobs -> mixed lanes
    // One topology per warp only pays when a warp can refill its lanes from the same job, i.e. when
    // the jobs are much larger than the machine (a warp that has to change topology drains first);
    // measured on the 3-component landscape (419 jobs x 10^4 trajectories): 5.7e10 steps/s one
    // topology per warp, 1.3e11 mixed.
    bool mixed = cfg->schedule == 2 || cfg->schedule == 3;
    const uint64_t resident = (uint64_t)eng->n_sm * eng->ki[0][0][1].blocks_per_sm * ct::kThreads;
    if (cfg->schedule == 0) mixed = n_jobs > 1 && total / (uint64_t)n_jobs < 4 * resident;
    if (dimers && !lane_progs_ok) mixed = false;   // per-lane programs need <= kMixK interactions per promoter
    // Dimer launches that cannot give every resident lane a trajectory of its own -- MCTS batches, single
    // get_reward calls -- last as long as their longest trajectory: spread each trajectory over a sub-group
    // of lanes instead (ssa_dimers_coop.cuh; identical results).
    bool coop = dimers && lane_progs_ok && cfg->schedule == 3;
    if (dimers && lane_progs_ok && cfg->schedule == 0)
        coop = total <= (uint64_t)(n_jobs > 1 ? kCoopAutoFactorMany : kCoopAutoFactorOne) * resident;
    if (coop) mixed = true;                        // (launch positions, not per-job tiles)
    if (mixed) A.n_tiles = (uint32_t)((total + ct::kTile - 1) / ct::kTile);   // tiles over launch positions

    // longest-expected-first order (cfg->lpt): cost pre-pass + radix sort of (job, cost) keys
    unsigned long long *d_keys = nullptr, *d_keys_out = nullptr;
    uint32_t *d_slots = nullptr, *d_order = nullptr;
    float* d_job_cost = nullptr;
    void* d_tmp = nullptr;
    if (cfg->lpt && total > 1) {
        if (total > 0x7fffffffull)
            return fail(CT_ERR_INVALID, "longest-first ordering (cfg.lpt) takes at most 2^31-1 trajectories per launch");
        CT_CUDA(scratch.alloc(&d_keys, sizeof(unsigned long long) * total));
        CT_CUDA(scratch.alloc(&d_keys_out, sizeof(unsigned long long) * total));
        CT_CUDA(scratch.alloc(&d_slots, sizeof(uint32_t) * total));
        CT_CUDA(scratch.alloc(&d_order, sizeof(uint32_t) * total));
        if (job_cost) {
            CT_CUDA(scratch.alloc(&d_job_cost, sizeof(float) * n_jobs));
            CT_CUDA(cudaMemcpyAsync(d_job_cost, job_cost, sizeof(float) * n_jobs, cudaMemcpyHostToDevice, stream));
        }
        ct::cost_key_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(A, d_job_cost, mixed ? 0 : 1, d_keys, d_slots);
        CT_CUDA(cudaGetLastError());
        int end_bit = 32;
        if (!mixed) { int b = 0; while ((1ll << b) < (long long)n_jobs) ++b; end_bit = 32 + b; }
        size_t tmp_bytes = 0;
        CT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys_out, d_slots, d_order, (int)total, 0,
                                                end_bit, stream));
        CT_CUDA(scratch.alloc(&d_tmp, tmp_bytes));
        CT_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys_out, d_slots, d_order, (int)total, 0,
                                                end_bit, stream));
        eng->launches += 2;
        A.order = d_order;
    }
    if (dimers) {
        // one topology per warp (big jobs) or per-lane programs (many small jobs); identical results
        void* d_progs = nullptr;
        const void* h_progs = mixed ? (const void*)lane_progs.data() : (const void*)progs.data();
        const size_t prog_bytes = mixed ? sizeof(ct::DimerLaneProg) * lane_progs.size() : sizeof(ct::DimerProg) * progs.size();
        CT_CUDA(scratch.alloc(&d_progs, prog_bytes));
        CT_CUDA(cudaMemcpyAsync(d_progs, h_progs, prog_bytes, cudaMemcpyHostToDevice, stream));
        if (coop) {
            const int G = (shape.n_tf <= 3 && shape.n_dim <= 4) ? 4 : 8;
            const int ser_species = shape.n_tf;
            const size_t csmem = ct::coop_warp_bytes(G, ser_species) * (ct::kCoopThreads / 32);
            int per_sm = 0;
            if (G == 8) CT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ct::ssa_dimers_coop_kernel<8>, ct::kCoopThreads, csmem));
            else CT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ct::ssa_dimers_coop_kernel<4>, ct::kCoopThreads, csmem));
            if (per_sm < 1) return fail(CT_ERR_CUDA, "cooperative dimer kernel does not fit on an SM (%zu bytes of shared memory)", csmem);
            const uint64_t per_cta = (uint64_t)ct::kCoopThreads / G;            // trajectories a CTA runs at a time
            const uint64_t want = (total + per_cta - 1) / per_cta;
            uint64_t cap = (uint64_t)eng->n_sm * per_sm;
            if (cfg->grid_limit > 0 && (uint64_t)cfg->grid_limit < cap) cap = (uint64_t)cfg->grid_limit;
            const unsigned cgrid = (unsigned)(want < cap ? want : cap);
            if (G == 8) ct::ssa_dimers_coop_kernel<8><<<cgrid, ct::kCoopThreads, csmem, stream>>>(A, (const ct::DimerLaneProg*)d_progs, ser_species);
            else ct::ssa_dimers_coop_kernel<4><<<cgrid, ct::kCoopThreads, csmem, stream>>>(A, (const ct::DimerLaneProg*)d_progs, ser_species);
            CT_CUDA(cudaGetLastError());
            eng->launches += 1;
            return CT_OK;
        }
        const size_t smem = (mixed ? shape.warp_bytes() : lay.warp_bytes()) * (ct::kDimThreads / 32);
        const size_t rec_lane = mixed ? shape.rec_lane_bytes() : lay.rec_lane_bytes();
        int per_sm = 0;
        if (mixed) CT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ct::ssa_dimers_mixed_kernel, ct::kDimThreads, smem));
        else CT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ct::ssa_dimers_kernel, ct::kDimThreads, smem));
        if (per_sm < 1) return fail(CT_ERR_CUDA, "dimer SSA kernel does not fit on an SM (%zu bytes of shared memory)", smem);
        const uint64_t want = (total + ct::kDimThreads - 1) / ct::kDimThreads;
        uint64_t cap = (uint64_t)eng->n_sm * per_sm;
        if (cfg->grid_limit > 0 && (uint64_t)cfg->grid_limit * (ct::kThreads / ct::kDimThreads) < cap)
            cap = (uint64_t)cfg->grid_limit * (ct::kThreads / ct::kDimThreads);   // grid_limit counts 128-thread CTAs
        const unsigned dgrid = (unsigned)(want < cap ? want : cap);
        uint8_t* d_records = nullptr;      // per-lane record scratch (see ssa_dimers.cuh)
        CT_CUDA(scratch.alloc(&d_records, (size_t)dgrid * ct::kDimThreads * rec_lane));
        if (mixed)
            ct::ssa_dimers_mixed_kernel<<<dgrid, ct::kDimThreads, smem, stream>>>(A, (const ct::DimerLaneProg*)d_progs, shape, d_records);
        else
            ct::ssa_dimers_kernel<<<dgrid, ct::kDimThreads, smem, stream>>>(A, (const ct::DimerProg*)d_progs, lay, d_records);
        CT_CUDA(cudaGetLastError());
        eng->launches += 1;
        // (pageable host memory: the copies above have been staged when cudaMemcpyAsync returns)
        return CT_OK;
    }
    // Pairwise launches that fit into ONE wave of the cooperative kernel (one trajectory per sub-group of lanes,
    // one gene per lane: ssa_pairwise_coop.cuh) last as long as their longest trajectory; its step is ~2x shorter.
    {
        const int G = max_nsp <= 4 ? 4 : 8;
        const size_t csmem = ct::pair_coop_warp_bytes(G, (int)max_nsp) * (ct::kPairCoopThreads / 32);
        int per_sm = 0;
        if (G == 4) CT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ct::ssa_pairwise_coop_kernel<4, 4>, ct::kPairCoopThreads, csmem));
        else CT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ct::ssa_pairwise_coop_kernel<8, 5>, ct::kPairCoopThreads, csmem));
        const uint64_t per_cta = (uint64_t)ct::kPairCoopThreads / G;
        uint64_t cap = (uint64_t)eng->n_sm * (per_sm > 0 ? per_sm : 1);
        // Automatic choice, in waves of the cooperative kernel (28,416 trajectories for <= 4 components, 14,208 for 5), from
        // profiles/r02_pairwise_coop_probe.txt: one job: 2^16 repressilator trajectories 206 ms vs 264 ms one-lane (2^8: 78 vs
        // 110 ms), the one-lane loops win from ~2^17; many small jobs (MCTS batches, 32 trajectories per leaf): 3-component
        // 955 leaves 1.06 s vs 1.61 s and 4096 leaves 1.61 s vs 2.02 s (lane utilisation 0.83 vs 0.53), 5-component 750 leaves
        // 1.65 s vs 2.81 s but 4096 leaves 3.73 s vs 3.46 s.
        const uint64_t wave = cap * per_cta;
        const uint64_t limit = n_jobs > 1 ? (G == 4 ? 5 * wave : 4 * wave) : (5 * wave) / 2;
        const bool pair_coop = per_sm > 0 && (cfg->schedule == 3 || (cfg->schedule == 0 && total <= limit));
        if (pair_coop) {
            if (cfg->grid_limit > 0 && (uint64_t)cfg->grid_limit < cap) cap = (uint64_t)cfg->grid_limit;
            const uint64_t want = (total + per_cta - 1) / per_cta;
            const unsigned cgrid = (unsigned)(want < cap ? want : cap);
            if (G == 4) ct::ssa_pairwise_coop_kernel<4, 4><<<cgrid, ct::kPairCoopThreads, csmem, stream>>>(A, (int)max_nsp);
            else ct::ssa_pairwise_coop_kernel<8, 5><<<cgrid, ct::kPairCoopThreads, csmem, stream>>>(A, (int)max_nsp);
            CT_CUDA(cudaGetLastError());
            eng->launches += 1;
            return CT_OK;
        }
    }
    const int mm = max_nsp <= 3 ? 3 : (int)max_nsp;
    bool all_full = true;
    for (int j = 0; j < n_jobs; ++j) all_full = all_full && (ct::topo_nspecies(handles[j]) == (uint32_t)mm || n_traj[j] == 0);
    // every gene of every job has at most one regulator -> the specialised hot loop (same results)
    bool fast = !mixed && all_full && cfg->fast_path != 0;
    for (int j = 0; j < n_jobs && fast; ++j) {
        if (n_traj[j] == 0) continue;
        for (int t = 0; t < mm; ++t) {
            int indeg = 0;
            for (int i = 0; i < mm; ++i) indeg += ct::topo_edge(handles[j], i, t) != 0u;
            if (indeg > 1) fast = false;
        }
    }
    const KernelInfo& ki = fast ? eng->ki_fast[mm - 3] : eng->ki[mm - 3][mixed ? 1 : 0][all_full ? 1 : 0];
    uint64_t want_blocks = (total + ct::kThreads - 1) / ct::kThreads;
    uint64_t max_blocks = (uint64_t)eng->n_sm * ki.blocks_per_sm;
    if (cfg->grid_limit > 0 && (uint64_t)cfg->grid_limit < max_blocks) max_blocks = (uint64_t)cfg->grid_limit;
    const unsigned grid = (unsigned)(want_blocks < max_blocks ? want_blocks : max_blocks);
    // per-lane record scratch: mm x 128 bytes per resident lane (a byte per species every `decim` grid points)
    CT_CUDA(scratch.alloc(&A.records, (size_t)grid * ct::kThreads * mm * ct::kSerStride));
#define CT_LAUNCH(MM, MIXED, ALL) ct::ssa_pairwise_kernel<MM, MIXED, ALL><<<grid, ct::kThreads, ki.smem, stream>>>(A)
#define CT_LAUNCH_M(MM)                                                          \
    do {                                                                         \
        if (fast) ct::ssa_pairwise_kernel<MM, false, true, true><<<grid, ct::kThreads, ki.smem, stream>>>(A);   \
        else if (mixed) { if (all_full) CT_LAUNCH(MM, true, true); else CT_LAUNCH(MM, true, false); }     \
        else { if (all_full) CT_LAUNCH(MM, false, true); else CT_LAUNCH(MM, false, false); }         \
    } while (0)
    if (mm == 3) CT_LAUNCH_M(3); else if (mm == 4) CT_LAUNCH_M(4); else CT_LAUNCH_M(5);
#undef CT_LAUNCH_M
#undef CT_LAUNCH
    CT_CUDA(cudaGetLastError());
    eng->launches += 1;
    return CT_OK;
}

int ct_reduce_jobs(ct_engine* eng, const uint32_t* n_traj, int32_t n_jobs, void* stream_, const float* d_reward,
                   const uint32_t* d_events, double* d_job_mean, uint64_t* d_job_events)
{
    NvtxRange nvtx_range("ct_reduce_jobs");
    if (!eng) return fail(CT_ERR_INVALID, "engine is NULL");
    if (n_jobs < 0 || (n_jobs && (!n_traj || !d_reward || !d_job_mean))) return fail(CT_ERR_INVALID, "bad arguments");
    if (n_jobs == 0) return CT_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    CT_CUDA(cudaSetDevice(eng->device));
    std::vector<uint32_t> off((size_t)n_jobs + 1);
    uint64_t total = 0;
    for (int j = 0; j < n_jobs; ++j) { off[j] = (uint32_t)total; total += n_traj[j]; }
    off[n_jobs] = (uint32_t)total;
    StreamScratch scratch(stream);
    uint32_t* d_off = nullptr;
    CT_CUDA(scratch.alloc(&d_off, sizeof(uint32_t) * off.size()));
    CT_CUDA(cudaMemcpyAsync(d_off, off.data(), sizeof(uint32_t) * off.size(), cudaMemcpyHostToDevice, stream));
    uint32_t biggest = 0;
    for (int j = 0; j < n_jobs; ++j) biggest = n_traj[j] > biggest ? n_traj[j] : biggest;
    if (biggest > 16384) {      // a warp per job would crawl through a big job: a block per job instead
        ct::reduce_big_jobs_kernel<<<(unsigned)n_jobs, 256, 0, stream>>>(d_off, n_jobs, d_reward, d_events, d_job_mean,
                                                                         (unsigned long long*)d_job_events);
    } else {
        const int warps_per_block = 8;
        const unsigned grid = (unsigned)((n_jobs + warps_per_block - 1) / warps_per_block);
        ct::reduce_jobs_kernel<<<grid, warps_per_block * 32, 0, stream>>>(d_off, n_jobs, d_reward, d_events, d_job_mean,
                                                                          (unsigned long long*)d_job_events);
    }
    CT_CUDA(cudaGetLastError());
    eng->launches += 1;
    return CT_OK;
}

int ct_rewards_host(ct_engine* eng, const uint64_t* handles, const uint32_t* n_traj, const uint64_t* traj_base,
                    const float* job_cost, int32_t n_jobs, const ct_sim_config* cfg, uint64_t seed, const float* h_params,
                    double* h_job_mean, uint64_t* h_job_events, float* h_reward, int32_t* h_lag)
{
    NvtxRange nvtx_range("ct_rewards_host");
    if (!eng) return fail(CT_ERR_INVALID, "engine is NULL");
    if (n_jobs < 0 || (n_jobs && (!handles || !n_traj || !traj_base))) return fail(CT_ERR_INVALID, "bad job arrays");
    if (n_jobs == 0) return CT_OK;
    CT_CUDA(cudaSetDevice(eng->device));
    cudaStream_t s = eng->stream;
    uint64_t total = 0;
    for (int j = 0; j < n_jobs; ++j) total += n_traj[j];
    if (total == 0) {
        for (int j = 0; j < n_jobs; ++j) { if (h_job_mean) h_job_mean[j] = 0.0; if (h_job_events) h_job_events[j] = 0; }
        return CT_OK;
    }
    float *d_params = nullptr, *d_reward = nullptr;
    int32_t* d_lag = nullptr;
    uint32_t* d_events = nullptr;
    double* d_mean = nullptr;
    uint64_t* d_jev = nullptr;
    int rc;
    if (h_params && (rc = check_host_params(h_params, (size_t)total, (size_t)params_per_trajectory(handles, n_jobs))) != CT_OK) return rc;
    {
        StreamScratch scratch(s);       // freed in stream order at the end of this block, on every path
        CT_CUDA(scratch.alloc(&d_reward, sizeof(float) * total));
        CT_CUDA(scratch.alloc(&d_lag, sizeof(int32_t) * total));
        CT_CUDA(scratch.alloc(&d_events, sizeof(uint32_t) * total));
        CT_CUDA(scratch.alloc(&d_mean, sizeof(double) * n_jobs));
        CT_CUDA(scratch.alloc(&d_jev, sizeof(uint64_t) * n_jobs));
        if (h_params) {
            const size_t np = (size_t)params_per_trajectory(handles, n_jobs);
            CT_CUDA(scratch.alloc(&d_params, sizeof(float) * total * np));
            CT_CUDA(cudaMemcpyAsync(d_params, h_params, sizeof(float) * total * np, cudaMemcpyHostToDevice, s));
        }
        rc = ct_launch_rewards(eng, handles, n_traj, traj_base, job_cost, n_jobs, cfg, seed, s, d_params, d_reward,
                               d_lag, d_events, nullptr, 0);
        if (rc == CT_OK) rc = ct_reduce_jobs(eng, n_traj, n_jobs, s, d_reward, d_events, d_mean, d_jev);
        if (rc == CT_OK) {
            if (h_job_mean) CT_CUDA(cudaMemcpyAsync(h_job_mean, d_mean, sizeof(double) * n_jobs, cudaMemcpyDeviceToHost, s));
            if (h_job_events) CT_CUDA(cudaMemcpyAsync(h_job_events, d_jev, sizeof(uint64_t) * n_jobs, cudaMemcpyDeviceToHost, s));
            if (h_reward) CT_CUDA(cudaMemcpyAsync(h_reward, d_reward, sizeof(float) * total, cudaMemcpyDeviceToHost, s));
            if (h_lag) CT_CUDA(cudaMemcpyAsync(h_lag, d_lag, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, s));
        }
    }
    cudaError_t e = cudaStreamSynchronize(s);      // (also on failure: the caller's host buffers must not be in use afterwards)
    if (rc != CT_OK) return rc;
    if (e != cudaSuccess) return fail(CT_ERR_CUDA, "kernel execution failed: %s", cudaGetErrorString(e));
    return CT_OK;
}

}  // extern "C"

namespace {
// Body of ct_rewards_submit after the stream exists: allocate, copy, launch, reduce.  On success the
// result buffers leave the scratch guard and belong to the ticket; on any failure the guard frees them.
int submit_on_stream(ct_engine* eng, Ticket& t, const uint64_t* handles, const uint32_t* n_traj, const uint64_t* traj_base,
                     const float* job_cost, int32_t n_jobs, const ct_sim_config* cfg, uint64_t seed, const float* h_params)
{
    cudaStream_t s = t.stream;
    StreamScratch scratch(s);
    CT_CUDA(scratch.alloc(&t.d_reward, sizeof(float) * t.total));
    CT_CUDA(scratch.alloc(&t.d_lag, sizeof(int32_t) * t.total));
    CT_CUDA(scratch.alloc(&t.d_events, sizeof(uint32_t) * t.total));
    CT_CUDA(scratch.alloc(&t.d_mean, sizeof(double) * n_jobs));
    CT_CUDA(scratch.alloc(&t.d_jev, sizeof(uint64_t) * n_jobs));
    if (h_params) {
        const size_t np = (size_t)params_per_trajectory(handles, n_jobs);
        CT_CUDA(scratch.alloc(&t.d_params, sizeof(float) * t.total * np));
        CT_CUDA(cudaMemcpyAsync(t.d_params, h_params, sizeof(float) * t.total * np, cudaMemcpyHostToDevice, s));
    }
    int rc = ct_launch_rewards(eng, handles, n_traj, traj_base, job_cost, n_jobs, cfg, seed, s, t.d_params, t.d_reward,
                               t.d_lag, t.d_events, nullptr, 0);
    if (rc == CT_OK) rc = ct_reduce_jobs(eng, n_traj, n_jobs, s, t.d_reward, t.d_events, t.d_mean, t.d_jev);
    if (rc != CT_OK) return rc;
    for (void* p : {(void*)t.d_reward, (void*)t.d_lag, (void*)t.d_events, (void*)t.d_mean, (void*)t.d_jev, (void*)t.d_params})
        if (p) scratch.keep(p);
    return CT_OK;
}

void release_ticket(Ticket& t)
{
    cudaStream_t s = t.stream;
    for (void* p : {(void*)t.d_reward, (void*)t.d_lag, (void*)t.d_events, (void*)t.d_mean, (void*)t.d_jev, (void*)t.d_params})
        if (p) cudaFreeAsync(p, s);
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
}
}  // namespace

extern "C" {

// Asynchronous pair: submit enqueues the batch on a stream of its own and returns at once, so
// the kernels of consecutive batches overlap (the tail of one launch -- a few long trajectories --
// runs next to the body of the following one).  wait copies the results out and releases the ticket.
int ct_rewards_submit(ct_engine* eng, const uint64_t* handles, const uint32_t* n_traj, const uint64_t* traj_base,
                      const float* job_cost, int32_t n_jobs, const ct_sim_config* cfg, uint64_t seed, const float* h_params,
                      int64_t* ticket)
{
    NvtxRange nvtx_range("ct_rewards_submit");
    if (!eng || !ticket) return fail(CT_ERR_INVALID, "engine or ticket is NULL");
    if (n_jobs < 1 || !handles || !n_traj || !traj_base) return fail(CT_ERR_INVALID, "bad job arrays");
    CT_CUDA(cudaSetDevice(eng->device));
    Ticket t;
    t.n_jobs = n_jobs;
    for (int j = 0; j < n_jobs; ++j) t.total += n_traj[j];
    if (t.total == 0) return fail(CT_ERR_INVALID, "empty batch");
    if (h_params) {
        const int rcp = check_host_params(h_params, (size_t)t.total, (size_t)params_per_trajectory(handles, n_jobs));
        if (rcp != CT_OK) return rcp;
    }
    CT_CUDA(cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking));
    cudaStream_t s = t.stream;
    int rc = submit_on_stream(eng, t, handles, n_traj, traj_base, job_cost, n_jobs, cfg, seed, h_params);
    if (rc != CT_OK) {
        // (the buffers were released in stream order by submit_on_stream's scratch guard)
        cudaStreamSynchronize(s);
        cudaStreamDestroy(s);
        return rc;
    }
    *ticket = eng->next_ticket++;
    eng->tickets[*ticket] = t;
    return CT_OK;
}

int ct_rewards_wait(ct_engine* eng, int64_t ticket, double* h_job_mean, uint64_t* h_job_events, float* h_reward,
                    int32_t* h_lag)
{
    NvtxRange nvtx_range("ct_rewards_wait");
    if (!eng) return fail(CT_ERR_INVALID, "engine is NULL");
    auto it = eng->tickets.find(ticket);
    if (it == eng->tickets.end()) return fail(CT_ERR_INVALID, "unknown ticket %lld", (long long)ticket);
    Ticket t = it->second;
    eng->tickets.erase(it);
    CT_CUDA(cudaSetDevice(eng->device));
    cudaStream_t s = t.stream;
    cudaError_t e = cudaSuccess, e2;
    if (h_job_mean && (e2 = cudaMemcpyAsync(h_job_mean, t.d_mean, sizeof(double) * t.n_jobs, cudaMemcpyDeviceToHost, s))) e = e2;
    if (h_job_events && (e2 = cudaMemcpyAsync(h_job_events, t.d_jev, sizeof(uint64_t) * t.n_jobs, cudaMemcpyDeviceToHost, s))) e = e2;
    if (h_reward && (e2 = cudaMemcpyAsync(h_reward, t.d_reward, sizeof(float) * t.total, cudaMemcpyDeviceToHost, s))) e = e2;
    if (h_lag && (e2 = cudaMemcpyAsync(h_lag, t.d_lag, sizeof(int32_t) * t.total, cudaMemcpyDeviceToHost, s))) e = e2;
    cudaFreeAsync(t.d_reward, s);
    cudaFreeAsync(t.d_lag, s);
    cudaFreeAsync(t.d_events, s);
    cudaFreeAsync(t.d_mean, s);
    cudaFreeAsync(t.d_jev, s);
    if (t.d_params) cudaFreeAsync(t.d_params, s);
    e2 = cudaStreamSynchronize(s);
    if (e2 != cudaSuccess) e = e2;
    cudaStreamDestroy(s);
    if (e != cudaSuccess) return fail(CT_ERR_CUDA, "asynchronous reward batch failed: %s", cudaGetErrorString(e));
    return CT_OK;
}

int ct_reward_from_series_host(ct_engine* eng, const uint8_t* h_series, int32_t n, int32_t n_species, int32_t n_dec,
                               float* h_reward, int32_t* h_lag)
{
    if (!eng) return fail(CT_ERR_INVALID, "engine is NULL");
    if (n < 0 || !h_series || !h_reward || !h_lag) return fail(CT_ERR_INVALID, "bad arguments");
    if (n_species < 1 || n_species > CT_MAX_SPECIES || n_dec < 2 || n_dec > CT_MAX_NDEC)
        return fail(CT_ERR_INVALID, "bad n_species/n_dec %d/%d", n_species, n_dec);
    if (n == 0) return CT_OK;
    CT_CUDA(cudaSetDevice(eng->device));
    cudaStream_t s = eng->stream;
    const size_t bytes = (size_t)n * n_species * n_dec;
    uint8_t* d_series = nullptr;
    float* d_reward = nullptr;
    int32_t* d_lag = nullptr;
    {
        StreamScratch scratch(s);       // freed in stream order at the end of this block, on every path
        CT_CUDA(scratch.alloc(&d_series, bytes));
        CT_CUDA(scratch.alloc(&d_reward, sizeof(float) * n));
        CT_CUDA(scratch.alloc(&d_lag, sizeof(int32_t) * n));
        CT_CUDA(cudaMemcpyAsync(d_series, h_series, bytes, cudaMemcpyHostToDevice, s));
        ct::reward_from_series_kernel<<<(n + 3) / 4, 128, 0, s>>>(d_series, n, n_species, n_dec, d_reward, d_lag);
        CT_CUDA(cudaGetLastError());
        eng->launches += 1;
        CT_CUDA(cudaMemcpyAsync(h_reward, d_reward, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
        CT_CUDA(cudaMemcpyAsync(h_lag, d_lag, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    }
    CT_CUDA(cudaStreamSynchronize(s));
    return CT_OK;
}

}  // extern "C"
