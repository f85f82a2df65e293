// Batched Gillespie SSA + in-kernel oscillation reward for the pairwise TF-network model
// (SPEC.md sections 2-8).  One trajectory per lane, one topology per warp at a time, persistent
// warps that pull 64-trajectory tiles off a global counter and refill lanes as trajectories end.
// The per-trajectory record (M x n_dec companded bytes) lives in shared memory and is reduced to
// (reward, lag) by the whole warp with shuffles when the lane finishes; it never reaches HBM.
//
// Reference boundary: this is what get_reward (reference circuitree/circuitree.py:135-152)
// computes for the topology parse_genotype (reference circuitree/models.py:441-480) describes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

namespace ct {

constexpr int kThreads = 128;        // 4 warps per CTA
constexpr int kTile = 64;            // trajectories per tile
constexpr int kSerStride = 128;      // bytes per species per lane (>= n_dec)
constexpr unsigned kFull = 0xffffffffu;
constexpr int kTopoStride = 5;       // edge (i,j) sits at bits 2*(i*5+j) of the topology code
constexpr float kSampleCap = 4095.0f;
constexpr int kMaxSpecies = 5;

struct Job {                         // 32 bytes, global memory
    uint64_t topo;                   // edge-type matrix + n_species (bits 56..59)
    uint64_t traj_base;              // trajectory id of the job's first trajectory
    uint32_t n_traj;
    uint32_t out_offset;             // output slot of the job's first trajectory
    uint32_t tile_start;             // index of the job's first tile
    uint32_t pad;
};

struct LaunchArgs {
    const Job* jobs;
    uint32_t n_jobs;
    uint32_t n_tiles;
    uint32_t* tile_counter;          // zeroed before launch
    const float* params;             // nullable [total][10]
    float* reward;
    int32_t* lag;                    // nullable
    uint32_t* events;                // nullable
    uint8_t* series;                 // nullable debug dump [total][series_stride]
    int32_t series_stride;
    float rng_lo[10];
    float rng_span[10];
    uint32_t log_mask;
    uint32_t seed_lo, seed_hi;
    int32_t nt, decim, n_dec;
    float dt, init_mean;
};

__host__ __device__ __forceinline__ uint32_t topo_edge(uint64_t topo, int i, int j)
{
    return (uint32_t)(topo >> (2 * (i * kTopoStride + j))) & 3u;
}
__host__ __device__ __forceinline__ uint32_t topo_nspecies(uint64_t topo) { return (uint32_t)(topo >> 56) & 15u; }

// ---- warp-cooperative reward (SPEC.md section 7) ---------------------------------------------
// ser: the finished lane's record [nsp][kSerStride] bytes in shared memory; z: per-warp scratch
// of kSerStride floats.  Every lane of the warp calls this; every lane returns the result.
__device__ __forceinline__ void warp_reward(const uint8_t* ser, float* z, int nsp, int n_dec, int lane,
                                            float& reward_out, int& lag_out)
{
    const int max_lag = n_dec >> 1;
    float best = 0.0f;
    int best_lag = 0;
    for (int s = 0; s < nsp; ++s) {
        const uint8_t* q = ser + s * kSerStride;
        float y[4];
        float sum = 0.0f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int k = lane + 32 * u;
            float v = (k < n_dec) ? (float)q[k] : 0.0f;
            y[u] = v * v;
            sum += y[u];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
        const float mu = sum / (float)n_dec;
        float c0 = 0.0f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int k = lane + 32 * u;
            float d = (k < n_dec) ? (y[u] - mu) : 0.0f;
            z[k] = d;
            c0 += d * d;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c0 += __shfl_xor_sync(kFull, c0, o);
        __syncwarp();
        if (c0 > 0.0f) {   // warp-uniform
            // lane handles lags l0 = lane and l1 = lane + 32 (0..63 covers 0..max_lag <= 64)
            float r0 = 0.0f, r1 = 0.0f;
            const int l0 = lane, l1 = lane + 32;
            for (int k = 0; k < n_dec; ++k) {
                float zk = z[k];
                int k0 = k + l0, k1 = k + l1;
                r0 += (k0 < n_dec) ? zk * z[k0] : 0.0f;
                r1 += (k1 < n_dec) ? zk * z[k1] : 0.0f;
            }
            const float inv = 1.0f / c0;
            r0 *= inv;
            r1 *= inv;
            // neighbours: r(l-1), r(l+1) for both halves
            float r0_prev = __shfl_up_sync(kFull, r0, 1);            // r(lane-1)
            float r0_next = __shfl_down_sync(kFull, r0, 1);          // r(lane+1), lane 31 -> r1 of lane 0
            float r1_first = __shfl_sync(kFull, r1, 0);              // r(32)
            float r0_last = __shfl_sync(kFull, r0, 31);              // r(31)
            float r1_prev = __shfl_up_sync(kFull, r1, 1);            // r(lane+31)
            float r1_next = __shfl_down_sync(kFull, r1, 1);          // r(lane+33)
            if (lane == 31) r0_next = r1_first;
            if (lane == 0) r1_prev = r0_last;
            float cand = 0.0f;
            int cand_lag = 0;
            if (l0 >= 1 && l0 < max_lag && r0 < r0_prev && r0 < r0_next && r0 < cand) { cand = r0; cand_lag = l0; }
            if (l1 < max_lag && r1 < r1_prev && r1 < r1_next && r1 < cand) { cand = r1; cand_lag = l1; }
            // arg-min over the warp: lowest value, ties -> smallest lag
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float oc = __shfl_xor_sync(kFull, cand, o);
                int ol = __shfl_xor_sync(kFull, cand_lag, o);
                if (oc < cand || (oc == cand && ol < cand_lag)) { cand = oc; cand_lag = ol; }
            }
            if (cand < best) { best = cand; best_lag = cand_lag; }   // strict: first species wins ties
        }
        __syncwarp();
    }
    reward_out = fminf(fmaxf(-best, 0.0f), 1.0f);
    lag_out = best_lag;
}

// ---- per-lane trajectory state ------------------------------------------------------------------
template <int MM>
struct Lane {
    float P[MM], R[MM], b[MM][MM];
    float n[MM], A[MM];      // bound molecules at promoter j: total, and on activating edges
    float blk[MM];
    float k_on, k_off1, k_off2, km0, km1, km2, km3, kp, gm, gp;
    float t_rem;
    int k, kd, dcount;
    uint32_t events;
    uint32_t ctr;           // next Philox block of stream 1
    uint32_t traj_lo, traj_hi;
    uint32_t out;           // output slot
};

// MUFU.LG2 / MUFU.RCP without the denormal and IEEE fix-up sequences (inputs are normal here)
__device__ __forceinline__ float fast_log2(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float compand_block(float s)
{
    return fminf(floorf(sqrtf(s) + 0.5f), 255.0f);
}

// One SSA step (SPEC.md section 2) executed by ALL lanes of the warp in lock-step; lanes without a
// trajectory carry zero rates (a0 = 0: nothing fires, nothing is recorded), which keeps the hot
// loop convergent so the topology tests below are warp-uniform branches.  `topo_lo/topo_hi/nsp`
// must be warp-uniform.  Returns false for a lane that has just recorded its last grid sample.
template <int MM>
__device__ __forceinline__ bool ssa_step(Lane<MM>& L, const bool active, const uint32_t topo_lo,
                                         const uint32_t topo_hi, const int nsp, uint32_t w_tau, uint32_t w_sel,
                                         uint8_t* ser, const int nt, const int decim, const float dt)
{
    auto edge = [&](int i, int j) -> uint32_t {
        const int sh = 2 * (i * kTopoStride + j);
        return sh < 32 ? (topo_lo >> sh) & 3u : (topo_hi >> (sh - 32)) & 3u;
    };
    // ---- propensities as a running sum in SPEC order; cum[r] is the sum up to reaction r ----
    float cum[4 * MM + 2 * MM * MM];
    float c = 0.0f;
#pragma unroll
    for (int j = 0; j < MM; ++j) {
        if (j < nsp) {
            const float n = L.n[j], na = L.A[j];
            const float km_a = (na > 0.0f) ? L.km1 : L.km0;
            const float km_i = (na > 0.0f) ? L.km3 : L.km2;
            c += (n > na) ? km_i : km_a;     cum[4 * j + 0] = c;
            c += L.gm * L.R[j];              cum[4 * j + 1] = c;
            c += L.kp * L.R[j];              cum[4 * j + 2] = c;
            c += L.gp * L.P[j];              cum[4 * j + 3] = c;
            const float koff = (n >= 2.0f) ? L.k_off2 : L.k_off1;
            const float kfree = L.k_on * (2.0f - n);
#pragma unroll
            for (int i = 0; i < MM; ++i) {
                if (edge(i, j)) {
                    asm volatile("");   // keep this a (warp-uniform) branch, not predicated code
                    const int r = 4 * MM + 2 * (j * MM + i);
                    c += kfree * L.P[i];     cum[r] = c;
                    c += koff * L.b[i][j];   cum[r + 1] = c;
                }
            }
        }
    }
    const float a0 = c;

    // ---- time to next event; grid samples crossed keep the pre-event state ----
    const float tau = (-0.69314718056f * fast_log2(u01(w_tau))) * fast_rcp(a0);   // a0 == 0 -> +inf
    L.t_rem -= tau;
    if (active && L.t_rem < 0.0f) {
        do {
#pragma unroll
            for (int j = 0; j < MM; ++j) L.blk[j] += fminf(L.P[j], kSampleCap);
            L.k += 1;
            if (--L.dcount == 0) {
                L.dcount = decim;
#pragma unroll
                for (int j = 0; j < MM; ++j) {
                    if (j < nsp) ser[j * kSerStride + L.kd] = (uint8_t)compand_block(L.blk[j]);
                    L.blk[j] = 0.0f;
                }
                L.kd += 1;
            }
            L.t_rem += dt;
        } while (L.t_rem < 0.0f && L.k < nt);
        if (L.k >= nt) return false;
    }

    // ---- select and fire: s_r = [target < cum_r] is monotone in r, so reaction r fires iff
    //      s_r - s_{r-1} = 1; no reaction fires when a0 = 0 or rounding leaves target >= a0 ----
    const float target = u01(w_sel) * a0;
    float s_prev = 0.0f;
#pragma unroll
    for (int j = 0; j < MM; ++j) {
        if (j < nsp) {
            const float s_tx = (target < cum[4 * j + 0]) ? 1.0f : 0.0f;
            const float s_rm = (target < cum[4 * j + 1]) ? 1.0f : 0.0f;
            const float s_tl = (target < cum[4 * j + 2]) ? 1.0f : 0.0f;
            const float s_dp = (target < cum[4 * j + 3]) ? 1.0f : 0.0f;
            L.R[j] += fmaf(2.0f, s_tx, -s_prev) - s_rm;
            L.P[j] += fmaf(2.0f, s_tl, -s_rm) - s_dp;
            s_prev = s_dp;
#pragma unroll
            for (int i = 0; i < MM; ++i) {
                const uint32_t e = edge(i, j);
                if (e) {
                    asm volatile("");
                    const int r = 4 * MM + 2 * (j * MM + i);
                    const float s_b = (target < cum[r]) ? 1.0f : 0.0f;
                    const float s_u = (target < cum[r + 1]) ? 1.0f : 0.0f;
                    const float d = fmaf(2.0f, s_b, -s_prev) - s_u;   // +1 bind, -1 unbind
                    L.b[i][j] += d;
                    L.P[i] -= d;
                    L.n[j] += d;
                    if (e == 1u) L.A[j] += d;
                    s_prev = s_u;
                }
            }
        }
    }
    L.events += active ? 1u : 0u;
    return true;
}

// A lane without a trajectory: zero rates and state so that ssa_step is a no-op for it.
template <int MM>
__device__ __forceinline__ void lane_park(Lane<MM>& L)
{
    L.k_on = L.k_off1 = L.k_off2 = L.km0 = L.km1 = L.km2 = L.km3 = L.kp = L.gm = L.gp = 0.0f;
#pragma unroll
    for (int j = 0; j < MM; ++j) {
        L.P[j] = L.R[j] = L.blk[j] = L.n[j] = L.A[j] = 0.0f;
#pragma unroll
        for (int i = 0; i < MM; ++i) L.b[i][j] = 0.0f;
    }
    L.t_rem = 0.0f;
}

// Start trajectory `idx` of `job` on this lane (SPEC.md sections 3-5).
template <int MM>
__device__ __forceinline__ void lane_init(Lane<MM>& L, const LaunchArgs& A, const Job& job, uint32_t idx, int nsp)
{
    const uint64_t traj = job.traj_base + idx;
    L.traj_lo = (uint32_t)traj;
    L.traj_hi = (uint32_t)(traj >> 32);
    L.out = job.out_offset + idx;
    float p[12];
#pragma unroll
    for (int blk = 0; blk < 3; ++blk) {
        Philox4 x = philox4x32_10(blk, 0u, L.traj_lo, L.traj_hi, A.seed_lo, A.seed_hi);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int i = 4 * blk + w;
            if (i < 10) {
                float v = A.rng_lo[i] + u01(x.w[w]) * A.rng_span[i];
                p[i] = ((A.log_mask >> i) & 1u) ? exp2f(v * 3.32192809489f) : v;
            }
        }
    }
    if (A.params) {
        const float* src = A.params + (size_t)L.out * 10;
#pragma unroll
        for (int i = 0; i < 10; ++i) p[i] = src[i];
    }
    L.k_on = p[0]; L.k_off1 = p[1]; L.k_off2 = p[2];
    L.km0 = p[3]; L.km1 = p[4]; L.km2 = p[5]; L.km3 = p[6];
    L.kp = p[7]; L.gm = p[8]; L.gp = p[9];

    const float e0 = expf(-A.init_mean);
#pragma unroll
    for (int blk = 0; blk < (MM + 3) / 4; ++blk) {
        Philox4 x = philox4x32_10(3 + blk, 0u, L.traj_lo, L.traj_hi, A.seed_lo, A.seed_hi);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int j = 4 * blk + w;
            if (j < MM) {
                float u = u01(x.w[w]);
                float pk = e0, F = e0;
                int kk = 0;
                while (u > F && kk < 1000) { ++kk; pk *= A.init_mean / (float)kk; F += pk; }
                L.P[j] = (j < nsp) ? (float)kk : 0.0f;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < MM; ++j) {
        L.R[j] = 0.0f;
        L.blk[j] = 0.0f;
        L.n[j] = 0.0f;
        L.A[j] = 0.0f;
#pragma unroll
        for (int i = 0; i < MM; ++i) L.b[i][j] = 0.0f;
    }
    L.t_rem = 0.0f;
    L.k = 0; L.kd = 0; L.dcount = A.decim;
    L.events = 0;
    L.ctr = 0;
}

// kMixed = false: a warp runs one topology at a time (uniform edge branches; a tile with another
// topology waits until the warp has drained) -- best for big jobs.  kMixed = true: every lane carries
// its own topology, tiles are handed out without draining -- best for MCTS batches of many small jobs.
// Results are identical in both modes.
template <int MM, bool kMixed>
__global__ void __launch_bounds__(kThreads, MM <= 3 ? 4 : (MM == 4 ? 3 : 2)) ssa_pairwise_kernel(const LaunchArgs A)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t* ser = smem + (size_t)threadIdx.x * (MM * kSerStride);
    uint8_t* warp_ser = smem + (size_t)(warp * 32) * (MM * kSerStride);
    float* z = reinterpret_cast<float*>(smem + (size_t)kThreads * (MM * kSerStride)) + warp * kSerStride;

    Lane<MM> L;
    lane_park<MM>(L);
    L.k = 0; L.kd = 0; L.dcount = 1; L.events = 0; L.ctr = 0; L.traj_lo = L.traj_hi = 0; L.out = 0;
    bool active = false, finished = false;
    uint32_t topo_lo = 0, topo_hi = 0;   // uniform mode: the warp's topology; mixed mode: this lane's topology
    // warp-uniform tile state
    Job job;                        // job of the current tile
    job.topo = 0; job.traj_base = 0; job.n_traj = 0; job.out_offset = 0; job.tile_start = 0; job.pad = 0;
    uint32_t pos = 0, end = 0;      // next / one-past-last trajectory index (within job) of the current tile
    bool have_pending = false, out_of_work = false;
    Job pend;                       // drawn tile waiting for the warp to drain (different topology)
    pend = job;
    uint32_t pend_pos = 0, pend_end = 0;
    uint64_t topo = 0;
    int nsp = 0;

    for (;;) {
        const unsigned idle = __ballot_sync(kFull, !active);
        if (idle) {
            // 1. reduce finished trajectories to (reward, lag), one lane at a time, whole warp helping
            unsigned fin = __ballot_sync(kFull, finished);
            while (fin) {
                const int src = __ffs(fin) - 1;
                fin &= fin - 1;
                float rw; int lg;
                const uint8_t* sser = warp_ser + (size_t)src * (MM * kSerStride);
                const int nsp_src = kMixed ? (int)((__shfl_sync(kFull, topo_hi, src) >> 24) & 15u) : nsp;
                warp_reward(sser, z, nsp_src, A.n_dec, lane, rw, lg);
                const uint32_t out = __shfl_sync(kFull, L.out, src);
                if (A.series) {
                    for (int s = 0; s < nsp_src; ++s)
                        for (int k = lane; k < A.n_dec; k += 32)
                            A.series[(size_t)out * A.series_stride + s * A.n_dec + k] = sser[s * kSerStride + k];
                }
                if (lane == src) {
                    A.reward[out] = rw;
                    if (A.lag) A.lag[out] = lg;
                    if (A.events) A.events[out] = L.events;
                    finished = false;
                }
            }
            // 2. hand out trajectories of the current tile; draw new tiles as needed
            unsigned need = idle;
            for (;;) {
                if (pos < end) {
                    const int rank = __popc(need & ((1u << lane) - 1u));
                    const uint32_t avail = end - pos;
                    const bool take = !active && (uint32_t)rank < avail;
                    if (take) {
                        lane_init<MM>(L, A, job, pos + rank, nsp);
                        active = true;
                        if (kMixed) { topo_lo = (uint32_t)topo; topo_hi = (uint32_t)(topo >> 32); }
                    }
                    const uint32_t n_need = __popc(need);
                    pos += (n_need < avail) ? n_need : avail;
                    need = __ballot_sync(kFull, !active);
                    if (!need) break;
                }
                // current tile exhausted and lanes still idle
                if (!have_pending && !out_of_work) {
                    uint32_t t = 0;
                    if (lane == 0) t = atomicAdd(A.tile_counter, 1u);
                    t = __shfl_sync(kFull, t, 0);
                    if (t >= A.n_tiles) {
                        out_of_work = true;
                    } else {
                        // binary search: last job with tile_start <= t
                        uint32_t lo = 0, hi = A.n_jobs - 1;
                        while (lo < hi) {
                            uint32_t mid = (lo + hi + 1) >> 1;
                            if (A.jobs[mid].tile_start <= t) lo = mid; else hi = mid - 1;
                        }
                        pend = A.jobs[lo];
                        pend_pos = (t - pend.tile_start) * kTile;
                        pend_end = min(pend_pos + (uint32_t)kTile, pend.n_traj);
                        have_pending = true;
                    }
                }
                if (have_pending) {
                    const bool same = (pend.topo == topo);
                    const bool drained = (__ballot_sync(kFull, active) == 0u);
                    if (kMixed || same || drained) {
                        job = pend; pos = pend_pos; end = pend_end;
                        topo = job.topo;
                        nsp = (int)topo_nspecies(topo);
                        if (!kMixed) { topo_lo = (uint32_t)topo; topo_hi = (uint32_t)(topo >> 32); }
                        have_pending = false;
                        continue;
                    }
                }
                break;   // wait for running lanes (or nothing left)
            }
            if (out_of_work && __ballot_sync(kFull, active) == 0u) break;
        }

        // two steps per Philox block (SPEC.md section 3); every lane runs them, parked lanes as no-ops
        const Philox4 x = philox4x32_10(L.ctr, 1u, L.traj_lo, L.traj_hi, A.seed_lo, A.seed_hi);
        L.ctr += 1;
        // A warp-wide OR of identical values: the result is warp-uniform by construction, which lets
        // the compiler keep the topology in uniform registers and branch uniformly on its bits.
        const uint32_t u_lo = kMixed ? topo_lo : __reduce_or_sync(kFull, topo_lo);
        const uint32_t u_hi = kMixed ? topo_hi : __reduce_or_sync(kFull, topo_hi);
        const int u_nsp = (int)((u_hi >> 24) & 15u);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const bool alive = ssa_step<MM>(L, active, u_lo, u_hi, u_nsp, h ? x.w[2] : x.w[0],
                                            h ? x.w[3] : x.w[1], ser, A.nt, A.decim, A.dt);
            if (__any_sync(kFull, !alive)) {   // rare, warp-uniform: keep the parking code off the hot path
                if (!alive) {
                    active = false;
                    finished = true;
                    lane_park<MM>(L);   // (mixed mode keeps topo_hi: the reward stage needs the species count)
                }
            }
        }
    }
}

// Stand-alone reward kernel for the test hook: one warp per series.
__global__ void reward_from_series_kernel(const uint8_t* series, int n, int nsp, int n_dec, float* reward, int* lag)
{
    __shared__ __align__(16) uint8_t ser[4][kMaxSpecies * kSerStride];
    __shared__ float zz[4][kSerStride];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int idx = blockIdx.x * 4 + warp;
    if (idx >= n) return;
    for (int s = 0; s < nsp; ++s)
        for (int k = lane; k < n_dec; k += 32) ser[warp][s * kSerStride + k] = series[((size_t)idx * nsp + s) * n_dec + k];
    __syncwarp();
    float rw; int lg;
    warp_reward(ser[warp], zz[warp], nsp, n_dec, lane, rw, lg);
    if (lane == 0) { reward[idx] = rw; lag[idx] = lg; }
}

// Deterministic per-job reduction: one warp per job, fixed summation order.
__global__ void reduce_jobs_kernel(const uint32_t* job_offset, int n_jobs, const float* reward, const uint32_t* events,
                                   double* job_mean, unsigned long long* job_events)
{
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= n_jobs) return;
    const uint32_t b = job_offset[j], e = job_offset[j + 1];
    double s = 0.0;
    unsigned long long ev = 0ull;
    for (uint32_t x = b + lane; x < e; x += 32) {
        s += (double)reward[x];
        if (events) ev += events[x];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(kFull, s, o);
        ev += __shfl_xor_sync(kFull, ev, o);
    }
    if (lane == 0) {
        job_mean[j] = (e > b) ? s / (double)(e - b) : 0.0;
        if (job_events) job_events[j] = ev;
    }
}

}  // namespace ct
