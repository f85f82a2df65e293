// Batched Gillespie SSA + in-kernel oscillation reward for the pairwise TF-network model
// (SPEC.md sections 2-8).  One trajectory per lane, one topology per warp at a time, persistent
// warps that pull 32-trajectory tiles off a global counter and refill lanes as trajectories end.
// The per-trajectory record (M x n_dec companded bytes, one byte per species every `decim` grid
// points) is parked in a per-lane scratch area in global memory -- a few hundred bytes written
// over ~10^5 events, L2-resident -- and staged back into shared memory when the lane finishes,
// where the whole warp reduces it to (reward, lag) with shuffles.  Keeping it in shared memory
// (round 1: 384 bytes per lane for M = 3) capped residency at 4 CTAs per SM; now registers alone
// decide (launch bounds below), and the full trajectories still never exist anywhere.
//
// Reference boundary: this is what get_reward (reference circuitree/circuitree.py:135-152)
// computes for the topology parse_genotype (reference circuitree/models.py:441-480) describes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

// Build-time variant of the hot loop (scripts/build_variants.sh + scripts/variants.sh time variants
// side by side on the GPU; results are identical).  Measured on B200 at 2^21 repressilator
// trajectories: rolled 1.495e11 steps/s, unrolled 1.572e11.  Tried and dropped: a predicated
// (branch-free) first grid crossing (-3 %), generating the next Philox block beside the current
// block's steps (+-0 %).
#ifndef CT_BLOCKS_PER_CHECK
#define CT_BLOCKS_PER_CHECK 8  // Philox blocks (two steps each) between two looks for idle lanes in the pairwise kernels
#endif
#ifndef CT_UNROLL_H
#define CT_UNROLL_H 1         // 1: the two steps of a Philox block are unrolled into one basic block
#endif
// Resident CTAs per SM the compiler has to make room for (registers per thread <= 65536 / (128 * n)),
// for M = 3, 4, 5.  With the record out of shared memory registers alone set the residency, and the
// compiler fits the hot variants into 96 (M = 3) / 128 (M = 4, 5) registers without spilling -- but
// more resident warps made every variant SLOWER on B200 (profiles/r02_occupancy_variants.txt, 2^21
// repressilator trajectories: 4 CTAs 1.57e11 steps/s, 5 CTAs 1.43e11, 6 CTAs 1.22e11; five-component
// MCTS batches: 2 CTAs 3.1e10, 3 CTAs 3.0e10, 4 CTAs 2.5e10): the loop is not short of eligible warps
// (ncu: 1.75 eligible-but-not-selected warps per issue), it is bound by dispatch and pipe conflicts that
// tighter register allocation makes worse.  Hence the round-1 residency.
#ifndef CT_MINB3
#define CT_MINB3 4
#endif
#ifndef CT_MINB4
#define CT_MINB4 3
#endif
#ifndef CT_MINB5
#define CT_MINB5 2
#endif

// -DCT_DEBUG_BOUNDS: every computed shared/global index of the hot loops is range-checked and the
// kernel traps on a violation (compute-sanitizer is closed on this pool; scripts/sanitize_smoke.py
// runs every kernel family on such a build, profiles/r02_bounds_checked_smoke.txt).
#ifdef CT_DEBUG_BOUNDS
#define CT_CHECK(cond) do { if (!(cond)) __trap(); } while (0)
#else
#define CT_CHECK(cond) do { } while (0)
#endif

namespace ct {

constexpr int kThreads = 128;        // 4 warps per CTA
constexpr int kTile = 32;            // trajectories per tile (one per lane: finest longest-first granularity)
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
    const float* params;             // nullable [total][n_params]
    float* reward;
    int32_t* lag;                    // nullable
    uint32_t* events;                // nullable
    uint8_t* records;                // pairwise kernels: per-lane record scratch [grid * kThreads][MM * kSerStride]
    uint8_t* series;                 // nullable debug dump [total][series_stride]
    int32_t series_stride;
    unsigned long long* warp_iters;  // nullable debug counter: hot-loop iterations (2 steps each) summed over warps
    const uint32_t* order;           // nullable: launch position -> output slot, longest expected first
    uint32_t total;                  // trajectories in the launch
    float rng_lo[12];
    float rng_span[12];
    uint32_t log_mask;
    int32_t n_params;                // 10 (pairwise, SPEC.md section 4) or 12 (dimers, section 9)
    uint32_t seed_lo, seed_hi;
    uint32_t rk[20];                 // Philox round keys (k0 + r*W0, k1 + r*W1), r = 0..9
    int32_t nt, decim, n_dec;
    float dt, init_mean;
};

__host__ __device__ __forceinline__ uint32_t topo_edge(uint64_t topo, int i, int j)
{
    return (uint32_t)(topo >> (2 * (i * kTopoStride + j))) & 3u;
}
__host__ __device__ __forceinline__ uint32_t topo_nspecies(uint64_t topo) { return (uint32_t)(topo >> 56) & 15u; }
// handle -> (presence mask | nsp << 28, activation mask), one bit per cell c = i*5+j
__host__ __device__ __forceinline__ void topo_masks(uint64_t topo, uint32_t& pres, uint32_t& actm)
{
    uint32_t p = 0, a = 0;
    for (int c = 0; c < 25; ++c) {
        const uint32_t e = (uint32_t)(topo >> (2 * c)) & 3u;
        p |= (e != 0u ? 1u : 0u) << c;
        a |= (e == 1u ? 1u : 0u) << c;
    }
    pres = p | (topo_nspecies(topo) << 28);
    actm = a;
}

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
    float h[MM][MM];         // kFast only: 1.0f where edge i->j exists (each gene has at most one regulator)
    float kmb[MM];           // kFast only: transcription rate of gene j with its regulator bound (km_act or km_rep)
    float blk[MM];
    float k_on, k_on2, k_off1, k_off2, km0, km1, km2, km3, kp, gm, gp;   // k_on2 = 2 k_on
    float t_rem;
    int kd, dcount;          // record block index; grid samples left in the block
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

// A lane without a trajectory: zero rates and state so that ssa_step is a no-op for it.
template <int MM>
__device__ __forceinline__ void lane_park(Lane<MM>& L)
{
    L.k_on = L.k_on2 = L.k_off1 = L.k_off2 = L.km0 = L.km1 = L.km2 = L.km3 = L.kp = L.gm = L.gp = 0.0f;
#pragma unroll
    for (int j = 0; j < MM; ++j) {
        L.P[j] = L.R[j] = L.blk[j] = L.n[j] = L.A[j] = L.kmb[j] = 0.0f;
#pragma unroll
        for (int i = 0; i < MM; ++i) L.b[i][j] = 0.0f;
    }
    L.t_rem = __int_as_float(0x7f800000);   // +inf: never "crosses a grid point" (ssa_step)
}

// One SSA step (SPEC.md section 2) executed by ALL lanes of the warp in lock-step; lanes without a
// trajectory carry zero rates (a0 = 0: nothing fires, nothing is recorded), which keeps the hot
// loop convergent.  `pres` has bit c = i*5+j set when edge i->j exists, `actm` when it activates;
// species count nsp.  Returns false for a lane that has just recorded its last grid sample.
//
// Reaction order (SPEC.md section 2), per gene j: transcription, mRNA decay, translation, protein
// decay, binding of regulator 0..M-1, unbinding of regulator 0..M-1.  The common path treats
// "some regulator binds at j" (propensity kfree_j * sum_i P_i) and "some regulator unbinds at j"
// (koff_j * n_j) as two lumped reactions, so it costs 6 running-sum entries per gene whatever the
// topology; which regulator it is gets resolved from the position of the target inside the lump,
// in a block that only runs when some lane of the warp fired a (un)binding at that gene.  This is
// the same direct method: the flat reaction is the one whose running sum first exceeds the target.
//
// The running sum is kept per gene (M independent chains of six instead of one chain of 6M): gene
// j's entries are lc[6j..6j+5], its total g_j, a0 = ((g_0 + g_1) + ...).  Selection walks the genes
// with t_0 = u a0, t_{j+1} = t_j - g_j and compares t_j with gene j's local sums.  IEEE subtraction
// keeps the sign exact (t_j < g_j  <=>  t_{j+1} < 0), so s_r = [t_j < lc_r] is monotone along the
// whole flat reaction list: reaction r fires iff s_r - s_{r-1} = 1, the state updates are
// differences of neighbouring s values, and a reaction with zero propensity can never fire.
//
// kFast (every gene has at most one regulator; one topology per warp; every job has MM species):
// the regulator of a lumped (un)binding is known, so there is no resolution block and no warp vote;
// the edge matrix is the 0/1 float matrix L.h (sum_i h_ij P_i, P_i -= h_ij d: FMA pipe, no
// predicates), b_ij == n_j and A_j == n_j or 0 are not tracked.  Bit-identical to the general path.
template <int MM, bool kAllSpecies, bool kFast>
__device__ __forceinline__ void ssa_step(Lane<MM>& L, bool& active, bool& finished, const uint32_t pres,
                                         const uint32_t actm, const int nsp_in, uint32_t w_tau, uint32_t w_sel,
                                         uint8_t* __restrict__ records, const uint32_t rec_off, const int nt, const int decim,
                                         const int n_dec, const float dt, const int h)
{
    const int nsp = kAllSpecies ? MM : nsp_in;   // kAllSpecies: every job has MM species, no per-gene tests
    auto has = [&](int i, int j) -> bool { return (pres >> (i * kTopoStride + j)) & 1u; };
    auto is_act = [&](int i, int j) -> bool { return (actm >> (i * kTopoStride + j)) & 1u; };
    // ---- propensities: one running sum per gene ----
    float lc[6 * MM], g[MM], S[MM], kfree[MM], koff[MM];
#pragma unroll
    for (int j = 0; j < MM; ++j) {
        g[j] = 0.0f;
        if (j < nsp) {   // (entries of absent genes stay unset: pass 2 never consults them)
            const float n = L.n[j];
            float c;
            if (kFast) {
                c = (n > 0.0f) ? L.kmb[j] : L.km0;
            } else {
                const float na = L.A[j];
                const float km_a = (na > 0.0f) ? L.km1 : L.km0;
                const float km_i = (na > 0.0f) ? L.km3 : L.km2;
                c = (n > na) ? km_i : km_a;
            }
            lc[6 * j + 0] = c;
            c = fmaf(L.gm, L.R[j], c);       lc[6 * j + 1] = c;
            c = fmaf(L.kp, L.R[j], c);       lc[6 * j + 2] = c;
            c = fmaf(L.gp, L.P[j], c);       lc[6 * j + 3] = c;
            float sp = 0.0f;
#pragma unroll
            for (int i = 0; i < MM; ++i) {
                if (kFast) sp = fmaf(L.h[i][j], L.P[i], sp);
                else if (has(i, j)) sp += L.P[i];
            }
            S[j] = sp;
            koff[j] = (n >= 2.0f) ? L.k_off2 : L.k_off1;
            kfree[j] = fmaf(-L.k_on, n, L.k_on2);    // k_on (2 - n), exact for n = 0, 1, 2
            c = fmaf(kfree[j], sp, c);       lc[6 * j + 4] = c;
            c = fmaf(koff[j], n, c);         lc[6 * j + 5] = c;
            g[j] = c;
        }
    }
    float a0 = g[0];
#pragma unroll
    for (int j = 1; j < MM; ++j) a0 += g[j];

    // ---- time to next event; grid samples crossed keep the pre-event state ----
    // A lane without a trajectory carries t_rem = +inf and a0 = 0: tau = +inf, t_rem turns NaN and stays NaN, so
    // "t_rem < 0" is false for it without a test of `active`.  (u01 is never 1, so tau of a live lane is never NaN.)
    const float tau = (-0.69314718056f * fast_log2(u01(w_tau))) * fast_rcp(a0);   // a0 == 0 -> +inf
    L.t_rem -= tau;
    float t = u01(w_sel) * a0;      // selection target (a lane that finishes below fires nothing: target beyond every sum)
    // No early return: every lane must reach the warp votes below.  In a 32-lane warp some lane crosses a grid point
    // in four steps out of ten, nearly always alone, so this divergent path is part of the hot loop and kept short:
    // the grid index is not counted per sample, only per record block (k = kd * decim + decim - dcount), and the end
    // of the trajectory can only be met at a block boundary.
    if (L.t_rem < 0.0f) {
        do {
#pragma unroll
            for (int j = 0; j < MM; ++j) L.blk[j] += fminf(L.P[j], kSampleCap);
            if (--L.dcount == 0) {
                // block boundary: record (a trailing block shorter than `decim` is simulated but not recorded)
                if (L.kd < n_dec) {
#pragma unroll
                    for (int j = 0; j < MM; ++j) {
                        CT_CHECK(L.kd >= 0 && L.kd < kSerStride);
                        if (j < nsp) records[rec_off + (uint32_t)(j * kSerStride + L.kd)] = (uint8_t)compand_block(L.blk[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < MM; ++j) L.blk[j] = 0.0f;
                L.kd += 1;
                const int left = nt - L.kd * decim;        // grid points after this boundary
                if (left <= 0) {
                    // last grid sample taken: the lane parks until the warp reduces its record.
                    // events fired = steps executed before this one (the crossing step fires nothing and is not
                    // counted, SPEC.md section 6): step h of Philox block ctr - 1 -- no per-step counter in the hot loop
                    L.events = 2u * (L.ctr - 1u) + (uint32_t)h;
                    active = false;
                    finished = true;
                    t = 3.0e38f;
                    lane_park<MM>(L);   // zero rates and state: the updates below add zeros to it
                    break;
                }
                L.dcount = min(decim, left);
            }
            L.t_rem += dt;
        } while (L.t_rem < 0.0f);
    }

    // ---- select and fire ----
    // s_r = [t < lc_r] as a float comes from the FMA pipe: saturate((lc_r - t) * 2^60) is 1 when
    // lc_r > t (the smallest positive difference of two such floats is far above 2^-60) and 0 when
    // lc_r <= t; the scaling by a power of two is exact, so the sign is that of the exact difference.
    // The compare/select pipe runs at half the FMA pipe's rate and is the busier one in this loop.
    constexpr float kBig = 1152921504606846976.0f;   // 2^60
    float s_prev = 0.0f;
#pragma unroll
    for (int j = 0; j < MM; ++j) {
        float d = 0.0f, d_b = 0.0f;
        if (j < nsp) {
            const float tb = -t * kBig;             // (-inf for a finished lane: every s is 0)
            const float s_tx = __saturatef(fmaf(lc[6 * j + 0], kBig, tb));
            const float s_rm = __saturatef(fmaf(lc[6 * j + 1], kBig, tb));
            const float s_tl = __saturatef(fmaf(lc[6 * j + 2], kBig, tb));
            const float s_dp = __saturatef(fmaf(lc[6 * j + 3], kBig, tb));
            const float s_b = __saturatef(fmaf(lc[6 * j + 4], kBig, tb));
            const float s_u = __saturatef(fmaf(lc[6 * j + 5], kBig, tb));
            L.R[j] += fmaf(2.0f, s_tx, -s_prev) - s_rm;
            L.P[j] += fmaf(2.0f, s_tl, -s_rm) - s_dp;
            d_b = s_b - s_dp;                        // 1 if some regulator binds at j
            d = fmaf(2.0f, s_b, -s_dp) - s_u;        // +1 bind, -1 unbind, 0 neither
            s_prev = s_u;
        }
        if (kFast) {
            // the one regulator of gene j (if any) is the one that (un)binds
            L.n[j] += d;
#pragma unroll
            for (int i = 0; i < MM; ++i) L.P[i] = fmaf(-L.h[i][j], d, L.P[i]);
        } else if (__any_sync(kFull, d != 0.0f)) {
            // Resolve which regulator (un)binds.  Warp-uniform vote: the hot loop is convergent here.
            const bool bind = d_b != 0.0f;
            const float base = bind ? lc[6 * j + 3] : lc[6 * j + 4];
            const float scale = bind ? kfree[j] : koff[j];
            const float tot = bind ? S[j] : L.n[j];                  // molecules in the lump (>= 1 if it fired)
            float y = (t - base) * fast_rcp(scale);                  // position inside the lump, in molecules
            y = fminf(y, __uint_as_float(__float_as_uint(tot) - 1u)); // strictly below the total: one regulator fires
            if (d == 0.0f) y = 3.0e38f;                              // lanes that fired nothing here
            float acc = 0.0f, sq = 0.0f;
#pragma unroll
            for (int i = 0; i < MM; ++i) {
                if (has(i, j)) {
                    acc += bind ? L.P[i] : L.b[i][j];
                    const float s = (y < acc) ? 1.0f : 0.0f;
                    const float dd = (s - sq) * d;                   // +-1 for the chosen regulator
                    sq = s;
                    L.b[i][j] += dd;
                    L.P[i] -= dd;
                    if (is_act(i, j)) L.A[j] += dd;
                }
            }
            L.n[j] += d;
        }
        t -= g[j];
    }
}

// Parameter vector of one trajectory (SPEC.md sections 3-4): explicit if given, else sampled.
__device__ __forceinline__ void trajectory_params(const LaunchArgs& A, uint32_t traj_lo, uint32_t traj_hi, uint32_t out,
                                                  float (&p)[12])
{
    if (A.params) {
        const float* src = A.params + (size_t)out * A.n_params;
#pragma unroll
        for (int i = 0; i < 12; ++i) p[i] = (i < A.n_params) ? src[i] : 0.0f;
        return;
    }
#pragma unroll
    for (int blk = 0; blk < 3; ++blk) {
        Philox4 x = philox4x32_10(blk, 0u, traj_lo, traj_hi, A.seed_lo, A.seed_hi);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int i = 4 * blk + w;
            {
                float v = A.rng_lo[i] + u01(x.w[w]) * A.rng_span[i];
                p[i] = ((A.log_mask >> i) & 1u) ? exp2f(v * 3.32192809489f) : v;
            }
        }
    }
}

// last job whose out_offset <= slot (jobs are in output order; empty jobs never match a slot)
__device__ __forceinline__ uint32_t job_of_slot(const LaunchArgs& A, uint32_t slot)
{
    uint32_t lo = 0, hi = A.n_jobs - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (A.jobs[mid].out_offset <= slot) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Pre-pass for longest-first scheduling: expected number of events of every trajectory, as a sort
// key.  The event rate of an unregulated gene is 2 km (1 + kp/gm) (transcription, translation and
// the two decays at steady state); job_cost[j] (nullable) scales it per topology, e.g. with the
// events per trajectory the caller has observed for that topology.  Ascending key order =
// descending cost; in uniform mode the job index is the major key so jobs stay contiguous.
__global__ void cost_key_kernel(const LaunchArgs A, const float* job_cost, int job_major, unsigned long long* keys,
                                uint32_t* slots)
{
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= A.total) return;
    const uint32_t j = job_of_slot(A, slot);
    const Job job = A.jobs[j];
    const uint64_t traj = job.traj_base + (slot - job.out_offset);
    float p[12];
    trajectory_params(A, (uint32_t)traj, (uint32_t)(traj >> 32), slot, p);
    float cost = p[3] * (1.0f + p[7] / p[8]);
    if (job_cost) cost *= job_cost[j];
    if (!(cost >= 0.0f)) cost = 0.0f;                       // NaN / negative explicit parameters
    const uint32_t bits = __float_as_uint(fminf(cost, 3.0e38f));
    uint32_t low = 0xffffffffu - bits;
    if (!job_major) {
        // mixed-topology warps: bucket the cost by octave (sign + exponent) and sort by topology
        // inside a bucket, so that neighbouring lanes mostly share their topology (less divergence)
        const uint32_t h = (uint32_t)((job.topo * 0x9E3779B97F4A7C15ull) >> 41);          // 23 bits
        low = ((0x1ffu - (bits >> 23)) << 23) | h;
    }
    keys[slot] = ((unsigned long long)(job_major ? j : 0u) << 32) | (unsigned long long)low;
    slots[slot] = slot;
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
    trajectory_params(A, L.traj_lo, L.traj_hi, L.out, p);
    L.k_on = p[0]; L.k_on2 = 2.0f * p[0]; L.k_off1 = p[1]; L.k_off2 = p[2];
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
    // kFast tables (dead code in the other variants: nothing reads them there)
#pragma unroll
    for (int j = 0; j < MM; ++j) {
        bool act = false;
#pragma unroll
        for (int i = 0; i < MM; ++i) {
            const uint32_t e = topo_edge(job.topo, i, j);
            L.h[i][j] = e ? 1.0f : 0.0f;
            act = act || e == 1u;
        }
        L.kmb[j] = act ? L.km1 : L.km2;
    }
    L.t_rem = 0.0f;
    L.kd = 0; L.dcount = A.decim;
    L.events = 0;
    L.ctr = 0;
}

// kMixed = false: a warp runs one topology at a time (uniform edge branches; a tile with another
// topology waits until the warp has drained) -- best for big jobs.  kMixed = true: every lane carries
// its own topology, tiles are handed out without draining -- best for MCTS batches of many small jobs.
// Results are identical in both modes.
// kFast: every job's topology gives each gene at most one regulator (requires !kMixed && kAllSpecies).
template <int MM, bool kMixed, bool kAllSpecies, bool kFast = false>
__global__ void __launch_bounds__(kThreads, MM <= 3 ? CT_MINB3 : (MM == 4 ? CT_MINB4 : CT_MINB5)) ssa_pairwise_kernel(const LaunchArgs A)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // shared memory: per warp the reward scratch z[kSerStride] and the staging copy of ONE finished record
    float* z = reinterpret_cast<float*>(smem) + warp * kSerStride;
    uint8_t* stage = smem + (size_t)(kThreads / 32) * kSerStride * sizeof(float) + (size_t)warp * (MM * kSerStride);
    // this lane's record and the first record of its warp, in the global scratch area
    // (a 32-bit offset the compiler cannot re-derive from the thread ids: otherwise it rebuilds the 64-bit address
    // of the record -- two special-register reads and half a dozen integer operations -- in every step)
    uint32_t rec_off;
    {
        const uint32_t v = (blockIdx.x * (uint32_t)kThreads + threadIdx.x) * (uint32_t)(MM * kSerStride);
        asm volatile("mov.u32 %0, %1;" : "=r"(rec_off) : "r"(v));
    }
    const uint32_t warp_rec_off = rec_off - (uint32_t)lane * (uint32_t)(MM * kSerStride);

    static_assert(!kFast || (!kMixed && kAllSpecies), "kFast needs one full-size topology per warp");
    Lane<MM> L;
    lane_park<MM>(L);
#pragma unroll
    for (int j = 0; j < MM; ++j)
#pragma unroll
        for (int i = 0; i < MM; ++i) L.h[i][j] = 0.0f;
    L.kd = 0; L.dcount = 1; L.events = 0; L.ctr = 0; L.traj_lo = L.traj_hi = 0; L.out = 0;
    bool active = false, finished = false;
    // topology as two 25-bit masks + species count in bits 28..31 of `pres`
    // (uniform mode: the warp's topology; mixed mode: this lane's topology)
    uint32_t pres = 0, actm = 0;
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
    unsigned long long n_iters = 0;

    for (;;) {
        const unsigned idle = __ballot_sync(kFull, !active);
        if (idle) {
            // 1. reduce finished trajectories to (reward, lag), one lane at a time, whole warp helping
            unsigned fin = __ballot_sync(kFull, finished);
            if (fin) __syncwarp();     // the finished lanes' record bytes are visible to the whole warp
            while (fin) {
                const int src = __ffs(fin) - 1;
                fin &= fin - 1;
                float rw; int lg;
                const int nsp_src = kMixed ? (int)(__shfl_sync(kFull, pres, src) >> 28) : nsp;
                // stage the record in shared memory (L2 reads: another lane wrote it), then reduce it
                const uint32_t* grec = reinterpret_cast<const uint32_t*>(A.records + warp_rec_off + (uint32_t)src * (uint32_t)(MM * kSerStride));
                for (int w = lane; w < nsp_src * (kSerStride / 4); w += 32) reinterpret_cast<uint32_t*>(stage)[w] = __ldcg(grec + w);
                __syncwarp();
                const uint8_t* sser = stage;
                warp_reward(sser, z, nsp_src, A.n_dec, lane, rw, lg);
                const uint32_t out = __shfl_sync(kFull, L.out, src);
                CT_CHECK(out < A.total && nsp_src >= 1 && nsp_src <= MM);
                if (A.series) {
                    for (int s = 0; s < nsp_src; ++s)
                        for (int k = lane; k < A.n_dec; k += 32)
                            A.series[(size_t)out * A.series_stride + s * A.n_dec + k] = sser[s * kSerStride + k];
                }
                __syncwarp();          // everybody is done with the staged copy before the next one overwrites it
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
                        if (kMixed) {
                            // tiles are ranges of launch positions; every lane looks up its own job
                            const uint32_t slot = A.order ? A.order[pos + rank] : pos + rank;
                            const Job mine = A.jobs[job_of_slot(A, slot)];
                            const int my_nsp = (int)topo_nspecies(mine.topo);
                            lane_init<MM>(L, A, mine, slot - mine.out_offset, my_nsp);
                            topo_masks(mine.topo, pres, actm);
                        } else {
                            // tiles are ranges of positions inside one job (jobs stay contiguous in `order`)
                            const uint32_t p = job.out_offset + pos + rank;
                            const uint32_t slot = A.order ? A.order[p] : p;
                            lane_init<MM>(L, A, job, slot - job.out_offset, nsp);
                        }
                        active = true;
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
                    } else if (kMixed) {
                        pend_pos = t * kTile;
                        pend_end = min(pend_pos + (uint32_t)kTile, A.total);
                        have_pending = true;
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
                    if (kMixed) {
                        pos = pend_pos; end = pend_end;
                        have_pending = false;
                        continue;
                    }
                    if (same || drained) {
                        job = pend; pos = pend_pos; end = pend_end;
                        topo = job.topo;
                        nsp = (int)topo_nspecies(topo);
                        topo_masks(topo, pres, actm);
                        have_pending = false;
                        continue;
                    }
                }
                break;   // wait for running lanes (or nothing left)
            }
            if (out_of_work && __ballot_sync(kFull, active) == 0u) break;
        }
        n_iters += CT_BLOCKS_PER_CHECK;

        // A warp-wide OR of identical values: the result is warp-uniform by construction, which lets
        // the compiler keep the topology in uniform registers and branch uniformly on its bits.
        // (kFast reads the per-lane 0/1 edge matrix instead: no topology bits in its hot loop)
        const uint32_t u_pres = (kMixed || kFast) ? pres : __reduce_or_sync(kFull, pres);
        const uint32_t u_act = (kMixed || kFast) ? actm : __reduce_or_sync(kFull, actm);
        const int u_nsp = (int)(u_pres >> 28);
        // CT_BLOCKS_PER_CHECK Philox blocks between two looks for idle lanes (the vote, its convergence check and
        // the branch are 8 instructions); two steps per block (SPEC.md section 3); every lane runs them, parked
        // lanes as no-ops, so a lane that finishes waits at most that many blocks for its reduction and refill.
#pragma unroll 1
        for (int rep = 0; rep < CT_BLOCKS_PER_CHECK; ++rep) {
            const Philox4 x = philox4x32_10_rk(L.ctr, 1u, L.traj_lo, L.traj_hi, A.rk);
            L.ctr += 1;
#if CT_UNROLL_H
#pragma unroll
#else
#pragma unroll 1
#endif
            for (int h = 0; h < 2; ++h)
                ssa_step<MM, kAllSpecies, kFast>(L, active, finished, u_pres, u_act, u_nsp, h ? x.w[2] : x.w[0], h ? x.w[3] : x.w[1],
                                          A.records, rec_off, A.nt, A.decim, A.n_dec, A.dt, h);
        }
    }
    if (A.warp_iters && lane == 0) atomicAdd(A.warp_iters, n_iters);
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

// The same for launches with big jobs: one 256-thread block per job, every thread a fixed strided share,
// then a fixed-shape tree in shared memory -- still deterministic, 30 ms -> ~0.2 ms on a 2^22-trajectory job.
__global__ void __launch_bounds__(256) reduce_big_jobs_kernel(const uint32_t* job_offset, int n_jobs, const float* reward,
                                                            const uint32_t* events, double* job_mean, unsigned long long* job_events)
{
    __shared__ double ss[256];
    __shared__ unsigned long long se[256];
    const int j = blockIdx.x;
    if (j >= n_jobs) return;
    const uint32_t b = job_offset[j], e = job_offset[j + 1];
    double s = 0.0;
    unsigned long long ev = 0ull;
    for (uint32_t x = b + threadIdx.x; x < e; x += 256) {
        s += (double)reward[x];
        if (events) ev += events[x];
    }
    ss[threadIdx.x] = s;
    se[threadIdx.x] = ev;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) { ss[threadIdx.x] += ss[threadIdx.x + w]; se[threadIdx.x] += se[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        job_mean[j] = (e > b) ? ss[0] / (double)(e - b) : 0.0;
        if (job_events) job_events[j] = se[0];
    }
}

}  // namespace ct
