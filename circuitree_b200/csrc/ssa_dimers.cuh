// Batched Gillespie SSA + oscillation reward for the dimer model (SPEC.md section 9), the
// simulator behind get_reward for DimersGrammar states (reference circuitree/models.py:844-884 is
// the topology input).  First, table-driven version: one trajectory per lane, lanes are
// independent (grid-stride over launch positions, longest-expected-first when `order` is given),
// per-lane state in registers / local memory with the topology table read through the read-only
// cache, the per-trajectory record in shared memory, and the reward computed by the lane itself.
// It shares Philox streams, parameter sampling, companding and reward definition with the pairwise
// kernel; it does not share its lock-step scheduling (config 3 is not the headline path).
#pragma once
#include "ssa_pairwise.cuh"

namespace ct {

constexpr int kMaxIxn = 10;
constexpr int kDimThreads = 128;

struct DimerTopo {                  // 56 bytes
    uint8_t n_tf, n_comp, n_ixn, n_dim;
    uint8_t dim_a[kMaxIxn], dim_b[kMaxIxn];
    uint8_t ix_dim[kMaxIxn], ix_tgt[kMaxIxn], ix_act[kMaxIxn];
    uint8_t pad[2];
};

// SPEC.md section 7 computed by one lane from its own record (rare: once per trajectory)
__device__ __forceinline__ void lane_reward(const uint8_t* ser, int nsp, int n_dec, float& reward_out, int& lag_out)
{
    const int max_lag = n_dec >> 1;
    float best = 0.0f;
    int best_lag = 0;
    for (int s = 0; s < nsp; ++s) {
        const uint8_t* q = ser + s * kSerStride;
        float sum = 0.0f;
        for (int k = 0; k < n_dec; ++k) { const float v = (float)q[k]; sum += v * v; }
        const float mu = sum / (float)n_dec;
        float c0 = 0.0f;
        for (int k = 0; k < n_dec; ++k) { const float v = (float)q[k]; const float d = v * v - mu; c0 += d * d; }
        if (!(c0 > 0.0f)) continue;
        const float inv = 1.0f / c0;
        float r_prev2 = 0.0f, r_prev = 1.0f;    // r(l-2), r(l-1)
        for (int l = 1; l <= max_lag; ++l) {
            float acc = 0.0f;
            for (int k = 0; k + l < n_dec; ++k) {
                const float a = (float)q[k], b = (float)q[k + l];
                acc += (a * a - mu) * (b * b - mu);
            }
            const float r = acc * inv;
            // r_prev = r(l-1) is a local minimum if r(l-2) > r(l-1) < r(l); candidates are lags 1..max_lag-1
            if (l >= 2 && r_prev < r_prev2 && r_prev < r && r_prev < best) { best = r_prev; best_lag = l - 1; }
            r_prev2 = r_prev;
            r_prev = r;
        }
    }
    reward_out = fminf(fmaxf(-best, 0.0f), 1.0f);
    lag_out = best_lag;
}

__global__ void __launch_bounds__(kDimThreads) ssa_dimers_kernel(const LaunchArgs A, const DimerTopo* __restrict__ topos)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* ser = smem + (size_t)threadIdx.x * (kMaxSpecies * kSerStride);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < A.total; pos += stride) {
        const uint32_t slot = A.order ? A.order[pos] : pos;
        const uint32_t jx = job_of_slot(A, slot);
        const Job job = A.jobs[jx];
        const DimerTopo T = topos[jx];            // per-launch table, one entry per job
        const uint64_t traj = job.traj_base + (slot - job.out_offset);
        const uint32_t t_lo = (uint32_t)traj, t_hi = (uint32_t)(traj >> 32);
        float p[12];
        trajectory_params(A, t_lo, t_hi, slot, p);
        const float k_on = p[0], k_off1 = p[1], k_off2 = p[2], kp = p[7], gm = p[8], gp = p[9], k_dim = p[10], k_undim = p[11];
        const float km[4] = {p[3], p[4], p[5], p[6]};
        const int M = T.n_tf;

        float P[kMaxSpecies], R[kMaxSpecies], blk[kMaxSpecies], D[kMaxIxn], b[kMaxIxn];
        {
            const float e0 = expf(-A.init_mean);
            for (int blk4 = 0; blk4 < 2; ++blk4) {
                const Philox4 x = philox4x32_10(3 + blk4, 0u, t_lo, t_hi, A.seed_lo, A.seed_hi);
                for (int w = 0; w < 4; ++w) {
                    const int s = 4 * blk4 + w;
                    if (s < kMaxSpecies) {
                        const float u = u01(x.w[w]);
                        float pk = e0, F = e0;
                        int kk = 0;
                        while (u > F && kk < 1000) { ++kk; pk *= A.init_mean / (float)kk; F += pk; }
                        P[s] = (s < M) ? (float)kk : 0.0f;
                        R[s] = 0.0f;
                        blk[s] = 0.0f;
                    }
                }
            }
            for (int e = 0; e < kMaxIxn; ++e) { D[e] = 0.0f; b[e] = 0.0f; }
        }
        float t_rem = 0.0f;
        int k = 0, kd = 0, dcount = A.decim;
        uint32_t events = 0, ctr = 0;
        bool done = false;
        while (!done) {
            const Philox4 x = philox4x32_10_rk(ctr++, 1u, t_lo, t_hi, A.rk);
            for (int h = 0; h < 2 && !done; ++h) {
                // ---- propensities, flat running sum in SPEC order ----
                float cum[4 * kMaxSpecies + 4 * kMaxIxn];
                int nr = 0;
                float c = 0.0f;
                for (int s = 0; s < M; ++s) {
                    float n = 0.0f, na = 0.0f;
                    for (int e = 0; e < T.n_ixn; ++e)
                        if (T.ix_tgt[e] == s) { n += b[e]; if (T.ix_act[e]) na += b[e]; }
                    const int sel = (na > 0.0f ? 1 : 0) | (n > na ? 2 : 0);
                    c += (s < T.n_comp) ? km[sel] : km[0];  cum[nr++] = c;
                    c = fmaf(gm, R[s], c);                  cum[nr++] = c;
                    c = fmaf(kp, R[s], c);                  cum[nr++] = c;
                    c = fmaf(gp, P[s], c);                  cum[nr++] = c;
                    if (s < T.n_comp) {
                        const float koff = (n >= 2.0f) ? k_off2 : k_off1;
                        const float kfree = k_on * (2.0f - n);
                        for (int e = 0; e < T.n_ixn; ++e)
                            if (T.ix_tgt[e] == s) { c = fmaf(kfree, D[T.ix_dim[e]], c); cum[nr++] = c; }
                        for (int e = 0; e < T.n_ixn; ++e)
                            if (T.ix_tgt[e] == s) { c = fmaf(koff, b[e], c); cum[nr++] = c; }
                    }
                }
                for (int d = 0; d < T.n_dim; ++d) {
                    const int xa = T.dim_a[d], xb = T.dim_b[d];
                    const float pairs = (xa != xb) ? P[xa] * P[xb] : 0.5f * P[xa] * (P[xa] - 1.0f);
                    c = fmaf(k_dim, pairs, c);              cum[nr++] = c;
                    c = fmaf(k_undim, D[d], c);             cum[nr++] = c;
                }
                const float a0 = c;
                const float tau = (-0.69314718056f * fast_log2(u01(h ? x.w[2] : x.w[0]))) * fast_rcp(a0);
                t_rem -= tau;
                while (t_rem < 0.0f && k < A.nt) {
                    for (int s = 0; s < kMaxSpecies; ++s) blk[s] += fminf(P[s], kSampleCap);
                    ++k;
                    if (--dcount == 0) {
                        dcount = A.decim;
                        for (int s = 0; s < kMaxSpecies; ++s) {
                            if (s < M) ser[s * kSerStride + kd] = (uint8_t)compand_block(blk[s]);
                            blk[s] = 0.0f;
                        }
                        ++kd;
                    }
                    t_rem += A.dt;
                }
                if (k >= A.nt) { done = true; break; }
                // ---- select: first reaction whose running sum exceeds the target ----
                const float target = u01(h ? x.w[3] : x.w[1]) * a0;
                int sel = -1;
                for (int r = 0; r < nr; ++r)
                    if (target < cum[r]) { sel = r; break; }
                ++events;
                if (sel < 0) continue;          // rounding left none: the step fires nothing
                // ---- apply (same walk as the propensity pass) ----
                int r = 0;
                bool applied = false;
                for (int s = 0; s < M && !applied; ++s) {
                    if (sel < r + 4) {
                        const int w = sel - r;
                        if (w == 0) R[s] += 1.0f; else if (w == 1) R[s] -= 1.0f; else if (w == 2) P[s] += 1.0f; else P[s] -= 1.0f;
                        applied = true;
                        break;
                    }
                    r += 4;
                    if (s < T.n_comp) {
                        for (int e = 0; e < T.n_ixn && !applied; ++e)
                            if (T.ix_tgt[e] == s) { if (sel == r) { D[T.ix_dim[e]] -= 1.0f; b[e] += 1.0f; applied = true; } ++r; }
                        for (int e = 0; e < T.n_ixn && !applied; ++e)
                            if (T.ix_tgt[e] == s) { if (sel == r) { D[T.ix_dim[e]] += 1.0f; b[e] -= 1.0f; applied = true; } ++r; }
                    }
                }
                for (int d = 0; d < T.n_dim && !applied; ++d) {
                    const int xa = T.dim_a[d], xb = T.dim_b[d];
                    if (sel == r) { P[xa] -= 1.0f; P[xb] -= 1.0f; D[d] += 1.0f; applied = true; }
                    else if (sel == r + 1) { P[xa] += 1.0f; P[xb] += 1.0f; D[d] -= 1.0f; applied = true; }
                    r += 2;
                }
            }
        }
        float rw; int lg;
        lane_reward(ser, M, A.n_dec, rw, lg);
        if (A.series)
            for (int s = 0; s < M; ++s)
                for (int q = 0; q < A.n_dec; ++q) A.series[(size_t)slot * A.series_stride + s * A.n_dec + q] = ser[s * kSerStride + q];
        A.reward[slot] = rw;
        if (A.lag) A.lag[slot] = lg;
        if (A.events) A.events[slot] = events;
    }
}

}  // namespace ct
