// Warp-cooperative Gillespie SSA for the pairwise model (SPEC.md sections 2-8): one trajectory per
// sub-group of G lanes, one GENE per lane.  The kernel for launches that cannot give every resident
// lane a trajectory of its own -- a single get_reward call (reference circuitree/circuitree.py:534),
// MCTS batches of a few thousand leaves, the small end of the throughput sweep: such launches last as
// long as their longest trajectory, and a step costs a lone warp ~9 cycles per instruction, so what
// shortens them is fewer instructions on the trajectory's serial path (~110 here against ~250 in the
// general one-lane loop of ssa_pairwise.cuh).
//
// Lane j of a sub-group owns gene j: its mRNA, free protein, promoter occupancy and the bound counts of
// its regulators live in that lane's registers.  Per step the lanes exchange their protein counts
// (M shuffles), every lane runs its gene's six-entry running sum, the gene totals are exchanged (M
// shuffles) and added in gene order, every lane derives the target's position relative to its own
// gene by the same sequence of subtractions the one-lane kernel performs, fires what falls into its
// gene and, for a (un)binding, resolves the regulator exactly like the one-lane kernel's resolution
// block; the one protein-count change that belongs to another lane travels by two shuffles.
// Time, grid crossing and decimation are computed redundantly; every lane records its own gene.
//
// The arithmetic is that of ssa_pairwise.cuh operation for operation, so results are bit-identical
// (tests/test_gpu_ssa.py::test_schedules_give_identical_results).
#pragma once
#include "ssa_pairwise.cuh"

namespace ct {

constexpr int kPairCoopThreads = 128;
constexpr int kPairCoopBlocksPerSm = 6;       // 80 registers: no spills on the serial path

__host__ __device__ inline size_t pair_coop_warp_bytes(int G, int n_rows)
{
    return kSerStride * sizeof(float) + (size_t)(32 / G) * n_rows * kSerStride;     // reward scratch | records
}

template <int G, int MM>      // (4, 4): up to four components; (8, 5): five
__global__ void __launch_bounds__(kPairCoopThreads, kPairCoopBlocksPerSm)
ssa_pairwise_coop_kernel(const LaunchArgs A, const int n_rows)
{
    static_assert((G == 4 && MM == 4) || (G == 8 && MM == 5), "sub-groups of 4 lanes (<= 4 genes) or 8 lanes (5 genes)");
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / G, l = lane % G, base = sub * G;
    const size_t rec_bytes = (size_t)n_rows * kSerStride;
    uint8_t* wbase = smem + (size_t)warp * pair_coop_warp_bytes(G, n_rows);
    float* z = reinterpret_cast<float*>(wbase);
    uint8_t* warp_rec = reinterpret_cast<uint8_t*>(z + kSerStride);
    uint8_t* rec = warp_rec + (size_t)sub * rec_bytes + (size_t)(l < n_rows ? l : 0) * kSerStride;    // my gene's row
    float lf;       // my position in the sub-group as a float the compiler cannot trace back to the integer
    asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(lf) : "r"(l));

    // my gene (a lane without one -- l >= species count, or a parked sub-group -- has zero rates and state)
    float P = 0.f, R = 0.f, n = 0.f, Ab = 0.f, blk = 0.f;
    float b[MM], hm[MM], am[MM];          // bound molecules of regulator i at my promoter; edge i -> me exists / activates
#pragma unroll
    for (int i = 0; i < MM; ++i) { b[i] = 0.f; hm[i] = 0.f; am[i] = 0.f; }
    float k_on = 0.f, k_on2 = 0.f, k_off1 = 0.f, k_off2 = 0.f, km0 = 0.f, km1 = 0.f, km2 = 0.f, km3 = 0.f, kp = 0.f, gm = 0.f, gp = 0.f;
    float t_rem = 0.0f;
    int k = 0, kd = 0, dcount = 1, nsp = 0;
    uint32_t events = 0, ctr = 0, traj_lo = 0, traj_hi = 0, out = 0;
    bool active = false, finished = false, out_of_work = false;
    uint32_t n_iters = 0;

    for (;;) {
        const unsigned idle = __ballot_sync(kFull, !active);
        if (idle) {
            // 1. reduce finished trajectories to (reward, lag), one at a time, whole warp helping
            unsigned fin = __ballot_sync(kFull, finished && l == 0);
            if (fin) __syncwarp();
            while (fin) {
                const int src = __ffs(fin) - 1;
                fin &= fin - 1;
                float rw; int lg;
                const int src_nsp = __shfl_sync(kFull, nsp, src);
                const uint8_t* srec = warp_rec + (size_t)(src / G) * rec_bytes;
                warp_reward(srec, z, src_nsp, A.n_dec, lane, rw, lg);
                const uint32_t o = __shfl_sync(kFull, out, src);
                CT_CHECK(o < A.total && src_nsp >= 1 && src_nsp <= n_rows);
                if (A.series) {
                    for (int s = 0; s < src_nsp; ++s)
                        for (int x = lane; x < A.n_dec; x += 32)
                            A.series[(size_t)o * A.series_stride + s * A.n_dec + x] = srec[s * kSerStride + x];
                }
                __syncwarp();
                if (lane == src) {
                    A.reward[o] = rw;
                    if (A.lag) A.lag[o] = lg;
                    if (A.events) A.events[o] = events;
                }
                if (sub == src / G) finished = false;
            }
            // 2. every idle sub-group draws one launch position (longest expected first)
            const bool want = !active && !out_of_work;
            uint32_t pos = 0xffffffffu;
            if (want && l == 0) pos = atomicAdd(A.tile_counter, 1u);
            pos = __shfl_sync(kFull, pos, base);
            if (want) {
                if (pos >= A.total) {
                    out_of_work = true;
                } else {
                    const uint32_t slot = A.order ? A.order[pos] : pos;
                    const Job mine = A.jobs[job_of_slot(A, slot)];
                    nsp = (int)topo_nspecies(mine.topo);
                    const bool has_gene = l < nsp;
                    const uint64_t traj = mine.traj_base + (slot - mine.out_offset);
                    traj_lo = (uint32_t)traj; traj_hi = (uint32_t)(traj >> 32);
                    out = slot;
                    float prm[12];
                    trajectory_params(A, traj_lo, traj_hi, slot, prm);
                    k_on = prm[0]; k_on2 = 2.0f * prm[0]; k_off1 = prm[1]; k_off2 = prm[2];
                    km0 = has_gene ? prm[3] : 0.0f; km1 = has_gene ? prm[4] : 0.0f;
                    km2 = has_gene ? prm[5] : 0.0f; km3 = has_gene ? prm[6] : 0.0f;
                    kp = prm[7]; gm = prm[8]; gp = prm[9];
#pragma unroll
                    for (int i = 0; i < MM; ++i) {
                        const uint32_t e = (has_gene && l < kMaxSpecies) ? topo_edge(mine.topo, i, l) : 0u;
                        hm[i] = e ? 1.0f : 0.0f;
                        am[i] = (e == 1u) ? 1.0f : 0.0f;
                        b[i] = 0.0f;
                    }
                    P = 0.0f;
                    if (has_gene) {     // SPEC.md section 5: gene j takes word j of the setup stream's blocks 3, 4
                        const Philox4 x = philox4x32_10(3u + (uint32_t)(l >> 2), 0u, traj_lo, traj_hi, A.seed_lo, A.seed_hi);
                        const int w = l & 3;
                        const uint32_t word = w == 0 ? x.w[0] : (w == 1 ? x.w[1] : (w == 2 ? x.w[2] : x.w[3]));
                        const float u = u01(word);
                        const float e0 = expf(-A.init_mean);
                        float pk = e0, F = e0;
                        int kk = 0;
                        while (u > F && kk < 1000) { ++kk; pk *= A.init_mean / (float)kk; F += pk; }
                        P = (float)kk;
                    }
                    R = 0.0f; n = 0.0f; Ab = 0.0f; blk = 0.0f;
                    t_rem = 0.0f;
                    k = 0; kd = 0; dcount = A.decim;
                    events = 0; ctr = 0;
                    active = true;
                }
            }
            __syncwarp();
            if (__ballot_sync(kFull, active) == 0u) break;      // an idle sub-group has no work left by now
        }
        ++n_iters;

        // one Philox block per lane: block (ctr + l) serves the sub-group's steps 2 (ctr + l), 2 (ctr + l) + 1
        const Philox4 x = philox4x32_10_rk(ctr + (uint32_t)l, 1u, traj_lo, traj_hi, A.rk);
        ctr += G;
#pragma unroll 1
        for (int pr = 0; pr < G; ++pr) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t w_tau = __shfl_sync(kFull, h ? x.w[2] : x.w[0], pr, G);
            const uint32_t w_sel = __shfl_sync(kFull, h ? x.w[3] : x.w[1], pr, G);

            // ---- my gene's running sum (ssa_pairwise.cuh, general path, operation for operation) ----
            float Pi[MM];
#pragma unroll
            for (int i = 0; i < MM; ++i) Pi[i] = __shfl_sync(kFull, P, i, G);
            const float km_a = (Ab > 0.0f) ? km1 : km0;
            const float km_i = (Ab > 0.0f) ? km3 : km2;
            const float lc0 = (n > Ab) ? km_i : km_a;
            const float lc1 = fmaf(gm, R, lc0);
            const float lc2 = fmaf(kp, R, lc1);
            const float lc3 = fmaf(gp, P, lc2);
            float sp = 0.0f;
#pragma unroll
            for (int i = 0; i < MM; ++i) sp = fmaf(hm[i], Pi[i], sp);
            const float koff = (n >= 2.0f) ? k_off2 : k_off1;
            const float kfree = fmaf(-k_on, n, k_on2);         // k_on (2 - n), exact for n = 0, 1, 2
            const float lc4 = fmaf(kfree, sp, lc3);
            const float lc5 = fmaf(koff, n, lc4);

            // ---- gene totals added in gene order; the target's position relative to my gene: t_0 = u a0, t_{j+1} = t_j - g_j ----
            float gb[MM];
#pragma unroll
            for (int j = 0; j < MM; ++j) gb[j] = __shfl_sync(kFull, lc5, j, G);
            float a0 = gb[0];
#pragma unroll
            for (int j = 1; j < MM; ++j) a0 += gb[j];

            // ---- time to next event; grid samples crossed keep the pre-event state ----
            const float tau = (-0.69314718056f * fast_log2(u01(w_tau))) * fast_rcp(a0);   // a0 == 0 -> +inf
            t_rem -= tau;
            bool alive = true;
            const bool was_active = active;
            if (active && t_rem < 0.0f) {
                do {
                    blk += fminf(P, kSampleCap);
                    k += 1;
                    if (--dcount == 0) {
                        dcount = A.decim;
                        CT_CHECK(kd >= 0 && kd < kSerStride);
                        if (l < nsp) rec[kd] = (uint8_t)compand_block(blk);
                        blk = 0.0f;
                        kd += 1;
                    }
                    t_rem += A.dt;
                } while (t_rem < 0.0f && k < A.nt);
                if (k >= A.nt) { alive = false; active = false; finished = true; }
            }

            // ---- select and fire inside my gene ----
            constexpr float kBig = 1152921504606846976.0f;   // 2^60 (as in ssa_pairwise.cuh)
            float t = alive ? u01(w_sel) * a0 : 3.0e38f;
#pragma unroll
            for (int j = 0; j < MM; ++j) t = fmaf(-__saturatef(lf - (float)j), gb[j], t);    // t -= g_j for the genes before mine
            const float tb = -t * kBig;
            const float s_prev = __saturatef(tb);               // [t < 0]: the reaction fired in a gene before mine
            const float s_tx = __saturatef(fmaf(lc0, kBig, tb));
            const float s_rm = __saturatef(fmaf(lc1, kBig, tb));
            const float s_tl = __saturatef(fmaf(lc2, kBig, tb));
            const float s_dp = __saturatef(fmaf(lc3, kBig, tb));
            const float s_b = __saturatef(fmaf(lc4, kBig, tb));
            const float s_u = __saturatef(fmaf(lc5, kBig, tb));
            R += fmaf(2.0f, s_tx, -s_prev) - s_rm;
            P += fmaf(2.0f, s_tl, -s_rm) - s_dp;
            const float d_b = s_b - s_dp;                        // 1 if some regulator binds at my promoter
            const float d = fmaf(2.0f, s_b, -s_dp) - s_u;        // +1 bind, -1 unbind, 0 neither
            if (__any_sync(kFull, d != 0.0f)) {
                // which regulator: the position of the target inside the lump (ssa_pairwise.cuh resolution block)
                const bool bind = d_b != 0.0f;
                const float basev = bind ? lc3 : lc4;
                const float scale = bind ? kfree : koff;
                const float tot = bind ? sp : n;
                float y = (t - basev) * fast_rcp(scale);
                y = fminf(y, __uint_as_float(__float_as_uint(tot) - 1u));
                if (d == 0.0f) y = 3.0e38f;
                float acc = 0.0f, sq = 0.0f, upd_i = 0.0f, upd_d = 0.0f;
#pragma unroll
                for (int i = 0; i < MM; ++i) {
                    // (an absent edge adds an exact 0 to acc: s == sq, dd == 0)
                    acc = fmaf(hm[i], bind ? Pi[i] : b[i], acc);
                    const float s = (y < acc) ? 1.0f : 0.0f;
                    const float dd = (s - sq) * d;               // +-1 for the chosen regulator
                    sq = s;
                    b[i] += dd;
                    Ab = fmaf(am[i], dd, Ab);
                    upd_d += dd;
                    upd_i = fmaf((float)i, fabsf(dd), upd_i);
                }
                n += d;
                // the chosen regulator's free protein count lives in its own lane
                const unsigned fired = (__ballot_sync(kFull, upd_d != 0.0f) >> base) & ((1u << G) - 1u);
                const int srcl = fired ? (__ffs(fired) - 1) : 0;
                const float ui = __shfl_sync(kFull, upd_i, srcl, G);
                const float ud = __shfl_sync(kFull, upd_d, srcl, G);
                if (ui == lf) P -= ud;
            }
            events += (was_active && alive) ? 1u : 0u;
            if (!alive) {   // park: zero rates and state so that the sub-group idles through the lock-step loop
                k_on = k_on2 = k_off1 = k_off2 = km0 = km1 = km2 = km3 = kp = gm = gp = 0.0f;
                P = R = n = Ab = blk = 0.0f;
#pragma unroll
                for (int i = 0; i < MM; ++i) { b[i] = 0.0f; hm[i] = 0.0f; am[i] = 0.0f; }
                t_rem = 0.0f;
            }
        }
        }
    }
    if (A.warp_iters && lane == 0) atomicAdd(A.warp_iters, (unsigned long long)n_iters);
}

}  // namespace ct
