// Warp-cooperative Gillespie SSA for the dimer model (SPEC.md section 9): G lanes share ONE
// trajectory, so a warp runs 32 / G trajectories and a step costs a few hundred cycles instead of
// the microseconds of the one-lane-per-trajectory kernels (ssa_dimers.cuh).  This is the kernel for
// launches that cannot fill the machine with one trajectory per lane, or whose length is set by
// their longest trajectory: MCTS batches over DimersGrammar states (BASELINE config 3), sequential
// get_reward calls (reference circuitree/circuitree.py:534).
//
// The reaction list of a dimer network is cut into BLOCKS of eight entries (SPEC.md section 9,
// "order of summation"): one block per TF -- transcription, mRNA decay, translation, monomer decay,
// binding of the promoter's interactions 0 and 1, their unbinding -- and one block per four dimers
// (dimerisation, dissociation of each).  Lane l of a sub-group owns block l:
//   G = 8: lanes 0..4 = TFs 0..4, lanes 5..7 = dimers 0..3, 4..7, 8..9      (any DimersGrammar state)
//   G = 4: lanes 0..2 = TFs 0..2, lane 3     = dimers 0..3                  (<= 3 TFs and <= 4 dimers)
// Per step every lane evaluates its block's running sum c_0..c_7 (eight FMAs, operands read from
// the trajectory's state vector in shared memory), the block totals are exchanged by shuffles and
// added up in block order (offsets o_l), every lane tests whether the target u * a0 falls into its
// block and which entry it hits there (tl = u a0 - o_l, first c_r > tl), and the ONE lane that was
// hit fires the reaction: four (slot, delta) updates of the state vector.  Time, grid crossing and
// decimation are computed redundantly by the sub-group's lanes; each TF lane accumulates and stores
// the record of its own TF.  One Philox block per lane serves the sub-group's next 2 G steps (the
// words are fetched by shuffle), so the generator costs 1 / G of what it costs a lane that owns a
// whole trajectory.
//
// The summation order above is the SPEC's, and ssa_dimers_kernel / ssa_dimers_mixed_kernel follow
// the same one, so all three kernels produce bit-identical results
// (tests/test_gpu_dimers.py::test_per_lane_programs_are_bit_identical).
#pragma once
#include <cuda_fp16.h>

#include "ssa_dimers.cuh"

namespace ct {

constexpr int kCoopThreads = 128;           // 4 warps per CTA
constexpr int kCoopStateStride = 72;        // floats per trajectory: 49 slots, padded so that sub-groups start 8 banks apart
constexpr int kCoopBlocksPerSm = 6;         // 24 warps / SM (shared memory: ~8 KB per warp at G = 8)

// One entry of a trajectory's stoichiometry table (shared memory, built when the trajectory starts):
// what firing entry r of block l does to the state vector -- four (byte offset, delta) fields on
// distinct slots; unused fields sit on the dummy slots with delta 0.
struct CoopStoich {                         // 16 bytes: one LDS.128
    uint32_t offs;                          // 4 x (slot * 4) >> 2 = slot index, one byte each
    __half2 d01, d23;                       // the four deltas (+-1, +-2, 0)
    uint32_t pad;
};

__host__ __device__ inline size_t coop_warp_bytes(int G, int ser_species)
{
    const int n_sub = 32 / G;
    return (size_t)n_sub * kCoopStateStride * sizeof(float) + kSerStride * sizeof(float) +
           (size_t)n_sub * G * 8 * sizeof(CoopStoich) +                                   // states | reward scratch | tables
           (size_t)n_sub * ser_species * kSerStride;                                      // | records
}

__device__ __forceinline__ CoopStoich coop_entry(int s0, float d0, int s1 = kSlotDummy + 1, float d1 = 0.0f,
                                                 int s2 = kSlotDummy + 2, float d2 = 0.0f, int s3 = kSlotDummy + 3, float d3 = 0.0f)
{
    CoopStoich e;
    e.offs = (uint32_t)s0 | ((uint32_t)s1 << 8) | ((uint32_t)s2 << 16) | ((uint32_t)s3 << 24);
    e.d01 = __floats2half2_rn(d0, d1);
    e.d23 = __floats2half2_rn(d2, d3);
    e.pad = 0u;
    return e;
}

template <int G>
__global__ void __launch_bounds__(kCoopThreads, kCoopBlocksPerSm)
ssa_dimers_coop_kernel(const LaunchArgs A, const DimerLaneProg* __restrict__ lprogs, const int ser_species)
{
    static_assert(G == 8 || G == 4, "sub-groups of 8 or 4 lanes");
    constexpr int kTf = (G == 8) ? 5 : 3;   // TF blocks (lanes 0..kTf-1); the other lanes own dimer blocks
    constexpr int kSub = 32 / G;
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / G, l = lane % G, base = sub * G;
    const size_t rec_bytes = (size_t)ser_species * kSerStride;
    uint8_t* wbase = smem + (size_t)warp * coop_warp_bytes(G, ser_species);
    float* Xs = reinterpret_cast<float*>(wbase) + sub * kCoopStateStride;          // this trajectory's state vector
    float* z = reinterpret_cast<float*>(wbase) + kSub * kCoopStateStride;
    CoopStoich* tab = reinterpret_cast<CoopStoich*>(z + kSerStride) + (size_t)(sub * G + l) * 8;   // my block's 8 entries
    uint8_t* warp_rec = reinterpret_cast<uint8_t*>(reinterpret_cast<CoopStoich*>(z + kSerStride) + kSub * G * 8);
    uint8_t* rec = warp_rec + (size_t)sub * rec_bytes + (size_t)(l < kTf ? l : 0) * kSerStride;   // my TF's row of the record
    const bool tf_role = l < kTf;
    const int q = l - kTf;                                                          // dimer block of a dimer lane
    float lf;       // my position in the sub-group as a float the compiler cannot trace back to the integer
    asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(lf) : "r"(l));

    // The lane's program, fixed while a trajectory runs: operand addresses and rates of its block.
    //   TF lane:    op = {N, A, R, P, D(slot 0), D(slot 1), B0, B1};
    //               rt = {km_unbound, km_act, km_rep, km_act_rep, gm, kp, gp, k_on, 2 k_on, k_off1, k_off2, present0, present1}
    //   dimer lane: op = {Pa[4], Pb[4], D[4]};  rt = {k_dim or 0 [4], k_undim or 0 [4]};  same4 = homodimer bits
    // A parked sub-group has every rate zero (nothing fires) and a zero state.
    float* op[12];
    float rt[13];
#pragma unroll
    for (int i = 0; i < 12; ++i) op[i] = Xs + kSlotDummy;
#pragma unroll
    for (int i = 0; i < 13; ++i) rt[i] = 0.0f;
    uint32_t same4 = 0;
    float t_rem = 0.0f, blk = 0.0f;
    int k = 0, kd = 0, dcount = 1, n_tf = 0;
    uint32_t events = 0, ctr = 0, traj_lo = 0, traj_hi = 0, out = 0;
    bool active = false, finished = false, out_of_work = false;
    uint32_t n_iters = 0;

    for (int s = l; s < kCoopStateStride; s += G) Xs[s] = 0.0f;
    __syncwarp();

    for (;;) {
        const unsigned idle = __ballot_sync(kFull, !active);
        if (idle) {
            // 1. reduce finished trajectories to (reward, lag), one at a time, whole warp helping
            unsigned fin = __ballot_sync(kFull, finished && l == 0);
            while (fin) {
                const int src = __ffs(fin) - 1;
                fin &= fin - 1;
                float rw; int lg;
                const int src_ntf = __shfl_sync(kFull, n_tf, src);
                const uint8_t* srec = warp_rec + (size_t)(src / G) * rec_bytes;
                __syncwarp();
                warp_reward(srec, z, src_ntf, A.n_dec, lane, rw, lg);
                const uint32_t o = __shfl_sync(kFull, out, src);
                if (A.series) {
                    for (int s = 0; s < src_ntf; ++s)
                        for (int x = lane; x < A.n_dec; x += 32)
                            A.series[(size_t)o * A.series_stride + s * A.n_dec + x] = srec[s * kSerStride + x];
                }
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
                    const uint32_t jx = job_of_slot(A, slot);
                    const Job mine = A.jobs[jx];
                    const uint32_t dmask_ntf = lprogs[jx].dmask_ntf;
                    n_tf = (int)(dmask_ntf >> 16) & 15;
                    const uint64_t traj = mine.traj_base + (slot - mine.out_offset);
                    traj_lo = (uint32_t)traj; traj_hi = (uint32_t)(traj >> 32);
                    out = slot;
                    float prm[12];
                    trajectory_params(A, traj_lo, traj_hi, slot, prm);
                    for (int s = l; s < kCoopStateStride; s += G) Xs[s] = 0.0f;
                    if (tf_role) {
                        const bool has_tf = l < n_tf;
                        const uint32_t iw = lprogs[jx].ixw[l];
                        const uint32_t f0 = iw, f1 = iw >> 6;
                        const int d0 = (int)((f0 >> 2) & 15u), d1 = (int)((f1 >> 2) & 15u);
                        op[0] = Xs + kSlotN + l; op[1] = Xs + kSlotA + l; op[2] = Xs + kSlotR + l; op[3] = Xs + kSlotP + l;
                        op[4] = Xs + kSlotD + d0; op[5] = Xs + kSlotD + d1;
                        op[6] = Xs + kSlotB + kMixK * l; op[7] = Xs + kSlotB + kMixK * l + 1;
                        // (a TF this topology does not have: no transcription, and nothing else can fire from a zero state)
                        rt[0] = has_tf ? prm[3] : 0.0f; rt[1] = has_tf ? prm[4] : 0.0f;
                        rt[2] = has_tf ? prm[5] : 0.0f; rt[3] = has_tf ? prm[6] : 0.0f;
                        rt[4] = prm[8]; rt[5] = prm[7]; rt[6] = prm[9];
                        rt[7] = prm[0]; rt[8] = 2.0f * prm[0]; rt[9] = prm[1]; rt[10] = prm[2];
                        rt[11] = (f0 & 1u) ? 1.0f : 0.0f; rt[12] = (f1 & 1u) ? 1.0f : 0.0f;
                        // stoichiometry of my block: transcription, mRNA decay, translation, monomer decay, ...
                        tab[0] = coop_entry(kSlotR + l, 1.0f);
                        tab[1] = coop_entry(kSlotR + l, -1.0f);
                        tab[2] = coop_entry(kSlotP + l, 1.0f);
                        tab[3] = coop_entry(kSlotP + l, -1.0f);
                        // ... binding of interaction slots 0, 1 (dimer -1, bound +1, promoter total +1, activating total +1), unbinding
#pragma unroll
                        for (int qq = 0; qq < kMixK; ++qq) {
                            const uint32_t f = iw >> (6 * qq);
                            const int dd = kSlotD + (int)((f >> 2) & 15u);
                            const int sa = (f & 2u) ? kSlotA + l : kSlotDummy + 3;
                            const float da = (f & 2u) ? 1.0f : 0.0f;
                            tab[4 + qq] = coop_entry(dd, -1.0f, kSlotB + kMixK * l + qq, 1.0f, kSlotN + l, 1.0f, sa, da);
                            tab[6 + qq] = coop_entry(dd, 1.0f, kSlotB + kMixK * l + qq, -1.0f, kSlotN + l, -1.0f, sa, -da);
                        }
                        if (has_tf) {   // SPEC.md section 5: TF l takes word l of the setup stream's blocks 3, 4 (slot l was zeroed by this lane)
                            const Philox4 x = philox4x32_10(3u + (uint32_t)(l >> 2), 0u, traj_lo, traj_hi, A.seed_lo, A.seed_hi);
                            const int w = l & 3;
                            const uint32_t word = w == 0 ? x.w[0] : (w == 1 ? x.w[1] : (w == 2 ? x.w[2] : x.w[3]));
                            const float u = u01(word);
                            const float e0 = expf(-A.init_mean);
                            float pk = e0, F = e0;
                            int kk = 0;
                            while (u > F && kk < 1000) { ++kk; pk *= A.init_mean / (float)kk; F += pk; }
                            Xs[kSlotP + l] = (float)kk;
                        }
                    } else {
                        const uint32_t pa4 = (lprogs[jx].dim_a >> (12 * q)) & 0xfffu, pb4 = (lprogs[jx].dim_b >> (12 * q)) & 0xfffu;
                        const uint32_t m4 = ((dmask_ntf & 0x3ffu) >> (4 * q)) & 0xfu;
                        same4 = 0;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            // (an absent dimer has both rates zero; the last block's slots for dimers 10, 11 do not exist)
                            const bool present = (m4 >> i) & 1u;
                            const int da = (int)((pa4 >> (3 * i)) & 7u), db = (int)((pb4 >> (3 * i)) & 7u);
                            const int dslot = present ? kSlotD + 4 * q + i : kSlotDummy;
                            op[i] = Xs + kSlotP + da; op[4 + i] = Xs + kSlotP + db; op[8 + i] = Xs + dslot;
                            rt[i] = present ? prm[10] : 0.0f; rt[4 + i] = present ? prm[11] : 0.0f;
                            same4 |= (da == db ? 1u : 0u) << i;
                            // dimerisation: monomers -1 each (a homodimer: -2 on its one monomer), dimer +1; dissociation: the reverse
                            if (da != db) {
                                tab[2 * i] = coop_entry(kSlotP + da, -1.0f, kSlotP + db, -1.0f, dslot, 1.0f);
                                tab[2 * i + 1] = coop_entry(kSlotP + da, 1.0f, kSlotP + db, 1.0f, dslot, -1.0f);
                            } else {
                                tab[2 * i] = coop_entry(kSlotP + da, -2.0f, kSlotDummy + 1, 0.0f, dslot, 1.0f);
                                tab[2 * i + 1] = coop_entry(kSlotP + da, 2.0f, kSlotDummy + 1, 0.0f, dslot, -1.0f);
                            }
                        }
                    }
                    t_rem = 0.0f; blk = 0.0f;
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
        for (int h = 0; h < 2; ++h) {       // (two steps per block, unrolled: no word selects in the loop)
            const uint32_t w_tau = __shfl_sync(kFull, h ? x.w[2] : x.w[0], pr, G);
            const uint32_t w_sel = __shfl_sync(kFull, h ? x.w[3] : x.w[1], pr, G);

            // ---- my block's running sum (starts at zero: SPEC.md section 9) ----
            float c[8];
            float pcur = 0.0f;
            if (tf_role) {
                const float n = *op[0], na = *op[1], r = *op[2];
                pcur = *op[3];
                const float dv0 = *op[4], dv1 = *op[5], b0 = *op[6], b1 = *op[7];
                const float km_a = (na > 0.0f) ? rt[1] : rt[0];
                const float km_i = (na > 0.0f) ? rt[3] : rt[2];
                c[0] = (n > na) ? km_i : km_a;
                c[1] = fmaf(rt[4], r, c[0]);
                c[2] = fmaf(rt[5], r, c[1]);
                c[3] = fmaf(rt[6], pcur, c[2]);
                const float kfree = fmaf(-rt[7], n, rt[8]);            // k_on (2 - n)
                const float koff = (n >= 2.0f) ? rt[10] : rt[9];
                c[4] = fmaf(kfree * rt[11], dv0, c[3]);                 // (an absent slot: rate * 0, exact)
                c[5] = fmaf(kfree * rt[12], dv1, c[4]);
                c[6] = fmaf(koff, b0, c[5]);
                c[7] = fmaf(koff, b1, c[6]);
            } else {
                float cc = 0.0f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float pa = *op[i], pb = *op[4 + i], dv = *op[8 + i];
                    const float pairs = ((same4 >> i) & 1u) ? 0.5f * pa * (pa - 1.0f) : pa * pb;
                    cc = fmaf(rt[i], pairs, cc);            c[2 * i] = cc;
                    cc = fmaf(rt[4 + i], dv, cc);           c[2 * i + 1] = cc;
                }
            }
            // ---- block offsets: totals added up in block order; a0 is the last offset ----
            // (my own offset is the same chain cut off at my block: the mask is 1 for blocks before mine, else 0,
            // and comes from the FMA pipe -- compares and selects run at half its rate and are the scarce resource here)
            float o = 0.0f, my_lo = 0.0f;
#pragma unroll
            for (int b = 0; b < G; ++b) {
                const float gb = __shfl_sync(kFull, c[7], b, G);
                my_lo = fmaf(__saturatef(lf - (float)b), gb, my_lo);
                o += gb;
            }
            const float a0 = o;
            const float my_hi = my_lo + c[7];

            // ---- time to next event; grid samples crossed keep the pre-event state ----
            const float tau = (-0.69314718056f * fast_log2(u01(w_tau))) * fast_rcp(a0);   // a0 == 0 -> +inf
            t_rem -= tau;
            bool alive = true;
            const bool was_active = active;
            if (active && t_rem < 0.0f) {
                do {
                    blk += fminf(pcur, kSampleCap);
                    k += 1;
                    if (--dcount == 0) {
                        dcount = A.decim;
                        CT_CHECK(kd >= 0 && kd < kSerStride && n_tf <= ser_species);
                        if (l < n_tf) rec[kd] = (uint8_t)compand_block(blk);
                        blk = 0.0f;
                        kd += 1;
                    }
                    t_rem += A.dt;
                } while (t_rem < 0.0f && k < A.nt);
                if (k >= A.nt) { alive = false; active = false; finished = true; }
            }

            // ---- the lane whose block holds the target fires its entry ----
            const float t = alive ? u01(w_sel) * a0 : 3.0e38f;
            const float tl = t - my_lo;
            // first entry with tl < c_r = number of entries with c_i <= tl; [c_i > tl] = saturate((c_i - tl) 2^80) is
            // exact for sums that are zero or above 2^-56 (the scaling is a power of two, the sign that of the exact difference)
            constexpr float kBig = 1208925819614629174706176.0f;      // 2^80
            const float tb = -tl * kBig;
            float above = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) above += __saturatef(fmaf(c[i], kBig, tb));
            const int r = 8 - (int)above;
            if (my_lo <= t && t < my_hi && r < 8) {
                const uint4 e = *reinterpret_cast<const uint4*>(tab + r);
                float* p0 = Xs + __byte_perm(e.x, 0u, 0x4440);
                float* p1 = Xs + __byte_perm(e.x, 0u, 0x4441);
                float* p2 = Xs + __byte_perm(e.x, 0u, 0x4442);
                float* p3 = Xs + __byte_perm(e.x, 0u, 0x4443);
                CT_CHECK(r >= 0 && r < 8 && (e.x & 0xffu) < kSlots && ((e.x >> 8) & 0xffu) < kSlots &&
                         ((e.x >> 16) & 0xffu) < kSlots && (e.x >> 24) < kSlots);
                const float2 d01 = __half22float2(*reinterpret_cast<const __half2*>(&e.y));
                const float2 d23 = __half22float2(*reinterpret_cast<const __half2*>(&e.z));
                const float v0 = *p0, v1 = *p1, v2 = *p2, v3 = *p3;     // distinct slots
                *p0 = v0 + d01.x;
                *p1 = v1 + d01.y;
                *p2 = v2 + d23.x;
                *p3 = v3 + d23.y;
            }
            events += (was_active && alive) ? 1u : 0u;
            if (!alive) {   // park: zero rates and state so that the sub-group idles through the lock-step loop
#pragma unroll
                for (int i = 0; i < 13; ++i) rt[i] = 0.0f;
                for (int s = l; s < kCoopStateStride; s += G) Xs[s] = 0.0f;
                t_rem = 0.0f; blk = 0.0f;
            }
            __syncwarp();       // the fired updates are visible to the whole sub-group before the next step reads
        }
        }
    }
    if (A.warp_iters && lane == 0) atomicAdd(A.warp_iters, (unsigned long long)n_iters);
}

}  // namespace ct
