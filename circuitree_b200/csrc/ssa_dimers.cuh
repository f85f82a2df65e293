// Batched Gillespie SSA + oscillation reward for the dimer model (SPEC.md section 9), the
// simulator behind get_reward for DimersGrammar states (reference circuitree/models.py:844-884 is
// the topology input).
//
// Same execution model as the pairwise kernel (ssa_pairwise.cuh): one trajectory per lane, one
// topology per warp at a time, all 32 lanes step in lock-step, persistent warps pull 32-trajectory
// tiles off a global counter and refill lanes as trajectories end, the finished lane's record is
// reduced to (reward, lag) by the whole warp.  What differs is where the state lives.  A dimer
// network has up to 5 TFs, 10 interactions and 10 dimer species whose wiring is data, not code, so
// the per-lane state vector and the running sum of propensities sit in shared memory as
// [slot][lane] float arrays (bank = lane: conflict-free for any slot, uniform or per-lane), and
// the warp walks the job's reaction program (a few packed words held in registers):
//   pass 1  running sum in SPEC order: per TF (transcription, mRNA decay, translation, monomer
//           decay, binding of each interaction at its promoter, unbinding of each), then per dimer
//           (dimerisation, dissociation); every partial sum is stored;
//   pass 2  each lane binary-searches its own target in the stored sums (same trip count for all
//           lanes) and applies the stoichiometry word of the reaction it found: four
//           (slot, +-1 or +-2) fields on distinct slots, loaded together and stored together.
// Nothing in the hot loop diverges except the rare grid-crossing block.  The kernel is bound by
// shared-memory latency, so residency matters more than anything else: the per-lane record (the
// companded block sums, 128 bytes per TF) would take 60 % of a warp's shared memory and is kept in a
// per-warp scratch area in global memory instead (one byte per TF every 16 grid points: ~500 bytes
// per ~5e5 events, L2-resident); that raises residency from 8 to ~20 warps per SM.
#pragma once
#include <string.h>

#include <algorithm>

#include "ssa_pairwise.cuh"

namespace ct {

constexpr int kMaxIxn = 10;                 // interactions, and distinct dimers, per topology
constexpr int kDimThreads = 64;             // 2 warps per CTA (shared memory per warp is what limits residency)
constexpr int kMaxEntries = 4 * kMaxSpecies + 4 * kMaxIxn;   // reactions in the running sum

// canonical description of a dimer topology (the key of the process-wide handle table)
struct DimerTopo {                  // 56 bytes
    uint8_t n_tf, n_comp, n_ixn, n_dim;
    uint8_t dim_a[kMaxIxn], dim_b[kMaxIxn];
    uint8_t ix_dim[kMaxIxn], ix_tgt[kMaxIxn], ix_act[kMaxIxn];
    uint8_t pad[2];
};

// state slots of one lane (floats holding exact small integers); four dummies so that the four
// fields of a stoichiometry word never share a slot
enum : int { kSlotP = 0, kSlotR = 5, kSlotD = 10, kSlotB = 20, kSlotN = 30, kSlotA = 35, kSlotBlk = 40, kSlotDummy = 45, kSlots = 49 };

// The job's reaction program, built on the host from a DimerTopo (dimer_program below).
struct DimerProg {                  // 272 bytes, one per job
    uint64_t pk_ix_dim;             // dimer of sorted interaction x at bits 4x (interactions sorted by target, stable)
    uint32_t pk_ix_start;           // TF s owns sorted interactions [start(s), start(s+1)), start(s) at bits 4s (s = 0..5)
    uint32_t pk_dim_a, pk_dim_b;    // monomers of dimer d at bits 3d
    uint8_t n_tf, n_ixn, n_dim, n_entries;
    uint8_t search_iters;           // ceil(log2(longest block + 1)): trip count of the search inside a block
    uint8_t pad[3];
    // reaction -> 4 x (slot | code << 6) on distinct slots; code 0 = +1, 1 = -1, 2 = +2, 3 = -2
    // (a dummy slot absorbs unused fields); entry n_entries = "no reaction"
    uint32_t stoich[kMaxEntries + 1];
};
static_assert(sizeof(DimerProg) % 8 == 0, "DimerProg is copied word by word and holds a 64-bit member");

inline DimerProg dimer_program(const DimerTopo& T)
{
    DimerProg G;
    memset(&G, 0, sizeof G);
    G.n_tf = T.n_tf; G.n_ixn = T.n_ixn; G.n_dim = T.n_dim;
    int sorted[kMaxIxn], start[kMaxSpecies + 1], nx = 0;
    for (int s = 0; s < T.n_tf; ++s) {
        start[s] = nx;
        for (int e = 0; e < T.n_ixn; ++e)
            if (T.ix_tgt[e] == s) sorted[nx++] = e;
    }
    for (int s = T.n_tf; s <= kMaxSpecies; ++s) start[s] = nx;
    for (int s = 0; s <= kMaxSpecies; ++s) G.pk_ix_start |= (uint32_t)start[s] << (4 * s);
    for (int x = 0; x < nx; ++x) G.pk_ix_dim |= (uint64_t)T.ix_dim[sorted[x]] << (4 * x);
    for (int d = 0; d < T.n_dim; ++d) {
        G.pk_dim_a |= (uint32_t)T.dim_a[d] << (3 * d);
        G.pk_dim_b |= (uint32_t)T.dim_b[d] << (3 * d);
    }
    auto word = [&](uint32_t f0, uint32_t f1 = kSlotDummy + 1, uint32_t f2 = kSlotDummy + 2, uint32_t f3 = kSlotDummy + 3) {
        return f0 | (f1 << 8) | (f2 << 16) | (f3 << 24);
    };
    auto inc = [](int slot) { return (uint32_t)slot; };
    auto dec = [](int slot) { return (uint32_t)slot | (1u << 6); };
    auto inc2 = [](int slot) { return (uint32_t)slot | (2u << 6); };
    auto dec2 = [](int slot) { return (uint32_t)slot | (3u << 6); };
    int r = 0;
    for (int s = 0; s < T.n_tf; ++s) {
        G.stoich[r++] = word(inc(kSlotR + s));
        G.stoich[r++] = word(dec(kSlotR + s));
        G.stoich[r++] = word(inc(kSlotP + s));
        G.stoich[r++] = word(dec(kSlotP + s));
        for (int x = start[s]; x < start[s + 1]; ++x) {
            const int e = sorted[x];
            G.stoich[r++] = word(dec(kSlotD + T.ix_dim[e]), inc(kSlotB + x), inc(kSlotN + s),
                                 T.ix_act[e] ? inc(kSlotA + s) : (uint32_t)(kSlotDummy + 3));
        }
        for (int x = start[s]; x < start[s + 1]; ++x) {
            const int e = sorted[x];
            G.stoich[r++] = word(inc(kSlotD + T.ix_dim[e]), dec(kSlotB + x), dec(kSlotN + s),
                                 T.ix_act[e] ? dec(kSlotA + s) : (uint32_t)(kSlotDummy + 3));
        }
    }
    for (int d = 0; d < T.n_dim; ++d) {
        const int a = T.dim_a[d], b = T.dim_b[d];
        if (a != b) {
            G.stoich[r++] = word(dec(kSlotP + a), dec(kSlotP + b), inc(kSlotD + d));
            G.stoich[r++] = word(inc(kSlotP + a), inc(kSlotP + b), dec(kSlotD + d));
        } else {
            G.stoich[r++] = word(dec2(kSlotP + a), (uint32_t)(kSlotDummy + 1), inc(kSlotD + d));
            G.stoich[r++] = word(inc2(kSlotP + a), (uint32_t)(kSlotDummy + 1), dec(kSlotD + d));
        }
    }
    G.n_entries = (uint8_t)r;
    G.stoich[r] = word((uint32_t)kSlotDummy);
    int longest = 8;                                   // a block of four dimers
    for (int s = 0; s < T.n_tf; ++s) longest = std::max(longest, 4 + 2 * (start[s + 1] - start[s]));
    int it = 0;
    while ((1 << it) < longest + 1) ++it;
    G.search_iters = (uint8_t)it;
    return G;
}

// Per-launch sizes.  Shared memory of one warp: state | running sums | reward scratch | program;
// the records live in global scratch, ser_species * 128 bytes per resident lane.
struct DimerLayout {
    int32_t ser_species;      // species rows in the per-lane record (max n_tf of the launch)
    int32_t cum_slots;        // max n_entries of the launch + 1 (the +inf sentinel)
    __host__ __device__ size_t rec_lane_bytes() const { return (size_t)ser_species * kSerStride; }
    __host__ __device__ size_t warp_bytes() const
    {
        return (size_t)(kSlots + cum_slots) * 32 * sizeof(float) + kSerStride * sizeof(float) + sizeof(DimerProg) +
               (size_t)kMaxSpecies * kSerStride;      // ... | staging copy of one finished record
    }
};

// delta of a stoichiometry field: code 0 = +1, 1 = -1, 2 = +2, 3 = -2 (built from the float's bits)
__device__ __forceinline__ float stoich_delta(uint32_t code)
{
    return __uint_as_float((0x3f800000u + ((code & 2u) << 22)) | (code << 31));
}

__global__ void __launch_bounds__(kDimThreads, 10) ssa_dimers_kernel(const LaunchArgs A, const DimerProg* __restrict__ progs,
                                                                 const DimerLayout lay, uint8_t* __restrict__ records)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t* wbase = smem + (size_t)warp * lay.warp_bytes();
    float* X = reinterpret_cast<float*>(wbase) + lane;        // slot s of this lane: X[s * 32]
    float* cum = X + kSlots * 32;                             // partial sum r of this lane: cum[r * 32]
    float* z = reinterpret_cast<float*>(wbase) + (kSlots + lay.cum_slots) * 32;
    DimerProg* G = reinterpret_cast<DimerProg*>(z + kSerStride);
    uint8_t* stage = reinterpret_cast<uint8_t*>(G) + sizeof(DimerProg);
    // records of this warp's 32 lanes (global scratch, see the header comment)
    const size_t ser_lane_bytes = lay.rec_lane_bytes();
    uint8_t* warp_ser = records + ((size_t)blockIdx.x * (kDimThreads / 32) + warp) * 32 * ser_lane_bytes;
    uint8_t* ser = warp_ser + (size_t)lane * ser_lane_bytes;

    // per-lane rate constants (zero = parked lane: nothing fires)
    float k_on = 0.f, k_on2 = 0.f, k_off1 = 0.f, k_off2 = 0.f, km0 = 0.f, km1 = 0.f, km2 = 0.f, km3 = 0.f, kp = 0.f, gm = 0.f,
          gp = 0.f, k_dim = 0.f, k_undim = 0.f;
    float t_rem = 0.0f;
    int k = 0, kd = 0, dcount = 1;
    uint32_t events = 0, ctr = 0, traj_lo = 0, traj_hi = 0, out = 0;
    bool active = false, finished = false;
#pragma unroll 1
    for (int s = 0; s < kSlots; ++s) X[s * 32] = 0.0f;

    // warp-uniform tile state (same protocol as ssa_pairwise_kernel, one topology per warp)
    Job job;
    job.topo = ~0ull; job.traj_base = 0; job.n_traj = 0; job.out_offset = 0; job.tile_start = 0; job.pad = 0;
    Job pend = job;
    uint32_t pos = 0, end = 0, pend_pos = 0, pend_end = 0, pend_job = 0;
    bool have_pending = false, out_of_work = false;
    // the current job's program, in registers (warp-uniform values)
    int n_tf = 0, n_dim = 0, n_entries = 0, search_iters = 0;
    uint32_t pk_ix_start = 0, pk_dim_a = 0, pk_dim_b = 0;
    uint64_t pk_ix_dim = 0;
    uint32_t n_iters = 0;

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
                // stage the record in shared memory (L2 reads: another lane wrote it), then reduce it
                const uint32_t* grec = reinterpret_cast<const uint32_t*>(warp_ser + (size_t)src * ser_lane_bytes);
                for (int w = lane; w < n_tf * (kSerStride / 4); w += 32) reinterpret_cast<uint32_t*>(stage)[w] = __ldcg(grec + w);
                __syncwarp();
                warp_reward(stage, z, n_tf, A.n_dec, lane, rw, lg);
                const uint32_t o = __shfl_sync(kFull, out, src);
                if (A.series) {
                    for (int s = 0; s < n_tf; ++s)
                        for (int q = lane; q < A.n_dec; q += 32)
                            A.series[(size_t)o * A.series_stride + s * A.n_dec + q] = stage[s * kSerStride + q];
                }
                __syncwarp();
                if (lane == src) {
                    A.reward[o] = rw;
                    if (A.lag) A.lag[o] = lg;
                    if (A.events) A.events[o] = events;
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
                        const uint32_t p = job.out_offset + pos + rank;
                        const uint32_t slot = A.order ? A.order[p] : p;
                        const uint64_t traj = job.traj_base + (slot - job.out_offset);
                        traj_lo = (uint32_t)traj; traj_hi = (uint32_t)(traj >> 32);
                        out = slot;
                        float prm[12];
                        trajectory_params(A, traj_lo, traj_hi, slot, prm);
                        k_on = prm[0]; k_on2 = 2.0f * prm[0]; k_off1 = prm[1]; k_off2 = prm[2];
                        km0 = prm[3]; km1 = prm[4]; km2 = prm[5]; km3 = prm[6];
                        kp = prm[7]; gm = prm[8]; gp = prm[9]; k_dim = prm[10]; k_undim = prm[11];
#pragma unroll 1
                        for (int s = 0; s < kSlots; ++s) X[s * 32] = 0.0f;
                        const float e0 = expf(-A.init_mean);
#pragma unroll 1
                        for (int blk4 = 0; blk4 < 2; ++blk4) {
                            const Philox4 x = philox4x32_10(3 + blk4, 0u, traj_lo, traj_hi, A.seed_lo, A.seed_hi);
#pragma unroll
                            for (int w = 0; w < 4; ++w) {
                                const int s = 4 * blk4 + w;
                                if (s < n_tf) {
                                    const float u = u01(x.w[w]);
                                    float pk = e0, F = e0;
                                    int kk = 0;
                                    while (u > F && kk < 1000) { ++kk; pk *= A.init_mean / (float)kk; F += pk; }
                                    X[(kSlotP + s) * 32] = (float)kk;
                                }
                            }
                        }
                        t_rem = 0.0f;
                        k = 0; kd = 0; dcount = A.decim;
                        events = 0; ctr = 0;
                        active = true;
                    }
                    const uint32_t n_need = __popc(need);
                    pos += (n_need < avail) ? n_need : avail;
                    need = __ballot_sync(kFull, !active);
                    if (!need) break;
                }
                if (!have_pending && !out_of_work) {
                    uint32_t t = 0;
                    if (lane == 0) t = atomicAdd(A.tile_counter, 1u);
                    t = __shfl_sync(kFull, t, 0);
                    if (t >= A.n_tiles) {
                        out_of_work = true;
                    } else {
                        uint32_t lo = 0, hi = A.n_jobs - 1;       // last job with tile_start <= t
                        while (lo < hi) {
                            const uint32_t mid = (lo + hi + 1) >> 1;
                            if (A.jobs[mid].tile_start <= t) lo = mid; else hi = mid - 1;
                        }
                        pend = A.jobs[lo];
                        pend_job = lo;
                        pend_pos = (t - pend.tile_start) * kTile;
                        pend_end = min(pend_pos + (uint32_t)kTile, pend.n_traj);
                        have_pending = true;
                    }
                }
                if (have_pending) {
                    const bool same = (pend.topo == job.topo);
                    const bool drained = (__ballot_sync(kFull, active) == 0u);
                    if (same || drained) {
                        if (!same) {
                            // switch topology: the stoichiometry table goes to shared memory, the rest to registers
                            const uint32_t* src = reinterpret_cast<const uint32_t*>(progs + pend_job);
                            uint32_t* dst = reinterpret_cast<uint32_t*>(G);
                            __syncwarp();
                            for (int w = lane; w < (int)(sizeof(DimerProg) / 4); w += 32) dst[w] = src[w];
                            __syncwarp();
                            n_tf = G->n_tf; n_dim = G->n_dim; n_entries = G->n_entries; search_iters = G->search_iters;
                            pk_ix_start = G->pk_ix_start; pk_dim_a = G->pk_dim_a; pk_dim_b = G->pk_dim_b; pk_ix_dim = G->pk_ix_dim;
                            cum[n_entries * 32] = 3.4e38f;    // sentinel (every lane): "no reaction" is always found
                        }
                        job = pend; pos = pend_pos; end = pend_end;
                        have_pending = false;
                        continue;
                    }
                }
                break;   // wait for running lanes (or nothing left)
            }
            if (out_of_work && __ballot_sync(kFull, active) == 0u) break;
        }
        ++n_iters;

        const Philox4 x = philox4x32_10_rk(ctr, 1u, traj_lo, traj_hi, A.rk);
        ctr += 1;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const uint32_t w_tau = h ? x.w[2] : x.w[0], w_sel = h ? x.w[3] : x.w[1];
            // ---- pass 1: block-wise running sums in SPEC order (section 9): one block per TF, one per four
            // dimers; every block starts at zero, every partial sum is stored, the block totals are added up
            // in block order (oe[b] = end offset of block b; an absent block adds an exact 0) ----
            float osum = 0.0f;
            float oe[kMaxSpecies + 3];
            float* cp = cum;
#pragma unroll
            for (int s = 0; s < kMaxSpecies; ++s) {
                if (s < n_tf) {      // warp-uniform
                    const float n = X[(kSlotN + s) * 32], na = X[(kSlotA + s) * 32];
                    const float r = X[(kSlotR + s) * 32], p = X[(kSlotP + s) * 32];
                    const float km_a = (na > 0.0f) ? km1 : km0;
                    const float km_i = (na > 0.0f) ? km3 : km2;
                    float c = (n > na) ? km_i : km_a;   cp[0] = c;
                    c = fmaf(gm, r, c);             cp[32] = c;
                    c = fmaf(kp, r, c);             cp[64] = c;
                    c = fmaf(gp, p, c);             cp[96] = c;
                    cp += 128;
                    const int x0 = (pk_ix_start >> (4 * s)) & 15, x1 = (pk_ix_start >> (4 * s + 4)) & 15;
                    if (x1 > x0) {
                        const float kfree = fmaf(-k_on, n, k_on2);            // k_on (2 - n)
                        const float koff = (n >= 2.0f) ? k_off2 : k_off1;
#pragma unroll 1
                        for (int q = x0; q < x1; ++q) {
                            const int d = (int)(pk_ix_dim >> (4 * q)) & 15;
                            c = fmaf(kfree, X[(kSlotD + d) * 32], c); *cp = c; cp += 32;
                        }
#pragma unroll 1
                        for (int q = x0; q < x1; ++q) { c = fmaf(koff, X[(kSlotB + q) * 32], c); *cp = c; cp += 32; }
                    }
                    osum += c;
                }
                oe[s] = osum;
            }
#pragma unroll
            for (int gq = 0; gq < 3; ++gq) {
                float c = 0.0f;
                const int d_end = min(n_dim, 4 * gq + 4);
#pragma unroll 1
                for (int d = 4 * gq; d < d_end; ++d) {
                    const int da = (pk_dim_a >> (3 * d)) & 7, db = (pk_dim_b >> (3 * d)) & 7;
                    const float pa = X[(kSlotP + da) * 32], pb = X[(kSlotP + db) * 32];
                    const float pairs = (da != db) ? pa * pb : 0.5f * pa * (pa - 1.0f);
                    c = fmaf(k_dim, pairs, c);                      cp[0] = c;
                    c = fmaf(k_undim, X[(kSlotD + d) * 32], c);     cp[32] = c;
                    cp += 64;
                }
                osum += c;
                oe[kMaxSpecies + gq] = osum;
            }
            const float a0 = osum;

            // ---- time to next event; grid samples crossed keep the pre-event state ----
            const float tau = (-0.69314718056f * fast_log2(u01(w_tau))) * fast_rcp(a0);   // a0 == 0 -> +inf
            t_rem -= tau;
            bool alive = true;
            const bool was_active = active;
            if (active && t_rem < 0.0f) {
                do {
#pragma unroll
                    for (int s = 0; s < kMaxSpecies; ++s)
                        if (s < n_tf) X[(kSlotBlk + s) * 32] += fminf(X[(kSlotP + s) * 32], kSampleCap);
                    k += 1;
                    if (--dcount == 0) {
                        dcount = A.decim;
#pragma unroll 1
                        CT_CHECK(kd >= 0 && kd < kSerStride);
                        for (int s = 0; s < n_tf; ++s) {
                            ser[s * kSerStride + kd] = (uint8_t)compand_block(X[(kSlotBlk + s) * 32]);
                            X[(kSlotBlk + s) * 32] = 0.0f;
                        }
                        kd += 1;
                    }
                    t_rem += A.dt;
                } while (t_rem < 0.0f && k < A.nt);
                if (k >= A.nt) { alive = false; active = false; finished = true; }
            }

            // ---- pass 2: each lane finds its reaction (first partial sum above the target) and fires it ----
            const float target = alive ? u01(w_sel) * a0 : 3.0e38f;
            // the block whose offsets bracket the target, then the first entry of that block whose local sum exceeds
            // the target's position inside the block; rounding can leave none: nothing fires (SPEC.md section 2)
            int bl = 0;
            float off = 0.0f;
#pragma unroll
            for (int b = 0; b < kMaxSpecies + 3; ++b) {
                const bool below = oe[b] <= target;
                bl += below ? 1 : 0;
                off = below ? oe[b] : off;
            }
            const float tl = target - off;
            const int tf_end = 4 * n_tf + 2 * (int)((pk_ix_start >> (4 * kMaxSpecies)) & 15u);
            int lo, hi;
            if (bl < kMaxSpecies) {
                lo = 4 * bl + 2 * (int)((pk_ix_start >> (4 * bl)) & 15u);
                hi = 4 * bl + 4 + 2 * (int)((pk_ix_start >> (4 * bl + 4)) & 15u);
            } else {
                lo = min(tf_end + 8 * (bl - kMaxSpecies), n_entries);
                hi = min(lo + 8, n_entries);
            }
            const int blk_end = hi;
#pragma unroll 1
            for (int it = 0; it < search_iters; ++it) {
                const int mid = (lo + hi) >> 1;
                const bool open = lo < hi;
                const bool lt = open && (tl < cum[mid * 32]);
                hi = lt ? mid : hi;
                lo = (open && !lt) ? mid + 1 : lo;
            }
            CT_CHECK(lo >= 0 && lo <= n_entries && n_entries <= kMaxEntries && (int)(cp - cum) == n_entries * 32);
            const uint32_t sw = G->stoich[(lo < blk_end) ? lo : n_entries];
            CT_CHECK((sw & 63u) < kSlots && ((sw >> 8) & 63u) < kSlots && ((sw >> 16) & 63u) < kSlots && ((sw >> 24) & 63u) < kSlots);
            // the four fields name distinct slots: load all, then store all
            float* f0 = X + (sw & 63u) * 32;
            float* f1 = X + ((sw >> 8) & 63u) * 32;
            float* f2 = X + ((sw >> 16) & 63u) * 32;
            float* f3 = X + ((sw >> 24) & 63u) * 32;
            const float v0 = *f0, v1 = *f1, v2 = *f2, v3 = *f3;
            *f0 = v0 + stoich_delta((sw >> 6) & 3u);
            *f1 = v1 + stoich_delta((sw >> 14) & 3u);
            *f2 = v2 + stoich_delta((sw >> 22) & 3u);
            *f3 = v3 + stoich_delta((sw >> 30) & 3u);
            events += (was_active && alive) ? 1u : 0u;
            if (!alive) {   // park: zero rates and state so that the lane idles through the lock-step loop
                k_on = k_on2 = k_off1 = k_off2 = km0 = km1 = km2 = km3 = kp = gm = gp = k_dim = k_undim = 0.0f;
#pragma unroll 1
                for (int s = 0; s < kSlots; ++s) X[s * 32] = 0.0f;
                t_rem = 0.0f;
            }
        }
    }
    if (A.warp_iters && lane == 0) atomicAdd(A.warp_iters, (unsigned long long)n_iters);
}


// =====================================================================================================
// Per-lane programs: the kernel for launches of MANY SMALL dimer jobs (MCTS batches).
//
// ssa_dimers_kernel above runs one topology per warp, so a job of 32 trajectories is one warp that
// waits for its slowest lane before it may change topology; a batch of a thousand distinct
// topologies then runs at a few percent of the machine.  Here every lane carries its own topology
// (eight packed words in registers) and lanes are refilled one by one from the launch-wide,
// longest-first order, exactly like the pairwise kernel's mixed mode.  To keep the warp in
// lock-step the reaction list has ONE shape for the whole launch: n_tf_max TFs x (4 + 2 K) entries
// (K = kMixK interaction slots per promoter) followed by 2 entries for each of n_dim_max dimers; what
// a lane's topology lacks has propensity zero and can never be selected.  Absent entries add an
// exact 0 to the running sum and the present ones come in SPEC order, so every partial sum, the
// chosen reaction and therefore every result is bit-identical to ssa_dimers_kernel
// (tests/test_gpu_dimers.py::test_per_lane_programs_are_bit_identical).  The stoichiometry of entry r
// follows from r and the lane's packed words; no per-lane table is needed.
// =====================================================================================================
constexpr int kMixK = 2;                       // interaction slots per promoter (the DimersGrammar default maximum)
constexpr int kMixPerTf = 4 + 2 * kMixK;       // 8 entries per TF: power of two, so entry -> (TF, kind) is a shift

struct DimerLaneProg {             // 32 bytes per job
    uint32_t ixw[kMaxSpecies];     // TF s: interaction slot k at bits 6k: present | activates << 1 | dimer << 2
    uint32_t dim_a, dim_b;         // monomers of dimer d at bits 3d
    uint32_t dmask_ntf;            // bit d: dimer d exists; bits 16..19: number of TFs
};

// false when some promoter has more than kMixK interactions (such launches use ssa_dimers_kernel)
inline bool dimer_lane_program(const DimerTopo& T, DimerLaneProg& L)
{
    memset(&L, 0, sizeof L);
    for (int s = 0; s < T.n_tf; ++s) {
        int k = 0;
        for (int e = 0; e < T.n_ixn; ++e) {
            if (T.ix_tgt[e] != s) continue;
            if (k == kMixK) return false;
            L.ixw[s] |= (1u | ((T.ix_act[e] ? 1u : 0u) << 1) | ((uint32_t)T.ix_dim[e] << 2)) << (6 * k);
            ++k;
        }
    }
    for (int d = 0; d < T.n_dim; ++d) {
        L.dim_a |= (uint32_t)T.dim_a[d] << (3 * d);
        L.dim_b |= (uint32_t)T.dim_b[d] << (3 * d);
        L.dmask_ntf |= 1u << d;
    }
    L.dmask_ntf |= (uint32_t)T.n_tf << 16;
    return true;
}

struct DimerMixShape {             // launch-wide shape of the reaction list
    int32_t n_tf, n_dim;           // maxima over the launch's jobs
    int32_t n_entries;             // n_tf * kMixPerTf + 2 * n_dim
    int32_t search_iters;          // ceil(log2(n_entries + 1))
    __host__ __device__ size_t rec_lane_bytes() const { return (size_t)n_tf * kSerStride; }
    __host__ __device__ size_t warp_bytes() const
    {
        return (size_t)(kSlots + n_entries + 1) * 32 * sizeof(float) + kSerStride * sizeof(float) +
               (size_t)kMaxSpecies * kSerStride;      // state | running sums | reward scratch | staged record
    }
};

__global__ void __launch_bounds__(kDimThreads, 8) ssa_dimers_mixed_kernel(const LaunchArgs A, const DimerLaneProg* __restrict__ lprogs,
                                                                          const DimerMixShape shape, uint8_t* __restrict__ records)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t* wbase = smem + (size_t)warp * shape.warp_bytes();
    float* X = reinterpret_cast<float*>(wbase) + lane;        // slot s of this lane: X[s * 32]
    float* cum = X + kSlots * 32;                             // partial sum r of this lane: cum[r * 32]
    float* z = reinterpret_cast<float*>(wbase) + (kSlots + shape.n_entries + 1) * 32;
    uint8_t* stage = reinterpret_cast<uint8_t*>(z + kSerStride);
    const size_t ser_lane_bytes = shape.rec_lane_bytes();
    uint8_t* warp_ser = records + ((size_t)blockIdx.x * (kDimThreads / 32) + warp) * 32 * ser_lane_bytes;
    uint8_t* ser = warp_ser + (size_t)lane * ser_lane_bytes;
    const int ntf_max = shape.n_tf, nd_max = shape.n_dim, n_entries = shape.n_entries;
    const int tf_entries = ntf_max * kMixPerTf;

    float k_on = 0.f, k_on2 = 0.f, k_off1 = 0.f, k_off2 = 0.f, km0 = 0.f, km1 = 0.f, km2 = 0.f, km3 = 0.f, kp = 0.f, gm = 0.f,
          gp = 0.f, k_dim = 0.f, k_undim = 0.f;
    float t_rem = 0.0f;
    int k = 0, kd = 0, dcount = 1;
    uint32_t events = 0, ctr = 0, traj_lo = 0, traj_hi = 0, out = 0;
    bool active = false, finished = false;
    // this lane's topology (a parked lane: no TFs, no dimers)
    uint32_t ixw[kMaxSpecies] = {0u, 0u, 0u, 0u, 0u};
    uint32_t dma = 0, dmb = 0, dmask = 0;
    int n_tf = 0;
#pragma unroll 1
    for (int s = 0; s < kSlots; ++s) X[s * 32] = 0.0f;
    cum[n_entries * 32] = 3.4e38f;        // sentinel: "no reaction" is always found

    uint32_t pos = 0, end = 0;            // launch positions of the warp's current tile
    bool out_of_work = false;
    uint32_t n_iters = 0;

    for (;;) {
        const unsigned idle = __ballot_sync(kFull, !active);
        if (idle) {
            // 1. reduce finished trajectories to (reward, lag), one lane at a time, whole warp helping
            unsigned fin = __ballot_sync(kFull, finished);
            if (fin) __syncwarp();
            while (fin) {
                const int src = __ffs(fin) - 1;
                fin &= fin - 1;
                float rw; int lg;
                const int src_ntf = __shfl_sync(kFull, n_tf, src);
                const uint32_t* grec = reinterpret_cast<const uint32_t*>(warp_ser + (size_t)src * ser_lane_bytes);
                for (int w = lane; w < src_ntf * (kSerStride / 4); w += 32) reinterpret_cast<uint32_t*>(stage)[w] = __ldcg(grec + w);
                __syncwarp();
                warp_reward(stage, z, src_ntf, A.n_dec, lane, rw, lg);
                const uint32_t o = __shfl_sync(kFull, out, src);
                if (A.series) {
                    for (int s = 0; s < src_ntf; ++s)
                        for (int q = lane; q < A.n_dec; q += 32)
                            A.series[(size_t)o * A.series_stride + s * A.n_dec + q] = stage[s * kSerStride + q];
                }
                __syncwarp();
                if (lane == src) {
                    A.reward[o] = rw;
                    if (A.lag) A.lag[o] = lg;
                    if (A.events) A.events[o] = events;
                    finished = false;
                }
            }
            // 2. refill idle lanes from the launch-wide order (tiles are ranges of launch positions)
            unsigned need = idle;
            for (;;) {
                if (pos < end) {
                    const int rank = __popc(need & ((1u << lane) - 1u));
                    const uint32_t avail = end - pos;
                    const bool take = !active && (uint32_t)rank < avail;
                    if (take) {
                        const uint32_t slot = A.order ? A.order[pos + rank] : pos + rank;
                        const uint32_t jx = job_of_slot(A, slot);
                        const Job mine = A.jobs[jx];
                        const DimerLaneProg P = lprogs[jx];
#pragma unroll
                        for (int s = 0; s < kMaxSpecies; ++s) ixw[s] = P.ixw[s];
                        dma = P.dim_a; dmb = P.dim_b; dmask = P.dmask_ntf & 0x3ffu; n_tf = (int)(P.dmask_ntf >> 16) & 15;
                        const uint64_t traj = mine.traj_base + (slot - mine.out_offset);
                        traj_lo = (uint32_t)traj; traj_hi = (uint32_t)(traj >> 32);
                        out = slot;
                        float prm[12];
                        trajectory_params(A, traj_lo, traj_hi, slot, prm);
                        k_on = prm[0]; k_on2 = 2.0f * prm[0]; k_off1 = prm[1]; k_off2 = prm[2];
                        km0 = prm[3]; km1 = prm[4]; km2 = prm[5]; km3 = prm[6];
                        kp = prm[7]; gm = prm[8]; gp = prm[9]; k_dim = prm[10]; k_undim = prm[11];
#pragma unroll 1
                        for (int s = 0; s < kSlots; ++s) X[s * 32] = 0.0f;
                        const float e0 = expf(-A.init_mean);
#pragma unroll 1
                        for (int blk4 = 0; blk4 < 2; ++blk4) {
                            const Philox4 x = philox4x32_10(3 + blk4, 0u, traj_lo, traj_hi, A.seed_lo, A.seed_hi);
#pragma unroll
                            for (int w = 0; w < 4; ++w) {
                                const int s = 4 * blk4 + w;
                                if (s < n_tf) {
                                    const float u = u01(x.w[w]);
                                    float pk = e0, F = e0;
                                    int kk = 0;
                                    while (u > F && kk < 1000) { ++kk; pk *= A.init_mean / (float)kk; F += pk; }
                                    X[(kSlotP + s) * 32] = (float)kk;
                                }
                            }
                        }
                        t_rem = 0.0f;
                        k = 0; kd = 0; dcount = A.decim;
                        events = 0; ctr = 0;
                        active = true;
                    }
                    const uint32_t n_need = __popc(need);
                    pos += (n_need < avail) ? n_need : avail;
                    need = __ballot_sync(kFull, !active);
                    if (!need) break;
                }
                if (out_of_work) break;
                uint32_t t = 0;
                if (lane == 0) t = atomicAdd(A.tile_counter, 1u);
                t = __shfl_sync(kFull, t, 0);
                if (t >= A.n_tiles) { out_of_work = true; break; }
                pos = t * kTile;
                end = min(pos + (uint32_t)kTile, A.total);
            }
            if (out_of_work && __ballot_sync(kFull, active) == 0u) break;
        }
        ++n_iters;

        const Philox4 x = philox4x32_10_rk(ctr, 1u, traj_lo, traj_hi, A.rk);
        ctr += 1;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const uint32_t w_tau = h ? x.w[2] : x.w[0], w_sel = h ? x.w[3] : x.w[1];
            // ---- pass 1: the launch-wide reaction list, block-wise (SPEC.md section 9: one block per TF, one per four
            // dimers, every block starts at zero, totals added in block order); what this lane's topology lacks
            // contributes an exact 0 ----
            float osum = 0.0f;
            float oe[kMaxSpecies + 3];
            float* cp = cum;
#pragma unroll
            for (int s = 0; s < kMaxSpecies; ++s) {
                if (s < ntf_max) {       // warp-uniform (launch constant)
                    const float n = X[(kSlotN + s) * 32], na = X[(kSlotA + s) * 32];
                    const float r = X[(kSlotR + s) * 32], p = X[(kSlotP + s) * 32];
                    const float km_a = (na > 0.0f) ? km1 : km0;
                    const float km_i = (na > 0.0f) ? km3 : km2;
                    const float km = (n > na) ? km_i : km_a;
                    float c = (s < n_tf) ? km : 0.0f;    cp[0] = c;      // (a TF this lane does not have: no transcription)
                    c = fmaf(gm, r, c);             cp[32] = c;
                    c = fmaf(kp, r, c);             cp[64] = c;
                    c = fmaf(gp, p, c);             cp[96] = c;
                    const float kfree = fmaf(-k_on, n, k_on2);            // k_on (2 - n)
                    const float koff = (n >= 2.0f) ? k_off2 : k_off1;
                    const uint32_t iw = ixw[s];
#pragma unroll
                    for (int q = 0; q < kMixK; ++q) {
                        const uint32_t f = iw >> (6 * q);
                        const float dv = X[(kSlotD + ((f >> 2) & 15u)) * 32];
                        c = fmaf((f & 1u) ? kfree : 0.0f, dv, c);       cp[(4 + q) * 32] = c;
                    }
#pragma unroll
                    for (int q = 0; q < kMixK; ++q) {
                        c = fmaf(koff, X[(kSlotB + kMixK * s + q) * 32], c);    cp[(4 + kMixK + q) * 32] = c;
                    }
                    cp += kMixPerTf * 32;
                    osum += c;
                }
                oe[s] = osum;
            }
#pragma unroll
            for (int gq = 0; gq < 3; ++gq) {
                float c = 0.0f;
                const int d_end = min(nd_max, 4 * gq + 4);
#pragma unroll 1
                for (int d = 4 * gq; d < d_end; ++d) {
                    const uint32_t da = (dma >> (3 * d)) & 7u, db = (dmb >> (3 * d)) & 7u;
                    const float pa = X[(kSlotP + da) * 32], pb = X[(kSlotP + db) * 32];
                    const float pairs = (da != db) ? pa * pb : 0.5f * pa * (pa - 1.0f);
                    c = fmaf(((dmask >> d) & 1u) ? k_dim : 0.0f, pairs, c);     cp[0] = c;
                    c = fmaf(k_undim, X[(kSlotD + d) * 32], c);                  cp[32] = c;
                    cp += 64;
                }
                osum += c;
                oe[kMaxSpecies + gq] = osum;
            }
            const float a0 = osum;

            // ---- time to next event; grid samples crossed keep the pre-event state ----
            const float tau = (-0.69314718056f * fast_log2(u01(w_tau))) * fast_rcp(a0);   // a0 == 0 -> +inf
            t_rem -= tau;
            bool alive = true;
            const bool was_active = active;
            if (active && t_rem < 0.0f) {
                do {
#pragma unroll
                    for (int s = 0; s < kMaxSpecies; ++s)
                        if (s < n_tf) X[(kSlotBlk + s) * 32] += fminf(X[(kSlotP + s) * 32], kSampleCap);
                    k += 1;
                    if (--dcount == 0) {
                        dcount = A.decim;
#pragma unroll 1
                        CT_CHECK(kd >= 0 && kd < kSerStride);
                        for (int s = 0; s < n_tf; ++s) {
                            ser[s * kSerStride + kd] = (uint8_t)compand_block(X[(kSlotBlk + s) * 32]);
                            X[(kSlotBlk + s) * 32] = 0.0f;
                        }
                        kd += 1;
                    }
                    t_rem += A.dt;
                } while (t_rem < 0.0f && k < A.nt);
                if (k >= A.nt) { alive = false; active = false; finished = true; }
            }

            // ---- pass 2: find the reaction (first partial sum above the target), derive its stoichiometry, fire ----
            const float target = alive ? u01(w_sel) * a0 : 3.0e38f;
            // the block whose offsets bracket the target, then the first of its (at most eight) entries whose local
            // sum exceeds the target's position inside the block; none (rounding): nothing fires
            int bl = 0;
            float off = 0.0f;
#pragma unroll
            for (int b = 0; b < kMaxSpecies + 3; ++b) {
                const bool below = oe[b] <= target;
                bl += below ? 1 : 0;
                off = below ? oe[b] : off;
            }
            const float tl = target - off;
            int lo = (bl < kMaxSpecies) ? kMixPerTf * bl : tf_entries + 8 * (bl - kMaxSpecies);
            lo = min(lo, n_entries);
            int hi = min(lo + 8, n_entries);
            const int blk_end = hi;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int mid = (lo + hi) >> 1;
                const bool open = lo < hi;
                const bool lt = open && (tl < cum[mid * 32]);
                hi = lt ? mid : hi;
                lo = (open && !lt) ? mid + 1 : lo;
            }
            lo = (lo < blk_end) ? lo : n_entries;
            // four (slot, delta) fields on distinct slots; unused fields sit on the four dummies
            int s0 = kSlotDummy, s1 = kSlotDummy + 1, s2 = kSlotDummy + 2, s3 = kSlotDummy + 3;
            float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
            if (lo < tf_entries) {
                const int s = lo >> 3, w = lo & 7;
                static_assert(kMixPerTf == 8 && kMixK == 2, "entry -> (TF, kind) decoding assumes 8 entries per TF");
                if (w < 4) {        // transcription, mRNA decay, translation, monomer decay
                    s0 = ((w & 2) ? kSlotP : kSlotR) + s;
                    d0 = (w & 1) ? -1.0f : 1.0f;
                } else {            // binding (w = 4, 5) / unbinding (w = 6, 7) of interaction slot q at promoter s
                    const int q = w & 1;
                    const float sg = (w >= 6) ? -1.0f : 1.0f;
                    uint32_t iw = ixw[0];
#pragma unroll
                    for (int j = 1; j < kMaxSpecies; ++j) iw = (s == j) ? ixw[j] : iw;
                    const uint32_t f = iw >> (6 * q);
                    s0 = kSlotD + (int)((f >> 2) & 15u);    d0 = -sg;
                    s1 = kSlotB + kMixK * s + q;            d1 = sg;
                    s2 = kSlotN + s;                        d2 = sg;
                    if (f & 2u) { s3 = kSlotA + s;          d3 = sg; }
                }
            } else if (lo < n_entries) {     // dimerisation (even) / dissociation (odd) of dimer d
                const int j = lo - tf_entries, d = j >> 1;
                const float sg = (j & 1) ? 1.0f : -1.0f;        // what happens to the monomers
                const int da = (int)((dma >> (3 * d)) & 7u), db = (int)((dmb >> (3 * d)) & 7u);
                s0 = kSlotP + da;   d0 = (da == db) ? 2.0f * sg : sg;
                if (da != db) { s1 = kSlotP + db;   d1 = sg; }
                s2 = kSlotD + d;    d2 = -sg;
            }
            CT_CHECK(s0 >= 0 && s0 < kSlots && s1 >= 0 && s1 < kSlots && s2 >= 0 && s2 < kSlots && s3 >= 0 && s3 < kSlots);
            CT_CHECK(lo >= 0 && lo <= n_entries && (int)(cp - cum) == n_entries * 32);
            float* f0 = X + s0 * 32;
            float* f1 = X + s1 * 32;
            float* f2 = X + s2 * 32;
            float* f3 = X + s3 * 32;
            const float v0 = *f0, v1 = *f1, v2 = *f2, v3 = *f3;
            *f0 = v0 + d0;
            *f1 = v1 + d1;
            *f2 = v2 + d2;
            *f3 = v3 + d3;
            events += (was_active && alive) ? 1u : 0u;
            if (!alive) {   // park: zero rates, state and topology so that the lane idles through the lock-step loop
                k_on = k_on2 = k_off1 = k_off2 = km0 = km1 = km2 = km3 = kp = gm = gp = k_dim = k_undim = 0.0f;
#pragma unroll 1
                for (int s = 0; s < kSlots; ++s) X[s * 32] = 0.0f;
                t_rem = 0.0f;
            }
        }
    }
    if (A.warp_iters && lane == 0) atomicAdd(A.warp_iters, (unsigned long long)n_iters);
}

}  // namespace ct
