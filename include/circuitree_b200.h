/*
 * circuitree_b200.h -- C ABI of the B200 rollout engine behind CircuiTree.get_reward.
 *
 * The reference has no FFI: its boundary for this path is the Python method
 *     get_reward(self, state, **kwargs) -> float      (reference circuitree/circuitree.py:135-152)
 * called from traverse (circuitree.py:534) and search_bfs (circuitree.py:687), fed by
 *     SimpleNetworkGrammar.parse_genotype             (reference circuitree/models.py:441-480)
 *     DimersGrammar.parse_genotype                    (reference circuitree/models.py:844-884)
 * The entry points below are what a ctypes/cffi binding on the reference side would bind
 * (see INTEGRATION.md).  Plain pointers and sizes only; no torch types; integer return codes
 * (0 = ok); nothing throws across the ABI.  ct_last_error() describes the last failure on the
 * calling thread.
 */
#ifndef CIRCUITREE_B200_H
#define CIRCUITREE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CT_N_PARAMS 10       /* pairwise model, SPEC.md section 4 */
#define CT_N_PARAMS_DIMER 12 /* dimer model, SPEC.md section 9: + k_dim, k_undim */
#define CT_MAX_PARAMS 12
#define CT_MAX_INTERACTIONS 10 /* dimer model: interactions (and distinct dimers) per topology */
#define CT_MAX_SPECIES 5     /* GPU path: components per state */
#define CT_MAX_NDEC 126      /* decimated samples kept per species (128 shared-memory bytes each) */

enum ct_status {
    CT_OK = 0,
    CT_ERR_INVALID = 1,      /* bad argument */
    CT_ERR_CUDA = 2,         /* CUDA runtime error (message in ct_last_error) */
    CT_ERR_UNSUPPORTED = 3,  /* e.g. more than CT_MAX_SPECIES components */
    CT_ERR_NO_DEVICE = 4
};

/* Simulation grid and reward settings (SPEC.md sections 5-7). */
typedef struct ct_sim_config {
    int32_t nt;              /* grid points, default 2000 */
    int32_t decim;           /* grid points per stored block, default 16; nt/decim <= CT_MAX_NDEC */
    float dt;                /* seconds between grid points, default 20 */
    float init_mean;         /* Poisson mean of initial protein counts, default 10 */
    int32_t schedule;        /* 0 auto; 1 one topology per warp (big jobs); 2 mixed topologies per warp
                                (many small jobs); 3 cooperative: a sub-group of 4 or 8 lanes shares one
                                trajectory (launches too small to fill the device: a shorter step instead of
                                more trajectories per warp).  Results are identical, only the speed differs. */
    int32_t lpt;             /* 1 (default): start trajectories longest-expected-first (cost pre-pass +
                                radix sort); 0: in id order.  Results are identical. */
    int32_t grid_limit;      /* 0 (default): a launch may fill every CTA slot of the device; k > 0: at most k CTAs,
                                so that several launches submitted on different streams (ct_rewards_submit)
                                share the device side by side instead of one after the other.  Results are identical. */
    int32_t fast_path;       /* 1 (default): launches whose topologies give every gene at most one regulator
                                use the specialised hot loop; 0: always the general one.  Results are identical. */
    float ranges[CT_MAX_PARAMS][3]; /* (lo, hi, is_log) per parameter, SPEC.md sections 4 and 9
                                (rows 10, 11 = k_dim, k_undim are used by dimer topologies only) */
} ct_sim_config;

typedef struct ct_engine ct_engine;   /* opaque; one per device */

const char *ct_last_error(void);
int ct_version(void);

/* Fill cfg with the SPEC.md defaults. */
int ct_default_config(ct_sim_config *cfg);

/* Create / destroy an engine bound to CUDA device `device`. */
int ct_engine_create(int device, ct_engine **out);
int ct_engine_destroy(ct_engine *eng);

/* Number of SMs and resident SSA lanes of the engine's device (for sizing batches). */
int ct_engine_info(ct_engine *eng, int32_t *n_sm, int32_t *resident_lanes);

/*
 * Compile one topology from parse_genotype output (models.py:441-480): `act` and `inh` are
 * int64 [n][k] row-major index arrays, k = 2 for SimpleNetworkGrammar ((regulator, target)).
 * The k = 2 handle is a self-contained 64-bit code.
 * k = 3 (DimersGrammar, models.py:844-884): rows are (monomer1 <= monomer2, target), indices into
 * components + regulators, n_species = number of components + regulators (a TF that is never a
 * target is expressed constitutively, which is what SPEC.md section 9 says of regulators).  The
 * handle (bit 63 set) names an entry of a process-wide, append-only topology table; equal
 * topologies get equal handles.  Neither kind of handle needs a release.
 * A launch takes pairwise handles only or dimer handles only (CT_ERR_INVALID otherwise); explicit
 * parameter rows are CT_N_PARAMS wide for pairwise and CT_N_PARAMS_DIMER wide for dimer launches.
 */
int ct_compile_topology(int32_t n_species, const int64_t *act, int32_t n_act,
                        const int64_t *inh, int32_t n_inh, int32_t k, uint64_t *handle);

/* Species (TFs) and reactions of a compiled topology (SPEC.md sections 2 and 9): the inputs of the
 * per-step work I_step that bench.py's roofline uses (SPEC.md section 8). */
int ct_topology_info(uint64_t handle, int32_t *n_species, int32_t *n_reactions);

/* The check ct_rewards_host / ct_rewards_submit apply to explicit parameter vectors (host float32
 * [n_trajectories][n_params]): CT_ERR_INVALID unless every value is finite and non-negative.  A caller that fills
 * the device buffer of ct_launch_rewards itself can run it on its host copy first.  Needs no GPU.
 * Not checked: magnitude.  A launch lasts as long as its longest trajectory, and the clock is float32: rates
 * that put more than ~10^6 events into one grid interval dt make a trajectory arbitrarily slow (SPEC.md's ranges
 * give 10^1..10^3). */
int ct_check_params(const float *h_params, int64_t n_trajectories, int32_t n_params);

/*
 * Device-buffer launch.  Job j simulates n_traj[j] trajectories of topology handles[j]; its
 * trajectory ids are traj_base[j] .. traj_base[j]+n_traj[j]-1 and its outputs occupy the slots
 * after those of jobs 0..j-1.  All d_* pointers are device pointers on the engine's device:
 *   d_params  (nullable) float32 [total][CT_N_PARAMS or CT_N_PARAMS_DIMER] explicit parameters, else sampled in-kernel
 *   d_reward  float32 [total]     d_lag  int32 [total] (nullable)   d_events uint32 [total] (nullable)
 *   d_series  (nullable, debug) uint8 [total][n_species_max][nt/decim] companded block sums
 * `stream` is a cudaStream_t (0 = default stream).  Asynchronous.
 * job_cost (host, nullable) float32 [n_jobs]: relative expected events per trajectory of each job's
 * topology (e.g. observed earlier); only steers the longest-first launch order, never the results.
 */
int ct_launch_rewards(ct_engine *eng, const uint64_t *handles, const uint32_t *n_traj,
                      const uint64_t *traj_base, const float *job_cost, int32_t n_jobs,
                      const ct_sim_config *cfg, uint64_t seed, void *stream, const float *d_params,
                      float *d_reward, int32_t *d_lag, uint32_t *d_events, uint8_t *d_series,
                      int32_t series_stride);

/*
 * Per-job reduction on the device (deterministic order): mean reward (float64), number of
 * events (uint64) for each job of the preceding launch layout.
 */
int ct_reduce_jobs(ct_engine *eng, const uint32_t *n_traj, int32_t n_jobs, void *stream,
                   const float *d_reward, const uint32_t *d_events,
                   double *d_job_mean, uint64_t *d_job_events);

/*
 * Host-buffer convenience used by get_reward / get_rewards: copies the job table in, runs
 * the kernel and the per-job reduction, copies job_mean[n_jobs] (float64), job_events[n_jobs]
 * (uint64, nullable) and optionally the per-trajectory reward / lag arrays (host, nullable)
 * back.  Synchronous.  h_params (nullable: sample per trajectory, SPEC.md section 4) holds CT_N_PARAMS
 * (CT_N_PARAMS_DIMER) float32 per trajectory; every value must be finite and non-negative, else CT_ERR_INVALID
 * (a NaN or infinite rate would stall the simulated clock).  ct_launch_rewards does not check its device buffer.
 */
int ct_rewards_host(ct_engine *eng, const uint64_t *handles, const uint32_t *n_traj,
                    const uint64_t *traj_base, const float *job_cost, int32_t n_jobs,
                    const ct_sim_config *cfg, uint64_t seed, const float *h_params, double *h_job_mean,
                    uint64_t *h_job_events, float *h_reward, int32_t *h_lag);

/*
 * Asynchronous form of ct_rewards_host: submit returns a ticket at once (the batch runs on a
 * stream of its own, so consecutive batches overlap on the GPU); wait blocks, copies the results
 * into host buffers (each nullable) and releases the ticket.  Used by the pipelined MCTS
 * scheduler: leaves of batch k+1 are selected (under virtual loss, reference circuitree.py:532-535)
 * while batch k is still simulating.
 * Lifetimes: handles / n_traj / traj_base / job_cost are consumed before submit returns.  h_params
 * (nullable) is copied with cudaMemcpyAsync: pageable memory has been staged when submit returns, but
 * PAGE-LOCKED (pinned) memory is read asynchronously and must stay valid and unmodified until
 * ct_rewards_wait has returned for this ticket.  A ticket nobody waits for is released by
 * ct_engine_destroy.  If submit fails, nothing stays allocated.  lpt launches take at most 2^31-1
 * trajectories.
 */
int ct_rewards_submit(ct_engine *eng, const uint64_t *handles, const uint32_t *n_traj,
                      const uint64_t *traj_base, const float *job_cost, int32_t n_jobs,
                      const ct_sim_config *cfg, uint64_t seed, const float *h_params, int64_t *ticket);
int ct_rewards_wait(ct_engine *eng, int64_t ticket, double *h_job_mean, uint64_t *h_job_events,
                    float *h_reward, int32_t *h_lag);

/* Test hook: reward and lag of companded series uint8 [n][n_species][n_dec] (host buffers). */
int ct_reward_from_series_host(ct_engine *eng, const uint8_t *h_series, int32_t n, int32_t n_species,
                               int32_t n_dec, float *h_reward, int32_t *h_lag);

/* Kernel launches issued by this engine since creation (bench.py's gpu_launches). */
int64_t ct_engine_launch_count(ct_engine *eng);
/* Debug counter: hot-loop iterations (2 lock-step SSA steps x 32 lanes) of all warps since the last
 * call (first call enables it); lane utilisation = events / (64 * iterations). */
int64_t ct_engine_warp_iterations(ct_engine *eng);

/* ---------------------------------------------------------------------------------------------
 * Native host side of the batched MCTS scheduler (csrc/native_tree.cpp): bit-packed
 * SimpleNetworkGrammar states and the tree statistics.  Each entry cites what it replaces:
 *   get_actions / do_action / get_unique_state   reference circuitree/models.py:164-270, :337-400
 *   select_and_expand + rollout + visit backprop reference circuitree/circuitree.py:315-384, :218-246, :468-482
 *   backpropagate_reward                         reference circuitree/circuitree.py:484-499
 *   mark_as_exhausted                            reference circuitree/circuitree.py:386-420
 *   grow_tree                                    reference circuitree/circuitree.py:539-586
 * The random stream is numpy's PCG64 (state handed over by the caller), consumed exactly like
 * Generator.shuffle / Generator.choice in the reference, so trees are identical for identical rewards.
 * Scope: root holds every component (sorted, e.g. "ABC::"), <= 5 components, <= 3 interaction types.
 * ------------------------------------------------------------------------------------------- */
typedef struct ct_tree ct_tree;

const char *ct_tree_last_error(void);
/* components / types: one character per component / interaction type, in declaration order.
 * max_interactions < 0 means n*n; n_exhausted = NaN means "never exhaust" (Python None). */
int ct_tree_create(const char *components, const char *types, int32_t max_interactions, const char *root,
                   double exploration_constant, double n_exhausted, ct_tree **out);
/* Same tree over DimersGrammar states (reference circuitree/models.py:524-939: get_actions :646-739, do_action
 * :741-757, get_unique_state :759-826, parse_genotype :844-884), e.g. root "ABC+DE::"; components / regulators /
 * types one character each in declaration order; at most 5 TFs in total.  ct_tree_key_to_handle registers the
 * state with ct_compile_topology(k = 3).  Their node keys are process-local ids of interned interaction lists; the replica exchange below
 * sends the lists themselves. */
int ct_tree_create_dimers(const char *components, const char *regulators, const char *types, int32_t max_interactions,
                          int32_t max_per_promoter, const char *root, double exploration_constant, double n_exhausted,
                          ct_tree **out);
int ct_tree_destroy(ct_tree *t);
/* numpy Generator(PCG64).bit_generator.state: 128-bit state and inc as (lo, hi) words. */
int ct_tree_set_rng(ct_tree *t, const uint64_t state[2], const uint64_t inc[2], int32_t has_uint32, uint32_t uinteger);
int ct_tree_get_rng(ct_tree *t, uint64_t state[2], uint64_t inc[2], int32_t *has_uint32, uint32_t *uinteger);
/* table[v] = log(v) as the caller's numpy computes it (keeps UCB scores bit-identical to numpy). */
int ct_tree_set_log_table(ct_tree *t, const double *table, int64_t n);
int ct_tree_rng_draw(ct_tree *t, int32_t kind, int64_t n, int64_t *out);
/* k MCTS selections under virtual loss; returns 100 with n_done < k when the tree is exhausted. */
int ct_tree_select(ct_tree *t, int32_t k, uint64_t *sim_keys, uint64_t *sim_handles, int32_t *n_done);
int ct_tree_backprop(ct_tree *t, int32_t k, const double *rewards);
int ct_tree_pending(ct_tree *t, int32_t *n_pending, int32_t *lengths, int32_t *ids);
int ct_tree_size(ct_tree *t, int64_t *n_nodes, int64_t *n_edges);
int ct_tree_export(ct_tree *t, uint64_t *node_keys, int64_t *node_visits, double *node_reward, uint8_t *node_exhausted,
                   int32_t *edge_parent, int32_t *edge_child, int64_t *edge_visits, double *edge_reward);
int ct_tree_key_to_string(ct_tree *t, uint64_t key, char *buf, int32_t buflen);
int ct_tree_string_to_key(ct_tree *t, const char *s, int32_t canonical, uint64_t *key);
int ct_tree_actions(ct_tree *t, uint64_t key, int32_t *ids, int32_t cap, int32_t *n_out);
int ct_tree_do_action(ct_tree *t, uint64_t key, int32_t action, uint64_t *out);
int ct_tree_key_to_handle(ct_tree *t, uint64_t key, uint64_t *handle);
int ct_tree_grow(ct_tree *t);
/* Root-parallel merge across GPUs (one tree per GPU, periodic exchange of statistics; the reference's
 * only parallel search shares one graph between greenlets, circuitree.py:786-893).  pack writes what
 * this replica did since the previous pack -- sparse (state key, d visits, d reward) records for
 * nodes and edges plus the keys of nodes it flagged exhausted (circuitree.py:386-420) -- into one
 * buffer of 64-bit words (layout: csrc/native_tree.cpp); with buf == NULL or cap too small it only
 * reports the size needed.  max_depth >= 0 limits the exchange to states with at most that many
 * interactions.  merge adds another replica's buffer and replays its exhaustion events; a buffer of a tree
 * with another key width (another grammar) or with a malformed state is refused and changes nothing. */
int ct_tree_pack_deltas(ct_tree *t, int32_t max_depth, int64_t *buf, int64_t cap, int64_t *n_words);
int ct_tree_merge_packed(ct_tree *t, const int64_t *buf, int64_t n_words, int32_t *tree_exhausted);

#ifdef __cplusplus
}
#endif
#endif /* CIRCUITREE_B200_H */
