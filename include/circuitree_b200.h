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

#define CT_N_PARAMS 10       /* SPEC.md section 4 */
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
    float ranges[CT_N_PARAMS][3]; /* (lo, hi, is_log) per parameter, SPEC.md section 4 */
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
 * k = 3 (DimersGrammar, models.py:844-884) returns CT_ERR_UNSUPPORTED in this round.
 * The handle is a self-contained 64-bit code; it needs no release.
 */
int ct_compile_topology(int32_t n_species, const int64_t *act, int32_t n_act,
                        const int64_t *inh, int32_t n_inh, int32_t k, uint64_t *handle);

/*
 * Device-buffer launch.  Job j simulates n_traj[j] trajectories of topology handles[j]; its
 * trajectory ids are traj_base[j] .. traj_base[j]+n_traj[j]-1 and its outputs occupy the slots
 * after those of jobs 0..j-1.  All d_* pointers are device pointers on the engine's device:
 *   d_params  (nullable) float32 [total][CT_N_PARAMS] explicit parameters, else sampled in-kernel
 *   d_reward  float32 [total]     d_lag  int32 [total] (nullable)   d_events uint32 [total] (nullable)
 *   d_series  (nullable, debug) uint8 [total][n_species_max][nt/decim] companded block sums
 * `stream` is a cudaStream_t (0 = default stream).  Asynchronous.
 */
int ct_launch_rewards(ct_engine *eng, const uint64_t *handles, const uint32_t *n_traj,
                      const uint64_t *traj_base, int32_t n_jobs, const ct_sim_config *cfg,
                      uint64_t seed, void *stream, const float *d_params, float *d_reward,
                      int32_t *d_lag, uint32_t *d_events, uint8_t *d_series, int32_t series_stride);

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
 * back.  Synchronous.
 */
int ct_rewards_host(ct_engine *eng, const uint64_t *handles, const uint32_t *n_traj,
                    const uint64_t *traj_base, int32_t n_jobs, const ct_sim_config *cfg,
                    uint64_t seed, const float *h_params, double *h_job_mean,
                    uint64_t *h_job_events, float *h_reward, int32_t *h_lag);

/* Test hook: reward and lag of companded series uint8 [n][n_species][n_dec] (host buffers). */
int ct_reward_from_series_host(ct_engine *eng, const uint8_t *h_series, int32_t n, int32_t n_species,
                               int32_t n_dec, float *h_reward, int32_t *h_lag);

/* Kernel launches issued by this engine since creation (bench.py's gpu_launches). */
int64_t ct_engine_launch_count(ct_engine *eng);

#ifdef __cplusplus
}
#endif
#endif /* CIRCUITREE_B200_H */
