/*
 * oracle/ssa_oracle.c -- CPU restatement (plain C, double precision) of the
 * stochastic reward path that sits behind CircuiTree.get_reward.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under circuitree_b200/ may import, link
 * or execute this file; it is the checker for tests/, __graft_entry__.smoke()
 * and the cpu_baseline / --impl reference legs of bench.py.
 *
 * PARITY UNPINNED.  The upstream package ships no simulator: get_reward is an
 * abstract method (reference circuitree/circuitree.py:135-152) and the only
 * thing the reference fixes for this path is the topology input produced by
 * SimpleNetworkGrammar.parse_genotype (reference circuitree/models.py:441-480:
 * activations / inhibitions as int[n,2] rows of (regulator, target) indices
 * into the components string).  Everything else here -- reaction set,
 * propensities, parameter ranges, time grid, reward -- follows this repo's own
 * SPEC.md section by section (cited below), not upstream code.
 *
 * Build: see oracle/Makefile (gcc -O2 -pthread -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define CT_MAX_SPECIES 8
#define CT_N_PARAMS 10
#define CT_N_PARAMS_DIMER 12
#define CT_MAX_IXN 10

/* ---- SPEC.md §3: Philox4x32-10 counter-based RNG ------------------------- */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                          uint32_t k0, uint32_t k1, uint32_t out[4])
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void ct_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

/* SPEC.md §3: 23-bit uniform strictly inside (0,1), exact in float32. */
static double u01(uint32_t w) { return ((double)(w >> 9) + 0.5) * (1.0 / 8388608.0); }

typedef struct {
    uint32_t k0, k1;      /* key  = seed */
    uint32_t t0, t1;      /* trajectory id */
    uint32_t stream;
    uint32_t ctr;         /* next block index */
    uint32_t buf[4];
    int have;             /* words left in buf */
} rng_t;

static void rng_init(rng_t *g, uint64_t seed, uint64_t traj, uint32_t stream)
{
    g->k0 = (uint32_t)seed; g->k1 = (uint32_t)(seed >> 32);
    g->t0 = (uint32_t)traj; g->t1 = (uint32_t)(traj >> 32);
    g->stream = stream; g->ctr = 0; g->have = 0;
}

static uint32_t rng_next(rng_t *g)
{
    if (g->have == 0) {
        philox4x32_10(g->ctr, g->stream, g->t0, g->t1, g->k0, g->k1, g->buf);
        g->ctr++; g->have = 4;
    }
    uint32_t w = g->buf[4 - g->have];
    g->have--;
    return w;
}

/* ---- SPEC.md §4: parameter vector from (lo, hi, is_log) ranges ----------- */
/* params: 0 k_on, 1 k_off1, 2 k_off2, 3 km_unbound, 4 km_act, 5 km_rep,
 *         6 km_act_rep, 7 kp, 8 gamma_m, 9 gamma_p                         */
static void sample_params_n(rng_t *g, const double *ranges /* [n][3] */, double *p, int n)
{
    for (int i = 0; i < n; ++i) {
        double u = u01(rng_next(g));
        double lo = ranges[3 * i], hi = ranges[3 * i + 1];
        double v = lo + u * (hi - lo);
        p[i] = ranges[3 * i + 2] != 0.0 ? pow(10.0, v) : v;
    }
    for (int i = n; i < 12; ++i) (void)rng_next(g);   /* the setup block ends at word 12 */
}

static void sample_params(rng_t *g, const double *ranges /* [10][3] */, double *p)
{
    for (int i = 0; i < CT_N_PARAMS; ++i) {
        double u = u01(rng_next(g));
        double lo = ranges[3 * i], hi = ranges[3 * i + 1];
        double v = lo + u * (hi - lo);
        p[i] = ranges[3 * i + 2] != 0.0 ? pow(10.0, v) : v;
    }
    /* words 10,11 of the setup stream are skipped so that the initial
     * condition starts on a block boundary (SPEC.md §3).                    */
    (void)rng_next(g); (void)rng_next(g);
}

/* SPEC.md §5: Poisson(lambda) by sequential inversion from one uniform. */
static int poisson_inv(double u, double lambda)
{
    double p = exp(-lambda), F = p;
    int k = 0;
    while (u > F && k < 1000) { ++k; p *= lambda / k; F += p; }
    return k;
}

/* ---- SPEC.md §7: reward from the companded, decimated series ------------ */
/* q: [n_species][n_dec] bytes.  Returns reward in [0,1]; *lag_out = lag (in
 * decimated samples) of the lowest interior ACF minimum over species, 0 if
 * none exists.                                                              */
double ct_oracle_reward_from_series(const uint8_t *q, int n_species, int n_dec, int *lag_out)
{
    int max_lag = n_dec / 2;
    double best = 0.0; int best_lag = 0;
    double *y = (double *)malloc(sizeof(double) * n_dec);
    double *r = (double *)malloc(sizeof(double) * (max_lag + 1));
    for (int s = 0; s < n_species; ++s) {
        double mu = 0.0;
        for (int k = 0; k < n_dec; ++k) { double v = q[s * n_dec + k]; y[k] = v * v; mu += y[k]; }
        mu /= n_dec;
        double c0 = 0.0;
        for (int k = 0; k < n_dec; ++k) { y[k] -= mu; c0 += y[k] * y[k]; }
        if (c0 <= 0.0) continue;
        for (int l = 0; l <= max_lag; ++l) {
            double acc = 0.0;
            for (int k = 0; k + l < n_dec; ++k) acc += y[k] * y[k + l];
            r[l] = acc / c0;
        }
        for (int l = 1; l < max_lag; ++l) {
            if (r[l] < r[l - 1] && r[l] < r[l + 1] && r[l] < best) { best = r[l]; best_lag = l; }
        }
    }
    free(y); free(r);
    if (lag_out) *lag_out = best_lag;
    double reward = -best;
    if (reward > 1.0) reward = 1.0;
    if (reward < 0.0) reward = 0.0;
    return reward;
}

/* SPEC.md §6: 8-bit square-root companding of a block sum. */
static uint8_t compand(double block_sum)
{
    double s = floor(sqrt(block_sum) + 0.5);
    if (s > 255.0) s = 255.0;
    return (uint8_t)s;
}

/* ---- SPEC.md §2,§5,§6: one trajectory of the pairwise TF-network model --- */
typedef struct {
    int n_species;
    int nt;              /* grid points            */
    int decim;           /* grid points per block  */
    double dt;
    double init_mean;
    double sample_cap;   /* saturating clamp on a grid sample */
} sim_cfg_t;

static uint64_t simulate_pairwise(const uint8_t etype[CT_MAX_SPECIES][CT_MAX_SPECIES],
                                  const sim_cfg_t *cfg, const double *prm,
                                  uint64_t seed, uint64_t traj, rng_t *setup,
                                  uint8_t *q /* [M][n_dec] */, double *final_state /* nullable [3][M] */)
{
    const int M = cfg->n_species;
    const int n_dec = cfg->nt / cfg->decim;
    const double k_on = prm[0], k_off1 = prm[1], k_off2 = prm[2];
    const double km[4] = { prm[3], prm[4], prm[5], prm[6] }; /* idx = act | rep<<1 */
    const double kp = prm[7], gm = prm[8], gp = prm[9];

    double P[CT_MAX_SPECIES], R[CT_MAX_SPECIES];
    int b[CT_MAX_SPECIES][CT_MAX_SPECIES];
    memset(b, 0, sizeof b);
    for (int j = 0; j < M; ++j) {
        R[j] = 0.0;
        P[j] = (double)poisson_inv(u01(rng_next(setup)), cfg->init_mean);
    }

    rng_t g; rng_init(&g, seed, traj, 1u);
    double block[CT_MAX_SPECIES] = {0};
    int k = 0;                   /* next grid point to record */
    double t = 0.0;
    uint64_t n_events = 0;
    double a[4 * CT_MAX_SPECIES + 2 * CT_MAX_SPECIES * CT_MAX_SPECIES];

    for (;;) {
        /* propensities, SPEC.md §2 order: per target gene j:
         * tx, mRNA decay, translation, protein decay, binding of each regulator i, unbinding of each i */
        int nr = 0; double a0 = 0.0;
        for (int j = 0; j < M; ++j) {
            int n = 0, na = 0, nrp = 0;
            for (int i = 0; i < M; ++i) {
                n += b[i][j];
                if (etype[i][j] == 1) na += b[i][j];
                if (etype[i][j] == 2) nrp += b[i][j];
            }
            a[nr++] = km[(na > 0) | ((nrp > 0) << 1)];
            a[nr++] = gm * R[j];
            a[nr++] = kp * R[j];
            a[nr++] = gp * P[j];
            double koff = (n >= 2) ? k_off2 : k_off1;
            for (int i = 0; i < M; ++i)
                if (etype[i][j]) a[nr++] = k_on * P[i] * (double)(2 - n);
            for (int i = 0; i < M; ++i)
                if (etype[i][j]) a[nr++] = koff * (double)b[i][j];
        }
        for (int r = 0; r < nr; ++r) a0 += a[r];

        double u1 = u01(rng_next(&g));
        double u2 = u01(rng_next(&g));
        double t_new = (a0 > 0.0) ? t - log(u1) / a0 : INFINITY;

        /* SPEC.md §6: grid sample k holds the state at time k*dt */
        while (k < cfg->nt && (double)k * cfg->dt < t_new) {
            for (int j = 0; j < M; ++j)
                block[j] += (P[j] > cfg->sample_cap) ? cfg->sample_cap : P[j];
            ++k;
            if (k % cfg->decim == 0) {
                int kd = k / cfg->decim - 1;
                if (kd < n_dec)
                    for (int j = 0; j < M; ++j) { q[j * n_dec + kd] = compand(block[j]); block[j] = 0.0; }
            }
        }
        if (k >= cfg->nt) break;
        t = t_new;

        /* select: first reaction r with target < cumulative sum */
        double target = u2 * a0, c = 0.0;
        int sel = -1;   /* rounding can leave none: the step then fires nothing (SPEC.md section 2) */
        for (int r = 0; r < nr; ++r) { c += a[r]; if (target < c) { sel = r; break; } }

        /* apply */
        nr = 0;
        for (int j = 0; j < M && sel >= 0; ++j) {
            if (sel == nr++) { R[j] += 1.0; sel = -1; break; }
            if (sel == nr++) { R[j] -= 1.0; sel = -1; break; }
            if (sel == nr++) { P[j] += 1.0; sel = -1; break; }
            if (sel == nr++) { P[j] -= 1.0; sel = -1; break; }
            for (int i = 0; i < M && sel >= 0; ++i)
                if (etype[i][j] && sel == nr++) { P[i] -= 1.0; b[i][j] += 1; sel = -1; }
            for (int i = 0; i < M && sel >= 0; ++i)
                if (etype[i][j] && sel == nr++) { P[i] += 1.0; b[i][j] -= 1; sel = -1; }
        }
        ++n_events;
    }
    if (final_state) {   /* state when the last grid sample was recorded: mRNA, free protein, molecules of j bound anywhere */
        for (int j = 0; j < M; ++j) {
            int bound = 0;
            for (int t2 = 0; t2 < M; ++t2) bound += b[j][t2];
            final_state[j] = R[j]; final_state[M + j] = P[j]; final_state[2 * M + j] = (double)bound;
        }
    }
    return n_events;
}

typedef struct {
    uint8_t etype[CT_MAX_SPECIES][CT_MAX_SPECIES];
    sim_cfg_t cfg;
    const double *ranges, *explicit_params;
    uint64_t seed, traj_base; int64_t n_traj;
    float *reward; int32_t *lag; uint64_t *n_events; uint8_t *series; double *params_out;
    double *final_state;
    atomic_llong next;
} job_t;

/* each worker thread pulls chunks of 4 trajectories off a shared counter */
static void *worker(void *arg)
{
    job_t *jb = (job_t *)arg;
    const int M = jb->cfg.n_species, n_dec = jb->cfg.nt / jb->cfg.decim;
    for (;;) {
        long long x0 = atomic_fetch_add(&jb->next, 4);
        if (x0 >= jb->n_traj) break;
        long long x1 = x0 + 4 < jb->n_traj ? x0 + 4 : jb->n_traj;
        for (long long x = x0; x < x1; ++x) {
            uint64_t traj = jb->traj_base + (uint64_t)x;
            rng_t setup; rng_init(&setup, jb->seed, traj, 0u);
            double prm[CT_N_PARAMS];
            sample_params(&setup, jb->ranges, prm);
            if (jb->explicit_params) memcpy(prm, jb->explicit_params + x * CT_N_PARAMS, sizeof prm);
            uint8_t qbuf[CT_MAX_SPECIES * 512];
            uint8_t *q = jb->series ? jb->series + (size_t)x * M * n_dec : qbuf;
            uint64_t ne = simulate_pairwise(jb->etype, &jb->cfg, prm, jb->seed, traj, &setup, q,
                                            jb->final_state ? jb->final_state + (size_t)x * 3 * M : NULL);
            int lg = 0;
            double rw = ct_oracle_reward_from_series(q, M, n_dec, &lg);
            if (jb->reward) jb->reward[x] = (float)rw;
            if (jb->lag) jb->lag[x] = lg;
            if (jb->n_events) jb->n_events[x] = ne;
            if (jb->params_out) memcpy(jb->params_out + x * CT_N_PARAMS, prm, sizeof prm);
        }
    }
    return NULL;
}

/* ---- public entry: a batch of trajectories on one topology --------------- */
/* act / inh: int32 [n][2] rows (regulator, target) exactly as parse_genotype
 * returns them.  ranges: [10][3] (lo, hi, is_log); explicit_params (nullable)
 * [n_traj][10] overrides sampling.  Outputs (each nullable): reward[n_traj],
 * lag[n_traj] (decimated samples), n_events[n_traj], series[n_traj][M][n_dec],
 * params_out[n_traj][10].  Returns 0 on success.                            */
int ct_oracle_run_pairwise_ex(int n_species,
                           const int32_t *act, int n_act,
                           const int32_t *inh, int n_inh,
                           const double *ranges, const double *explicit_params,
                           uint64_t seed, uint64_t traj_base, int64_t n_traj,
                           int nt, int decim, double dt, double init_mean,
                           int n_threads,
                           float *reward, int32_t *lag, uint64_t *n_events,
                           uint8_t *series, double *params_out, double *final_state);

int ct_oracle_run_pairwise(int n_species,
                           const int32_t *act, int n_act,
                           const int32_t *inh, int n_inh,
                           const double *ranges, const double *explicit_params,
                           uint64_t seed, uint64_t traj_base, int64_t n_traj,
                           int nt, int decim, double dt, double init_mean,
                           int n_threads,
                           float *reward, int32_t *lag, uint64_t *n_events,
                           uint8_t *series, double *params_out)
{
    return ct_oracle_run_pairwise_ex(n_species, act, n_act, inh, n_inh, ranges, explicit_params, seed, traj_base, n_traj,
                                     nt, decim, dt, init_mean, n_threads, reward, lag, n_events, series, params_out, NULL);
}

/* Same, plus final_state (nullable) [n_traj][3][M]: mRNA counts, free protein counts and the number
 * of molecules of each species bound at promoters when the last grid sample was recorded (used to
 * check the simulator against exact master-equation moments, tests/test_oracle.py). */
int ct_oracle_run_pairwise_ex(int n_species,
                           const int32_t *act, int n_act,
                           const int32_t *inh, int n_inh,
                           const double *ranges, const double *explicit_params,
                           uint64_t seed, uint64_t traj_base, int64_t n_traj,
                           int nt, int decim, double dt, double init_mean,
                           int n_threads,
                           float *reward, int32_t *lag, uint64_t *n_events,
                           uint8_t *series, double *params_out, double *final_state)
{
    if (n_species < 1 || n_species > CT_MAX_SPECIES || nt < decim || decim < 1 || nt / decim > 512) return 1;
    uint8_t etype[CT_MAX_SPECIES][CT_MAX_SPECIES];
    memset(etype, 0, sizeof etype);
    for (int e = 0; e < n_act; ++e) {
        int i = act[2 * e], j = act[2 * e + 1];
        if (i < 0 || j < 0 || i >= n_species || j >= n_species) return 2;
        etype[i][j] = 1;
    }
    for (int e = 0; e < n_inh; ++e) {
        int i = inh[2 * e], j = inh[2 * e + 1];
        if (i < 0 || j < 0 || i >= n_species || j >= n_species) return 2;
        etype[i][j] = 2;
    }
    job_t job;
    memcpy(job.etype, etype, sizeof etype);
    job.cfg = (sim_cfg_t){ n_species, nt, decim, dt, init_mean, 4095.0 };
    job.ranges = ranges; job.explicit_params = explicit_params;
    job.seed = seed; job.traj_base = traj_base; job.n_traj = n_traj;
    job.reward = reward; job.lag = lag; job.n_events = n_events;
    job.series = series; job.params_out = params_out; job.final_state = final_state;
    atomic_init(&job.next, 0);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if (n_threads == 1) { worker(&job); return 0; }
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < n_threads; ++i)
        if (pthread_create(&th[started], NULL, worker, &job) == 0) ++started;
    if (started == 0) worker(&job);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
    return 0;
}


/* ---- SPEC.md section 9: dimer model ------------------------------------------------------------ */
typedef struct {
    int n_tf, n_comp, n_ixn, n_dim;
    int dim_a[CT_MAX_IXN], dim_b[CT_MAX_IXN];
    int ix_dim[CT_MAX_IXN], ix_tgt[CT_MAX_IXN], ix_act[CT_MAX_IXN];
} dimer_topo_t;

static uint64_t simulate_dimers(const dimer_topo_t *T, const sim_cfg_t *cfg, const double *prm, uint64_t seed,
                                uint64_t traj, rng_t *setup, uint8_t *q)
{
    const int M = T->n_tf, n_dec = cfg->nt / cfg->decim;
    const double k_on = prm[0], k_off1 = prm[1], k_off2 = prm[2];
    const double km[4] = { prm[3], prm[4], prm[5], prm[6] };
    const double kp = prm[7], gm = prm[8], gp = prm[9], k_dim = prm[10], k_undim = prm[11];
    double P[CT_MAX_SPECIES], R[CT_MAX_SPECIES], D[CT_MAX_IXN];
    int b[CT_MAX_IXN];
    for (int s = 0; s < M; ++s) { R[s] = 0.0; P[s] = (double)poisson_inv(u01(rng_next(setup)), cfg->init_mean); }
    for (int d = 0; d < CT_MAX_IXN; ++d) { D[d] = 0.0; b[d] = 0; }
    rng_t g; rng_init(&g, seed, traj, 1u);
    double block[CT_MAX_SPECIES] = {0};
    int k = 0; double t = 0.0; uint64_t n_events = 0;
    double a[4 * CT_MAX_SPECIES + 4 * CT_MAX_IXN];
    for (;;) {
        int nr = 0; double a0 = 0.0;
        for (int s = 0; s < M; ++s) {
            int n = 0, na = 0, ni = 0;
            for (int e = 0; e < T->n_ixn; ++e) if (T->ix_tgt[e] == s) { n += b[e]; if (T->ix_act[e]) na += b[e]; else ni += b[e]; }
            a[nr++] = (s < T->n_comp) ? km[(na > 0) | ((ni > 0) << 1)] : km[0];
            a[nr++] = gm * R[s]; a[nr++] = kp * R[s]; a[nr++] = gp * P[s];
            if (s < T->n_comp) {
                double koff = (n >= 2) ? k_off2 : k_off1;
                for (int e = 0; e < T->n_ixn; ++e) if (T->ix_tgt[e] == s) a[nr++] = k_on * D[T->ix_dim[e]] * (double)(2 - n);
                for (int e = 0; e < T->n_ixn; ++e) if (T->ix_tgt[e] == s) a[nr++] = koff * (double)b[e];
            }
        }
        for (int d = 0; d < T->n_dim; ++d) {
            int x = T->dim_a[d], y = T->dim_b[d];
            a[nr++] = (x != y) ? k_dim * P[x] * P[y] : k_dim * P[x] * (P[x] - 1.0) * 0.5;
            a[nr++] = k_undim * D[d];
        }
        for (int r = 0; r < nr; ++r) a0 += a[r];
        double u1 = u01(rng_next(&g)), u2 = u01(rng_next(&g));
        double t_new = (a0 > 0.0) ? t - log(u1) / a0 : INFINITY;
        while (k < cfg->nt && (double)k * cfg->dt < t_new) {
            for (int s = 0; s < M; ++s) block[s] += (P[s] > cfg->sample_cap) ? cfg->sample_cap : P[s];
            ++k;
            if (k % cfg->decim == 0) {
                int kd = k / cfg->decim - 1;
                if (kd < n_dec) for (int s = 0; s < M; ++s) { q[s * n_dec + kd] = compand(block[s]); block[s] = 0.0; }
            }
        }
        if (k >= cfg->nt) break;
        t = t_new;
        double target = u2 * a0, c = 0.0; int sel = -1;
        for (int r = 0; r < nr; ++r) { c += a[r]; if (target < c) { sel = r; break; } }
        nr = 0;
        for (int s = 0; s < M && sel >= 0; ++s) {
            if (sel == nr++) { R[s] += 1.0; sel = -1; break; }
            if (sel == nr++) { R[s] -= 1.0; sel = -1; break; }
            if (sel == nr++) { P[s] += 1.0; sel = -1; break; }
            if (sel == nr++) { P[s] -= 1.0; sel = -1; break; }
            if (s < T->n_comp) {
                for (int e = 0; e < T->n_ixn && sel >= 0; ++e)
                    if (T->ix_tgt[e] == s && sel == nr++) { D[T->ix_dim[e]] -= 1.0; b[e] += 1; sel = -1; }
                for (int e = 0; e < T->n_ixn && sel >= 0; ++e)
                    if (T->ix_tgt[e] == s && sel == nr++) { D[T->ix_dim[e]] += 1.0; b[e] -= 1; sel = -1; }
            }
        }
        for (int d = 0; d < T->n_dim && sel >= 0; ++d) {
            int x = T->dim_a[d], y = T->dim_b[d];
            if (sel == nr++) { P[x] -= 1.0; P[y] -= 1.0; D[d] += 1.0; sel = -1; break; }
            if (sel == nr++) { P[x] += 1.0; P[y] += 1.0; D[d] -= 1.0; sel = -1; break; }
        }
        ++n_events;
    }
    return n_events;
}

/* act / inh: int32 [n][3] rows (monomer1, monomer2, target) as DimersGrammar.parse_genotype returns
 * them (reference circuitree/models.py:844-884); ranges [12][3].  Single-threaded. */
int ct_oracle_run_dimers(int n_tf, int n_comp, const int32_t *act, int n_act, const int32_t *inh, int n_inh,
                         const double *ranges, const double *explicit_params, uint64_t seed, uint64_t traj_base,
                         int64_t n_traj, int nt, int decim, double dt, double init_mean,
                         float *reward, int32_t *lag, uint64_t *n_events, uint8_t *series)
{
    if (n_tf < 1 || n_tf > 5 || n_comp < 1 || n_comp > n_tf || n_act + n_inh > CT_MAX_IXN || nt / decim > 512) return 1;
    dimer_topo_t T; memset(&T, 0, sizeof T);
    T.n_tf = n_tf; T.n_comp = n_comp;
    for (int pass = 0; pass < 2; ++pass) {
        const int32_t *x = pass ? inh : act; int n = pass ? n_inh : n_act;
        for (int e = 0; e < n; ++e) {
            int m1 = x[3 * e], m2 = x[3 * e + 1], tg = x[3 * e + 2];
            if (m1 < 0 || m2 < m1 || m2 >= n_tf || tg < 0 || tg >= n_comp) return 2;
            int d = -1;
            for (int y = 0; y < T.n_dim; ++y) if (T.dim_a[y] == m1 && T.dim_b[y] == m2) d = y;
            if (d < 0) { d = T.n_dim++; T.dim_a[d] = m1; T.dim_b[d] = m2; }
            T.ix_dim[T.n_ixn] = d; T.ix_tgt[T.n_ixn] = tg; T.ix_act[T.n_ixn] = pass ? 0 : 1; T.n_ixn++;
        }
    }
    sim_cfg_t cfg = { n_tf, nt, decim, dt, init_mean, 4095.0 };
    const int n_dec = nt / decim;
    for (int64_t x = 0; x < n_traj; ++x) {
        uint64_t traj = traj_base + (uint64_t)x;
        rng_t setup; rng_init(&setup, seed, traj, 0u);
        double prm[CT_N_PARAMS_DIMER];
        sample_params_n(&setup, ranges, prm, CT_N_PARAMS_DIMER);
        if (explicit_params) memcpy(prm, explicit_params + x * CT_N_PARAMS_DIMER, sizeof prm);
        uint8_t qbuf[CT_MAX_SPECIES * 512];
        uint8_t *q = series ? series + (size_t)x * n_tf * n_dec : qbuf;
        uint64_t ne = simulate_dimers(&T, &cfg, prm, seed, traj, &setup, q);
        int lg = 0;
        double rw = ct_oracle_reward_from_series(q, n_tf, n_dec, &lg);
        if (reward) reward[x] = (float)rw;
        if (lag) lag[x] = lg;
        if (n_events) n_events[x] = ne;
    }
    return 0;
}

int ct_oracle_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
