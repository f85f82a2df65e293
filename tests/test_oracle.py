"""The CPU checker itself: Philox known answers, determinism, and reward known answers.

The SSA model has no upstream implementation (SPEC.md, "parity unpinned"); what can be pinned is
the RNG (Random123 known-answer vectors) and the reward arithmetic (numpy restatement here)."""
import numpy as np
import pytest

from oracle import ssa_oracle as so

RANGES = [(-2.3, -1.7, 1), (-1.0, -0.4, 1), (-2.7, -2.1, 1), (-0.3, 0.3, 1), (0.0, 0.6, 1),
          (-3.5, -2.5, 1), (-1.3, -0.7, 1), (-1.9, -1.3, 1), (-2.0, -1.4, 1), (-2.7, -2.1, 1)]
REPRESSILATOR = ([], [(0, 1), (1, 2), (2, 0)])


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert so.philox([0, 0, 0, 0], [0, 0]).tolist() == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert so.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2).tolist() == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert so.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]).tolist() == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def numpy_reward(q):
    """SPEC.md section 7 in numpy."""
    best, best_lag = 0.0, 0
    n = q.shape[1]
    L = n // 2
    for s in range(q.shape[0]):
        y = q[s].astype(np.float64) ** 2
        z = y - y.mean()
        c0 = (z * z).sum()
        if c0 <= 0:
            continue
        r = np.array([(z[: n - l] * z[l:]).sum() / c0 for l in range(L + 1)])
        for l in range(1, L):
            if r[l] < r[l - 1] and r[l] < r[l + 1] and r[l] < best:
                best, best_lag = r[l], l
    return min(max(-best, 0.0), 1.0), best_lag


def test_reward_matches_numpy_restatement():
    rg = np.random.default_rng(0)
    for _ in range(50):
        q = rg.integers(0, 256, size=(3, 125), dtype=np.uint8)
        r, lag = so.reward_from_series(q)
        r2, lag2 = numpy_reward(q)
        assert abs(r - r2) < 1e-12 and lag == lag2


def test_reward_known_answers():
    k = np.arange(125)
    sine = np.round(100 + 60 * np.sin(2 * np.pi * k / 20)).astype(np.uint8)
    flat = np.full(125, 80, dtype=np.uint8)
    r, lag = so.reward_from_series(np.stack([flat, sine, flat]))
    assert lag == 10 and 0.85 < r <= 1.0          # half a period of 20 samples
    assert so.reward_from_series(np.stack([flat, flat, flat])) == (0.0, 0)
    ramp = np.arange(125, dtype=np.uint8)
    assert so.reward_from_series(ramp[None, :])[0] == 0.0     # monotone ACF: no interior minimum below 0 ... or none at all


def test_trajectories_depend_only_on_seed_and_id():
    act, inh = REPRESSILATOR
    a = so.run_pairwise(3, act, inh, RANGES, 24, seed=5, n_threads=1, want_series=True, want_params=True)
    b = so.run_pairwise(3, act, inh, RANGES, 24, seed=5, n_threads=4, want_series=True, want_params=True)
    for key in ("reward", "lag", "n_events", "series", "params"):
        assert np.array_equal(a[key], b[key]), key
    c = so.run_pairwise(3, act, inh, RANGES, 8, seed=5, traj_base=16, n_threads=2, want_series=True)
    assert np.array_equal(c["series"], a["series"][16:24])
    d = so.run_pairwise(3, act, inh, RANGES, 8, seed=6, n_threads=2)
    assert not np.array_equal(d["n_events"], a["n_events"][:8])


def test_sampled_parameters_respect_ranges():
    act, inh = REPRESSILATOR
    out = so.run_pairwise(3, act, inh, RANGES, 64, seed=1, nt=64, decim=16, n_threads=2, want_params=True)
    logp = np.log10(out["params"])
    for i, (lo, hi, _) in enumerate(RANGES):
        assert logp[:, i].min() >= lo - 1e-9 and logp[:, i].max() <= hi + 1e-9
    lin = [(v, v, 0) for v in [0.01, 0.2, 0.004, 1.0, 1.0, 0.001, 0.001, 0.05, 0.02, 0.004]]
    out = so.run_pairwise(3, act, inh, lin, 4, seed=1, nt=64, decim=16, want_params=True)
    assert np.allclose(out["params"], [r[0] for r in lin])


def test_explicit_parameters_and_dead_network():
    # all rates zero: nothing ever fires, the record repeats the initial state
    out = so.run_pairwise(2, [], [], RANGES, 3, seed=2, nt=64, decim=16, explicit_params=np.zeros((3, 10)),
                          want_series=True)
    assert out["n_events"].tolist() == [0, 0, 0]
    s = out["series"]
    assert (s == s[:, :, :1]).all()
    assert out["reward"].tolist() == [0.0, 0.0, 0.0]


def test_unregulated_gene_reaches_expected_level():
    # constitutive expression: <P> = km*kp/(gm*gp) = 0.5*0.05/(0.02*0.005) = 250
    p = [0.01, 0.2, 0.004, 0.5, 0.5, 0.001, 0.001, 0.05, 0.02, 0.005]
    out = so.run_pairwise(1, [], [], RANGES, 64, seed=3, explicit_params=np.tile(p, (64, 1)), n_threads=4, want_series=True)
    mean_late = (out["series"][:, 0, 60:].astype(float) ** 2 / 16).mean()
    assert 230 < mean_late < 270


def test_invalid_inputs():
    with pytest.raises(ValueError):
        so.run_pairwise(3, [(0, 5)], [], RANGES, 1)
    with pytest.raises(ValueError):
        so.run_pairwise(9, [], [], RANGES, 1)


# ---- dimer model (SPEC.md section 9) ---------------------------------------------------------------------
DIMER_RANGES = RANGES + [(-5.3, -4.7, 1), (-2.8, -2.2, 1)]
HOMODIMER_REPRESSILATOR = ([], [(0, 0, 1), (1, 1, 2), (2, 2, 0)])


def test_dimer_checker_is_deterministic_and_thread_independent():
    act, inh = HOMODIMER_REPRESSILATOR
    a = so.run_dimers(3, 3, act, inh, DIMER_RANGES, 24, seed=5, nt=320, want_series=True)
    b = so.run_dimers(3, 3, act, inh, DIMER_RANGES, 24, seed=5, nt=320, want_series=True, n_threads=4)
    for k in ("reward", "lag", "n_events", "series"):
        assert np.array_equal(a[k], b[k]), k
    c = so.run_dimers(3, 3, act, inh, DIMER_RANGES, 8, seed=5, traj_base=16, nt=320)
    assert np.array_equal(c["n_events"], a["n_events"][16:])


def test_dimer_model_without_interactions_is_constitutive_expression():
    """No interactions: no dimer species exist, every TF sits at km kp / (gm gp); and the pairwise
    and dimer models then describe the same process (same streams: identical records)."""
    p = np.tile([0.0, 0.2, 0.004, 0.5, 0.5, 0.5, 0.5, 0.05, 0.02, 0.005, 2e-5, 2e-3], (48, 1))
    d = so.run_dimers(4, 3, [], [], DIMER_RANGES, 48, seed=2, explicit_params=p, want_series=True)
    level = (d["series"][:, :, 60:].astype(float) ** 2 / 16).mean(axis=(0, 2))
    assert np.all((230 < level) & (level < 270)), level
    q = so.run_pairwise(4, [], [], RANGES, 48, seed=2, explicit_params=p[:, :10], want_series=True)
    assert np.array_equal(d["series"], q["series"]) and np.array_equal(d["n_events"], q["n_events"])


def test_dimer_pool_and_oscillation():
    p = np.tile([0.0, 0.2, 0.004, 0.5, 0.5, 0.5, 0.5, 0.05, 0.02, 0.005, 2e-5, 2e-3], (48, 1))
    d = so.run_dimers(3, 3, [], [(0, 0, 1)], DIMER_RANGES, 48, seed=3, explicit_params=p, want_series=True, n_threads=4)
    level = (d["series"][:, :, 60:].astype(float) ** 2 / 16).mean(axis=(0, 2))
    # k_on = 0 and free dimers do not decay (SPEC.md section 9): the AA pool is in detailed balance with the
    # monomers, so every mean monomer level stays at 250; the pool only adds (un)dimerisation events
    assert np.all((230 < level) & (level < 270)), level
    base = so.run_dimers(3, 3, [], [], DIMER_RANGES, 48, seed=3, explicit_params=p)
    assert d["n_events"].mean() > 1.08 * base["n_events"].mean()
    act, inh = HOMODIMER_REPRESSILATOR
    osc = so.run_dimers(3, 3, act, inh, DIMER_RANGES, 64, seed=4, n_threads=8)
    flat = so.run_dimers(3, 3, [], [], DIMER_RANGES, 64, seed=4, n_threads=8)
    assert osc["reward"].mean() > 0.6 and flat["reward"].mean() < 0.3


def test_dimer_checker_rejects_bad_topologies():
    with pytest.raises(ValueError):
        so.run_dimers(3, 3, [(1, 0, 2)], [], DIMER_RANGES, 1)      # monomers not sorted
    with pytest.raises(ValueError):
        so.run_dimers(3, 2, [(0, 1, 2)], [], DIMER_RANGES, 1)      # target is not a component


def test_recorded_dimer_results_are_what_the_checker_computes():
    """tests/golden/dimer_oracle.npz (used by the GPU parity tests) against a fresh run of a slice."""
    import importlib.util
    from pathlib import Path

    here = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("make_dimer_oracle", here / "make_dimer_oracle.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gold = np.load(here / "dimer_oracle.npz")
    for name, case in mod.CASES.items():
        assert len(gold[name + "/reward"]) == case[2]
        r = mod.run_case(name, n=12, n_threads=4)
        assert np.array_equal(r["reward"], gold[name + "/reward"][:12]), name
        assert np.array_equal(r["lag"], gold[name + "/lag"][:12]), name
        assert np.array_equal(r["n_events"], gold[name + "/n_events"][:12]), name


def test_recorded_sweep_results_are_what_the_checker_computes():
    """tests/golden/sweep_oracle.npz (the randomised GPU parity sweep) against a fresh run of a slice."""
    import importlib.util
    from pathlib import Path

    here = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("make_sweep_oracle", here / "make_sweep_oracle.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gold = np.load(here / "sweep_oracle.npz")
    states = mod.states()
    assert [str(s) for s in gold["states"]] == states and len(states) == 200
    assert sum(s.startswith("*ABCD::") for s in states) == 100
    for i in (0, 57, 100, 133, 199):
        r = mod.run_state(i, states[i], n=mod.N_RAW, n_threads=4)
        assert np.array_equal(r["reward"], gold["raw_reward"][i]), states[i]
        assert np.array_equal(r["lag"], gold["raw_lag"][i]), states[i]
        assert np.array_equal(r["n_events"], gold["raw_events"][i]), states[i]
    assert (gold["reward_var"] > 0).all() and (gold["events_mean"] > 1e3).all()


def test_unregulated_gene_matches_birth_death_theory():
    """An external anchor for the unpinned model: one unregulated gene is the textbook two-stage
    birth-death process -- stationary means R = km/gm, P = km kp/(gm gp), and 2 km (1 + kp/gm) events
    per second (each of transcription, mRNA decay, translation, protein decay at its stationary flux)."""
    km, kp, gm, gp = 0.5, 0.05, 0.02, 0.005
    p = np.tile([0.01, 0.2, 0.004, km, km, km, km, kp, gm, gp], (96, 1))
    o = so.run_pairwise(1, [], [], RANGES, 96, seed=8, explicit_params=p, want_series=True, n_threads=so.max_threads())
    T = 2000 * 20.0
    relax = 5.0 / gp                                    # the protein level needs ~5/gp = 1000 s to settle
    level = (o["series"][:, 0, 60:].astype(float) ** 2 / 16).mean()
    assert abs(level - km * kp / (gm * gp)) / (km * kp / (gm * gp)) < 0.03
    rate = 2.0 * km * (1.0 + kp / gm)
    expected = rate * (T - relax) + rate * relax * 0.5  # ramp-up from an (almost) empty cell
    assert abs(o["n_events"].mean() - expected) / expected < 0.02


# ---- external anchor that exercises promoter binding: exact moments of promoter-switching models ----
def switching_moments(Q, k, gamma):
    """Exact stationary mean and variance of a birth-death species whose production rate k[s]
    follows a finite Markov chain s(t) with generator Q (rows sum to 0), degraded at rate gamma.
    Standard moment equations of the master equation: with pi the stationary law of s and
    m_s = E[R 1{s}],  0 = Q^T m + k*pi - gamma m,  E[R(R-1)] = sum_s k_s m_s / gamma.
    For two states this is the telegraph model of Peccoud & Ycart (1995)."""
    Q = np.asarray(Q, dtype=float); k = np.asarray(k, dtype=float)
    n = len(k)
    A = np.vstack([Q.T, np.ones(n)])
    pi = np.linalg.lstsq(A, np.concatenate([np.zeros(n), [1.0]]), rcond=None)[0]
    m = np.linalg.solve(gamma * np.eye(n) - Q.T, k * pi)
    mean = m.sum()
    var = (k * m).sum() / gamma + mean - mean ** 2
    return mean, var


def test_switching_moments_reduce_to_the_telegraph_closed_form():
    a, b, k0, k1, g = 0.03, 0.2, 1.0, 0.05, 0.02          # unbound -> bound at a, back at b
    mean, var = switching_moments([[-a, a], [b, -b]], [k0, k1], g)
    p_u = b / (a + b)
    mean_cf = (k0 * p_u + k1 * (1 - p_u)) / g
    var_cf = mean_cf + (k0 - k1) ** 2 * p_u * (1 - p_u) / (g * (g + a + b))
    assert abs(mean - mean_cf) < 1e-9 * mean_cf and abs(var - var_cf) < 1e-9 * var_cf


def _telegraph_run(n_species, act, inh, prm, n, seed):
    # k_p = gamma_p = 0 freezes every protein count at its Poisson(1) initial value except for
    # (un)binding, so a regulator that starts with ONE molecule toggles one promoter site:
    # binding 0 -> 1 at k_on * 1 * (2 - 0), no second binding (no free molecule left), unbinding at k_off1.
    params = np.tile(np.asarray(prm, dtype=np.float64), (n, 1))
    out = so.run_pairwise(n_species, act, inh, RANGES, n, seed=seed, explicit_params=params, nt=500, decim=20, dt=20.0,
                          init_mean=1.0, n_threads=so.max_threads(), want_final=True)
    f = out["final"]
    return f[:, 0, :], f[:, 1, :] + f[:, 2, :]           # final mRNA counts, conserved protein totals


def _check_moments(r, mean, var, label):
    n = len(r)
    assert n > 3000, (label, n)
    se_mean = np.sqrt(var / n)
    assert abs(r.mean() - mean) < 4.0 * se_mean, (label, r.mean(), mean)
    # the sample variance of n values has standard error ~ var * sqrt((kurt - 1) / n); these
    # over-dispersed counts stay below kurtosis 4.5, so 4 s.e. is within 7.5 % at n > 10^4 / 4
    assert abs(r.var(ddof=1) - var) < 4.0 * var * np.sqrt(3.5 / n), (label, r.var(ddof=1), var)
    assert var / mean > 1.5, "the test must be sensitive to promoter noise (Fano factor well above Poisson)"


def test_promoter_binding_matches_exact_telegraph_moments():
    """Pins the checker's binding / unbinding / regulated-transcription arithmetic to an external
    result (SPEC.md section 2): mean and Fano factor of the target's mRNA for an inhibitor and for an
    activator with a single molecule (two-state telegraph), and for one activator plus one inhibitor
    molecule competing for the two sites (four promoter states, all four transcription rates)."""
    k_on, k_off1, k_off2 = 0.004, 0.012, 0.005
    km0, km_a, km_r, km_ar = 0.4, 1.5, 0.02, 0.15
    gm = 0.02
    prm = [k_on, k_off1, k_off2, km0, km_a, km_r, km_ar, 0.0, gm, 0.0]
    a, b = 2 * k_on, k_off1
    # A inhibits B
    r, tot = _telegraph_run(2, [], [(0, 1)], prm, 25_000, seed=71)
    mean, var = switching_moments([[-a, a], [b, -b]], [km0, km_r], gm)
    _check_moments(r[tot[:, 0] == 1, 1], mean, var, "inhibitor")
    # the unregulated gene A is a plain birth-death process: Poisson with mean km0 / gm
    assert abs(r[:, 0].mean() - km0 / gm) < 4 * np.sqrt(km0 / gm / len(r))
    assert abs(r[:, 0].var(ddof=1) / r[:, 0].mean() - 1.0) < 0.05
    # A activates B
    r, tot = _telegraph_run(2, [(0, 1)], [], prm, 25_000, seed=72)
    mean, var = switching_moments([[-a, a], [b, -b]], [km0, km_a], gm)
    _check_moments(r[tot[:, 0] == 1, 1], mean, var, "activator")
    # A activates C, B inhibits C, one molecule each: promoter states (bA, bB) = 00, 10, 01, 11
    r, tot = _telegraph_run(3, [(0, 2)], [(1, 2)], prm, 80_000, seed=73)
    Q = np.zeros((4, 4))
    Q[0, 1] = 2 * k_on; Q[0, 2] = 2 * k_on          # first molecule: two free sites
    Q[1, 3] = k_on; Q[2, 3] = k_on                  # second molecule: one free site
    Q[1, 0] = k_off1; Q[2, 0] = k_off1              # n = 1: k_off1 per bound molecule
    Q[3, 1] = k_off2; Q[3, 2] = k_off2              # n = 2: k_off2 per bound molecule (B leaves -> 10, A leaves -> 01)
    Q -= np.diag(Q.sum(axis=1))
    mean, var = switching_moments(Q, [km0, km_a, km_r, km_ar], gm)
    sel = (tot[:, 0] == 1) & (tot[:, 1] == 1)
    _check_moments(r[sel, 2], mean, var, "activator + inhibitor")
