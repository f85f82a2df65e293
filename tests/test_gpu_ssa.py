"""GPU parity tests (run with -m gpu on a B200): the CUDA path through the C ABI vs the CPU checker.

Bars (BASELINE.json north_star): reward arithmetic equal to float32 rounding on identical series;
initial conditions identical; per-topology mean reward within 3 standard errors over 10^4
trajectories; two-sample KS on the period (lag) distribution p > 0.01; plus size-independent
properties (determinism, independence of launch shape and scheduling) at large batch sizes.
The SSA model itself is this repo's SPEC.md -- upstream has none ("parity unpinned").
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from scipy import stats

from circuitree_b200 import OscillationTree, RewardEngine, SimpleNetworkGrammar, _capi
from oracle import ssa_oracle as so

AI = ["activates", "inhibits"]
THREADS = so.max_threads()


@pytest.fixture(scope="module")
def eng():
    return _capi.Engine(0)


@pytest.fixture(scope="module")
def cfg():
    return _capi.default_config()


def small_cfg(nt=320, decim=16):
    c = _capi.default_config()
    c.nt, c.decim = nt, decim
    return c


def oracle_run(state, n, cfg, seed, **kw):
    comps, act, inh = SimpleNetworkGrammar.parse_genotype(state)
    return so.run_pairwise(len(comps), act, inh, cfg.ranges_array(), n, seed=seed, nt=cfg.nt, decim=cfg.decim,
                           dt=cfg.dt, init_mean=cfg.init_mean, n_threads=THREADS, **kw)


def handle_of(state):
    comps, act, inh = SimpleNetworkGrammar.parse_genotype(state)
    return _capi.compile_topology(len(comps), act, inh)


def gpu_run(eng, states, n, cfg, seed, bases=None, series=False, params=None, job_cost=None):
    handles = [handle_of(s) for s in states]
    counts = [n] * len(states) if np.isscalar(n) else list(n)
    if bases is None:
        bases = np.concatenate(([0], np.cumsum(counts)[:-1]))
    total = int(np.sum(counts))
    nsp = max((h >> 56) & 15 for h in handles)
    d_reward = torch.zeros(total, device="cuda")
    d_lag = torch.zeros(total, dtype=torch.int32, device="cuda")
    d_events = torch.zeros(total, dtype=torch.int32, device="cuda")
    stride = nsp * cfg.n_dec
    d_series = torch.zeros((total, stride), dtype=torch.uint8, device="cuda") if series else None
    d_params = torch.as_tensor(np.asarray(params, dtype=np.float32), device="cuda") if params is not None else None
    eng.launch_rewards(handles, counts, bases, cfg, seed, 0, d_reward.data_ptr(), d_lag=d_lag.data_ptr(),
                       d_events=d_events.data_ptr(), d_params=d_params.data_ptr() if d_params is not None else 0,
                       d_series=d_series.data_ptr() if series else 0, series_stride=stride, job_cost=job_cost)
    torch.cuda.synchronize()
    out = {"reward": d_reward.cpu().numpy(), "lag": d_lag.cpu().numpy(), "events": d_events.cpu().numpy().astype(np.int64)}
    if series:
        out["series"] = d_series.cpu().numpy().reshape(total, nsp, cfg.n_dec)
    return out


# ---- reward stage -------------------------------------------------------------------------------------
def test_reward_stage_matches_checker_on_random_series(eng):
    rg = np.random.default_rng(0)
    for nsp, n_dec in ((3, 125), (1, 125), (5, 125), (3, 20), (2, 126), (4, 3)):
        q = rg.integers(0, 256, size=(300, nsp, n_dec), dtype=np.uint8)
        q[0] = 7                                   # constant series: c0 = 0 -> reward 0
        q[1, :, :] = np.arange(n_dec)[None, :]     # ramp
        rw, lg = eng.reward_from_series(q)
        ref = np.array([so.reward_from_series(x) for x in q])
        assert np.abs(rw - ref[:, 0]).max() < 2e-6, (nsp, n_dec)
        assert (lg == ref[:, 1]).mean() > 0.995, (nsp, n_dec)   # float32 vs float64 ties on near-equal minima
        assert rw[0] == 0.0 and lg[0] == 0


def test_reward_stage_on_simulated_series(eng, cfg):
    o = oracle_run("*ABC::ABi_BCi_CAi", 1500, cfg, seed=11, want_series=True)
    rw, lg = eng.reward_from_series(o["series"])
    assert np.abs(rw - o["reward"]).max() < 2e-6
    assert (lg == o["lag"]).mean() > 0.998


# ---- initial conditions and sampling share the RNG streams ---------------------------------------------
def test_first_block_and_dead_network_are_bit_identical(eng):
    c = small_cfg(nt=64, decim=1)      # every grid sample stored: sample 0 is the Poisson initial condition
    g = gpu_run(eng, ["*ABC::ABi_BCi_CAi"], 512, c, seed=3, series=True)
    o = oracle_run("*ABC::ABi_BCi_CAi", 512, c, seed=3, want_series=True)
    assert np.array_equal(g["series"][:, :, 0], o["series"][:, :, 0])
    # zero rates: no event ever fires; the whole record is the initial state on both sides
    z = np.zeros((64, 10), dtype=np.float32)
    g = gpu_run(eng, ["*AB::ABa"], 64, c, seed=4, series=True, params=z)
    o = so.run_pairwise(2, [(0, 1)], [], c.ranges_array(), 64, seed=4, nt=c.nt, decim=c.decim, dt=c.dt,
                        init_mean=c.init_mean, explicit_params=z.astype(np.float64), want_series=True)
    assert np.array_equal(g["series"], o["series"])
    assert g["events"].sum() == 0 and np.all(g["reward"] == 0)


def test_constitutive_expression_level(eng, cfg):
    p = np.tile(np.array([0.01, 0.2, 0.004, 0.5, 0.5, 0.001, 0.001, 0.05, 0.02, 0.005], np.float32), (256, 1))
    g = gpu_run(eng, ["*A::"], 256, cfg, seed=5, series=True, params=p)
    level = (g["series"][:, 0, 60:].astype(float) ** 2 / 16).mean()
    assert 235 < level < 265          # km*kp/(gm*gp) = 250
    # textbook birth-death fluxes: 2 km (1 + kp/gm) = 3.5 events/s once settled (~5/gp = 1000 s of ramp-up)
    expected = 3.5 * (2000 * 20.0 - 1000.0) + 3.5 * 1000.0 * 0.5
    assert abs(g["events"].mean() - expected) / expected < 0.02


# ---- statistical equivalence, 10^4 trajectories per topology -------------------------------------------
TOPOLOGIES = [
    "*ABC::ABi_BCi_CAi",                 # repressilator
    "*AB::ABa_BAi",                      # two-component negative feedback
    "*ABC::AAi_ABi_ACi_BAi_BBi_BCi_CAi_CBi_CCi",   # fully repressed (cheap)
    "*ABC::ABa_BCi_CAa",
    "*ABCD::ABi_BCi_CDi_DAi",
    "*ABCDE::ABi_BCi_CDi_DEi_EAi",
    # promoters with an activator AND an inhibitor, and with two different regulators bound: the
    # km_act_rep rate, the A_j bookkeeping and the mixed-regulator branch of the resolution block
    "*ABC::AAa_ABi_BBa_BCi_CAi_CCa",     # top hit of the 3-component landscape
    "*ABC::ABa_CBi_BCi_CAa",
    "*ABCD::ABa_CBi_BCa_DCi_ADi_CAa_DDa",
]


@pytest.mark.parametrize("state", TOPOLOGIES)
def test_statistical_equivalence(eng, cfg, state):
    n = 10_000
    g = gpu_run(eng, [state], n, cfg, seed=101)
    o = oracle_run(state, n, cfg, seed=202)        # independent streams: a real two-sample test
    se = np.sqrt(g["reward"].var(ddof=1) / n + o["reward"].var(ddof=1) / n)
    z = (g["reward"].mean() - o["reward"].mean()) / se
    assert abs(z) < 3.0, f"{state}: mean reward {g['reward'].mean():.4f} vs {o['reward'].mean():.4f} (z={z:.2f})"
    ks = stats.ks_2samp(g["lag"], o["lag"])
    assert ks.pvalue > 0.01, f"{state}: KS on lags p={ks.pvalue:.4f}"
    se_ev = np.sqrt(g["events"].var(ddof=1) / n + o["n_events"].astype(float).var(ddof=1) / n)
    z_ev = (g["events"].mean() - o["n_events"].mean()) / se_ev
    assert abs(z_ev) < 3.5, f"{state}: events/trajectory {g['events'].mean():.0f} vs {o['n_events'].mean():.0f} (z={z_ev:.2f})"


def test_random_topology_sweep(eng):
    """200 random terminal topologies (three and four components) x 1000 trajectories against the
    recorded checker results (tests/golden/make_sweep_oracle.py; independent seeds): the per-topology
    z-scores of mean reward and of events per trajectory must look like N(0,1) draws -- no outlier
    beyond 4.5, Kolmogorov-Smirnov p > 0.01 -- and the pooled lag histogram must agree."""
    from pathlib import Path

    gold = np.load(Path(__file__).parent / "golden" / "sweep_oracle.npz")
    states = [str(s) for s in gold["states"]]
    n = 1000
    c = small_cfg(nt=640, decim=16)
    g = gpu_run(eng, states, n, c, seed=777)
    rw = g["reward"].astype(np.float64).reshape(len(states), n)
    ev = g["events"].astype(np.float64).reshape(len(states), n)
    z_r = (rw.mean(axis=1) - gold["reward_mean"]) / np.sqrt(rw.var(axis=1, ddof=1) / n + gold["reward_var"] / n)
    z_e = (ev.mean(axis=1) - gold["events_mean"]) / np.sqrt(ev.var(axis=1, ddof=1) / n + gold["events_var"] / n)
    for name, z in (("reward", z_r), ("events", z_e)):
        assert np.abs(z).max() < 4.5, (name, states[int(np.abs(z).argmax())], float(np.abs(z).max()))
        p = stats.kstest(z, "norm").pvalue
        assert p > 0.01, f"{name}: z-scores are not N(0,1): KS p={p:.4f}, mean {z.mean():.3f}, sd {z.std():.3f}"
        assert abs(z.mean()) < 4.0 / np.sqrt(len(z)), f"{name}: systematic shift, mean z = {z.mean():.3f}"
    lag_hist = np.bincount(g["lag"], minlength=gold["lag_hist"].shape[1])
    gold_hist = gold["lag_hist"].sum(axis=0)
    keep = (lag_hist + gold_hist) >= 10
    chi = stats.chi2_contingency(np.vstack([lag_hist[keep], gold_hist[keep]]))
    assert chi.pvalue > 0.01, f"pooled lag histograms differ: p={chi.pvalue:.4f}"


def test_event_count_has_no_bias_at_high_power(eng):
    """2^20 trajectories on a short grid, same streams on both sides (paired): float32 arithmetic,
    lg2.approx / rcp.approx in tau and the float32 time-to-next-sample accumulation must not shift
    the number of events per trajectory by more than 5e-4 (relative), for a single-regulator topology
    (specialised loop) and for one with mixed promoters (general loop)."""
    n = 1 << 20
    c = small_cfg(nt=64, decim=16)
    for state in ("*ABC::ABi_BCi_CAi", "*ABC::AAa_ABi_BBa_BCi_CAi_CCa"):
        g = gpu_run(eng, [state], n, c, seed=31)
        o = oracle_run(state, n, c, seed=31)
        d = g["events"].astype(np.float64) - o["n_events"].astype(np.float64)
        rel = d.mean() / o["n_events"].mean()
        assert abs(rel) < 5e-4, f"{state}: events/trajectory biased by {rel:.2e} (gpu {g['events'].mean():.2f}, cpu {o['n_events'].mean():.2f})"
        # the paired differences are what float32 rounding does to individual trajectories: small and centred
        assert abs(d.mean()) < 5.0 * d.std(ddof=1) / np.sqrt(n) + 1e-4 * o["n_events"].mean()


def test_same_stream_trajectories_track_the_checker(eng, cfg):
    """With identical seeds the two sides share parameters and initial conditions; trajectories
    decorrelate only through float32 rounding, so event counts stay strongly correlated."""
    n = 2000
    g = gpu_run(eng, ["*ABC::ABi_BCi_CAi"], n, cfg, seed=9)
    o = oracle_run("*ABC::ABi_BCi_CAi", n, cfg, seed=9)
    r = np.corrcoef(g["events"], o["n_events"].astype(float))[0, 1]
    assert r > 0.99


# ---- size-independent properties at scale ----------------------------------------------------------------
def test_determinism_and_launch_shape_independence(eng, cfg):
    state = "*ABC::ABi_BCi_CAi"
    c = small_cfg(nt=480, decim=16)
    n = 1 << 17
    a = gpu_run(eng, [state], n, c, seed=77)
    b = gpu_run(eng, [state], n, c, seed=77)
    for k in ("reward", "lag", "events"):
        assert np.array_equal(a[k], b[k]), k
    # the same trajectory ids, cut into 97 ragged jobs in shuffled order
    rg = np.random.default_rng(1)
    cuts = np.sort(rg.choice(np.arange(1, n), size=96, replace=False))
    starts = np.concatenate(([0], cuts)); ends = np.concatenate((cuts, [n]))
    order = rg.permutation(97)
    counts = (ends - starts)[order]; bases = starts[order]
    s = gpu_run(eng, [state] * 97, counts, c, seed=77, bases=bases)
    off = np.concatenate(([0], np.cumsum(counts)[:-1]))
    for j in range(97):
        sl = slice(off[j], off[j] + counts[j])
        ref = slice(bases[j], bases[j] + counts[j])
        assert np.array_equal(s["reward"][sl], a["reward"][ref])
        assert np.array_equal(s["events"][sl], a["events"][ref])


def test_mixed_topology_batch_equals_separate_launches(eng):
    c = small_cfg(nt=320, decim=16)
    states = ["*ABC::ABi_BCi_CAi", "*AB::ABa_BAi", "*ABC::ABa_BCi_CAa", "*A::AAi", "*ABC::", "*ABC::ABi_BCi_CAi"]
    counts = [700, 33, 64, 1, 257, 5]
    bases = [0, 1000, 2000, 3000, 4000, 5000]
    mixed = gpu_run(eng, states, counts, c, seed=8, bases=bases)
    off = np.concatenate(([0], np.cumsum(counts)))
    for j, s in enumerate(states):
        single = gpu_run(eng, [s], counts[j], c, seed=8, bases=[bases[j]])
        assert np.array_equal(mixed["reward"][off[j]:off[j + 1]], single["reward"]), s
        assert np.array_equal(mixed["events"][off[j]:off[j + 1]], single["events"]), s


def test_schedules_give_identical_results(eng):
    """One topology per warp (schedule 1) vs mixed topologies per warp (schedule 2): same bits."""
    rg = np.random.default_rng(3)
    pool = ["*ABC::ABi_BCi_CAi", "*ABC::ABa_BCi_CAa", "*ABC::AAi_BBi_CCi", "*AB::ABa_BAi", "*ABC::", "*ABC::ABi_BAi_CCa_ACa"]
    states = [pool[i] for i in rg.integers(0, len(pool), size=150)]
    counts = rg.integers(1, 40, size=150)
    bases = np.arange(150) * 1000
    outs = []
    for schedule in (1, 2, 0, 3):
        c = small_cfg(nt=320, decim=16)
        c.schedule = schedule
        outs.append(gpu_run(eng, states, counts, c, seed=13, bases=bases))
    for k in ("reward", "lag", "events"):
        for other in outs[1:]:
            assert np.array_equal(outs[0][k], other[k]), k
    # longest-first ordering (and any cost hint) changes the launch order only
    for schedule in (1, 2):
        c = small_cfg(nt=320, decim=16)
        c.schedule, c.lpt = schedule, 0
        plain = gpu_run(eng, states, counts, c, seed=13, bases=bases)
        c.lpt = 1
        hinted = gpu_run(eng, states, counts, c, seed=13, bases=bases, job_cost=rg.random(150) * 100)
        for k in ("reward", "lag", "events"):
            assert np.array_equal(plain[k], outs[0][k]) and np.array_equal(hinted[k], outs[0][k]), (schedule, k)
    bad = _capi.default_config(); bad.schedule = 7
    with pytest.raises(_capi.EngineError):
        eng.rewards_host([handle_of("*AB::ABi")], [4], [0], bad, 1)


def test_cooperative_pairwise_kernel_is_bit_identical_on_the_full_grid(eng, cfg):
    """ssa_pairwise_coop_kernel (schedule 3: one trajectory per sub-group of 4 / 8 lanes, one gene per
    lane; what schedule 0 picks for launches of at most one wave) against the one-lane kernels over whole
    2000-point trajectories: single-regulator, mixed promoters, self-loops, 2-5 components, series too."""
    cases = [["*ABC::ABi_BCi_CAi", "*ABC::AAa_ABi_BBa_BCi_CAi_CCa", "*AB::ABa_BAi", "*A::AAi", "*ABC::"],
             ["*ABCD::ABa_CBi_BCa_DCi_ADi_CAa_DDa", "*ABCD::ABi_BCi_CDi_DAi"],
             ["*ABCDE::ABi_BCi_CDi_DEi_EAi", "*ABCDE::AAa_ABi_BAa_BBi_CAi_DCa_EDi_AEa_CCi", "*ABC::ABa_CBi_BCi_CAa"]]
    for states in cases:
        counts = [37 + 11 * k for k in range(len(states))]
        outs = []
        for schedule in (3, 2, 1):
            c = _capi.default_config()
            c.schedule = schedule
            outs.append(gpu_run(eng, states, counts, c, seed=41, bases=np.arange(len(states)) * 1000, series=True))
        for k in ("reward", "lag", "events", "series"):
            assert np.array_equal(outs[0][k], outs[1][k]) and np.array_equal(outs[0][k], outs[2][k]), (states[0], k)
        assert outs[0]["events"].min() > 1000
    # the automatic schedule takes the cooperative kernel for a launch of one wave and gives the same bits
    c = _capi.default_config()
    auto = gpu_run(eng, cases[0], [64] * 5, c, seed=43)
    c.schedule = 1
    one = gpu_run(eng, cases[0], [64] * 5, c, seed=43)
    for k in ("reward", "lag", "events"):
        assert np.array_equal(auto[k], one[k]), k


def test_single_regulator_fast_path_is_bit_identical(eng):
    """Launches whose topologies give every gene at most one regulator run a specialised hot loop
    (ct_sim_config.fast_path); it must reproduce the general loop bit for bit."""
    cases = [["*ABC::ABi_BCi_CAi"], ["*ABC::AAi_BAa_CBi", "*ABC::", "*ABC::CAa"], ["*ABCD::ABi_BCi_CDi_DAa"],
             ["*ABCDE::ABi_BCa_CDi_DEi_EAi", "*ABCDE::AAa_EBi"]]
    for states in cases:
        outs = []
        for fast in (1, 0):
            c = small_cfg(nt=640, decim=16)
            c.fast_path, c.schedule = fast, 1
            outs.append(gpu_run(eng, states, 3000, c, seed=17))
        c = small_cfg(nt=640, decim=16)
        c.schedule = 2                                  # mixed lanes: always the general loop
        outs.append(gpu_run(eng, states, 3000, c, seed=17))
        for k in ("reward", "lag", "events"):
            assert np.array_equal(outs[0][k], outs[1][k]), (states, k)
            assert np.array_equal(outs[0][k], outs[2][k]), (states, k)
        assert outs[0]["events"].sum() > 0


def test_empty_and_ragged_jobs(eng):
    c = small_cfg(nt=160, decim=16)
    out = eng.rewards_host([handle_of("*AB::ABi"), handle_of("*AB::ABa"), handle_of("*AB::ABi")], [0, 3, 0], [0, 0, 0], c, 1,
                           per_trajectory=True)
    assert out["job_mean"][0] == 0.0 and out["job_mean"][2] == 0.0 and len(out["reward"]) == 3
    out = eng.rewards_host([], [], [], c, 1)
    assert len(out["job_mean"]) == 0


def test_host_api_matches_device_api(eng):
    c = small_cfg(nt=320, decim=16)
    states = ["*ABC::ABi_BCi_CAi", "*AB::ABa_BAi"]
    dev = gpu_run(eng, states, [500, 300], c, seed=21, bases=[10, 5000])
    host = eng.rewards_host([handle_of(s) for s in states], [500, 300], [10, 5000], c, 21, per_trajectory=True)
    assert np.array_equal(host["reward"], dev["reward"]) and np.array_equal(host["lag"], dev["lag"])
    assert host["job_events"].tolist() == [int(dev["events"][:500].sum()), int(dev["events"][500:].sum())]
    assert np.allclose(host["job_mean"], [dev["reward"][:500].astype(np.float64).mean(), dev["reward"][500:].astype(np.float64).mean()], atol=1e-12)


def test_error_codes(eng, cfg):
    bad = _capi.default_config(); bad.nt, bad.decim = 2000, 8          # 250 stored samples > CT_MAX_NDEC
    with pytest.raises(_capi.EngineError):
        eng.rewards_host([handle_of("*AB::ABi")], [4], [0], bad, 1)
    with pytest.raises(_capi.EngineError):
        eng.rewards_host([12345], [4], [0], cfg, 1)                    # not a compiled handle (0 species)
    params = np.full((4, _capi.N_PARAMS), 0.1, dtype=np.float32)
    params[2, 3] = np.nan                                              # would stall the clock of trajectory 2
    with pytest.raises(_capi.EngineError, match="parameter 3 of trajectory 2"):
        eng.rewards_host([handle_of("*AB::ABi")], [4], [0], cfg, 1, params=params)


# ---- the reference-facing API ---------------------------------------------------------------------------
def test_get_reward_through_circuitree_api():
    grammar = SimpleNetworkGrammar(list("ABC"), AI, root="ABC::")
    tree = OscillationTree(grammar=grammar, root="ABC::", seed=1, n_exhausted=np.inf, trajectories_per_reward=256)
    r_osc = tree.get_reward("*ABC::ABi_BCi_CAi")
    r_flat = tree.get_reward("*ABC::ABa_BCa_CAa")
    assert 0.36 < r_osc < 0.50 and 0.12 < r_flat < 0.28 and r_osc > r_flat + 0.1
    rs = tree.get_rewards(["*ABC::ABi_BCi_CAi", "*ABC::ABa_BCa_CAa", "*AB::ABa_BAi"])
    assert len(rs) == 3 and abs(rs[0] - r_osc) < 0.06
    tree.search_mcts(40)
    assert tree.graph.nodes["ABC::"]["visits"] == 40
    tree.search_mcts_batched(256, 64)
    assert tree.graph.nodes["ABC::"]["visits"] == 296
    assert 0.0 < tree.graph.nodes["ABC::"]["reward"] / 296 < 1.0
    assert tree.reward_engine.total_trajectories == (1 + 1 + 3 + 40 + 256) * 256


def test_async_submit_wait_matches_sync(eng):
    c = small_cfg(nt=320, decim=16)
    hs = [handle_of(s) for s in ["*ABC::ABi_BCi_CAi", "*AB::ABa_BAi", "*ABC::ABa_BCi_CAa"]]
    t1 = eng.submit(hs, [300, 200, 100], [0, 1000, 2000], c, 5)
    t2 = eng.submit(hs[:2], [64, 64], [5000, 6000], c, 5)          # two batches in flight
    b = eng.wait(t2, per_trajectory=True)
    a = eng.wait(t1, per_trajectory=True)
    sa = eng.rewards_host(hs, [300, 200, 100], [0, 1000, 2000], c, 5, per_trajectory=True)
    sb = eng.rewards_host(hs[:2], [64, 64], [5000, 6000], c, 5, per_trajectory=True)
    for x, y in ((a, sa), (b, sb)):
        assert np.array_equal(x["reward"], y["reward"]) and np.array_equal(x["lag"], y["lag"])
        assert np.array_equal(x["job_events"], y["job_events"]) and np.array_equal(x["job_mean"], y["job_mean"])
    with pytest.raises(_capi.EngineError):
        eng.wait(t1)                                                # tickets are single-use


def test_native_tree_pipelined_gpu_search():
    from circuitree_b200.native import NativeTree

    def run(depth):
        t = NativeTree(SimpleNetworkGrammar(list("ABC"), AI, root="ABC::"), "ABC::", seed=3, n_exhausted=np.inf,
                       trajectories_per_reward=8)
        t.search_mcts(512, batch_size=64, pipeline_depth=depth)
        g = t.to_networkx()
        return g, t

    g, t = run(3)
    assert g.nodes["ABC::"]["visits"] == 512
    assert t.reward_engine.total_trajectories == 512 * 8
    mean = g.nodes["ABC::"]["reward"] / 512
    assert 0.1 < mean < 0.5
    g2, _ = run(3)
    assert sorted((n, d["visits"], d["reward"]) for n, d in g.nodes(data=True)) == \
        sorted((n, d["visits"], d["reward"]) for n, d in g2.nodes(data=True))


def test_search_bfs_batched_on_gpu():
    """search_bfs (reference circuitree.py:601-691) with the GPU reward, visits grouped into launches."""
    grammar = SimpleNetworkGrammar(list("AB"), AI, root="AB::")
    tree = OscillationTree(grammar=grammar, root="AB::", seed=2, n_exhausted=np.inf, trajectories_per_reward=4)
    tree.search_bfs(n_cycles=1, batch_size=256)
    terminals = [n for n in tree.graph.nodes if n.startswith("*")]
    assert len(terminals) > 20
    assert all(tree.graph.nodes[n]["visits"] == 1 for n in terminals)
    assert all(0.0 <= tree.graph.nodes[n]["reward"] <= 1.0 for n in terminals)
    assert tree.reward_engine.total_trajectories == 4 * len(terminals)


def test_reward_landscape_is_sharding_independent():
    from circuitree_b200 import reward_landscape

    grammar = SimpleNetworkGrammar(list("ABC"), AI, root="ABC::")
    states = ["*ABC::ABi_BCi_CAi", "*ABC::ABa_BCa_CAa", "*ABC::AAi", "*ABC::ABi_BAi", "*ABC::ABa_BCi_CAa"]
    cfg = small_cfg(nt=480, decim=16)
    whole = reward_landscape(RewardEngine(grammar, device=0, sim_config=cfg, seed=4), states, 128)
    parts = [reward_landscape(RewardEngine(grammar, device=0, sim_config=cfg, seed=4), states, 128, chunk=256, rank=r, world=2)
             for r in range(2)]
    merged = np.zeros(len(states))
    for p in parts:
        merged[p["index"]] = p["mean"]
    assert np.array_equal(merged, whole["mean"])
    assert whole["mean"][0] > whole["mean"][1]          # repressilator oscillates, the activation ring does not
    assert (whole["sem"] > 0).all() and (whole["success"] <= 1).all()


def test_same_seed_same_search():
    def run():
        grammar = SimpleNetworkGrammar(list("ABC"), AI, root="ABC::")
        t = OscillationTree(grammar=grammar, root="ABC::", seed=5, n_exhausted=np.inf, trajectories_per_reward=16)
        t.search_mcts_batched(128, 32)
        return sorted((n, d["visits"], d["reward"]) for n, d in t.graph.nodes(data=True))
    assert run() == run()
