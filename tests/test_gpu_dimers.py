"""GPU parity tests for the dimer model (SPEC.md section 9; run with -m gpu on a B200): the CUDA path
through the C ABI vs the CPU checker, for DimersGrammar topologies (the index arrays of
DimersGrammar.parse_genotype, reference circuitree/models.py:844-884).

Same bars as tests/test_gpu_ssa.py: shared RNG streams give identical initial conditions; mean
reward within 3 standard errors over 10^4 trajectories with independent seeds; KS on the lag
distribution p > 0.01; events per trajectory within 3.5 s.e.; results independent of launch shape.
The model is this repo's SPEC.md -- upstream has no simulator ("parity unpinned").
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from scipy import stats

from pathlib import Path

from circuitree_b200 import DimersGrammar, OscillationTree, _capi
from oracle import ssa_oracle as so

# per-trajectory results of the CPU checker for the cases below, recorded by
# tests/golden/make_dimer_oracle.py (minutes of CPU time; tests/test_oracle.py re-checks slices of them)
GOLDEN = np.load(Path(__file__).parent / "golden" / "dimer_oracle.npz")


def recorded(name):
    return {"reward": GOLDEN[name + "/reward"], "lag": GOLDEN[name + "/lag"].astype(np.int64),
            "n_events": GOLDEN[name + "/n_events"].astype(np.int64)}


AI = ["activates", "inhibits"]
THREADS = so.max_threads()
GRAMMAR = DimersGrammar(components=["A", "B", "C"], regulators=["R"], interactions=AI)


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


def parse(state, grammar=GRAMMAR):
    comps, regs, act, inh = grammar.parse_genotype(state)
    return len(comps) + len(regs), len(comps), act, inh


def handle_of(state, grammar=GRAMMAR):
    n_tf, _, act, inh = parse(state, grammar)
    return _capi.compile_topology(n_tf, act, inh, k=3)


def oracle_run(state, n, cfg, seed, grammar=GRAMMAR, **kw):
    n_tf, n_comp, act, inh = parse(state, grammar)
    return so.run_dimers(n_tf, n_comp, act, inh, cfg.ranges_array(12), n, seed=seed, nt=cfg.nt, decim=cfg.decim,
                         dt=cfg.dt, init_mean=cfg.init_mean, n_threads=THREADS, **kw)


def gpu_run(eng, states, n, cfg, seed, bases=None, series=False, params=None, grammar=GRAMMAR):
    handles = [handle_of(s, grammar) for s in states]
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
                       d_series=d_series.data_ptr() if series else 0, series_stride=stride)
    torch.cuda.synchronize()
    out = {"reward": d_reward.cpu().numpy(), "lag": d_lag.cpu().numpy(), "events": d_events.cpu().numpy().astype(np.int64)}
    if series:
        out["series"] = d_series.cpu().numpy().reshape(total, nsp, cfg.n_dec)
    return out


def test_first_block_and_dead_network_are_bit_identical(eng):
    c = small_cfg(nt=64, decim=1)      # every grid sample stored: sample 0 is the Poisson initial condition
    state = "*ABC+R::AAiB_BBiC_CCiA"
    g = gpu_run(eng, [state], 512, c, seed=3, series=True)
    o = oracle_run(state, 512, c, seed=3, want_series=True)
    assert np.array_equal(g["series"][:, :, 0], o["series"][:, :, 0])
    z = np.zeros((64, 12), dtype=np.float32)           # zero rates: nothing ever fires
    g = gpu_run(eng, [state], 64, c, seed=4, series=True, params=z)
    o = oracle_run(state, 64, c, seed=4, want_series=True, explicit_params=z.astype(np.float64))
    assert np.array_equal(g["series"], o["series"])
    assert g["events"].sum() == 0 and np.all(g["reward"] == 0)


def test_constitutive_level_and_dimer_pool(eng, cfg):
    """Explicit parameters: without interactions every TF sits at km kp / (gm gp) = 250; a
    homodimer interaction with binding switched off adds a dimer pool in detailed balance (free
    dimers do not decay, so the mean monomer level stays) and its dimerisation events."""
    p = np.tile(np.array([0.0, 0.2, 0.004, 0.5, 0.5, 0.5, 0.5, 0.05, 0.02, 0.005, 2e-5, 2e-3], np.float32), (256, 1))
    g = gpu_run(eng, ["*ABC+R::"], 256, cfg, seed=5, series=True, params=p)
    level = (g["series"][:, :, 60:].astype(float) ** 2 / 16).mean(axis=(0, 2))
    assert np.all((235 < level) & (level < 265)), level
    g = gpu_run(eng, ["*ABC+R::AAiB"], 256, cfg, seed=5, series=True, params=p)
    o = oracle_run("*ABC+R::AAiB", 256, cfg, seed=6, want_series=True, explicit_params=p.astype(np.float64))
    lg = (g["series"][:, :, 60:].astype(float) ** 2 / 16).mean(axis=(0, 2))
    lo = (o["series"][:, :, 60:].astype(float) ** 2 / 16).mean(axis=(0, 2))
    assert np.all(np.abs(lg - lo) / lo < 0.03), (lg, lo)
    ev_g, ev_o = g["events"].mean(), o["n_events"].mean()
    assert abs(ev_g - ev_o) / ev_o < 0.01 and ev_g > 1.05 * 4 * 104_000     # 4 unregulated TFs ~ 104k events each


TOPOLOGIES = [
    "*ABC+R::AAiB_BBiC_CCiA",            # homodimer repressilator
    "*ABC+R::ARaB_BBiC_CCiA",            # heterodimer with the regulator, one activation
    "*ABC+R::",                          # no interactions
    "*ABC+R::ABiA_ABiB_CRaC_CCiA",       # shared heterodimer on two promoters, two interactions on promoter A
    "*ABC+R::AAaA_BRiA_ABiC",            # an activating AND an inhibiting dimer on promoter A (km_act_rep, the A bookkeeping)
]
CASE_NAMES = ["homodimer_repressilator", "heterodimer_regulator", "no_interactions", "shared_heterodimer", "mixed_promoter"]


@pytest.mark.parametrize("case", range(5))
def test_statistical_equivalence(eng, cfg, case):
    # the bar's 10^4 trajectories at the full grid for the oscillator; the others (1.7e6 events per
    # trajectory) at 2500 trajectories on a 800-point grid.  Checker side: recorded, seed 202.
    state = TOPOLOGIES[case]
    n = 10_000 if case == 0 else 2500
    if case != 0:
        cfg = small_cfg(nt=800, decim=16)
    g = gpu_run(eng, [state], n, cfg, seed=101)
    o = recorded(CASE_NAMES[case])                 # independent streams: a real two-sample test
    assert len(o["reward"]) == n
    se = np.sqrt(g["reward"].var(ddof=1) / n + o["reward"].var(ddof=1) / n)
    z = (g["reward"].mean() - o["reward"].mean()) / se
    assert abs(z) < 3.0, f"{state}: mean reward {g['reward'].mean():.4f} vs {o['reward'].mean():.4f} (z={z:.2f})"
    ks = stats.ks_2samp(g["lag"], o["lag"])
    assert ks.pvalue > 0.01, f"{state}: KS on lags p={ks.pvalue:.4f}"
    se_ev = np.sqrt(g["events"].var(ddof=1) / n + o["n_events"].astype(float).var(ddof=1) / n)
    z_ev = (g["events"].mean() - o["n_events"].mean()) / se_ev
    assert abs(z_ev) < 3.5, f"{state}: events/trajectory {g['events'].mean():.0f} vs {o['n_events'].mean():.0f} (z={z_ev:.2f})"


def test_five_component_dimer_network(eng, cfg):
    """BASELINE config 3 shape: five TFs, ten interactions (two per promoter)."""
    g5 = DimersGrammar(components=list("ABCDE"), regulators=[], interactions=AI)
    state = "*ABCDE+::AAiB_BBiC_CCiD_DDiE_EEiA_ABaA_BCaB_CDaC_DEaD_AEaE"
    n = 1000
    cfg = small_cfg(nt=800, decim=16)
    g = gpu_run(eng, [state], n, cfg, seed=31, grammar=g5)
    o = recorded("five_tf_ten_interactions")       # checker seed 32
    se = np.sqrt(g["reward"].var(ddof=1) / n + o["reward"].var(ddof=1) / n)
    assert abs(g["reward"].mean() - o["reward"].mean()) < 3.0 * se
    se_ev = np.sqrt(g["events"].var(ddof=1) / n + o["n_events"].astype(float).var(ddof=1) / n)
    assert abs(g["events"].mean() - o["n_events"].mean()) < 3.5 * se_ev


def test_same_stream_trajectories_track_the_checker(eng, cfg):
    n = 1000
    g = gpu_run(eng, [TOPOLOGIES[0]], n, cfg, seed=9)
    o = recorded("same_stream")                    # checker seed 9: same parameters and initial conditions
    assert np.corrcoef(g["events"], o["n_events"].astype(float))[0, 1] > 0.99


def test_launch_shape_independence_and_mixed_batches(eng):
    c = small_cfg(nt=320, decim=16)
    counts = [700, 33, 64, 1, 40]
    bases = [0, 1000, 2000, 3000, 4000]
    mixed = gpu_run(eng, TOPOLOGIES, counts, c, seed=8, bases=bases)
    again = gpu_run(eng, TOPOLOGIES, counts, c, seed=8, bases=bases)
    off = np.concatenate(([0], np.cumsum(counts)))
    for k in ("reward", "lag", "events"):
        assert np.array_equal(mixed[k], again[k])
    for j, s in enumerate(TOPOLOGIES):
        single = gpu_run(eng, [s], counts[j], c, seed=8, bases=[bases[j]])
        assert np.array_equal(mixed["reward"][off[j]:off[j + 1]], single["reward"]), s
        assert np.array_equal(mixed["events"][off[j]:off[j + 1]], single["events"]), s
    c.lpt = 0
    plain = gpu_run(eng, TOPOLOGIES, counts, c, seed=8, bases=bases)
    for k in ("reward", "lag", "events"):
        assert np.array_equal(mixed[k], plain[k])


def test_per_lane_programs_are_bit_identical(eng):
    """Launches of many small jobs give every lane its own topology (ssa_dimers_mixed_kernel, schedule 2);
    big jobs run one topology per warp (ssa_dimers_kernel, schedule 1); launches too small to fill the
    device spread every trajectory over a sub-group of 8 or 4 lanes (ssa_dimers_coop_kernel, schedule 3,
    and what schedule 0 picks here).  All follow the summation order of SPEC.md section 9: same bits."""
    g5 = DimersGrammar(components=list("ABCDE"), regulators=[], interactions=AI)
    cases = [
        (GRAMMAR, TOPOLOGIES + ["*ABC+R::CCaB"], [96, 40, 33, 64, 50, 7]),
        (g5, ["*ABCDE+::AAiB_BBiC_CCiD_DDiE_EEiA_ABaA_BCaB_CDaC_DEaD_AEaE", "*ABCDE+::AEaC", "*ABCDE+::"], [64, 30, 9]),
        # three interactions on promoter A: more than the per-lane kernel has slots -> it falls back, still equal
        (GRAMMAR, ["*ABC+R::AAaA_BBiA_CCiA_ABiB", "*ABC+R::AAiB"], [40, 24]),
    ]
    for grammar, states, counts in cases:
        outs = []
        for schedule in (1, 2, 3, 0):
            c = small_cfg(nt=400, decim=16)
            c.schedule = schedule
            outs.append(gpu_run(eng, states, counts, c, seed=23, bases=np.arange(len(states)) * 1000, grammar=grammar, series=True))
        for key in ("reward", "lag", "events", "series"):
            for other in outs[1:]:
                assert np.array_equal(outs[0][key], other[key]), (states[0], key)
        assert outs[0]["events"].sum() > 0
    # three TFs and at most four dimers: sub-groups of 4 lanes (the other cases above use 8)
    g3 = DimersGrammar(components=list("ABC"), regulators=[], interactions=AI)
    states = ["*ABC+::AAiB_BBiC_CCiA", "*ABC+::ABaA_ABiB_BCaC_ACiC", "*ABC+::", "*ABC+::CCaC"]
    outs = []
    for schedule in (1, 2, 3):
        c = small_cfg(nt=400, decim=16)
        c.schedule = schedule
        outs.append(gpu_run(eng, states, [70, 33, 5, 64], c, seed=29, bases=np.arange(4) * 1000, grammar=g3, series=True))
    for key in ("reward", "lag", "events", "series"):
        assert np.array_equal(outs[0][key], outs[1][key]) and np.array_equal(outs[0][key], outs[2][key]), key
    assert outs[0]["events"].sum() > 0


def test_cooperative_kernel_statistics_and_full_grid(eng, cfg):
    """The cooperative kernel on its own, on the full grid: the oscillator's recorded checker
    statistics (same bars as test_statistical_equivalence), and bit-equality with the one-lane kernels
    over whole 2000-point trajectories (hundreds of thousands of events each)."""
    state = TOPOLOGIES[0]
    n = 10_000
    coop = _capi.default_config(); coop.schedule = 3
    g = gpu_run(eng, [state], n, coop, seed=101)
    o = recorded(CASE_NAMES[0])
    se = np.sqrt(g["reward"].var(ddof=1) / n + o["reward"].var(ddof=1) / n)
    assert abs(g["reward"].mean() - o["reward"].mean()) < 3.0 * se
    assert stats.ks_2samp(g["lag"], o["lag"]).pvalue > 0.01
    se_ev = np.sqrt(g["events"].var(ddof=1) / n + o["n_events"].astype(float).var(ddof=1) / n)
    assert abs(g["events"].mean() - o["n_events"].mean()) < 3.5 * se_ev
    uni = _capi.default_config(); uni.schedule = 1
    u = gpu_run(eng, [state], n, uni, seed=101)
    for key in ("reward", "lag", "events"):
        assert np.array_equal(g[key], u[key]), key
    # a single small call (the reference's sequential get_reward): one launch of 16 trajectories
    few = gpu_run(eng, [TOPOLOGIES[3]], 16, coop, seed=5)
    ref = gpu_run(eng, [TOPOLOGIES[3]], 16, uni, seed=5)
    for key in ("reward", "lag", "events"):
        assert np.array_equal(few[key], ref[key]), key


def test_one_model_per_launch(eng):
    c = small_cfg(nt=160, decim=16)
    pair = _capi.compile_topology(3, [], [(0, 1), (1, 2), (2, 0)])
    dim = handle_of(TOPOLOGIES[0])
    with pytest.raises(_capi.EngineError):
        eng.rewards_host([pair, dim], [4, 4], [0, 4], c, 1)
    with pytest.raises(_capi.EngineError):
        eng.rewards_host([dim, pair], [4, 4], [0, 4], c, 1)
    with pytest.raises(_capi.EngineError):
        eng.rewards_host([_capi.DIMER_TAG | (3 << 56) | 0xFFFFFF], [4], [0], c, 1)      # unknown table entry
    out = eng.rewards_host([dim, dim], [0, 5], [0, 0], c, 1, per_trajectory=True)       # empty + ragged jobs
    assert out["job_mean"][0] == 0.0 and len(out["reward"]) == 5


def test_get_reward_through_circuitree_api_with_dimers_grammar():
    tree = OscillationTree(grammar=GRAMMAR, root="ABC+R::", seed=1, n_exhausted=np.inf, trajectories_per_reward=256)
    r_osc = tree.get_reward("*ABC+R::AAiB_BBiC_CCiA")
    r_flat = tree.get_reward("*ABC+R::")
    assert r_osc > 0.55 and r_flat < 0.3
    # the search itself on a short grid (random dimer networks fire ~2e6 events per full trajectory)
    short = OscillationTree(grammar=GRAMMAR, root="ABC+R::", seed=1, n_exhausted=np.inf, trajectories_per_reward=32,
                            sim_config=small_cfg(nt=320, decim=16))
    short.search_mcts_batched(64, 32)
    assert short.graph.nodes["ABC+R::"]["visits"] == 64
    assert 0.0 < short.graph.nodes["ABC+R::"]["reward"] / 64 < 1.0


def test_pipelined_search_with_dimers_grammar():
    """search_mcts_batched(pipeline_depth=3): asynchronous dimer-kernel batches behind the Python tree."""
    def run():
        c = small_cfg(nt=160, decim=16)
        t = OscillationTree(grammar=GRAMMAR, root="ABC+R::", seed=4, n_exhausted=np.inf, trajectories_per_reward=8, sim_config=c)
        t.search_mcts_batched(96, 16, pipeline_depth=3)
        return t
    a, b = run(), run()
    assert a.graph.nodes["ABC+R::"]["visits"] == 96
    assert a.reward_engine.total_trajectories == 96 * 8
    assert 0.0 < a.graph.nodes["ABC+R::"]["reward"] / 96 < 1.0
    assert sorted((n, d["visits"], d["reward"]) for n, d in a.graph.nodes(data=True)) == \
        sorted((n, d["visits"], d["reward"]) for n, d in b.graph.nodes(data=True))
