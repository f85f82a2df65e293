"""Host-side behaviour of OscillationTree / RewardEngine that needs no GPU."""
import numpy as np

from circuitree_b200 import DimerNetworkTree, OscillationTree, SimpleNetworkGrammar, SimpleNetworkTree

AI = ["activates", "inhibits"]


def test_simulator_stream_position_survives_a_checkpoint(tmp_path):
    """to_file/from_file keep reward_seed AND the first unused trajectory id, so a restored tree
    continues with fresh Philox streams instead of replaying the ones it already used."""
    g = SimpleNetworkGrammar(list("ABC"), AI, root="ABC::")
    t = OscillationTree(grammar=g, root="ABC::", seed=5, n_exhausted=np.inf, trajectories_per_reward=16)
    t.reward_engine.next_traj = 123_456_789_012
    gml, js = t.to_file(tmp_path / "tree", tmp_path / "tree")
    r = OscillationTree.from_file(gml, js)
    assert r.reward_engine.next_traj == 123_456_789_012
    assert r.reward_seed == t.reward_seed and r.trajectories_per_reward == 16
    assert r.rg.bit_generator.state == t.rg.bit_generator.state


def test_explicit_trajectory_ids_advance_the_stream():
    """run(traj_base=...) (reward_landscape) moves next_traj past the ids it used."""
    from circuitree_b200.oscillation import RewardEngine

    class FakeEngine:
        def rewards_host(self, handles, counts, bases, cfg, seed, params=None, per_trajectory=False):
            return {"job_mean": np.zeros(len(handles)), "job_events": np.zeros(len(handles), dtype=np.uint64)}

    e = RewardEngine(SimpleNetworkGrammar(list("AB"), AI, root="AB::"), device=0, sim_config=object(), seed=1)
    e._engine = FakeEngine()
    e.run(["*AB::ABi", "*AB::ABa"], 100, traj_base=[5000, 700])
    assert e.next_traj == 5100
    out = e.run(["*AB::ABi"], 10)
    assert int(out["traj_base"][0]) == 5100 and e.next_traj == 5110


def test_convenience_trees_build_their_grammar():
    class S(SimpleNetworkTree):
        def get_reward(self, state, **kw):
            return 0.25

    class D(DimerNetworkTree):
        def get_reward(self, state, **kw):
            return 0.25

    s = S(components=["A", "B"], interactions=AI, root="AB::", n_exhausted=np.inf)
    assert s.seed == 2023 and s.grammar.components == ["A", "B"]
    s.search_mcts(10)
    assert s.graph.nodes["AB::"]["visits"] == 10
    d = D(components=["A", "B"], regulators=["R"], interactions=AI, root="AB+R::", n_exhausted=np.inf, seed=4)
    assert d.seed == 4 and d.grammar.regulators == ["R"]
    d.search_mcts(10)
    assert d.graph.nodes["AB+R::"]["visits"] == 10


class _RecordingEngine:
    """Stands in for _capi.Engine: per-trajectory reward = its global trajectory id (mod 2^24, exact in float32),
    events = 7 per trajectory, so the tests below can see which Philox streams a result was built from."""

    resident_lanes = 75776

    def __init__(self):
        self.calls = []
        self._tickets = {}

    def _result(self, counts, bases, per_trajectory):
        counts = np.asarray(counts, dtype=np.int64)
        ids = np.concatenate([np.arange(int(b), int(b) + int(c)) for b, c in zip(bases, counts)]) if len(counts) else np.zeros(0)
        reward = (ids % (1 << 24)).astype(np.float32)
        ends = np.cumsum(counts)
        means = np.array([reward[e - c:e].astype(np.float64).mean() if c else 0.0 for e, c in zip(ends, counts)])
        out = {"job_mean": means, "job_events": (7 * counts).astype(np.uint64)}
        if per_trajectory:
            out["reward"] = reward
            out["lag"] = np.zeros(len(reward), dtype=np.int32)
        return out

    def rewards_host(self, handles, counts, bases, cfg, seed, params=None, per_trajectory=False):
        self.calls.append(("host", np.array(handles), np.array(counts), np.array(bases)))
        return self._result(counts, bases, per_trajectory)

    def submit(self, handles, counts, bases, cfg, seed, job_cost=None):
        self.calls.append(("submit", np.array(handles), np.array(counts), np.array(bases), np.array(job_cost), cfg.grid_limit))
        self._tickets[len(self._tickets) + 1] = (np.array(counts), np.array(bases))
        return len(self._tickets)

    def wait(self, ticket, per_trajectory=False):
        counts, bases = self._tickets.pop(ticket)
        return self._result(counts, bases, per_trajectory)


def _fake_reward_engine(components="ABC"):
    from circuitree_b200 import _capi
    from circuitree_b200.oscillation import RewardEngine

    cfg = _capi.default_config()
    e = RewardEngine(SimpleNetworkGrammar(list(components), AI, root=f"{components}::"), device=0, sim_config=cfg, seed=1)
    e._engine = _RecordingEngine()
    return e


def test_every_leaf_of_a_batch_gets_its_own_block_of_trajectories():
    """submit_leaves groups the leaves by topology into one job each; collect_leaves hands leaf x the mean of
    its own B trajectories -- the r-th leaf of a topology (in selection order) gets the r-th block of that job."""
    e = _fake_reward_engine()
    e.next_traj = 1000
    states = ["*ABC::ABi", "*ABC::ABa_BCi", "*ABC::ABi", "*ABC::CAi", "*ABC::ABi", "*ABC::ABa_BCi"]
    handles = np.array([e.handle(s) for s in states], dtype=np.uint64)
    B = 4
    job = e.submit_leaves(handles, B)
    kind, uniq, counts, bases, job_cost, _ = e.engine.calls[-1]
    assert kind == "submit" and sorted(counts.tolist()) == [4, 8, 12] and int(counts.sum()) == len(states) * B
    assert e.next_traj == 1000 + len(states) * B
    assert np.all(job_cost == 1.0e6)                       # nothing seen yet: every topology assumed expensive
    means = e.collect_leaves(job)
    # expected: job u starts at bases[u]; the r-th leaf with that topology owns ids [base + r B, base + (r+1) B)
    seen = {}
    for x, h in enumerate(handles):
        u = int(np.nonzero(uniq == h)[0][0])
        r = seen.get(u, 0)
        seen[u] = r + 1
        first = int(bases[u]) + r * B
        assert means[x] == np.mean(np.arange(first, first + B)), (x, u, r)
    # all blocks distinct and inside the ids handed out
    assert len(set(means.tolist())) == len(states)
    assert e.total_trajectories == len(states) * B and e.total_events == 7 * len(states) * B
    # second batch: the observed events per trajectory are the cost hints; an unseen topology gets the largest one
    job2 = e.submit_leaves(np.array([handles[0], e.handle("*ABC::BAi")], dtype=np.uint64), B)
    assert np.all(e.engine.calls[-1][4] == 7.0)
    e.collect_leaves(job2)


def test_batches_in_flight_share_the_cta_slots():
    e = _fake_reward_engine()
    e.share = 4
    h = np.array([e.handle("*ABC::ABi")], dtype=np.uint64)
    e.submit_leaves(h, 2)
    assert e.engine.calls[-1][5] == 148 and e.cfg.grid_limit == 0        # 592 CTA slots / 4; restored afterwards
    e.share = 1
    e.submit_leaves(h, 2)
    assert e.engine.calls[-1][5] == 0


def test_reward_landscape_trajectory_ids_do_not_depend_on_the_sharding():
    """State i always runs trajectory ids [i n, (i + 1) n): a rank-sharded landscape equals the one-rank one."""
    from circuitree_b200.oscillation import reward_landscape

    states = [f"*ABC::{x}" for x in ("ABi", "ABa", "BCi", "BCa_CAi", "CAa", "ABi_BCi_CAi", "AAa")]
    n = 6
    whole = reward_landscape(_fake_reward_engine(), states, n, chunk=4 * n)       # several launches
    assert whole["index"].tolist() == list(range(len(states)))
    assert np.allclose(whole["mean"], [i * n + (n - 1) / 2 for i in range(len(states))])
    assert np.all(whole["events"] == 7 * n)
    assert whole["success"][0] == (n - 1) / n and np.all(whole["success"][1:] == 1.0)    # only "reward" 0 is below 0.4
    parts = [reward_landscape(_fake_reward_engine(), states, n, chunk=2 * n, rank=r, world=3) for r in range(3)]
    merged = np.zeros(len(states))
    for p in parts:
        merged[p["index"]] = p["mean"]
    assert sorted(np.concatenate([p["index"] for p in parts]).tolist()) == list(range(len(states)))
    assert np.array_equal(merged, whole["mean"])
    e = _fake_reward_engine()
    reward_landscape(e, states, n)
    assert e.next_traj == len(states) * n          # a search on the same engine afterwards uses fresh ids


def test_issue_slot_accounting_follows_the_spec_formula():
    """total_issue_slots = sum over jobs of events x I_step(R, M) with R = 4 M + 2 edges (SPEC.md section 8);
    the shape decoded from the handle agrees with the library's ct_topology_info."""
    from circuitree_b200 import _capi, _capi_defaults
    from circuitree_b200.oscillation import pairwise_shape

    e = _fake_reward_engine("ABCD")
    states = ["*ABCD::ABi_BCi_CAi", "*ABCD::AAa_ABi_BDi_CAa_DCi", "*ABCD::ABa"]
    e.run(states, 10)
    hs = np.array([e.handle(s) for s in states], dtype=np.uint64)
    r, m = pairwise_shape(hs)
    assert m.tolist() == [4, 4, 4] and r.tolist() == [16 + 6, 16 + 10, 16 + 2]
    for h, rr, mm in zip(hs, r, m):
        assert _capi.topology_info(int(h)) == (int(mm), int(rr))
    want = sum(70.0 * _capi_defaults.i_step(int(rr), int(mm)) for rr, mm in zip(r, m))
    assert abs(e.total_issue_slots - want) < 1e-6 * want
    assert _capi_defaults.i_step(18, 3) == 113.5      # the repressilator figure quoted in bench.py and DESIGN.md
