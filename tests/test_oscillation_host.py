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
