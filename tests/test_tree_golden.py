"""CircuiTree host mirror vs traces recorded from the unmodified reference.

Covers SURVEY.md section 8 rows a4-a7 and a10-a12: with identical (deterministic) rewards the
selection paths, rollouts, rewards and the final search graph are identical step for step,
including the exhaustion bookkeeping of reference circuitree.py:377-420.
"""
import hashlib
import zlib

import numpy as np
import pytest

from circuitree_b200 import CircuiTree, ExhaustionError, SimpleNetworkGrammar

AI = ["activates", "inhibits"]


def crc_reward(state: str) -> float:
    return (zlib.crc32(state.encode()) % 1000) / 1000.0


class CrcTree(CircuiTree):
    def get_reward(self, state, **kw):
        return crc_reward(state)


def step_digest(path, sim_node, reward):
    return zlib.crc32(("|".join(path) + "#" + sim_node + "#" + f"{reward:.12g}").encode())


def graph_digest(graph):
    h = hashlib.sha256()
    for n in sorted(graph.nodes):
        a = graph.nodes[n]
        h.update(f"{n}|{a.get('visits')}|{float(a.get('reward')):.12g}|{int(bool(a.get('is_exhausted', False)))}\n".encode())
    for u, v in sorted(graph.edges):
        a = graph.edges[u, v]
        h.update(f"{u}>{v}|{a.get('visits')}|{float(a.get('reward')):.12g}\n".encode())
    return h.hexdigest()


def grammar(n):
    return SimpleNetworkGrammar(list("ABCDE"[:n]), AI)


CASES = {
    "mcts3": (3, "ABC::", 3000, 0, np.inf, None),
    "mcts3_c05": (3, "ABC::", 500, 7, np.inf, 0.5),
    "mcts5": (5, "ABCDE::", 150, 1, np.inf, None),
    "mcts2_exhaust": (2, "AB::", 5000, 3, 3, None),
    "mcts3_exhaust": (3, "ABC::", 2500, 5, 2, None),
}


@pytest.mark.parametrize("key", list(CASES))
def test_mcts_trace_is_bit_exact(ref_vectors, key):
    n, root, steps, seed, n_exh, c = CASES[key]
    gold = ref_vectors[key]
    tree = CrcTree(grammar=grammar(n), root=root, seed=seed, n_exhausted=n_exh, exploration_constant=c)
    digests, first, err = [], [], None
    try:
        for step in range(steps):
            path, reward, sim = tree.traverse()
            digests.append(step_digest(path, sim, reward))
            if step < 12:
                first.append([path, sim, reward])
    except ExhaustionError:
        err = "ExhaustionError"
    assert first == gold["first"]
    assert digests == gold["digests"]
    assert err == gold["error"]
    assert tree.graph.number_of_nodes() == gold["n_nodes"]
    assert tree.graph.number_of_edges() == gold["n_edges"]
    assert tree.graph.nodes[root]["visits"] == gold["root_visits"]
    assert tree.graph.nodes[root]["reward"] == gold["root_reward"]
    assert graph_digest(tree.graph) == gold["graph_digest"]


def test_batched_with_batch_one_equals_sequential(ref_vectors):
    gold = ref_vectors["mcts3"]
    tree = CrcTree(grammar=grammar(3), root="ABC::", seed=0, n_exhausted=np.inf)
    digests = []
    tree.search_mcts_batched(3000, 1, callback=lambda t, i, p, s, r: digests.append(step_digest(p, s, r)) if i >= 0 else None)
    assert digests == gold["digests"]
    assert graph_digest(tree.graph) == gold["graph_digest"]


BATCHED = ["b8_d1_exh3", "b16_d3_exh4", "b8_d2_exh2_to_the_end", "b5_d2_two_components", "b1_d1_is_sequential"]


def batched_grammar(name):
    return {"g3": grammar(3), "g2": grammar(2), "g3m3": SimpleNetworkGrammar(list("ABC"), AI, max_interactions=3)}[name]


@pytest.mark.parametrize("key", BATCHED)
def test_exhaustion_with_leaves_in_flight_matches_reference_schedule(batched_vectors, key):
    """finite n_exhausted under batching / pipelining."""
    gname, root, steps, seed, n_exh, batch, depth = batched_vectors[key]["args"]
    gold = batched_vectors[key]["res"]
    tree = CrcTree(grammar=batched_grammar(gname), root=root, seed=seed, n_exhausted=n_exh)
    digests, err = [], None
    try:
        tree.search_mcts_batched(steps, batch, pipeline_depth=depth, callback_before_start=False,
                                 callback=lambda t, i, p, s, r: digests.append(step_digest(p, s, r)))
    except ExhaustionError:
        err = "ExhaustionError"
    assert err == gold["error"]
    assert digests == gold["digests"]
    assert (tree.graph.number_of_nodes(), tree.graph.number_of_edges()) == (gold["n_nodes"], gold["n_edges"])
    assert graph_digest(tree.graph) == gold["graph_digest"]
    assert tree.graph.nodes[root]["visits"] == gold["root_visits"] and tree.graph.nodes[root]["reward"] == gold["root_reward"]


def test_batched_virtual_loss_conserves_counts():
    tree = CrcTree(grammar=grammar(3), root="ABC::", seed=3, n_exhausted=np.inf)
    tree.search_mcts_batched(1000, 64)
    g = tree.graph
    assert g.nodes["ABC::"]["visits"] == 1000
    for n in g.nodes:
        out_visits = sum(g.edges[n, c]["visits"] for c in g.successors(n))
        assert out_visits <= g.nodes[n]["visits"]
    assert abs(sum(g.edges["ABC::", c]["reward"] for c in g.successors("ABC::")) - g.nodes["ABC::"]["reward"]) < 1e-9


def test_n_exhausted_none_does_not_crash():
    # the reference checkout raises TypeError here; we treat None as "never"
    tree = CrcTree(grammar=grammar(2), root="AB::", seed=0)
    tree.search_mcts(300)
    assert tree.graph.nodes["AB::"]["visits"] == 300


def test_grow_tree_and_enumeration(simple3_space):
    tree = CrcTree(grammar=grammar(3), root="ABC::", seed=0)
    tree.grow_tree()
    nodes = sorted(tree.graph.nodes)
    assert nodes == simple3_space["nodes"]
    index = {n: i for i, n in enumerate(nodes)}
    assert sorted([index[u], index[v]] for u, v in tree.graph.edges) == simple3_space["edges"]
    assert len(nodes) == 6651 and tree.graph.number_of_edges() == 22193
    terms = CrcTree(grammar=grammar(3), root="ABC::", seed=0).enumerate_terminal_states()
    assert sorted(terms) == [n for n in nodes if n.startswith("*")]
    assert len(terms) == simple3_space["n_terminal"] == 3325
    assert sorted(grammar(3).iter_states("ABC::", terminal_only=True)) == sorted(terms)


def test_two_component_space(ref_vectors):
    tree = CrcTree(grammar=grammar(2), root="AB::", seed=0)
    tree.grow_tree()
    assert sorted(tree.graph.nodes) == ref_vectors["simple2_nodes"]
    assert sorted([u, v] for u, v in tree.graph.edges) == ref_vectors["simple2_edges"]


def test_random_rollouts_consume_rng_like_reference(ref_vectors):
    t3 = CrcTree(grammar=grammar(3), root="ABC::", seed=42)
    assert [t3.get_random_terminal_descendant("ABC::") for _ in range(200)] == ref_vectors["rollouts3"]
    t5 = CrcTree(grammar=grammar(5), root="ABCDE::", seed=42)
    assert [t5.get_random_terminal_descendant("ABCDE::") for _ in range(40)] == ref_vectors["rollouts5"]


def test_search_bfs(ref_vectors):
    t = CrcTree(grammar=grammar(2), root="AB::", seed=0)
    t.search_bfs(n_cycles=2)
    assert graph_digest(t.graph) == ref_vectors["bfs2"]["graph_digest"]
    assert list(t.bfs_iterator()) == ref_vectors["bfs2"]["order"]
    t = CrcTree(grammar=grammar(2), root="AB::", seed=11)
    t.search_bfs(n_steps=100, n_repeats=3, shuffle=True)
    assert graph_digest(t.graph) == ref_vectors["bfs2_shuffle"]["graph_digest"]
    # batching changes nothing observable
    t = CrcTree(grammar=grammar(2), root="AB::", seed=11)
    t.search_bfs(n_steps=100, n_repeats=3, shuffle=True, batch_size=16)
    assert graph_digest(t.graph) == ref_vectors["bfs2_shuffle"]["graph_digest"]
    with pytest.raises(ValueError):
        t.search_bfs()


def test_file_roundtrip(tmp_path):
    tree = CrcTree(grammar=SimpleNetworkGrammar(list("ABC"), AI, root="ABC::"), root="ABC::", seed=5, n_exhausted=np.inf)
    tree.search_mcts(200)
    gml, js = tree.to_file(tmp_path / "tree", tmp_path / "tree")
    loaded = CrcTree.from_file(gml, js)
    assert graph_digest(loaded.graph) == graph_digest(tree.graph)
    assert loaded.root == "ABC::" and loaded.grammar.components == ["A", "B", "C"]
    assert loaded.seed == 5
    gz = tree.to_file(tmp_path / "tree2", compress=True)
    assert str(gz).endswith(".gml.gz")
    with pytest.raises(ValueError):
        tree.get_attributes(["graph"])


def test_resume_from_checkpoint_is_reproducible(tmp_path):
    """RNG state travels with the checkpoint (the reference drops it, circuitree.py:129-133)."""
    g = SimpleNetworkGrammar(list("ABC"), AI, root="ABC::")
    full = CrcTree(grammar=g, root="ABC::", seed=8, n_exhausted=np.inf)
    full.search_mcts(300)
    half = CrcTree(grammar=SimpleNetworkGrammar(list("ABC"), AI, root="ABC::"), root="ABC::", seed=8, n_exhausted=np.inf)
    half.search_mcts(150)
    gml, js = half.to_file(tmp_path / "half", tmp_path / "half")
    resumed = CrcTree.from_file(gml, js)
    resumed.search_mcts(150)
    assert graph_digest(resumed.graph) == graph_digest(full.graph)


def test_reference_can_load_our_checkpoint(tmp_path):
    """GML + JSON written here load in the unmodified reference (only where it is mounted)."""
    import importlib.util
    import sys

    ref_root = "/root/reference"
    if not importlib.util.find_spec("networkx") or not __import__("os").path.isdir(ref_root):
        pytest.skip("reference checkout not available on this machine")
    tree = CrcTree(grammar=SimpleNetworkGrammar(list("ABC"), AI, root="ABC::"), root="ABC::", seed=5, n_exhausted=1e9)
    tree.search_mcts(120)
    gml, js = tree.to_file(tmp_path / "t", tmp_path / "t")
    code = f"""
import sys, json
sys.path.insert(0, {ref_root!r})
import circuitree
from circuitree.models import SimpleNetworkGrammar
class T(circuitree.CircuiTree):
    def get_reward(self, s, **k): return 0.0
t = T.from_file({str(gml)!r}, {str(js)!r}, grammar_cls=SimpleNetworkGrammar)
print(json.dumps([t.graph.number_of_nodes(), t.graph.number_of_edges(), t.graph.nodes["ABC::"]["visits"], t.root]))
"""
    import subprocess

    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env={"PYTHONDONTWRITEBYTECODE": "1", "PATH": "/usr/bin:/bin"})
    assert res.returncode == 0, res.stderr[-1500:]
    import json

    nn, ne, visits, root = json.loads(res.stdout.strip().splitlines()[-1])
    assert (nn, ne) == (tree.graph.number_of_nodes(), tree.graph.number_of_edges())
    assert visits == 120 and root == "ABC::"


def test_missing_root_in_supplied_graph():
    import networkx as nx

    with pytest.raises(ValueError):
        CrcTree(grammar=grammar(2), root="AB::", graph=nx.DiGraph())


def test_pipelined_batched_search_keeps_virtual_loss_bookkeeping():
    """search_mcts_batched(pipeline_depth > 1): several reward batches in flight (submit_rewards /
    collect_rewards); every selected leaf is visited once and rewarded once, deterministically."""
    from circuitree_b200 import CircuiTree, SimpleNetworkGrammar

    class Stub(CircuiTree):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.submitted, self.collected, self.handed_out = 0, 0, 0.0

        def get_reward(self, state, **kw):
            r = (sum(map(ord, state)) % 7) / 7.0
            self.handed_out += r
            return r

        def submit_rewards(self, states, **kw):
            self.submitted += 1
            assert self.submitted - self.collected <= 3          # at most pipeline_depth batches in flight
            return super().submit_rewards(states, **kw)

        def collect_rewards(self, ticket):
            self.collected += 1
            return super().collect_rewards(ticket)

    def run(depth):
        g = SimpleNetworkGrammar(list("ABC"), ["activates", "inhibits"], root="ABC::")
        t = Stub(grammar=g, root="ABC::", seed=3, n_exhausted=np.inf)
        t.search_mcts_batched(500, 32, pipeline_depth=depth)
        return t

    a, b, c = run(3), run(3), run(1)
    assert a.submitted == a.collected == 16 and c.submitted == 0
    for t in (a, c):
        assert t.graph.nodes["ABC::"]["visits"] == 500
        assert abs(t.handed_out - t.graph.nodes["ABC::"]["reward"]) < 1e-9      # every reward reached the root once
    assert sorted((n, d["visits"], d["reward"]) for n, d in a.graph.nodes(data=True)) == \
        sorted((n, d["visits"], d["reward"]) for n, d in b.graph.nodes(data=True))
    with pytest.raises(ValueError):
        c.search_mcts_batched(10, 4, pipeline_depth=0)
