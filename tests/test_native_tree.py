"""Native (C++) grammar + MCTS scheduler vs the reference's golden vectors and vs the Python mirror.

Everything here runs on the CPU: the native tree is host code inside libcircuitree_b200.so.
Covers SURVEY.md section 8 rows a4-a10, a12 for the batched scheduler: bit-exact state strings,
action order, canonical forms, RNG consumption, tree statistics and exhaustion bookkeeping.
"""
import hashlib
import zlib

import numpy as np
import pytest

from circuitree_b200 import CircuiTree, DimersGrammar, SimpleNetworkGrammar
from circuitree_b200.native import NativeExhaustionError, NativeTree

AI = ["activates", "inhibits"]


def crc_reward(state: str) -> float:
    return (zlib.crc32(state.encode()) % 1000) / 1000.0


def crc_rewards(states):
    return [crc_reward(s) for s in states]


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


def native(n, root, seed, n_exhausted=np.inf, c=None, **kw):
    return NativeTree(grammar(n), root, seed=seed, n_exhausted=n_exhausted, exploration_constant=c, reward_fn=crc_rewards, **kw)


def test_rng_matches_numpy_generator():
    t = native(3, "ABC::", 123)
    rg = np.random.default_rng(123)
    from circuitree_b200.native import _lib, _check
    import ctypes

    sizes = [18, 17, 2, 3, 5, 50, 7, 1, 33, 64, 65, 100, 2, 2, 9]
    for rep in range(40):
        for n in sizes:
            if (rep + n) % 3 == 0:
                ref = list(range(n)); rg.shuffle(ref)
                out = np.zeros(n, dtype=np.int64)
                _check(_lib().ct_tree_rng_draw(t._h, 0, n, out.ctypes.data))
                assert out.tolist() == ref
            else:
                ref = int(rg.choice(np.arange(n)))
                out = np.zeros(1, dtype=np.int64)
                _check(_lib().ct_tree_rng_draw(t._h, 1, n, out.ctypes.data))
                assert int(out[0]) == ref
    t._pull_rng()
    assert t.rg.bit_generator.state == rg.bit_generator.state


def test_actions_and_unique_state_match_reference(ref_vectors):
    t3, t5 = native(3, "ABC::", 0), native(5, "ABCDE::", 0)
    for state, actions in ref_vectors["actions3"].items():
        assert t3.get_actions(state) == actions, state
    for state, actions in ref_vectors["actions5"].items():
        assert t5.get_actions(state) == actions, state
    t4 = native(4, "ABCD::", 0)
    for state, unique in ref_vectors["unique"].items():
        n = len(state.strip("*").split("::")[0])
        t = {3: t3, 4: t4, 5: t5}[n]
        assert t.get_unique_state(state) == unique, state
    tm = NativeTree(SimpleNetworkGrammar(list("ABC"), AI, max_interactions=4), "ABC::", seed=0, reward_fn=crc_rewards)
    for state, actions in ref_vectors["actions3_max4"].items():
        assert tm.get_actions(state) == actions, state


def test_declaration_order_is_respected():
    g = SimpleNetworkGrammar(["C", "A", "B"], ["inhibits", "activates"])
    t = NativeTree(g, "ABC::", seed=0, reward_fn=crc_rewards)
    for s in ["ABC::", "ABC::CAi", "ABC::ABa_BCi", "ABC::AAa_ABi_BBa"]:
        assert t.get_actions(s) == g.get_actions(s), s


@pytest.mark.parametrize("key,n,root,steps,seed,n_exh,c", [
    ("mcts3", 3, "ABC::", 3000, 0, np.inf, None),
    ("mcts3_c05", 3, "ABC::", 500, 7, np.inf, 0.5),
    ("mcts5", 5, "ABCDE::", 150, 1, np.inf, None),
    ("mcts2_exhaust", 2, "AB::", 5000, 3, 3, None),
    ("mcts3_exhaust", 3, "ABC::", 2500, 5, 2, None),
])
def test_native_mcts_trace_is_bit_exact(ref_vectors, key, n, root, steps, seed, n_exh, c):
    gold = ref_vectors[key]
    t = native(n, root, seed, n_exh, c)
    digests, err = [], None
    try:
        t.search_mcts(steps, batch_size=1, callback=lambda tr, i, p, s, r: digests.append(step_digest(p, s, r)))
    except NativeExhaustionError:
        err = "ExhaustionError"
    assert digests == gold["digests"][: len(digests)]
    assert len(digests) == gold["n_done"]
    assert err == gold["error"]
    g = t.to_networkx()
    assert (g.number_of_nodes(), g.number_of_edges()) == (gold["n_nodes"], gold["n_edges"])
    assert graph_digest(g) == gold["graph_digest"]
    assert g.nodes[root]["visits"] == gold["root_visits"] and g.nodes[root]["reward"] == gold["root_reward"]


class CrcTree(CircuiTree):
    def get_reward(self, state, **kw):
        return crc_reward(state)


@pytest.mark.parametrize("key", ["b8_d1_exh3", "b16_d3_exh4", "b8_d2_exh2_to_the_end", "b5_d2_two_components",
                                 "b1_d1_is_sequential"])
def test_native_exhaustion_with_leaves_in_flight(batched_vectors, key):
    """Row f3 on the native tree: same schedule, same bits as the reference's own methods."""
    gname, root, steps, seed, n_exh, batch, depth = batched_vectors[key]["args"]
    gold = batched_vectors[key]["res"]
    g = {"g3": grammar(3), "g2": grammar(2), "g3m3": SimpleNetworkGrammar(list("ABC"), AI, max_interactions=3)}[gname]
    t = NativeTree(g, root, seed=seed, n_exhausted=n_exh, reward_fn=crc_rewards)
    digests, err = [], None
    try:
        t.search_mcts(steps, batch_size=batch, pipeline_depth=depth,
                      callback=lambda tr, i, p, s, r: digests.append(step_digest(p, s, r)))
    except NativeExhaustionError:
        err = "ExhaustionError"
    assert err == gold["error"]
    assert digests == gold["digests"]
    nx_g = t.to_networkx()
    assert (nx_g.number_of_nodes(), nx_g.number_of_edges()) == (gold["n_nodes"], gold["n_edges"])
    assert graph_digest(nx_g) == gold["graph_digest"]
    assert nx_g.nodes[root]["visits"] == gold["root_visits"] and nx_g.nodes[root]["reward"] == gold["root_reward"]


@pytest.mark.parametrize("batch", [7, 64])
def test_native_batched_equals_python_batched(batch):
    """Virtual-loss batching: same leaves, same statistics, same RNG state as the Python scheduler."""
    py = CrcTree(grammar=grammar(4), root="ABCD::", seed=9, n_exhausted=np.inf)
    py.search_mcts_batched(600, batch)
    nt = native(4, "ABCD::", 9)
    nt.search_mcts(600, batch_size=batch)
    assert graph_digest(nt.to_networkx()) == graph_digest(py.graph)
    assert nt.rg.bit_generator.state == py.rg.bit_generator.state


def test_pipelined_search_conserves_counts():
    """Several reward batches in flight (virtual loss across batches): every selection is
    back-propagated exactly once, in selection order."""
    seen = []
    t = native(3, "ABC::", 4)
    t.search_mcts(1000, batch_size=64, pipeline_depth=3, callback=lambda tr, i, p, s, r: seen.append((i, r, crc_reward(s))))
    assert [i for i, _, _ in seen] == list(range(1000))
    assert all(r == ref for _, r, ref in seen)
    g = t.to_networkx()
    assert g.nodes["ABC::"]["visits"] == 1000
    assert abs(g.nodes["ABC::"]["reward"] - sum(r for _, r, _ in seen)) < 1e-9
    # depth 1 with the same seed selects differently (less virtual loss) but conserves counts too
    t1 = native(3, "ABC::", 4)
    t1.search_mcts(1000, batch_size=64, pipeline_depth=1)
    assert t1.to_networkx().nodes["ABC::"]["visits"] == 1000


def test_native_grow_tree(simple3_space):
    t = native(3, "ABC::", 0)
    t.grow_tree()
    g = t.to_networkx()
    nodes = sorted(g.nodes)
    assert nodes == simple3_space["nodes"]
    index = {n: i for i, n in enumerate(nodes)}
    assert sorted([index[u], index[v]] for u, v in g.edges) == simple3_space["edges"]


def test_four_component_space_size():
    t = native(4, "ABCD::", 0)
    t.grow_tree()
    nn, ne = t.size()
    keys = t.export()["node_keys"]
    n_term = int(((keys >> np.uint64(62)) & np.uint64(1)).sum())
    assert nn == 2 * n_term + 1          # every non-root state has exactly one terminal twin
    assert n_term > 100_000
    # spot-check canonical forms against the Python grammar
    g4 = grammar(4)
    rg = np.random.default_rng(0)
    for k in rg.choice(keys, size=300, replace=False):
        s = t.key_to_string(int(k))
        assert g4.get_unique_state(s) == s


def test_handles_match_compile_topology():
    from circuitree_b200 import _capi

    t = native(3, "ABC::", 0)
    for s in ["*ABC::ABi_BCi_CAi", "*ABC::AAa_BAi_CCa", "*ABC::ABa"]:
        comps, act, inh = SimpleNetworkGrammar.parse_genotype(s)
        assert t.handle_of_key(t.string_to_key(s)) == _capi.compile_topology(3, act, inh)


def test_delta_merge_equals_single_tree_sums():
    """Root-parallel merge semantics: after exchanging deltas every replica holds the sum of all
    replicas' increments on every node and edge."""
    reps = [native(3, "ABC::", s) for s in (1, 2, 3)]
    for rounds in range(3):
        for r in reps:
            r.search_mcts(200, batch_size=16)
        deltas = [r.collect_deltas() for r in reps]
        for i, r in enumerate(reps):
            for j, d in enumerate(deltas):
                if i != j:
                    r.merge_deltas(d)
    digs = {graph_digest(r.to_networkx()) for r in reps}
    g = reps[0].to_networkx()
    assert g.nodes["ABC::"]["visits"] == 3 * 3 * 200
    # replicas agree on visits everywhere; rewards agree up to summation order
    g1 = reps[1].to_networkx()
    assert sorted(g.nodes) == sorted(g1.nodes)
    for n in g.nodes:
        assert g.nodes[n]["visits"] == g1.nodes[n]["visits"]
        assert abs(g.nodes[n]["reward"] - g1.nodes[n]["reward"]) < 1e-9
    assert len(digs) <= 3


def test_delta_merge_with_finite_n_exhausted_keeps_replicas_consistent():
    """Four replicas on ABC::, n_exhausted=4, merging every 128 steps: exhaustion travels as an event
    (the un-counting is replayed once per replica, never exchanged as an increment), so visit counts
    stay non-negative, the replicas agree node for node, and the search ends with the reference's
    ExhaustionError on every replica instead of a dead end."""
    from circuitree_b200.native import NativeExhaustionError

    reps = [NativeTree(SimpleNetworkGrammar(list("ABC"), AI, max_interactions=3), "ABC::", seed=10 + s, n_exhausted=4,
                       reward_fn=crc_rewards) for s in range(4)]
    exhausted = [False] * 4
    rounds = 0
    while not all(exhausted) and rounds < 2000:
        for i, r in enumerate(reps):
            if exhausted[i]:
                continue
            try:
                r.search_mcts(128, batch_size=16)
            except NativeExhaustionError:
                exhausted[i] = True
        bufs = [r.pack_deltas().copy() for r in reps]
        for i, r in enumerate(reps):
            for j, b in enumerate(bufs):
                if i != j and r.merge_packed(b):
                    exhausted[i] = True
        rounds += 1
        exports = [r.export() for r in reps]
        assert all((e["node_visits"] >= 0).all() and (e["edge_visits"] >= 0).all() for e in exports), rounds
        ref = {int(k): (int(v), bool(x)) for k, v, x in zip(exports[0]["node_keys"], exports[0]["node_visits"], exports[0]["node_exhausted"])}
        for e in exports[1:]:
            other = {int(k): (int(v), bool(x)) for k, v, x in zip(e["node_keys"], e["node_visits"], e["node_exhausted"])}
            assert other == ref, rounds
    assert all(exhausted) and rounds > 1
    g = reps[0].to_networkx()
    assert g.nodes["ABC::"].get("is_exhausted")
    assert all(d.get("is_exhausted") for n, d in g.nodes(data=True) if n.startswith("*"))
    for r in reps:      # an exhausted tree selects nothing more
        with pytest.raises(NativeExhaustionError):
            r.search_mcts(1)


def test_depth_limited_exchange():
    a, twin, b = native(3, "ABC::", 1), native(3, "ABC::", 1), native(3, "ABC::", 2)
    a.search_mcts(300, batch_size=8)
    twin.search_mcts(300, batch_size=8)
    full = twin.collect_deltas()
    d = a.collect_deltas(max_depth=2)
    assert 0 < len(d["node_key"]) < len(full["node_key"])
    b.merge_deltas(d)
    gb = b.to_networkx()
    assert gb.nodes["ABC::"]["visits"] == 300

    def depth(name):
        body = name.split("::")[1]
        return (len(body.split("_")) if body else 0) + name.startswith("*")

    assert max(depth(n) for n in gb.nodes) == 2
    # what was held back is still pending: a later unrestricted exchange delivers the rest exactly once
    b.merge_deltas(a.collect_deltas())
    ga, gb = a.to_networkx(), b.to_networkx()
    assert {n: ga.nodes[n]["visits"] for n in ga.nodes} == {n: gb.nodes[n]["visits"] for n in gb.nodes}


def test_unsupported_configurations_are_refused():
    from circuitree_b200 import _capi

    with pytest.raises(_capi.EngineError):
        NativeTree(SimpleNetworkGrammar(list("ABC"), AI), "AB::", seed=0, reward_fn=crc_rewards)      # partial root
    with pytest.raises(_capi.EngineError):
        NativeTree(SimpleNetworkGrammar(list("ABCDEF"), AI), "ABCDEF::", seed=0, reward_fn=crc_rewards)
    with pytest.raises(_capi.EngineError):
        NativeTree(SimpleNetworkGrammar(list("ABC"), AI, fixed_components=["A"]), "ABC::", seed=0, reward_fn=crc_rewards)


# ---- DimersGrammar on the native tree (reference models.py:524-939; parity target = the Python mirror, which is
# pinned to the reference wherever the reference runs and restates its documented intent elsewhere) ----
def dimers_grammar(comps="ABC", regs="DE", **kw):
    return DimersGrammar(components=list(comps), regulators=list(regs), interactions=AI, **kw)


def test_native_dimers_actions_and_canonical_forms_match_the_mirror():
    for comps, regs, kw in (("ABC", "DE", {}), ("AB", "R", {}), ("ABCDE", "", {}), ("BCA", "ED", {"max_interactions_per_promoter": 1}),
                            ("ABC", "R", {"max_interactions": 3})):
        g = dimers_grammar(comps, regs, **kw)
        root = "".join(sorted(comps)) + "+" + "".join(sorted(regs)) + "::"
        t = NativeTree(g, root, seed=0, reward_fn=crc_rewards)
        py = CrcTree(grammar=g, root=root, seed=5, n_exhausted=np.inf)
        rg = np.random.default_rng(7)
        state = root
        seen = 0
        for _ in range(400):
            acts = g.get_actions(state)
            assert t.get_actions(state) == acts, state
            if not acts:
                state = root
                continue
            state = py._do_action(state, acts[int(rg.integers(len(acts)))])
            # a scrambled spelling of the same circuit canonicalises to the same string on both sides
            body = [x for x in state.split("::")[1].split("_") if x]
            if body:
                cp = dict(zip(sorted(comps), rg.permutation(sorted(comps))))
                rp = dict(zip(sorted(regs), rg.permutation(sorted(regs)))) if regs else {}
                m = {**cp, **rp}
                sc = ["".join(m.get(ch, ch) for ch in x) for x in body]
                sc = [(x[1] + x[0] + x[2:]) if rg.random() < 0.5 else x for x in sc]
                rg.shuffle(sc)
                raw = ("*" if state.startswith("*") else "") + root + "_".join(sc)
                assert t.get_unique_state(raw) == g.get_unique_state(raw), raw
                seen += 1
        assert seen > 100


@pytest.mark.parametrize("comps,regs,steps,seed,n_exh,batch,depth", [
    ("ABC", "DE", 600, 1, np.inf, 1, 1),
    ("AB", "R", 1500, 2, np.inf, 1, 1),
    ("ABC", "R", 800, 3, np.inf, 16, 2),
    ("AB", "", 3000, 4, 3, 1, 1),
    ("AB", "R", 4000, 5, 2, 8, 2),
])
def test_native_dimers_mcts_is_bit_exact_with_the_mirror(comps, regs, steps, seed, n_exh, batch, depth):
    """Same leaves, same statistics, same RNG state as the Python tree on DimersGrammar states --
    sequential, batched / pipelined, and with finite n_exhausted."""
    from circuitree_b200 import ExhaustionError

    g = dimers_grammar(comps, regs)
    root = comps + "+" + regs + "::"
    py = CrcTree(grammar=g, root=root, seed=seed, n_exhausted=n_exh)
    nt = NativeTree(dimers_grammar(comps, regs), root, seed=seed, n_exhausted=n_exh, reward_fn=crc_rewards)
    dp, dn, ep, en = [], [], None, None
    try:
        py.search_mcts_batched(steps, batch, pipeline_depth=depth, callback_before_start=False,
                               callback=lambda t, i, p, s, r: dp.append(step_digest(p, s, r)))
    except ExhaustionError:
        ep = "ExhaustionError"
    try:
        nt.search_mcts(steps, batch_size=batch, pipeline_depth=depth, callback=lambda t, i, p, s, r: dn.append(step_digest(p, s, r)))
    except NativeExhaustionError:
        en = "ExhaustionError"
    assert ep == en
    assert dn == dp and len(dn) > 100
    assert graph_digest(nt.to_networkx()) == graph_digest(py.graph)
    if ep is None:
        assert nt.rg.bit_generator.state == py.rg.bit_generator.state


def test_native_dimers_handles_and_limits():
    from circuitree_b200 import _capi

    g = dimers_grammar("ABC", "DE")
    t = NativeTree(g, "ABC+DE::", seed=0, reward_fn=crc_rewards)
    for s in ["*ABC+DE::ADiB_BEiC_CCiA", "*ABC+DE::AAaA_ABiA_BDaC", "*ABC+DE::"]:
        comps, regs, act, inh = g.parse_genotype(s)
        assert t.handle_of_key(t.string_to_key(s, canonical=False)) == _capi.compile_topology(5, act, inh, k=3), s
    with pytest.raises(_capi.EngineError):
        NativeTree(dimers_grammar("ABCD", "EF"), "ABCD+EF::", seed=0, reward_fn=crc_rewards)       # six TFs
    with pytest.raises(_capi.EngineError):
        NativeTree(g, "AB+DE::", seed=0, reward_fn=crc_rewards)                                     # partial root


def _node_stats(tree):
    g = tree.to_networkx()
    return ({n: (d["visits"], round(d["reward"], 9), bool(d.get("is_exhausted", False))) for n, d in g.nodes(data=True)},
            {(u, v): (d["visits"], round(d["reward"], 9)) for u, v, d in g.edges(data=True)})


@pytest.mark.parametrize("comps,regs,kw", [("ABC", "DE", {}), ("ABCD", "E", {}), ("ABC", "", {"max_interactions_per_promoter": 3})])
def test_native_dimers_delta_merge_equals_single_tree_sums(comps, regs, kw):
    """Replica exchange on DimersGrammar trees: their keys are process-local ids of interned interaction lists, so
    the lists themselves travel (1 + ceil((cap - 6) / 7) words per state).  After exchanging, every replica holds the
    sum of all replicas' increments, by state string, although the three processes interned their states in
    different orders."""
    root = comps + "+" + regs + "::"
    reps = [NativeTree(dimers_grammar(comps, regs, **kw), root, seed=s, n_exhausted=np.inf, reward_fn=crc_rewards) for s in (1, 2, 3)]
    solo = [NativeTree(dimers_grammar(comps, regs, **kw), root, seed=s, n_exhausted=np.inf, reward_fn=crc_rewards) for s in (1, 2, 3)]
    for r, t in zip(reps, solo):
        r.search_mcts(300, batch_size=16)
        t.search_mcts(300, batch_size=16)
    bufs = [r.pack_deltas().copy() for r in reps]
    d = NativeTree.unpack_deltas(bufs[0])
    cap = min(reps[0].grammar.max_interactions, reps[0].grammar.max_interactions_per_promoter * len(comps))
    assert d["key_words"] == 1 + max(0, -(-(cap - 6) // 7)) == (1 if comps == "ABC" and regs else 2)
    assert d["node_key"].size == len(d["node_dv"]) * d["key_words"]
    for i, r in enumerate(reps):
        for j, b in enumerate(bufs):
            if i != j:
                assert r.merge_packed(b) is False
    # expected: the three independent searches added up, state by state
    want_n, want_e = {}, {}
    for t in solo:
        n, e = _node_stats(t)
        for k, (v, w, _) in n.items():
            a = want_n.get(k, (0, 0.0))
            want_n[k] = (a[0] + v, a[1] + w)
        for k, (v, w) in e.items():
            a = want_e.get(k, (0, 0.0))
            want_e[k] = (a[0] + v, a[1] + w)
    for r in reps:
        n, e = _node_stats(r)
        assert set(n) == set(want_n) and set(e) == set(want_e)
        assert all(n[k][0] == want_n[k][0] and abs(n[k][1] - want_n[k][1]) < 1e-6 for k in n)
        assert all(e[k][0] == want_e[k][0] and abs(e[k][1] - want_e[k][1]) < 1e-6 for k in e)
    assert reps[0].to_networkx().nodes[root]["visits"] == 900
    # nothing is sent twice, and a search continues on the merged tree
    assert len(NativeTree.unpack_deltas(reps[0].pack_deltas())["node_dv"]) == 0
    reps[0].search_mcts(50, batch_size=8)
    assert reps[0].to_networkx().nodes[root]["visits"] == 950


def test_native_dimers_exchange_with_finite_n_exhausted_and_foreign_buffers():
    """Exhaustion events travel as interaction lists too; replicas agree state by state and end exhausted.  A buffer
    of another grammar (other key width) or a malformed list is refused without touching the tree."""
    from circuitree_b200 import _capi

    mk = lambda s: NativeTree(dimers_grammar("AB", "R"), "AB+R::", seed=20 + s, n_exhausted=2, reward_fn=crc_rewards)
    reps = [mk(0), mk(1)]
    exhausted = [False, False]
    rounds = 0
    while not all(exhausted) and rounds < 3000:
        for i, r in enumerate(reps):
            if not exhausted[i]:
                try:
                    r.search_mcts(64, batch_size=8)
                except NativeExhaustionError:
                    exhausted[i] = True
        bufs = [r.pack_deltas().copy() for r in reps]
        for i, r in enumerate(reps):
            if r.merge_packed(bufs[1 - i]):
                exhausted[i] = True
        rounds += 1
        (n0, _), (n1, _) = _node_stats(reps[0]), _node_stats(reps[1])
        assert {k: (v[0], v[2]) for k, v in n0.items()} == {k: (v[0], v[2]) for k, v in n1.items()}, rounds
        assert all(v[0] >= 0 for v in n0.values())
    assert all(exhausted) and rounds > 1
    assert reps[0].to_networkx().nodes["AB+R::"].get("is_exhausted")
    # foreign / malformed buffers
    simple = native(3, "ABC::", 1)
    simple.search_mcts(20)
    wide = NativeTree(dimers_grammar("ABCD", "E"), "ABCD+E::", seed=1, reward_fn=crc_rewards)
    wide.search_mcts(20)
    before = _node_stats(wide)
    with pytest.raises(_capi.EngineError):
        wide.merge_packed(simple.pack_deltas())                 # one word per key vs two
    buf = NativeTree(dimers_grammar("ABCD", "E"), "ABCD+E::", seed=2, reward_fn=crc_rewards)
    buf.search_mcts(20)
    b = buf.pack_deltas().copy()
    d = NativeTree.unpack_deltas(b)
    assert d["key_words"] == 2
    x = int(np.argmax((d["node_key"][:, 0] >> np.uint64(1)) & np.uint64(127)))      # a state with >= 1 interaction
    b[NativeTree.PACK_HEADER + 2 * x] |= np.int64(511) << np.int64(8)                # code 511 does not exist
    with pytest.raises(_capi.EngineError):
        wide.merge_packed(b)
    assert _node_stats(wide) == before
