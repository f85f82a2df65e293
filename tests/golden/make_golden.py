"""Generate golden vectors from the UNMODIFIED reference package.

Run once in the build container (the reference is not available on the GPU box):

    PYTHONHASHSEED=0 PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference \
        python tests/golden/make_golden.py

Writes small JSON(.gz) fixtures next to this file.  Everything recorded here is an output of
reference code (circuitree 0.11.1 at /root/reference), never of this repository's code.
The deterministic pseudo-reward ``crc_reward`` lets both sides compute identical rewards
without sharing an RNG.
"""
import gzip
import hashlib
import json
import sys
import zlib
from pathlib import Path

import numpy as np

import circuitree
from circuitree import CircuiTree, SimpleNetworkGrammar, DimersGrammar
from circuitree.circuitree import ExhaustionError

assert "/root/reference" in circuitree.__file__, circuitree.__file__
HERE = Path(__file__).resolve().parent


def crc_reward(state: str) -> float:
    return (zlib.crc32(state.encode()) % 1000) / 1000.0


class CrcTree(CircuiTree):
    def get_reward(self, state, **kw):
        return crc_reward(state)


def dump(name, obj, compress=False):
    data = json.dumps(obj, indent=None, separators=(",", ":"), sort_keys=True)
    if compress:
        with gzip.GzipFile(HERE / (name + ".json.gz"), "wb", mtime=0) as f:
            f.write(data.encode())
    else:
        (HERE / (name + ".json")).write_text(data)
    print(name, len(data), "bytes")


def graph_digest(graph):
    h = hashlib.sha256()
    for n in sorted(graph.nodes):
        a = graph.nodes[n]
        h.update(f"{n}|{a.get('visits')}|{float(a.get('reward')):.12g}|{int(bool(a.get('is_exhausted', False)))}\n".encode())
    for u, v in sorted(graph.edges):
        a = graph.edges[u, v]
        h.update(f"{u}>{v}|{a.get('visits')}|{float(a.get('reward')):.12g}\n".encode())
    return h.hexdigest()


def step_digest(path, sim_node, reward):
    return zlib.crc32(("|".join(path) + "#" + sim_node + "#" + f"{reward:.12g}").encode())


def random_state(rg, comps, ixn_types, n_edges, terminal):
    pairs = [(a, b) for a in comps for b in comps]
    idx = rg.choice(len(pairs), size=n_edges, replace=False)
    ixns = [pairs[i][0] + pairs[i][1] + ixn_types[rg.integers(len(ixn_types))] for i in idx]
    c = list(comps)
    rg.shuffle(c)
    return ("*" if terminal else "") + "".join(c) + "::" + "_".join(ixns)


def main():
    out = {}

    # ---- 1. full enumeration, 3 components -------------------------------------------------------
    g3 = SimpleNetworkGrammar(components=["A", "B", "C"], interactions=["activates", "inhibits"])
    t = CrcTree(grammar=g3, root="ABC::", seed=0, n_exhausted=np.inf)
    t.grow_tree()
    nodes = sorted(t.graph.nodes)
    index = {n: i for i, n in enumerate(nodes)}
    edges = sorted((index[u], index[v]) for u, v in t.graph.edges)
    dump("simple3_space", {"root": "ABC::", "nodes": nodes, "edges": edges,
                           "n_terminal": sum(n.startswith("*") for n in nodes)}, compress=True)
    t2 = CrcTree(grammar=g3, root="ABC::", seed=0, n_exhausted=np.inf)
    terms = sorted(t2.enumerate_terminal_states())
    assert terms == [n for n in nodes if n.startswith("*")]

    # counts for 2 components and for a/i/no-self-loops variants
    g2 = SimpleNetworkGrammar(components=["A", "B"], interactions=["activates", "inhibits"])
    t = CrcTree(grammar=g2, root="AB::", seed=0, n_exhausted=np.inf)
    t.grow_tree()
    out["simple2_nodes"] = sorted(t.graph.nodes)
    out["simple2_edges"] = sorted([u, v] for u, v in t.graph.edges)

    # ---- 2. get_actions (order preserved) ----------------------------------------------------------
    rg = np.random.default_rng(123)
    acts = {}
    sample = [nodes[i] for i in rg.choice(len(nodes), size=150, replace=False)] + ["ABC::"]
    for s in sample:
        acts[s] = g3.get_actions(s)
    g3m = SimpleNetworkGrammar(components=["A", "B", "C"], interactions=["activates", "inhibits"], max_interactions=4)
    acts_max4 = {s: g3m.get_actions(s) for s in sample[:60]}
    g5 = SimpleNetworkGrammar(components=["A", "B", "C", "D", "E"], interactions=["activates", "inhibits"])
    acts5 = {}
    for _ in range(60):
        s = random_state(rg, "ABCDE", "ai", int(rg.integers(0, 12)), False)
        acts5[s] = g5.get_actions(s)
    # partial component sets: compare as sets for the component-add part (hash order upstream)
    gp = SimpleNetworkGrammar(components=["A", "B", "C"], interactions=["activates", "inhibits"], root="A::")
    acts_partial = {s: gp.get_actions(s) for s in ["A::", "AB::", "AB::ABa", "AC::CAi", "ABC::ABa", "B::BBa", "BA::BBa_ABi"]}
    out["actions3"] = acts
    out["actions3_max4"] = acts_max4
    out["actions5"] = acts5
    out["actions_partial"] = acts_partial

    # ---- 3. get_unique_state ------------------------------------------------------------------------
    uniq = {}
    for comps, g in (("ABC", g3), ("ABCDE", g5)):
        for _ in range(150):
            n_e = int(rg.integers(0, len(comps) ** 2 + 1))
            s = random_state(rg, comps, "ai", n_e, bool(rg.integers(2)))
            uniq[s] = g.get_unique_state(s)
    g4 = SimpleNetworkGrammar(components=["A", "B", "C", "D"], interactions=["activates", "inhibits"])
    for _ in range(100):
        s = random_state(rg, "ABCD", "ai", int(rg.integers(0, 17)), bool(rg.integers(2)))
        uniq[s] = g4.get_unique_state(s)
    out["unique"] = uniq
    gfix = SimpleNetworkGrammar(components=["A", "B", "C", "D"], interactions=["activates", "inhibits"], fixed_components=["A"])
    out["unique_fixedA"] = {}
    for _ in range(80):
        s = random_state(rg, "ABCD", "ai", int(rg.integers(0, 17)), bool(rg.integers(2)))
        out["unique_fixedA"][s] = gfix.get_unique_state(s)
    # subsets of components
    out["unique_partial"] = {s: g3.get_unique_state(s) for s in
                             ["B::", "CB::CBa", "BC::BCi_CCa", "*CA::ACi", "C::CCi", "AC::CAa_ACa", "BA::"]}

    # ---- 4. do_action / has_pattern / parse_genotype / to_dict ------------------------------------
    out["do_action"] = [[s, a, g3.do_action(s, a)] for s, a in [
        ("ABC::", "ABa"), ("ABC::ABa", "BCi"), ("ABC::ABa_BCi", "*terminate*"), ("AB::", "C"), ("AB::ABi", "C"),
        ("MP::", "MPi"), ("MP::MPi", "Q")]]
    pats = ["AAa", "ABi_BAi", "ABa_BAa", "ABa_BCa_CAa", "ABa_BCi_CAi", "ABi_BCi_CAi", "AAi_ABa"]
    hp = []
    for s in [nodes[i] for i in rg.choice(len(nodes), size=80, replace=False)]:
        hp.append([s, [bool(g3.has_pattern(s, p)) for p in pats]])
    out["has_pattern"] = {"patterns": pats, "cases": hp}
    pg = {}
    for s in [n for n in nodes if n.startswith("*")][::97]:
        c, a, i = SimpleNetworkGrammar.parse_genotype(s)
        pg[s] = [c, a.tolist(), i.tolist()]
    c, a, i = SimpleNetworkGrammar.parse_genotype("ACB::CAa_BBi_ABi", nonterminal_ok=True)
    pg["ACB::CAa_BBi_ABi"] = [c, a.tolist(), i.tolist()]
    out["parse_genotype"] = pg
    gfresh = SimpleNetworkGrammar(components=["A", "B", "C"], interactions=["activates", "inhibits"], root="ABC::")
    out["to_dict_simple"] = gfresh.to_dict()

    # ---- 5. MCTS traces (tree updates given fixed rewards) -----------------------------------------
    def run_mcts(grammar, root, n, seed, n_exhausted, c=None):
        tr = CrcTree(grammar=grammar, root=root, seed=seed, n_exhausted=n_exhausted, exploration_constant=c)
        digests, first = [], []
        err = None
        try:
            for step in range(n):
                path, reward, sim = tr.traverse()
                digests.append(step_digest(path, sim, reward))
                if step < 12:
                    first.append([path, sim, reward])
        except ExhaustionError as e:
            err = "ExhaustionError"
        return {"n_done": len(digests), "digests": digests, "first": first, "error": err,
                "n_nodes": tr.graph.number_of_nodes(), "n_edges": tr.graph.number_of_edges(),
                "graph_digest": graph_digest(tr.graph),
                "root_visits": tr.graph.nodes[root]["visits"], "root_reward": tr.graph.nodes[root]["reward"]}

    out["mcts3"] = run_mcts(g3, "ABC::", 3000, 0, np.inf)
    out["mcts3_c05"] = run_mcts(g3, "ABC::", 500, 7, np.inf, c=0.5)
    out["mcts5"] = run_mcts(g5, "ABCDE::", 150, 1, np.inf)
    out["mcts2_exhaust"] = run_mcts(g2, "AB::", 5000, 3, 3)
    out["mcts3_exhaust"] = run_mcts(g3, "ABC::", 2500, 5, 2)

    # ---- 6. search_bfs -----------------------------------------------------------------------------
    tb = CrcTree(grammar=g2, root="AB::", seed=0, n_exhausted=np.inf)
    tb.search_bfs(n_cycles=2)
    out["bfs2"] = {"graph_digest": graph_digest(tb.graph), "n_nodes": tb.graph.number_of_nodes(),
                   "order": list(tb.bfs_iterator())}
    tb = CrcTree(grammar=g2, root="AB::", seed=11, n_exhausted=np.inf)
    tb.search_bfs(n_steps=100, n_repeats=3, shuffle=True)
    out["bfs2_shuffle"] = {"graph_digest": graph_digest(tb.graph)}

    # ---- 7. random rollouts (rg.choice consumption) ------------------------------------------------
    tr = CrcTree(grammar=g3, root="ABC::", seed=42, n_exhausted=np.inf)
    out["rollouts3"] = [tr.get_random_terminal_descendant("ABC::") for _ in range(200)]
    tr = CrcTree(grammar=g5, root="ABCDE::", seed=42, n_exhausted=np.inf)
    out["rollouts5"] = [tr.get_random_terminal_descendant("ABCDE::") for _ in range(40)]

    # ---- 8. DimersGrammar: the parts that run upstream ---------------------
    dg = DimersGrammar(components=["A", "B"], regulators=["R"], interactions=["activates", "inhibits"])
    d = {}
    d["do_action"] = [[s, a, dg.do_action(s, a)] for s, a in [
        ("A+::", "B"), ("AB+::", "+R"), ("AB+R::", "ARiB"), ("AB+R::ARiB", "BBaB"), ("AB+R::ARiB_BBaB", "*terminate*")]]
    d["genotype_parts"] = [[s, list(dg.get_genotype_parts(s))] for s in ["*AB+R::ARiB_BBaB", "AB+::", "A+R::AAaA"]]
    pgd = {}
    for s in ["*AB+R::ARiB_BBaB", "*AB+R::RAiB_BBaB_ABaA", "AB+R::", "*BA+R::BRaA"]:
        c, r, a, i = dg.parse_genotype(s)
        pgd[s] = [c, r, a.tolist(), i.tolist()]
    d["parse_genotype"] = pgd
    d["dimer_options"] = dg.dimer_options
    d["edge_options_sorted"] = sorted(sorted(g) for g in dg.edge_options)
    d["actions_sorted"] = {s: sorted(dg.get_actions(s)) for s in
                           ["AB+R::ARiB", "AB+R::ARiB_BBaB", "AB+R::ARiB_BBaB_AAaA", "A+::AAaA", "AB+::ABiA", "*AB+R::ARiB"]}
    dg1 = DimersGrammar(components=["A", "B"], regulators=["R"], interactions=["activates", "inhibits"], max_interactions=2)
    d["actions_sorted_max2"] = {s: sorted(dg1.get_actions(s)) for s in ["AB+R::ARiB", "AB+R::ARiB_BBaB"]}
    d["to_dict"] = DimersGrammar(components=["A", "B"], regulators=["R"], interactions=["activates", "inhibits"]).to_dict()
    out["dimers"] = d

    out["_meta"] = {"circuitree_version": "0.11.1", "numpy": np.__version__, "python": sys.version.split()[0],
                    "reward": "crc32(state) % 1000 / 1000"}
    dump("reference_vectors", out, compress=True)

    # ---- 9. exhaustion with leaves in flight (separate fixture: batched_vectors.json.gz) -----------
    # The reference's parallel mode (circuitree.py:786-893) needs gevent, which is not installed; what
    # it does is interleave greenlets at get_reward.  The schedule below is the interleaving that
    # results when `batch` greenlets all block in get_reward together and `depth` such groups are
    # outstanding: the reference's OWN select_and_expand / rollout / backpropagate_visit run for a whole
    # group (virtual loss, :529-533) before any of its rewards is back-propagated (:534-535), in
    # selection order.  On ExhaustionError the complete groups still in flight are finished; the
    # partially selected group keeps its visits and never gets rewards (greenlets that died).
    def run_batched(grammar, root, n, seed, n_exhausted, batch, depth):
        tr = CrcTree(grammar=grammar, root=root, seed=seed, n_exhausted=n_exhausted)
        inflight, digests, err, selected = [], [], None, 0

        def retire():
            for path, sim in inflight.pop(0):
                reward = tr.get_reward(sim)
                tr.backpropagate_reward(path, reward)
                digests.append(step_digest(path, sim, reward))

        try:
            while selected < n:
                k = min(batch, n - selected)
                group = []
                for _ in range(k):
                    path = tr.select_and_expand()
                    sim = tr.get_random_terminal_descendant(path[-1])
                    tr.backpropagate_visit(path)
                    group.append((path, sim))
                    selected += 1
                inflight.append(group)
                while len(inflight) >= depth:
                    retire()
        except ExhaustionError:
            err = "ExhaustionError"
        while inflight:
            retire()
        return {"n_selected": selected, "n_done": len(digests), "digests": digests, "error": err,
                "n_nodes": tr.graph.number_of_nodes(), "n_edges": tr.graph.number_of_edges(),
                "graph_digest": graph_digest(tr.graph),
                "n_flagged": sum(bool(a.get("is_exhausted", False)) for _, a in tr.graph.nodes(data=True)),
                "min_node_visits": min(a["visits"] for _, a in tr.graph.nodes(data=True)),
                "root_visits": tr.graph.nodes[root]["visits"], "root_reward": tr.graph.nodes[root]["reward"]}

    g3m3 = SimpleNetworkGrammar(components=["A", "B", "C"], interactions=["activates", "inhibits"], max_interactions=3)
    bat = {
        "b8_d1_exh3": dict(args=["g3", "ABC::", 4000, 5, 3, 8, 1], res=run_batched(g3, "ABC::", 4000, 5, 3, 8, 1)),
        "b16_d3_exh4": dict(args=["g3", "ABC::", 6000, 6, 4, 16, 3], res=run_batched(g3, "ABC::", 6000, 6, 4, 16, 3)),
        "b8_d2_exh2_to_the_end": dict(args=["g3m3", "ABC::", 100000, 7, 2, 8, 2], res=run_batched(g3m3, "ABC::", 100000, 7, 2, 8, 2)),
        "b5_d2_two_components": dict(args=["g2", "AB::", 100000, 3, 3, 5, 2], res=run_batched(g2, "AB::", 100000, 3, 3, 5, 2)),
        "b1_d1_is_sequential": dict(args=["g3", "ABC::", 2500, 5, 2, 1, 1], res=run_batched(g3, "ABC::", 2500, 5, 2, 1, 1)),
    }
    assert bat["b1_d1_is_sequential"]["res"]["graph_digest"] == out["mcts3_exhaust"]["graph_digest"]
    bat["_meta"] = dict(out["_meta"], schedule="groups of `batch` leaves selected under virtual loss, `depth` groups in flight")
    dump("batched_vectors", bat, compress=True)


if __name__ == "__main__":
    main()
