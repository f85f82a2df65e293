"""World-size-2 gloo test of the root-parallel MCTS merge (CPU only, injected rewards)."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

WORKER = textwrap.dedent("""
    import os, sys, zlib, json, hashlib
    sys.path.insert(0, %r)
    import numpy as np
    import torch.distributed as dist
    from circuitree_b200 import SimpleNetworkGrammar
    from circuitree_b200.native import NativeTree
    from circuitree_b200.parallel import RootParallelSearch

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    crc = lambda states: [(zlib.crc32(s.encode()) %% 1000) / 1000.0 for s in states]
    n_exh = float(os.environ["CT_TEST_NEXH"])
    max_ixn = int(os.environ["CT_TEST_MAXIXN"])
    ROOT_STATE = os.environ.get("CT_TEST_ROOT", "ABC::")
    if "+" in ROOT_STATE:       # DimersGrammar: components + regulators
        from circuitree_b200 import DimersGrammar
        comps, regs = ROOT_STATE[:-2].split("+")
        grammar = DimersGrammar(components=list(comps), regulators=list(regs), interactions=["activates", "inhibits"])
    else:
        grammar = SimpleNetworkGrammar(list("ABC"), ["activates", "inhibits"], max_interactions=max_ixn)
    tree = NativeTree(grammar, ROOT_STATE, seed=100 + rank, n_exhausted=n_exh, reward_fn=crc)
    search = RootParallelSearch(tree, sync_every=64, batch_size=16, pipeline_depth=int(os.environ["CT_TEST_DEPTH"]))
    n_done = search.run(int(os.environ["CT_TEST_STEPS"]))
    g = tree.to_networkx()
    h = hashlib.sha256()
    for n in sorted(g.nodes):
        h.update(f"{n}|{g.nodes[n]['visits']}|{g.nodes[n]['reward']:.9f}\\n".encode())
    for u, v in sorted(g.edges):
        h.update(f"{u}>{v}|{g.edges[u, v]['visits']}|{g.edges[u, v]['reward']:.9f}\\n".encode())
    ex = tree.export()
    out = {"rank": rank, "root_visits": g.nodes[ROOT_STATE]["visits"], "digest": h.hexdigest(), "syncs": search.n_syncs,
           "nodes": g.number_of_nodes(), "bytes": search.bytes_exchanged, "n_done": n_done, "exhausted": search.exhausted,
           "min_visits": int(min(ex["node_visits"].min(), ex["edge_visits"].min())),
           "n_flagged": int(ex["node_exhausted"].sum()),
           "edge_sum": sum(g.edges[ROOT_STATE, c]["visits"] for c in g.successors(ROOT_STATE))}
    open(os.path.join(os.environ["CT_TEST_OUT"], f"rank{rank}.json"), "w").write(json.dumps(out))
    dist.destroy_process_group()
""") % str(ROOT)


def run_world2(tmp_path, port, n_exhausted="inf", max_ixn=9, steps=256, depth=1, root="ABC::"):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1", CT_TEST_OUT=str(tmp_path), CT_TEST_NEXH=str(n_exhausted),
               CT_TEST_MAXIXN=str(max_ixn), CT_TEST_STEPS=str(steps), CT_TEST_DEPTH=str(depth), CT_TEST_ROOT=root)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    import json

    results = [json.loads((tmp_path / f"rank{r}.json").read_text()) for r in range(2)]
    return sorted(results, key=lambda r: r["rank"])


def test_root_parallel_merge_world2(tmp_path):
    a, b = run_world2(tmp_path, 29517, depth=2)      # batches stay in flight across the exchanges
    # both replicas hold the sum of both ranks' work and agree on every statistic
    assert a["root_visits"] == b["root_visits"] == 512
    assert a["edge_sum"] == 512
    assert a["digest"] == b["digest"]
    assert a["syncs"] == b["syncs"] == 4 and a["bytes"] > 0
    assert a["n_done"] == b["n_done"] == 256


def test_root_parallel_merge_world2_finite_n_exhausted(tmp_path):
    """Finite n_exhausted (reference circuitree.py:377-420) across replicas: exhaustion events are
    exchanged and replayed, statistics never go negative, both replicas end exhausted in the same
    exchange, agree on every statistic, and keep the collectives matched until the planned end."""
    a, b = run_world2(tmp_path, 29519, n_exhausted=4, max_ixn=3, steps=64 * 400)
    assert a["exhausted"] and b["exhausted"]
    assert a["syncs"] == b["syncs"] == 400
    assert a["min_visits"] >= 0 and b["min_visits"] >= 0
    assert a["digest"] == b["digest"] and a["n_flagged"] == b["n_flagged"] == a["nodes"]
    assert 0 < a["n_done"] < 64 * 400


def test_root_parallel_merge_world2_dimers_grammar(tmp_path):
    """The same exchange on DimersGrammar trees (five TFs): node keys are process-local there, so the interaction
    lists travel; both replicas end with the sum of both ranks' work, state by state."""
    a, b = run_world2(tmp_path, 29521, depth=2, root="ABCD+E::")
    assert a["root_visits"] == b["root_visits"] == 512 and a["edge_sum"] == 512
    assert a["digest"] == b["digest"] and a["nodes"] == b["nodes"] > 100
    assert a["syncs"] == b["syncs"] == 4 and a["bytes"] > 0
