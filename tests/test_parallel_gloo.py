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
    tree = NativeTree(SimpleNetworkGrammar(list("ABC"), ["activates", "inhibits"]), "ABC::", seed=100 + rank,
                      n_exhausted=np.inf, reward_fn=crc)
    search = RootParallelSearch(tree, sync_every=64, batch_size=16)
    search.run(256)
    g = tree.to_networkx()
    h = hashlib.sha256()
    for n in sorted(g.nodes):
        h.update(f"{n}|{g.nodes[n]['visits']}|{g.nodes[n]['reward']:.9f}\\n".encode())
    for u, v in sorted(g.edges):
        h.update(f"{u}>{v}|{g.edges[u, v]['visits']}|{g.edges[u, v]['reward']:.9f}\\n".encode())
    out = {"rank": rank, "root_visits": g.nodes["ABC::"]["visits"], "digest": h.hexdigest(), "syncs": search.n_syncs,
           "nodes": g.number_of_nodes(), "bytes": search.bytes_exchanged,
           "edge_sum": sum(g.edges["ABC::", c]["visits"] for c in g.successors("ABC::"))}
    open(os.path.join(os.environ["CT_TEST_OUT"], f"rank{rank}.json"), "w").write(json.dumps(out))
    dist.destroy_process_group()
""") % str(ROOT)


def test_root_parallel_merge_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1", CT_TEST_OUT=str(tmp_path))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    import json

    results = [json.loads((tmp_path / f"rank{r}.json").read_text()) for r in range(2)]
    a, b = sorted(results, key=lambda r: r["rank"])
    # both replicas hold the sum of both ranks' work and agree on every statistic
    assert a["root_visits"] == b["root_visits"] == 512
    assert a["edge_sum"] == 512
    assert a["digest"] == b["digest"]
    assert a["syncs"] == b["syncs"] == 4 and a["bytes"] > 0
