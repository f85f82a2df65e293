"""Where does an MCTS reward batch spend its time?  (mixed-topology launch anatomy)"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from circuitree_b200 import SimpleNetworkGrammar, _capi
from circuitree_b200.native import NativeTree

K = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
g = SimpleNetworkGrammar(list("ABC"), ["activates", "inhibits"], root="ABC::")
t = NativeTree(g, "ABC::", seed=1, n_exhausted=np.inf, trajectories_per_reward=B)
t.search_mcts(1024, batch_size=1024)
for rep in range(2):
    t0 = time.perf_counter(); keys, handles = t.select(K); t1 = time.perf_counter()
    eng = t.reward_engine
    uniq, inv, counts = np.unique(handles, return_inverse=True, return_counts=True)
    cost = np.array([eng.cost.get(int(h), np.nan) for h in uniq])
    print(f"select {t1 - t0:.3f}s  leaves {K} distinct topologies {len(uniq)} known cost {np.isfinite(cost).sum()}")
    for sched in (2, 1):
        eng.cfg.schedule = sched
        ev0 = eng.total_events
        t2 = time.perf_counter(); r = t._rewards(keys, handles); t3 = time.perf_counter()
        ev = eng.total_events - ev0
        print(f"  schedule {sched}: rewards {t3 - t2:.3f}s events {ev:.3e} -> {ev / (t3 - t2):.3e} steps/s, {K / (t3 - t2):.0f} rollouts/s")
    eng.cfg.schedule = 0
    t.backprop(r)

# pipelined search: several batches in flight
for depth in (1, 2, 3, 4):
    t = NativeTree(g, "ABC::", seed=2, n_exhausted=np.inf, trajectories_per_reward=B)
    t.search_mcts(1024, batch_size=1024)
    ev0 = t.reward_engine.total_events
    n = 6 * K // 2
    t0 = time.perf_counter(); t.search_mcts(n, batch_size=K // 2, pipeline_depth=depth); dt = time.perf_counter() - t0
    ev = t.reward_engine.total_events - ev0
    print(f"depth {depth}: {n} rollouts in {dt:.2f}s -> {n / dt:.0f} rollouts/s, {ev / dt:.3e} steps/s, host {t.timing}")
