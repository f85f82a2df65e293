"""Which dimer kernel for which launch?  Times schedule 1 (one topology per warp), 2 (per-lane programs)
and 3 (cooperative sub-groups) on (a) one topology at 2^k trajectories and (b) MCTS-like batches of
random five-TF DimersGrammar terminals x `traj` trajectories each.  Checks the results are identical.

    python scripts/dimer_sched_probe.py [--nt 2000] [--sizes 10,12,14,16] [--leaves 256,1024] [--traj 32]
"""
import argparse, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from circuitree_b200 import CircuiTree, DimersGrammar, _capi

ap = argparse.ArgumentParser()
ap.add_argument("--nt", type=int, default=2000)
ap.add_argument("--sizes", default="8,10,12,14,16")
ap.add_argument("--leaves", default="256,1024")
ap.add_argument("--traj", type=int, default=32)
ap.add_argument("--schedules", default="1,2,3")
ap.add_argument("--wide-sizes", default="", help="launch sizes (log2) for the five-TF topology (default: --sizes)")
args = ap.parse_args()
AI = ["activates", "inhibits"]
eng = _capi.Engine(0)
scheds = [int(x) for x in args.schedules.split(",")]


def handle(g, s):
    comps, regs, act, inh = g.parse_genotype(s)
    return _capi.compile_topology(len(comps) + len(regs), act, inh, k=3)


def run(handles, counts, schedule):
    cfg = _capi.default_config()
    cfg.nt, cfg.schedule = args.nt, schedule
    bases = np.arange(len(handles), dtype=np.uint64) * 100_000
    eng.warp_iterations()
    t0 = time.perf_counter()
    out = eng.rewards_host(handles, counts, bases, cfg, 3, per_trajectory=True)
    dt = time.perf_counter() - t0
    it = eng.warp_iterations()
    ev = int(out["job_events"].sum())
    return dt, ev, ev / (64 * max(it, 1)), out


G3 = DimersGrammar(components=list("ABC"), regulators=[], interactions=AI)
G5 = DimersGrammar(components=list("ABC"), regulators=list("DE"), interactions=AI)
one = [(G3, "*ABC+::AAiB_BBiC_CCiA"), (G5, "*ABC+DE::ADiB_BEiC_CCiA_ABaA_DEaB_BCaC")]
run([handle(*one[0])], [64], 1)   # warm-up
for g, s in one:
    h = handle(g, s)
    sizes = args.wide_sizes if (g is G5 and args.wide_sizes) else args.sizes
    for k in [int(x) for x in sizes.split(",")]:
        ref = None
        for sc in scheds:
            dt, ev, util, out = run([h], [1 << k], sc)
            same = "" if ref is None else (" identical" if np.array_equal(ref["reward"], out["reward"]) and np.array_equal(ref["job_events"], out["job_events"]) else " DIFFERENT")
            ref = ref or out
            print(f"{s} 2^{k} schedule {sc}: {dt*1e3:9.1f} ms {ev/dt:.3e} steps/s util {util:.3f}{same}", flush=True)


class T(CircuiTree):
    def get_reward(self, state, **kw):
        return 0.0


t = T(grammar=G5, root="ABC+DE::", seed=1, n_exhausted=np.inf)
for n_leaves in [int(x) for x in args.leaves.split(",")]:
    states = [t.get_random_terminal_descendant("ABC+DE::") for _ in range(n_leaves)]
    hs = np.array([handle(G5, s) for s in states], dtype=np.uint64)
    uniq, counts = np.unique(hs, return_counts=True)
    ref = None
    for sc in scheds:
        dt, ev, util, out = run(uniq, counts * args.traj, sc)
        same = "" if ref is None else (" identical" if np.array_equal(ref["reward"], out["reward"]) else " DIFFERENT")
        ref = ref or out
        print(f"batch of {n_leaves} random leaves x {args.traj} ({len(uniq)} topologies) schedule {sc}: {dt*1e3:9.1f} ms "
              f"{ev/dt:.3e} steps/s {n_leaves/dt:.1f} rollouts/s util {util:.3f}{same}", flush=True)
