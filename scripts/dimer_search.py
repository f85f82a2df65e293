"""BASELINE config 3: DimersGrammar oscillation search on one B200.

    python scripts/dimer_search.py [--rollouts 100000] [--batch 8192] [--traj 8] [--depth 3]

Five transcription factors (three components A, B, C and two regulators D, E), activating and
inhibiting dimer interactions, at most two per promoter (the DimersGrammar default).  The tree is
the Python mirror of the reference (its DimersGrammar state strings and actions); every batch of
leaves is one launch of the dimer-model kernel (SPEC.md section 9), several batches in flight.
Prints one JSON line: rollouts/s, SSA reaction-steps/s, and the best terminal states found.
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from circuitree_b200 import DimersGrammar, OscillationTree

ap = argparse.ArgumentParser()
ap.add_argument("--rollouts", type=int, default=100_000)
ap.add_argument("--batch", type=int, default=8192)
ap.add_argument("--traj", type=int, default=32, help="SSA trajectories per rollout (a warp runs one topology: multiples of 32 fill it)")
ap.add_argument("--depth", type=int, default=3, help="reward batches in flight")
ap.add_argument("--components", default="ABC")
ap.add_argument("--regulators", default="DE")
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--native", action="store_true", help="select on the native (C++) DimersGrammar tree instead of the Python mirror")
args = ap.parse_args()

grammar = DimersGrammar(components=list(args.components), regulators=list(args.regulators),
                        interactions=["activates", "inhibits"])
root = f"{args.components}+{args.regulators}::"
if args.native:
    from circuitree_b200.native import NativeTree

    nt = NativeTree(grammar, root, seed=args.seed, n_exhausted=np.inf, trajectories_per_reward=args.traj)
    nt.search_mcts(min(256, args.rollouts), batch_size=256)                   # warm-up (engine, handles)
    eng = nt.reward_engine
    ev0, tr0, sl0 = eng.total_events, eng.total_trajectories, eng.total_issue_slots
    nt.timing = {k: 0.0 for k in nt.timing}
    t0 = time.perf_counter()
    nt.search_mcts(args.rollouts, batch_size=args.batch, pipeline_depth=args.depth)
    dt = time.perf_counter() - t0
    ex = nt.export()
    v, r = ex["node_visits"].astype(float), ex["node_reward"]
    term = ((ex["node_keys"] >> np.uint64(62)) & np.uint64(1)).astype(bool) & (v >= 20)
    order = np.argsort(-np.where(term, r / np.maximum(v, 1), -1.0))[:5]
    print(json.dumps({
        "config": f"DimersGrammar {root} activates/inhibits on the native tree, {args.rollouts} rollouts, batch {args.batch}, "
                  f"{args.traj} trajectories per rollout, {args.depth} batches in flight",
        "seconds": round(dt, 2), "rollouts_per_s": args.rollouts / dt, "reaction_steps_per_s": (eng.total_events - ev0) / dt,
        "issue_roofline_frac": (eng.total_issue_slots - sl0) / dt / (eng.engine.n_sm * 128 * 1.965e9),
        "trajectories": eng.total_trajectories - tr0, "events_per_trajectory": (eng.total_events - ev0) / max(1, eng.total_trajectories - tr0),
        "tree_nodes": int(len(v)), "host_seconds": {k: round(x, 2) for k, x in nt.timing.items()},
        "best": [[nt.key_to_string(int(ex["node_keys"][i])), round(float(r[i] / v[i]), 4), int(v[i])] for i in order if term[i]],
    }))
    sys.exit(0)
tree = OscillationTree(grammar=grammar, root=root, seed=args.seed, n_exhausted=np.inf, trajectories_per_reward=args.traj)
tree.search_mcts_batched(min(256, args.rollouts), 256)                      # warm-up (engine, handles)
eng = tree.reward_engine
ev0, tr0, v0, sl0 = eng.total_events, eng.total_trajectories, tree.graph.nodes[root]["visits"], eng.total_issue_slots
t0 = time.perf_counter()


def progress(t, step, path, sim, reward):
    if step >= 0 and (step + 1) % args.batch == 0:
        el = time.perf_counter() - t0
        print(f"[{el:7.1f}s] {step + 1} rollouts back-propagated, {eng.total_events - ev0:.3e} events, "
              f"{(eng.total_events - ev0) / max(1, eng.total_trajectories - tr0):.3e} events/trajectory", file=sys.stderr, flush=True)


tree.search_mcts_batched(args.rollouts, args.batch, pipeline_depth=args.depth, callback=progress, callback_before_start=False)
dt = time.perf_counter() - t0
terminals = [(n, d["reward"] / d["visits"], d["visits"]) for n, d in tree.graph.nodes(data=True)
             if isinstance(n, str) and n.startswith("*") and d["visits"] >= 20]
terminals.sort(key=lambda x: -x[1])
print(json.dumps({
    "config": f"DimersGrammar {root} activates/inhibits, {args.rollouts} rollouts, batch {args.batch}, "
              f"{args.traj} trajectories per rollout, {args.depth} batches in flight",
    "seconds": round(dt, 2), "rollouts_per_s": args.rollouts / dt,
    "reaction_steps_per_s": (eng.total_events - ev0) / dt, "trajectories": eng.total_trajectories - tr0,
    "issue_roofline_frac": (eng.total_issue_slots - sl0) / dt / (eng.engine.n_sm * 128 * 1.965e9),   # SPEC.md section 8 dimer I_step
    "events_per_trajectory": (eng.total_events - ev0) / max(1, eng.total_trajectories - tr0),
    "tree_nodes": tree.graph.number_of_nodes(), "root_visits": tree.graph.nodes[root]["visits"] - v0,
    "distinct_topologies_simulated": len(eng._handles),
    "best": [[n, round(r, 4), v] for n, r, v in terminals[:5]],
}))
