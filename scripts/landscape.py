"""BASELINE config 5: reward landscape of every 3-component topology (and, sampled, 4-component ones).

    python scripts/landscape.py [--n 10000] [--components 3] [--limit K]
    torchrun --nproc-per-node 8 scripts/landscape.py --param-sets 10000        # sharded, no collective

Enumerates the terminal states natively (ct_tree_grow), shards them round-robin over ranks and
simulates n parameter sets per topology on the GPU; prints throughput and the top topologies.
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from circuitree_b200 import RewardEngine, SimpleNetworkGrammar, landscape_motifs, reward_landscape
from circuitree_b200.native import NativeTree

ap = argparse.ArgumentParser()
ap.add_argument("--n", "--param-sets", dest="n", type=int, default=10000, help="parameter sets per topology (use --param-sets under torchrun: its own parser claims --n)")
ap.add_argument("--components", type=int, default=3)
ap.add_argument("--limit", type=int, default=0, help="only the first K terminal states (0 = all)")
ap.add_argument("--states-file", default="", help="gzipped list of terminal states to simulate instead of the enumeration "
                                                  "(profiles/landscape4_sample_states.txt.gz: scripts/make_landscape4_sample.py)")
ap.add_argument("--no-motifs", action="store_true", help="skip the motif-enrichment table")
ap.add_argument("--out", default="", help="write this rank's JSON line to <out>.rank<r>.json as well")
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
comps = list("ABCDE"[: args.components])
grammar = SimpleNetworkGrammar(comps, ["activates", "inhibits"], root="".join(comps) + "::")
if args.states_file:
    import gzip
    terminals = gzip.open(args.states_file, "rt").read().split()
    args.components = len(terminals[0].lstrip("*").split("::")[0])
    comps = list("ABCDE"[: args.components])
    grammar = SimpleNetworkGrammar(comps, ["activates", "inhibits"], root="".join(comps) + "::")
    t_enum = 0.0
else:
    tree = NativeTree(grammar, grammar.root, seed=0, reward_fn=lambda s: [0.0] * len(s))
    t0 = time.perf_counter(); tree.grow_tree(); t_enum = time.perf_counter() - t0
    keys = tree.export()["node_keys"]
    terminals = sorted(tree.key_to_string(int(k)) for k in keys if (int(k) >> 62) & 1)
if args.limit:
    terminals = terminals[: args.limit]
eng = RewardEngine(grammar, device=int(os.environ.get("LOCAL_RANK", 0)), seed=0)
t0 = time.perf_counter()
res = reward_landscape(eng, terminals, args.n, rank=rank, world=world)
dt = time.perf_counter() - t0
best = np.argsort(-res["mean"])[:5]
# which motifs are enriched among the (topology, parameter set) pairs that oscillate? (this rank's share)
MOTIFS = ["ABi_BCi_CAi", "ABi_BAi", "AAa", "AAi", "ABa_BAi", "ABa_BCa_CAa", "AAa_ABi_BAi", "ABi_BCi_CAa"]
motifs = None if args.no_motifs else landscape_motifs(grammar, terminals, res, MOTIFS, args.n)
line = json.dumps({"rank": rank, "world": world, "terminal_states": len(terminals), "enumeration_s": round(t_enum, 3),
                  "states_this_rank": len(res["index"]), "trajectories": len(res["index"]) * args.n,
                  "events": int(res["events"].sum()), "seconds": round(dt, 2),
                  "reaction_steps_per_s": float(res["events"].sum() / dt),
                  "top": [[terminals[res["index"][i]], round(float(res["mean"][i]), 4), round(float(res["success"][i]), 3)] for i in best],
                  "motif_odds_ratios": None if motifs is None else
                  {r["pattern"]: round(float(r["odds_ratio"]), 3) for _, r in motifs.iterrows()}})
print(line)
if args.out:
    open(f"{args.out}.rank{rank}.json", "w").write(line + "\n")
