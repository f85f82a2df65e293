"""Draws the seeded, uniform sample of four-component terminal topologies used by the 8-GPU landscape
run (BASELINE config 5): enumerates the whole ABCD space natively (3,603,807 states, 1,801,903
terminals, ~35 s) and writes `--count` of the terminals, chosen without replacement by
numpy default_rng(`--seed`), to profiles/landscape4_sample_states.txt.gz (one state string per line,
sorted).  Run here once; the GPU box only reads the file."""
import argparse, gzip, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from circuitree_b200 import SimpleNetworkGrammar
from circuitree_b200.native import NativeTree

ap = argparse.ArgumentParser()
ap.add_argument("--count", type=int, default=10_000)
ap.add_argument("--seed", type=int, default=2024)
args = ap.parse_args()
g = SimpleNetworkGrammar(list("ABCD"), ["activates", "inhibits"], root="ABCD::")
t = NativeTree(g, "ABCD::", seed=0, reward_fn=lambda s: [0.0] * len(s))
t.grow_tree()
keys = t.export()["node_keys"]
terminals = np.sort(keys[((keys >> np.uint64(62)) & np.uint64(1)) == 1])
pick = np.sort(np.random.default_rng(args.seed).choice(len(terminals), size=args.count, replace=False))
states = sorted(t.key_to_string(int(terminals[i])) for i in pick)
out = Path(__file__).resolve().parents[1] / "profiles" / "landscape4_sample_states.txt.gz"
with gzip.GzipFile(out, "wb", mtime=0) as f:
    f.write(("\n".join(states) + "\n").encode())
print(f"{len(terminals)} terminals of {len(keys)} states; wrote {len(states)} to {out}")
