"""Times the pairwise kernel variants a build offers: the single-regulator loop and the general loop on
the repressilator, a mixed-lane batch of random 3-component terminals and one of random 5-component
terminals (MCTS-like: `traj` trajectories each).  usage: pairwise_probe.py [log2 n] [leaves] [traj]"""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from circuitree_b200 import CircuiTree, SimpleNetworkGrammar, _capi

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 21
leaves = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
traj = int(sys.argv[3]) if len(sys.argv) > 3 else 32
eng = _capi.Engine(0)
print("lib", _capi.library_path().name, "resident lanes", eng.resident_lanes, flush=True)
h = _capi.compile_topology(3, [], [(0, 1), (1, 2), (2, 0)])


def timed(handles, counts, cfg, label, reps=2):
    total = int(np.sum(counts))
    d_reward = torch.zeros(total, device="cuda")
    d_events = torch.zeros(total, dtype=torch.int32, device="cuda")
    bases = np.arange(len(handles), dtype=np.uint64) * (1 << 24)
    best = 0.0
    for r in range(reps):
        eng.warp_iterations()
        st = torch.cuda.Event(enable_timing=True); en = torch.cuda.Event(enable_timing=True)
        st.record()
        eng.launch_rewards(handles, counts, bases + np.uint64(r * (1 << 40)), cfg, 7, torch.cuda.current_stream().cuda_stream,
                           d_reward.data_ptr(), d_events=d_events.data_ptr())
        en.record(); torch.cuda.synchronize()
        ev = int(d_events.to(torch.int64).sum())
        it = eng.warp_iterations()
        ms = st.elapsed_time(en)
        best = max(best, ev / ms * 1e3)
        print(f"{label}: {ms:8.1f} ms {ev / ms * 1e3:.4e} steps/s util {ev / (64 * max(it, 1)):.3f} reward {float(d_reward.mean()):.4f}", flush=True)
    return best


cfg = _capi.default_config()
for k in (8, 12, 14, 16):            # small launches: one-lane kernels against the cooperative one
    for sched in (1, 3):
        c = _capi.default_config(); c.schedule = sched
        timed([h], [1 << k], c, f"repressilator 2^{k} schedule {sched}", reps=1)
timed([h], [1 << logn], cfg, f"fast loop 2^{logn}")
gen = _capi.default_config(); gen.fast_path = 0
timed([h], [1 << logn], gen, f"general loop 2^{logn}")


class T(CircuiTree):
    def get_reward(self, state, **kw):
        return 0.0


for comps in ("ABC", "ABCDE"):
    g = SimpleNetworkGrammar(list(comps), ["activates", "inhibits"], root=comps + "::")
    t = T(grammar=g, root=comps + "::", seed=3, n_exhausted=np.inf)
    hs = []
    for _ in range(leaves):
        c, a, i = g.parse_genotype(t.get_random_terminal_descendant(comps + "::"))
        hs.append(_capi.compile_topology(len(c), a, i))
    uniq, counts = np.unique(np.array(hs, dtype=np.uint64), return_counts=True)
    for sched in (2, 3):
        c = _capi.default_config(); c.schedule = sched
        for frac in (8, 1):
            sel = slice(0, max(1, len(uniq) // frac))
            timed(uniq[sel], (counts[sel] * traj).astype(np.uint32), c,
                  f"mixed batch {comps}: {int(counts[sel].sum())} leaves x {traj} ({len(uniq[sel])} topologies) schedule {sched}", reps=1)
