"""One launch of 2^logn repressilator trajectories (used under ncu)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from circuitree_b200 import _capi

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = _capi.Engine(0)
cfg = _capi.default_config()
cfg.schedule = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cfg.lpt = int(sys.argv[4]) if len(sys.argv) > 4 else 1
h = _capi.compile_topology(3, [], [(0, 1), (1, 2), (2, 0)])
N = 1 << logn
d_reward = torch.zeros(N, device="cuda")
d_events = torch.zeros(N, dtype=torch.int32, device="cuda")
eng.warp_iterations()
for r in range(reps):
    st = torch.cuda.Event(enable_timing=True); en = torch.cuda.Event(enable_timing=True)
    st.record()
    eng.launch_rewards([h], [N], [0], cfg, 2 + r, torch.cuda.current_stream().cuda_stream, d_reward.data_ptr(), d_events=d_events.data_ptr())
    en.record(); torch.cuda.synchronize()
    ev = int(d_events.to(torch.int64).sum())
    it = eng.warp_iterations()
    print(f"lane utilisation {ev / (64 * it):.3f}  ({it} warp iterations)")
    print(f"N=2^{logn} ms {st.elapsed_time(en):.1f} events {ev} steps/s {ev / st.elapsed_time(en) * 1e3:.4e} reward {float(d_reward.mean()):.4f}")
