"""One launch of the dimer kernel (homodimer repressilator, SPEC.md section 9) for ncu / timing.
usage: prof_dimers.py [log2 n] [nt] [reps] [big|small] [schedule]
(big = five TFs, ten interactions, ten dimers; schedule 1 one topology per warp, 2 per lane, 3 cooperative sub-groups)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from circuitree_b200 import _capi

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 12
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 320
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
eng = _capi.Engine(0)
cfg = _capi.default_config()
cfg.nt = nt
cfg.schedule = int(sys.argv[5]) if len(sys.argv) > 5 else 0
if len(sys.argv) > 4 and sys.argv[4] == "big":
    # *ABCDE+::AAiB_BBiC_CCiD_DDiE_EEiA_ABaA_BCaB_CDaC_DEaD_AEaE
    h = _capi.compile_topology(5, [(0, 1, 0), (1, 2, 1), (2, 3, 2), (3, 4, 3), (0, 4, 4)],
                               [(0, 0, 1), (1, 1, 2), (2, 2, 3), (3, 3, 4), (4, 4, 0)], k=3)
else:
    h = _capi.compile_topology(4, [], [(0, 0, 1), (1, 1, 2), (2, 2, 0)], k=3)
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
    print(f"dimers N=2^{logn} nt={nt} ms {st.elapsed_time(en):.1f} events {ev} steps/s {ev / st.elapsed_time(en) * 1e3:.4e} "
          f"lane utilisation {ev / (64 * max(it, 1)):.3f} reward {float(d_reward.mean()):.4f}")
