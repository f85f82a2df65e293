"""First GPU sanity run: reward-stage parity, SSA statistics vs the CPU checker, raw timing."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import torch
from circuitree_b200 import _capi
from oracle import ssa_oracle as so

eng = _capi.Engine(0)
cfg = _capi.default_config()
print("n_sm", eng.n_sm, "resident lanes", eng.resident_lanes)
ranges = cfg.ranges_array()
inh = [(0, 1), (1, 2), (2, 0)]
h = _capi.compile_topology(3, [], inh)
print("handle", hex(h))

# 1. reward stage on oracle-generated series
o = so.run_pairwise(3, [], inh, ranges, 512, seed=1, n_threads=so.max_threads(), want_series=True)
rw, lg = eng.reward_from_series(o["series"])
print("reward stage: max |d reward|", np.abs(rw - o["reward"]).max(), "lag mismatches", int((lg != o["lag"]).sum()))
rs = np.random.default_rng(0).integers(0, 256, size=(256, 3, 125), dtype=np.uint8)
rw2, lg2 = eng.reward_from_series(rs)
orw = np.array([so.reward_from_series(x) for x in rs])
print("random series: max |d reward|", np.abs(rw2 - orw[:, 0]).max(), "lag mismatches", int((lg2 != orw[:, 1]).sum()))

# 2. SSA statistics
N = 4096
t0 = time.time()
g = eng.rewards_host([h], [N], [0], cfg, seed=1, per_trajectory=True)
t1 = time.time()
print("gpu N", N, "time", t1 - t0, "mean reward", g["reward"].mean(), "events/traj", g["job_events"][0] / N)
t0 = time.time()
o = so.run_pairwise(3, [], inh, ranges, N, seed=1, n_threads=so.max_threads())
t1 = time.time()
print("cpu N", N, "time", t1 - t0, "threads", so.max_threads(), "mean reward", o["reward"].mean(), "events/traj", o["n_events"].mean(),
      "steps/s", o["n_events"].sum() / (t1 - t0))
se = np.sqrt(g["reward"].var() / N + o["reward"].var() / N)
print("delta mean reward / se", (g["reward"].mean() - o["reward"].mean()) / se)
from scipy import stats
print("KS lag", stats.ks_2samp(g["lag"], o["lag"]))
print("KS reward", stats.ks_2samp(g["reward"], o["reward"]))

# 3. series dump parity of the first grid block: identical initial conditions
d_reward = torch.zeros(256, device="cuda"); d_series = torch.zeros(256, 3 * 125, dtype=torch.uint8, device="cuda")
eng.launch_rewards([h], [256], [0], cfg, 1, 0, d_reward.data_ptr(), d_series=d_series.data_ptr(), series_stride=375)
torch.cuda.synchronize()
o2 = so.run_pairwise(3, [], inh, ranges, 256, seed=1, n_threads=8, want_series=True)
gs = d_series.cpu().numpy().reshape(256, 3, 125)
print("series[:, :, 0] equal frac", (gs[:, :, 0] == o2["series"][:, :, 0]).mean(), " mean |d| over all", np.abs(gs.astype(int) - o2["series"].astype(int)).mean())
print("mean series gpu", gs.mean(), "cpu", o2["series"].mean())

# 4. timing
for logn in (14, 16, 18, 20):
    N = 1 << logn
    d_reward = torch.zeros(N, device="cuda"); d_events = torch.zeros(N, dtype=torch.int32, device="cuda")
    st = torch.cuda.Event(enable_timing=True); en = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    st.record()
    eng.launch_rewards([h], [N], [0], cfg, 2, torch.cuda.current_stream().cuda_stream, d_reward.data_ptr(), d_events=d_events.data_ptr())
    en.record(); torch.cuda.synchronize()
    ms = st.elapsed_time(en)
    ev = int(d_events.to(torch.int64).sum())
    print(f"N=2^{logn} ms {ms:.1f} events {ev} steps/s {ev / ms * 1e3:.3e} mean reward {float(d_reward.mean()):.4f}")
