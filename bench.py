"""Headline benchmark: batched Gillespie SSA reaction-steps/s (BASELINE.json metric, config 2) and,
in the same JSON line, every other BASELINE configuration at a bounded size.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--log2-traj T] [--impl reference]

One "step" = one pass of the hot path (SSA + in-kernel reward) over one batch of 2^T repressilator
trajectories x 2000 grid points per GPU, parameters drawn per trajectory from the SPEC.md ranges.
N > 1 is launched by torchrun (one rank per GPU); trajectories are independent, so ranks shard
trajectory ids with no data-path collective ("weak" scaling: per-GPU work fixed).

`value` times the kernel with inputs resident in HBM (CUDA events on the launch stream); `e2e`
times the same batches through the public host-buffer API (RewardEngine.run -> ct_rewards_host):
explicit parameter vectors copied host->device, per-trajectory rewards and lags copied back, over
all K steps.  Further keys of the line (each with its own config):
  sweep      config 2: 2^14 .. 2^22 trajectories of the same topology, steps/s, lane utilisation, roofline fraction
  variants   the general hot loop (what arbitrary topologies run) beside the single-regulator one
  mcts3      config 1: 3-component MCTS, 10^4 iterations per GPU, rollouts/s (+ CPU baseline on rank 0)
  mcts       config 4: root-parallel 5-component MCTS with periodic NCCL statistic merges inside the
             timed region, and the same budget without merges as the search-quality control
  dimers     config 3's kernel on one big launch; `dimer_search` a bounded DimersGrammar search
  landscape  config 5: a seeded sample of the exhaustive 3-component landscape, sharded over ranks
`--impl reference` times the CPU restatement (oracle/, all host threads) on a bounded sample; the
upstream package has no simulator of its own (SPEC.md), so that port is the only CPU path there is.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

REPRESSILATOR = "*ABC::ABi_BCi_CAi"
DIMER_REPRESSILATOR = "*ABC+::AAiB_BBiC_CCiA"
METRIC = "gillespie_reaction_steps_per_s"
UNIT = "reaction-steps/s"


def workload_name(log2_traj: int, nt: int = 2000) -> str:
    """The benchmark workload, named identically by both arms."""
    return (f"repressilator {REPRESSILATOR}: 2^{log2_traj} trajectories x {nt} grid points per GPU per step, "
            f"parameters sampled per trajectory (SPEC.md ranges)")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows: list[list[str]] = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
            except ValueError:
                continue
            for name, flag in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(n_traj: int, seed: int, threads: int):
    """Time the CPU restatement on `n_traj` repressilator trajectories."""
    from circuitree_b200 import SimpleNetworkGrammar, _capi_defaults
    from oracle import ssa_oracle

    comps, act, inh = SimpleNetworkGrammar.parse_genotype(REPRESSILATOR)
    d = _capi_defaults.SPEC_DEFAULTS
    t0 = time.perf_counter()
    out = ssa_oracle.run_pairwise(len(comps), act, inh, d["ranges"], n_traj, seed=seed, nt=d["nt"], decim=d["decim"],
                                  dt=d["dt"], init_mean=d["init_mean"], n_threads=threads)
    dt = time.perf_counter() - t0
    return int(out["n_events"].sum()), dt, float(out["reward"].mean())


def cpu_mcts_sample(n_iter: int, n_traj: int, threads: int) -> dict:
    """BASELINE.md plan item 1: sequential `search_mcts` (reference circuitree.py:693-784) on one
    host thread with the CPU SSA as its reward (`n_traj` trajectories per rollout on `threads` threads).
    /root/reference is not on the GPU box, so the tree is this repo's Python mirror of CircuiTree
    (bit-exact with the reference on the golden traces); the reward is the oracle (checker) code."""
    from circuitree_b200 import CircuiTree, SimpleNetworkGrammar, _capi_defaults
    from oracle import ssa_oracle

    d = _capi_defaults.SPEC_DEFAULTS
    counter = {"events": 0, "next": 0}

    class CpuOscillationTree(CircuiTree):
        def get_reward(self, state, **kw):
            comps, act, inh = SimpleNetworkGrammar.parse_genotype(state)
            out = ssa_oracle.run_pairwise(len(comps), act, inh, d["ranges"], n_traj, seed=4321, traj_base=counter["next"],
                                          nt=d["nt"], decim=d["decim"], dt=d["dt"], init_mean=d["init_mean"], n_threads=threads)
            counter["next"] += n_traj
            counter["events"] += int(out["n_events"].sum())
            return float(out["reward"].mean())

    grammar = SimpleNetworkGrammar(["A", "B", "C"], ["activates", "inhibits"], root="ABC::")
    tree = CpuOscillationTree(grammar=grammar, root="ABC::", seed=0, n_exhausted=np.inf)
    t0 = time.perf_counter()
    tree._run_mcts(tree, range(n_iter))
    dt = time.perf_counter() - t0
    return {"value": n_iter / dt, "unit": "rollouts/s", "cores": threads, "kind": "port",
            "reaction_steps_per_s": counter["events"] / dt,
            "sample": f"{n_iter} sequential search_mcts iterations on ABC::, {n_traj} CPU SSA trajectories per rollout "
                      f"({counter['events']} events in {dt:.1f} s); tree = this repo's Python mirror of the reference's "
                      "CircuiTree (the reference is not on the GPU box), reward = oracle/ CPU restatement"}


def run_reference(args) -> None:
    """CPU arm: bounded sample per step on all host threads (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ssa_oracle

    ssa_oracle.build()
    threads = ssa_oracle.max_threads()
    n_sample = args.cpu_traj
    for w in range(args.warmup):
        cpu_sample(max(64, n_sample // 8), 1000 + w, threads)
    events, secs = 0, 0.0
    for k in range(args.steps):
        ev, dt, _ = cpu_sample(n_sample, 2000 + k, threads)
        events += ev
        secs += dt
    value = events / secs
    sample = f"{n_sample} repressilator trajectories x 2000 grid points per step, {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.log2_traj),
                   "sample": f"bounded CPU sample of that workload: {sample}",
                   "note": "upstream circuitree ships no SSA; this is the repo's CPU restatement of SPEC.md (oracle/): a plain "
                           "direct method that re-evaluates every propensity each step, gcc -O2, one trajectory per thread"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def tree_quality(tree, root: str, min_visits: int = 48) -> dict:
    """What the search has found so far, from one replica's statistics."""
    ex = tree.export()
    v, r = ex["node_visits"].astype(np.float64), ex["node_reward"]
    root_id = int(np.nonzero(ex["node_keys"] == np.uint64(tree.string_to_key(root)))[0][0])
    ok = v >= min_visits
    ok[root_id] = False
    best = int(np.argmax(np.where(ok, r / np.maximum(v, 1.0), -1.0))) if ok.any() else root_id
    return {"mean_rollout_reward": float(r[root_id] / max(v[root_id], 1.0)), "root_visits": int(v[root_id]),
            "tree_nodes": int(len(v)), "nodes_with_min_visits": int(ok.sum()), "min_visits": min_visits,
            "best_node": tree.key_to_string(int(ex["node_keys"][best])), "best_node_mean_reward": float(r[best] / max(v[best], 1.0)),
            "best_node_visits": int(v[best])}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log2-traj", type=int, default=22, help="trajectories per GPU per step = 2^T")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-traj", type=int, default=4096, help="trajectories of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--legs", default="all", help="comma list of extra legs: sweep,variants,mcts3,mcts,dimers,dimer_search,landscape (or all / none)")
    ap.add_argument("--mcts-batch", type=int, default=2048, help="config 4: leaves per reward launch")
    ap.add_argument("--mcts-batches", type=int, default=12, help="config 4: reward launches per rank")
    ap.add_argument("--mcts-sync-batches", type=int, default=3, help="config 4: reward launches between statistic merges")
    ap.add_argument("--mcts-depth", type=int, default=4, help="reward batches in flight (virtual loss), sharing the device")
    ap.add_argument("--mcts-traj", type=int, default=32, help="SSA trajectories per rollout in the MCTS legs")
    ap.add_argument("--mcts-max-depth", type=int, default=-1, help="config 4: exchange only states with <= this many interactions")
    ap.add_argument("--mcts3-iters", type=int, default=10_000)
    ap.add_argument("--mcts3-batch", type=int, default=1024)
    ap.add_argument("--dimer-log2-traj", type=int, default=20, help="trajectories of the dimer-model leg = 2^T")
    ap.add_argument("--dimer-search-rollouts", type=int, default=4096)
    ap.add_argument("--landscape-topologies", type=int, default=3325, help="config 5 leg: sample size (3325 = every 3-component terminal)")
    ap.add_argument("--landscape-param-sets", type=int, default=512)
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return
    all_legs = ["sweep", "variants", "mcts3", "mcts", "dimers", "dimer_search", "landscape"]
    legs = all_legs if args.legs == "all" else ([] if args.legs == "none" else args.legs.split(","))

    import torch
    import torch.distributed as dist

    from circuitree_b200 import RewardEngine, SimpleNetworkGrammar, _capi, _capi_defaults
    from circuitree_b200._capi_defaults import i_step

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the reward path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def over_ranks(times: list[float], sums: list[float]) -> tuple[list[float], list[float]]:
        """max over ranks of every entry of `times`, sum over ranks of every entry of `sums`."""
        if world == 1:
            return list(times), list(sums)
        t = torch.tensor(times, dtype=torch.float64, device=dev)
        s = torch.tensor(sums, dtype=torch.float64, device=dev)
        if len(times):
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if len(sums):
            dist.all_reduce(s, op=dist.ReduceOp.SUM)
        return t.tolist(), s.tolist()

    grammar = SimpleNetworkGrammar(["A", "B", "C"], ["activates", "inhibits"], root="ABC::")
    re = RewardEngine(grammar, device=local_rank, seed=0)
    eng, cfg = re.engine, re.cfg
    handle = re.handle(REPRESSILATOR)
    n = 1 << args.log2_traj
    handles = np.array([handle], dtype=np.uint64)
    d_reward = torch.zeros(n, dtype=torch.float32, device=dev)
    d_lag = torch.zeros(n, dtype=torch.int32, device=dev)
    d_events = torch.zeros(n, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream()
    n_sm = eng.n_sm
    sm_max = 1965.0
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        sm_max = json.loads(peaks_file.read_text()).get("sm_max_mhz", sm_max)
    peak = n_sm * 128 * sm_max * 1e6 / 1e12          # T lane-instructions/s at the max SM clock
    I_REP = i_step(18, 3)
    next_id = [0]

    def launch(count: int, config=None):
        """One device-buffer launch of `count` repressilator trajectories with fresh ids on every rank."""
        base = np.array([(next_id[0] * world + rank) << 23], dtype=np.uint64)
        next_id[0] += 1
        eng.launch_rewards(handles, np.array([count], dtype=np.uint32), base, config or cfg, 0, stream.cuda_stream,
                           d_reward.data_ptr(), d_lag=d_lag.data_ptr(), d_events=d_events.data_ptr())

    def timed_launch(count: int, config=None) -> tuple[float, int]:
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        launch(count, config)
        b.record(stream)
        b.synchronize()
        return a.elapsed_time(b), int(d_events[:count].to(torch.int64).sum().item())

    # ---- headline (config 2 at 2^T): device-resident, CUDA events, L2 flushed between steps ---------
    for w in range(args.warmup):
        flush.fill_(w)
        launch(n)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count
    kernel_ms, total_events = 0.0, 0
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        ms, ev = timed_launch(n)
        kernel_ms += ms
        total_events += ev
    barrier()
    wall = time.perf_counter() - t_wall0
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    mean_reward = float(d_reward.mean().item())

    # ---- e2e: public host-buffer API, explicit host parameters in, per-trajectory results out, K steps ----
    rg = np.random.default_rng(1234 + rank)
    ranges = cfg.ranges_array()
    host_params = torch.empty((n, _capi.N_PARAMS), dtype=torch.float32).pin_memory()
    u = rg.random((n, _capi.N_PARAMS), dtype=np.float32)
    vals = ranges[:, 0] + u * (ranges[:, 1] - ranges[:, 0])
    host_params.numpy()[:] = np.where(ranges[:, 2] != 0, 10.0 ** vals, vals).astype(np.float32)
    del u, vals
    re.run([REPRESSILATOR], n, per_trajectory=True, params=host_params.numpy())   # warm-up
    barrier()
    e2e_events, t0 = 0, time.perf_counter()
    for k in range(args.steps):
        out = re.run([REPRESSILATOR], n, per_trajectory=True, params=host_params.numpy())
        e2e_events += int(out["job_events"].sum())
    barrier()
    e2e_secs = time.perf_counter() - t0
    h2d = n * _capi.N_PARAMS * 4 + 32 + 8
    d2h = n * 8 + 16
    del host_params

    (kernel_ms, e2e_secs), (total_events, e2e_events) = over_ranks([kernel_ms, e2e_secs], [total_events, e2e_events])
    value = total_events / (kernel_ms * 1e-3)
    achieved = (value / world) * I_REP / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": kernel_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.log2_traj, cfg.nt),
                   "trajectories_per_gpu": n, "l2": "256 MiB buffer written between timed iterations",
                   "sharding": "independent trajectory ids per rank, no data-path collective"},
        "clocks": clocks,
        "e2e": {"value": e2e_events / e2e_secs, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": args.steps, "api": "RewardEngine.run -> ct_rewards_host (explicit host parameter vectors)"},
        "gpu_launches": int(launches),   # kernel launches inside the timed region of `value` (SSA + launch-order pre-pass)
        "roofline": {"bound": "alu_issue", "achieved": achieved, "peak": peak, "unit": "Tlane-instr/s",
                     "frac": achieved / peak, "traffic": None,
                     "note": f"algorithmic I_step(R=18, M=3) = {I_REP:.1f} issue slots/step (SPEC.md section 8) x steps/s per GPU; "
                             f"peak = {n_sm} SMs x 128 lanes x {sm_max:.0f} MHz (issue bound: no HBM or tensor roofline applies); "
                             "DRAM traffic is not measurable inside this run (null): ncu puts one 2^20 launch at 46 MB read+written "
                             "(profiles/r02_final_ssa_pairwise_M3_2p20.txt; outputs 12.6 MB + launch order 12.6 MB + the L2-resident "
                             f"record scratch), i.e. ~0.06 TB/s against {(n * 12) / 1e6:.0f} MB of outputs per 2^{args.log2_traj} launch"},
        "mean_reward": mean_reward, "wall_s_timed_region": wall,
    }

    # ---- config 2 sweep: 2^14 .. 2^22 trajectories of the same topology in one launch ----------------
    if "sweep" in legs:
        rows = []
        eng.warp_iterations()                      # switch the hot-loop iteration counter on
        for t in (14, 16, 18, 20, 22):
            if t > args.log2_traj:
                continue
            cnt = 1 << t
            launch(cnt)                             # warm-up at this size
            torch.cuda.synchronize()
            eng.warp_iterations()
            barrier()
            reps = 3 if t <= 18 else 1
            ms = ev = 0
            for _ in range(reps):
                a, b = timed_launch(cnt)
                ms += a; ev += b
            iters = eng.warp_iterations()
            (ms,), (ev, iters) = over_ranks([ms], [ev, iters])
            sps = ev / (ms * 1e-3)
            rows.append({"log2_traj": t, "reaction_steps_per_s": sps, "ms_per_launch": ms / reps,
                         "lane_utilisation": ev / (64.0 * iters) if iters else None,
                         "roofline_frac": (sps / world) * I_REP / 1e12 / peak})
        line["sweep"] = {"metric": METRIC, "unit": UNIT, "rows": rows,
                         "config": {"workload": f"{REPRESSILATOR}, one launch of 2^T trajectories x {cfg.nt} grid points per GPU "
                                                "(device buffers, CUDA events, L2 flushed; 3 launches averaged up to 2^18)"}}

    # ---- the general hot loop (arbitrary topologies) on the same workload ---------------------------------
    if "variants" in legs:
        general = _capi.default_config()
        general.fast_path = 0
        cnt = min(n, 1 << 21)
        launch(cnt, general)
        barrier()
        ms, ev = timed_launch(cnt, general)
        (ms,), (ev,) = over_ranks([ms], [ev])
        sps = ev / (ms * 1e-3)
        line["variants"] = {"general_loop": {"reaction_steps_per_s": sps, "roofline_frac": (sps / world) * I_REP / 1e12 / peak,
                                             "config": f"same workload at {cnt} trajectories with ct_sim_config.fast_path = 0: the "
                                                       "loop every topology with two regulators on some promoter runs"}}

    from circuitree_b200.native import NativeTree
    from circuitree_b200.parallel import RootParallelSearch

    # ---- config 1: 3-component MCTS, 10^4 iterations per GPU ------------------------------------------------
    if "mcts3" in legs:
        tree3 = NativeTree(grammar, "ABC::", seed=1000 + rank, n_exhausted=np.inf, trajectories_per_reward=args.mcts_traj,
                           device=local_rank, reward_seed=77 + rank)
        tree3.search_mcts(256, batch_size=256)                                   # warm-up
        e0, s0 = tree3.reward_engine.total_events, tree3.reward_engine.total_issue_slots
        barrier()
        t0 = time.perf_counter()
        tree3.search_mcts(args.mcts3_iters, batch_size=args.mcts3_batch, pipeline_depth=args.mcts_depth)
        barrier()
        secs = time.perf_counter() - t0
        (secs,), (ev, slots) = over_ranks([secs], [tree3.reward_engine.total_events - e0, tree3.reward_engine.total_issue_slots - s0])
        line["mcts3"] = {"metric": "mcts_rollouts_per_s", "value": world * args.mcts3_iters / secs, "unit": "rollouts/s",
                         "reaction_steps_per_s": ev / secs, "roofline_frac": slots / secs / world / 1e12 / peak,
                         "config": {"workload": f"SimpleNetworkGrammar ABC (activates/inhibits) oscillation-reward MCTS, "
                                                f"{args.mcts3_iters} iterations per GPU (independent trees), {args.mcts_traj} SSA "
                                                f"trajectories per rollout, {args.mcts3_batch} leaves per launch, {args.mcts_depth} launches in flight",
                                    "quality_rank0": tree_quality(tree3, "ABC::")}}
        del tree3

    # ---- config 4: root-parallel 5-component MCTS, periodic NCCL merges inside the timed region ----------
    if "mcts" in legs:
        g5 = SimpleNetworkGrammar(list("ABCDE"), ["activates", "inhibits"], root="ABCDE::")
        K, NB, SB = args.mcts_batch, args.mcts_batches, args.mcts_sync_batches

        def run_mcts5(sync_batches: int, tag: int) -> dict:
            tree = NativeTree(g5, "ABCDE::", seed=5000 + rank, n_exhausted=np.inf, trajectories_per_reward=args.mcts_traj,
                              device=local_rank, reward_seed=900 + 16 * tag + rank)
            search = RootParallelSearch(tree, sync_every=K * sync_batches, batch_size=K, device=dev if world > 1 else None,
                                        pipeline_depth=args.mcts_depth, max_depth=args.mcts_max_depth)
            tree.search_mcts(min(K, 512), batch_size=min(K, 512))              # warm-up (kernels, handles)
            tree.timing = {k: 0.0 for k in tree.timing}
            e0, s0 = tree.reward_engine.total_events, tree.reward_engine.total_issue_slots
            barrier()
            t0 = time.perf_counter()
            search.run(K * NB)
            barrier()
            secs = time.perf_counter() - t0
            host = dict(tree.timing, **search.timing)
            (secs, *host_max), (ev, slots, sent) = over_ranks(
                [secs] + [host[k] for k in sorted(host)],
                [tree.reward_engine.total_events - e0, tree.reward_engine.total_issue_slots - s0, search.bytes_sent])
            return {"value": world * K * NB / secs, "seconds": secs, "reaction_steps_per_s": ev / secs,
                    "roofline_frac": slots / secs / world / 1e12 / peak, "n_syncs": search.n_syncs,
                    "merge_bytes_sent_all_ranks": int(sent), "merge_bytes_received_rank0": search.bytes_exchanged,
                    "records_merged_rank0": search.records_merged,
                    "host_seconds_max_over_ranks": {k: round(v, 3) for k, v in zip(sorted(host), host_max)},
                    "quality_rank0": tree_quality(tree, "ABCDE::")}

        merged = run_mcts5(SB, 0)
        control = run_mcts5(NB, 1)            # same budget and seeds, one merge at the very end
        line["mcts"] = {"metric": "mcts_rollouts_per_s", "value": merged["value"], "unit": "rollouts/s",
                        "reaction_steps_per_s": merged["reaction_steps_per_s"], "roofline_frac": merged["roofline_frac"],
                        "config": {"workload": "root-parallel MCTS, SimpleNetworkGrammar ABCDE (activates/inhibits), one tree per GPU, "
                                               f"{NB} launches x {K} leaves per rank, {args.mcts_traj} SSA trajectories per rollout, "
                                               f"leaves grouped by topology, {args.mcts_depth} launches in flight, sparse "
                                               f"(key, d visits, d reward) all-gather + native merge every {SB} launches "
                                               f"({K * SB} iterations) inside the timed region",
                                   "n_exhausted": "inf", "exchange_max_depth": args.mcts_max_depth},
                        "with_merges": merged,
                        "control_one_merge_at_end": control,
                        "note": "host seconds: select/submit/backprop/pack/merge run on the host while launches are in flight; "
                                "reward_s is time blocked waiting for the GPU, comm_s the all-gathers (incl. waiting for the slowest rank)"}

    # ---- dimer model (config 3's kernel): one big launch through the public host API ----------------------
    if "dimers" in legs or "dimer_search" in legs:
        from circuitree_b200 import DimersGrammar
        from circuitree_b200._capi_defaults import i_step_dimers

    if "dimers" in legs:
        dgrammar = DimersGrammar(components=["A", "B", "C"], regulators=[], interactions=["activates", "inhibits"])
        dre = RewardEngine(dgrammar, device=local_rank, seed=0)
        n_dimer = 1 << args.dimer_log2_traj
        dre.run([DIMER_REPRESSILATOR], 4096, traj_base=[1 << 40])                # warm-up
        dh = np.array([dre.handle(DIMER_REPRESSILATOR)], dtype=np.uint64)
        dn = np.array([n_dimer], dtype=np.uint32)
        dr = torch.zeros(n_dimer, dtype=torch.float32, device=dev)
        de = torch.zeros(n_dimer, dtype=torch.int32, device=dev)
        runs, host_runs = [], []
        for rep in range(3):                 # same trajectory ids on every rank and repeat: only the machine varies
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            dre.engine.launch_rewards(dh, dn, np.array([0], dtype=np.uint64), dre.cfg, 0, stream.cuda_stream, dr.data_ptr(),
                                      d_events=de.data_ptr())
            b.record(stream)
            b.synchronize()
            ev = int(de.to(torch.int64).sum().item())
            (ms,), (ev_all,) = over_ranks([a.elapsed_time(b)], [ev])
            runs.append(ev_all / (ms * 1e-3))
            barrier()
            t0 = time.perf_counter()
            ev = int(dre.run([DIMER_REPRESSILATOR], n_dimer, traj_base=[0])["job_events"].sum())
            torch.cuda.synchronize()
            (secs,), (ev_all,) = over_ranks([time.perf_counter() - t0], [ev])
            host_runs.append(ev_all / secs)
        I_DIM = i_step_dimers(3 * 4 + 3 * 2 + 3 * 2, 3)
        line["dimers"] = {"metric": METRIC, "value": float(np.median(runs)), "unit": UNIT, "runs": runs,
                          "host_api_runs": host_runs, "roofline_frac": (float(np.median(runs)) / world) * I_DIM / 1e12 / peak,
                          "config": {"workload": f"dimer model (SPEC.md section 9), homodimer repressilator {DIMER_REPRESSILATOR}: "
                                                 f"2^{args.dimer_log2_traj} trajectories x {cfg.nt} grid points per GPU, parameters "
                                                 "sampled in-kernel, identical trajectory ids on every rank, 3 repeats: `runs` = device "
                                                 "buffers + CUDA events (max over ranks), `host_api_runs` = the same launch through "
                                                 "RewardEngine.run, wall clock; value = median of `runs`",
                                     "i_step": I_DIM}}
        del dr, de

    # ---- config 3 (bounded): DimersGrammar search, five TFs -------------------------------------------------
    if "dimer_search" in legs:
        dg = DimersGrammar(components=list("ABC"), regulators=list("DE"), interactions=["activates", "inhibits"])
        dtree = NativeTree(dg, "ABC+DE::", seed=rank, n_exhausted=np.inf, trajectories_per_reward=args.mcts_traj, device=local_rank,
                           reward_seed=300 + rank)
        dtree.search_mcts(64, batch_size=64)                                     # warm-up
        dtree.timing = {k: 0.0 for k in dtree.timing}
        e0, s0 = dtree.reward_engine.total_events, dtree.reward_engine.total_issue_slots
        barrier()
        t0 = time.perf_counter()
        dtree.search_mcts(args.dimer_search_rollouts, batch_size=1024, pipeline_depth=4)
        barrier()
        secs = time.perf_counter() - t0
        (secs,), (ev, slots) = over_ranks([secs], [dtree.reward_engine.total_events - e0, dtree.reward_engine.total_issue_slots - s0])
        line["dimer_search"] = {"metric": "mcts_rollouts_per_s", "value": world * args.dimer_search_rollouts / secs, "unit": "rollouts/s",
                                "reaction_steps_per_s": ev / secs, "roofline_frac": slots / secs / world / 1e12 / peak,
                                "host_seconds_rank0": {k: round(v, 3) for k, v in dtree.timing.items()},
                                "config": {"workload": f"DimersGrammar ABC+DE (five TFs) oscillation search on the native tree, "
                                                       f"{args.dimer_search_rollouts} rollouts per GPU (independent trees), "
                                                       f"{args.mcts_traj} trajectories per rollout, 1024 leaves per launch, 4 launches "
                                                       "in flight: ONE round of launches, i.e. as long as its longest trajectory; the "
                                                       "sustained rate is the 10^5-rollout run of scripts/dimer_search.py "
                                                       "(profiles/r02_dimer_search_100k.json: 225 rollouts/s)"}}
        del dtree

    # ---- config 5 (bounded sample): exhaustive 3-component landscape, sharded over ranks -------------------
    if "landscape" in legs:
        from circuitree_b200 import reward_landscape

        enum = NativeTree(grammar, "ABC::", seed=0, reward_fn=lambda s: [0.0] * len(s))
        enum.grow_tree()
        terminals = sorted(enum.key_to_string(int(k)) for k in enum.export()["node_keys"] if (int(k) >> 62) & 1)
        pick = np.random.default_rng(2024).choice(len(terminals), size=min(args.landscape_topologies, len(terminals)), replace=False)
        sample = [terminals[i] for i in sorted(pick)]
        lre = RewardEngine(grammar, device=local_rank, seed=5)
        reward_landscape(lre, sample[:8], 64, rank=0, world=1)                   # warm-up
        s0 = lre.total_issue_slots
        barrier()
        t0 = time.perf_counter()
        res = reward_landscape(lre, sample, args.landscape_param_sets, rank=rank, world=world)
        barrier()
        secs = time.perf_counter() - t0
        (secs,), (ev, slots) = over_ranks([secs], [float(res["events"].sum()), lre.total_issue_slots - s0])
        line["landscape"] = {"metric": METRIC, "value": ev / secs, "unit": UNIT, "seconds": secs,
                             "roofline_frac": slots / secs / world / 1e12 / peak, "scaling": "strong",
                             "config": {"workload": f"seeded sample of {len(sample)} of the {len(terminals)} three-component terminal "
                                                    f"topologies x {args.landscape_param_sets} parameter sets, dealt round-robin to "
                                                    "the ranks, no collective (total work fixed); the full sets are "
                                                    "scripts/landscape.py (profiles/)"}}

    if rank == 0:
        if not args.no_cpu_baseline:
            from oracle import ssa_oracle

            ssa_oracle.build()
            threads = ssa_oracle.max_threads()
            ev, dt, _ = cpu_sample(args.cpu_traj, 99, threads)
            line["cpu_baseline"] = {"value": ev / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{args.cpu_traj} repressilator trajectories x {cfg.nt} grid points "
                                              f"({ev} events in {dt:.1f} s); CPU restatement of SPEC.md (plain direct method, every "
                                              "propensity re-evaluated each step, gcc -O2), upstream has no SSA"}
            if "mcts3" in line:
                line["mcts3"]["cpu_baseline"] = cpu_mcts_sample(40, args.mcts_traj, threads)
        print(json.dumps(line))
    if world > 1:
        barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
