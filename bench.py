"""Headline benchmark: batched Gillespie SSA reaction-steps/s (BASELINE.json metric, config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--log2-traj T] [--impl reference]

One "step" = one pass of the hot path (SSA + in-kernel reward) over one batch of 2^T repressilator
trajectories x 2000 grid points per GPU, parameters drawn per trajectory from the SPEC.md ranges.
N > 1 is launched by torchrun (one rank per GPU); trajectories are independent, so ranks shard
trajectory ids with no data-path collective ("weak" scaling: per-GPU work fixed).

`value` times the kernel with inputs resident in HBM (CUDA events on the launch stream); `e2e`
times the same batch through the public host-buffer API (RewardEngine.run -> ct_rewards_host):
explicit parameter vectors copied host->device, per-trajectory rewards and lags copied back.
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
# dram__bytes_read.sum + dram__bytes_write.sum of one ssa_pairwise_kernel<3,0,1,1> launch over 2^22 trajectories
# (ncu, round 1: 44,168,960 + 26,891,520; profiles/r01b_dram_2p22.txt)
DRAM_BYTES_PER_LAUNCH_2P22 = 71060480
METRIC = "gillespie_reaction_steps_per_s"
UNIT = "reaction-steps/s"
# Algorithmic issue slots per SSA step for R reactions and M species (DESIGN.md "Roofline"):
# Philox4x32-10 shared by two steps (30) + per reaction: propensity fused into the running sum,
# one compare, 0.75 state updates (2.75 R) + rate selects (7 M) + tau/grid/loop (13).
def i_step(n_reactions: int, n_species: int) -> float:
    return 30.0 + 2.75 * n_reactions + 7.0 * n_species + 13.0


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
                   "note": "upstream circuitree ships no SSA; this is the repo's CPU restatement of SPEC.md (oracle/)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log2-traj", type=int, default=22, help="trajectories per GPU per step = 2^T")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-traj", type=int, default=4096, help="trajectories of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mcts-batch", type=int, default=8192, help="leaves per reward launch in the MCTS leg")
    ap.add_argument("--mcts-batches", type=int, default=12)
    ap.add_argument("--mcts-depth", type=int, default=4, help="reward batches in flight (virtual loss), sharing the device")
    ap.add_argument("--mcts-traj", type=int, default=16, help="SSA trajectories per rollout in the MCTS leg")
    ap.add_argument("--dimer-log2-traj", type=int, default=20, help="trajectories of the dimer-model leg = 2^T (0 = skip)")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from circuitree_b200 import RewardEngine, SimpleNetworkGrammar, _capi, _capi_defaults

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the reward path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    grammar = SimpleNetworkGrammar(["A", "B", "C"], ["activates", "inhibits"], root="ABC::")
    re = RewardEngine(grammar, device=local_rank, seed=0)
    eng, cfg = re.engine, re.cfg
    handle = re.handle(REPRESSILATOR)
    n = 1 << args.log2_traj
    handles = np.array([handle], dtype=np.uint64)
    counts = np.array([n], dtype=np.uint32)
    dev = torch.device("cuda", local_rank)
    d_reward = torch.zeros(n, dtype=torch.float32, device=dev)
    d_lag = torch.zeros(n, dtype=torch.int32, device=dev)
    d_events = torch.zeros(n, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream()

    def launch(step_id: int):
        base = np.array([(step_id * world + rank) * n], dtype=np.uint64)   # disjoint trajectory ids per rank/step
        eng.launch_rewards(handles, counts, base, cfg, 0, stream.cuda_stream, d_reward.data_ptr(),
                           d_lag=d_lag.data_ptr(), d_events=d_events.data_ptr())

    for w in range(args.warmup):
        flush.fill_(w)
        launch(w)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count
    ev_start = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev_end = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    total_events = 0
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k)                  # L2 flush between timed iterations (not timed)
        ev_start[k].record(stream)
        launch(args.warmup + k)
        ev_end[k].record(stream)
        ev_end[k].synchronize()
        total_events += int(d_events.to(torch.int64).sum().item())
    barrier()
    wall = time.perf_counter() - t_wall0
    kernel_ms = sum(s.elapsed_time(e) for s, e in zip(ev_start, ev_end))
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    mean_reward = float(d_reward.mean().item())

    # ---- e2e: public host-buffer API, explicit host parameters in, per-trajectory results out ----
    rg = np.random.default_rng(1234 + rank)
    ranges = cfg.ranges_array()
    u = rg.random((n, _capi.N_PARAMS), dtype=np.float32)
    host_params = torch.empty((n, _capi.N_PARAMS), dtype=torch.float32).pin_memory()
    vals = ranges[:, 0] + u * (ranges[:, 1] - ranges[:, 0])
    host_params.numpy()[:] = np.where(ranges[:, 2] != 0, 10.0 ** vals, vals).astype(np.float32)
    e2e_steps = max(1, min(args.steps, 2))
    re.run([REPRESSILATOR], n, per_trajectory=True, params=host_params.numpy())   # warm-up
    barrier()
    e2e_events, t0 = 0, time.perf_counter()
    for k in range(e2e_steps):
        out = re.run([REPRESSILATOR], n, per_trajectory=True, params=host_params.numpy())
        e2e_events += int(out["job_events"].sum())
    barrier()
    e2e_secs = time.perf_counter() - t0
    h2d = n * _capi.N_PARAMS * 4 + 32 + 8
    d2h = n * 8 + 16

    # ---- second half of the metric: MCTS rollouts/s (BASELINE config 1/4 shape) -------------------
    # Root-parallel search on the 3-component grammar: native selection under virtual loss, one reward
    # launch per batch of leaves (grouped by topology), per-rank trees merged by a sparse all-gather.
    from circuitree_b200.native import NativeTree
    from circuitree_b200.parallel import RootParallelSearch

    mcts_B, mcts_K, mcts_batches = args.mcts_traj, args.mcts_batch, args.mcts_batches
    tree = NativeTree(grammar, "ABC::", seed=1000 + rank, n_exhausted=np.inf, trajectories_per_reward=mcts_B,
                      device=local_rank, reward_seed=77 + rank)
    search = RootParallelSearch(tree, sync_every=mcts_K * mcts_batches, batch_size=mcts_K,
                                device=dev if world > 1 else None, pipeline_depth=args.mcts_depth)
    tree.search_mcts(min(mcts_K, 1024), batch_size=min(mcts_K, 1024))        # warm-up
    tree.timing = {k: 0.0 for k in tree.timing}
    ev0 = tree.reward_engine.total_events
    barrier()
    t0 = time.perf_counter()
    search.run(mcts_K * mcts_batches)
    barrier()
    mcts_secs = time.perf_counter() - t0
    mcts_events = tree.reward_engine.total_events - ev0
    mcts_launches = tree.reward_engine.engine.launch_count

    # ---- dimer model (BASELINE config 3's kernel): homodimer repressilator through the same public API ----
    dimer_secs, dimer_events, n_dimer = 1.0, 0, 0
    if args.dimer_log2_traj > 0:
        from circuitree_b200 import DimersGrammar

        dgrammar = DimersGrammar(components=["A", "B", "C"], regulators=[], interactions=["activates", "inhibits"])
        dre = RewardEngine(dgrammar, device=local_rank, seed=rank)
        n_dimer = 1 << args.dimer_log2_traj
        dre.run([DIMER_REPRESSILATOR], 2048)                                     # warm-up
        barrier()
        t0 = time.perf_counter()
        dimer_events = int(dre.run([DIMER_REPRESSILATOR], n_dimer)["job_events"].sum())
        barrier()
        dimer_secs = time.perf_counter() - t0

    # ---- reduce over ranks: max time, summed events ----
    stats = torch.tensor([kernel_ms, e2e_secs, float(total_events), float(e2e_events), mcts_secs, float(mcts_events),
                          dimer_secs, float(dimer_events)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        kernel_ms, e2e_secs, mcts_secs, dimer_secs = float(mx[0]), float(mx[1]), float(mx[4]), float(mx[6])
        total_events, e2e_events, mcts_events, dimer_events = float(sm[2]), float(sm[3]), float(sm[5]), float(sm[7])
    value = total_events / (kernel_ms * 1e-3)
    e2e_value = e2e_events / e2e_secs

    if rank == 0:
        n_react, n_spec = 18, 3
        sm_max = (clocks or {}).get("sm_max_mhz") or 1965.0
        peaks_file = ROOT / "MEASURED_PEAKS.json"
        if peaks_file.exists():
            sm_max = json.loads(peaks_file.read_text()).get("sm_max_mhz", sm_max)
        n_sm = eng.n_sm
        peak = n_sm * 128 * sm_max * 1e6 / 1e12          # T lane-instructions/s at the max SM clock
        achieved = (value / world) * i_step(n_react, n_spec) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kernel_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.log2_traj, cfg.nt),
                       "trajectories_per_gpu": n, "l2": "256 MiB buffer written between timed iterations",
                       "sharding": "independent trajectory ids per rank, no data-path collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "RewardEngine.run -> ct_rewards_host (explicit host parameter vectors)"},
            "gpu_launches": int(launches),   # SSA kernel launches inside the timed region of `value`
            "roofline": {"bound": "alu_issue", "achieved": achieved, "peak": peak, "unit": "Tlane-instr/s",
                         "frac": achieved / peak, "traffic": DRAM_BYTES_PER_LAUNCH_2P22 if args.log2_traj == 22 else None,
                         "note": f"algorithmic {i_step(n_react, n_spec):.1f} issue slots/step (DESIGN.md) x steps/s per GPU; "
                                 f"peak = {n_sm} SMs x 128 lanes x {sm_max:.0f} MHz (issue bound: no HBM or tensor roofline applies); "
                                 f"traffic = dram bytes read+written by one 2^22 launch (ncu, profiles/r01b_dram_2p22.txt) "
                                 f"against {(n * 12) / 1e6:.0f} MB of outputs + {(n * 4) / 1e6:.0f} MB launch order"},
            "mean_reward": mean_reward, "wall_s_timed_region": wall,
            "mcts": {"metric": "mcts_rollouts_per_s", "value": world * mcts_K * mcts_batches / mcts_secs, "unit": "rollouts/s",
                     "reaction_steps_per_s": mcts_events / mcts_secs,
                     "config": {"workload": "root-parallel MCTS, SimpleNetworkGrammar ABC (activates/inhibits), "
                                            f"{mcts_batches} batches x {mcts_K} leaves per rank, {mcts_B} SSA trajectories "
                                            "per rollout, leaves grouped by topology, "
                                            f"{args.mcts_depth} reward batches in flight, one sparse stats merge per rank",
                                "n_exhausted": "inf", "tree_nodes_rank0": tree.size()[0],
                                "host_seconds_rank0": {k: round(v, 3) for k, v in tree.timing.items()},
                                "merge_bytes_rank0": search.bytes_exchanged}},
        }
        if n_dimer:
            line["dimers"] = {"metric": METRIC, "value": dimer_events / dimer_secs, "unit": UNIT,
                              "config": {"workload": f"dimer model (SPEC.md section 9), homodimer repressilator {DIMER_REPRESSILATOR}: "
                                                     f"2^{args.dimer_log2_traj} trajectories x {cfg.nt} grid points per GPU, "
                                                     "host API (RewardEngine.run), parameters sampled in-kernel"}}
        if not args.no_cpu_baseline:
            from oracle import ssa_oracle

            ssa_oracle.build()
            threads = ssa_oracle.max_threads()
            ev, dt, _ = cpu_sample(args.cpu_traj, 99, threads)
            line["cpu_baseline"] = {"value": ev / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{args.cpu_traj} repressilator trajectories x {cfg.nt} grid points "
                                              f"({ev} events in {dt:.1f} s); CPU restatement of SPEC.md, upstream has no SSA"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
