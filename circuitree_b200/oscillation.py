"""``get_reward`` on the B200: batched Gillespie SSA + in-kernel oscillation score.

This is the piece the reference leaves abstract (``CircuiTree.get_reward``, reference
circuitree/circuitree.py:135-152, called at :534 and :687).  ``OscillationTree`` is a drop-in
``CircuiTree`` subclass: ``get_reward(state)`` turns the state into the index arrays of
``parse_genotype`` (reference circuitree/models.py:441-480), hands them to the CUDA engine
through the C ABI (include/circuitree_b200.h) and returns the mean reward of
``trajectories_per_reward`` stochastic trajectories with freshly sampled parameters (SPEC.md).
``get_rewards(states)`` does the same for a batch of leaves in ONE kernel launch.

There is no CPU fallback: without the built library or without a CUDA device these calls raise.
"""
from __future__ import annotations

import os
from typing import Hashable, Optional, Sequence

import numpy as np

from . import _capi, _capi_defaults
from .circuitree import CircuiTree

__all__ = ["OscillationTree", "RewardEngine", "reward_landscape", "landscape_motifs"]


def pairwise_shape(handles: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """(reactions R = 4 M + 2 edges, species M) of pairwise topology handles (SPEC.md section 2)."""
    h = np.asarray(handles, dtype=np.uint64)
    m = ((h >> np.uint64(56)) & np.uint64(15)).astype(np.int64)
    edges = np.zeros(len(h), dtype=np.int64)
    for c in range(25):
        edges += ((h >> np.uint64(2 * c)) & np.uint64(3)) != 0
    return 4 * m + 2 * edges, m


class RewardEngine:
    """Thin host wrapper: state strings -> topology handles -> one launch per batch."""

    def __init__(self, grammar, device: Optional[int] = None, sim_config: Optional[_capi.SimConfig] = None,
                 seed: int = 0):
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.grammar = grammar
        self.device = device
        self.cfg = sim_config if sim_config is not None else _capi.default_config()
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._engine: Optional[_capi.Engine] = None
        self._handles: dict[Hashable, int] = {}
        self._dimer_shape: dict[int, tuple[int, int]] = {}    # dimer handle -> (reactions, TFs)
        self.next_traj = 0           # trajectory ids handed out so far (fresh Philox streams per call)
        self.total_events = 0
        self.total_trajectories = 0
        self.total_issue_slots = 0.0  # sum over jobs of events x I_step (SPEC.md section 8): roofline numerator
        self.cost: dict[int, float] = {}   # topology handle -> events per trajectory observed so far
        self.share = 1               # asynchronous batches kept in flight: each gets 1/share of the CTA slots

    @property
    def engine(self) -> _capi.Engine:
        if self._engine is None:
            self._engine = _capi.Engine(self.device)
        return self._engine

    def handle(self, state: Hashable) -> int:
        h = self._handles.get(state)
        if h is None:
            parsed = self.grammar.parse_genotype(state)
            if len(parsed) == 3:
                comps, act, inh = parsed
                h = _capi.compile_topology(len(comps), act, inh, k=2)
            else:
                comps, regs, act, inh = parsed
                h = _capi.compile_topology(len(comps) + len(regs), act, inh, k=3)
                # reactions of the dimer model (SPEC.md section 9): 4 per TF, 2 per interaction, 2 per distinct dimer
                rows = np.concatenate([np.asarray(act).reshape(-1, 3), np.asarray(inh).reshape(-1, 3)])
                n_tf = len(comps) + len(regs)
                self._dimer_shape[h] = (4 * n_tf + 2 * len(rows) + 2 * len({(int(a), int(b)) for a, b, _ in rows}), n_tf)
            self._handles[state] = h
        return h

    def run(self, states: Sequence[Hashable], n_traj: int | Sequence[int], per_trajectory: bool = False,
            params: Optional[np.ndarray] = None, traj_base: Optional[Sequence[int]] = None) -> dict:
        """Simulate ``n_traj`` trajectories per state; returns per-state mean reward / event counts
        and optionally the per-trajectory reward and lag arrays."""
        handles = np.array([self.handle(s) for s in states], dtype=np.uint64)
        counts = np.full(len(handles), n_traj, dtype=np.uint32) if np.isscalar(n_traj) else np.asarray(n_traj, np.uint32)
        if traj_base is None:
            offsets = np.concatenate(([0], np.cumsum(counts, dtype=np.uint64)[:-1])).astype(np.uint64)
            bases = offsets + np.uint64(self.next_traj)
            self.next_traj += int(counts.sum())
        else:
            # explicit ids (reward_landscape): later automatic ids start beyond them, so no stream is replayed
            bases = np.asarray(traj_base, dtype=np.uint64)
            if len(bases):
                self.next_traj = max(self.next_traj, int((bases + counts.astype(np.uint64)).max()))
        out = self.engine.rewards_host(handles, counts, bases, self.cfg, self.seed, params=params,
                                       per_trajectory=per_trajectory)
        self.total_events += int(out["job_events"].sum())
        self.total_trajectories += int(counts.sum())
        self._count_issue_slots(handles, out["job_events"])
        out["n_traj"] = counts
        out["traj_base"] = bases
        return out


    def _count_issue_slots(self, handles: np.ndarray, job_events: np.ndarray) -> None:
        if not len(handles):
            return
        if not _capi.is_dimer_handle(int(handles[0])):
            r, m = pairwise_shape(handles)
            self.total_issue_slots += float((job_events.astype(np.float64) * _capi_defaults.i_step(r, m)).sum())
        else:
            for h in handles:
                if int(h) not in self._dimer_shape:       # (handles that came from the native tree)
                    n_tf, n_r = _capi.topology_info(int(h))
                    self._dimer_shape[int(h)] = (n_r, n_tf)
            shape = np.array([self._dimer_shape[int(h)] for h in handles], dtype=np.float64)
            self.total_issue_slots += float((job_events.astype(np.float64) *
                                             _capi_defaults.i_step_dimers(shape[:, 0], shape[:, 1])).sum())

    # -- asynchronous batches of leaves (pipelined MCTS) ---------------------------------------------------
    def submit_leaves(self, handles: np.ndarray, trajectories_per_leaf: int) -> dict:
        """Start the reward batch of a set of selected leaves (one topology handle per leaf) and
        return at once.  Leaves are grouped by topology: one job per distinct simulated state, with
        the events per trajectory seen so far for that topology as the cost hint of the
        longest-first launch order (unseen topologies: assume expensive)."""
        handles = np.asarray(handles, dtype=np.uint64)
        uniq, inverse, counts = np.unique(handles, return_inverse=True, return_counts=True)
        cost = np.array([self.cost.get(int(h), np.inf) for h in uniq])
        known = cost[np.isfinite(cost)]
        job_cost = np.where(np.isfinite(cost), cost, float(known.max()) if len(known) else 1.0e6)
        B = int(trajectories_per_leaf)
        n_traj = (counts * B).astype(np.uint32)
        offsets = np.concatenate(([0], np.cumsum(n_traj, dtype=np.uint64)[:-1])).astype(np.uint64)
        bases = offsets + np.uint64(self.next_traj)
        self.next_traj += int(n_traj.sum())
        # batches in flight run side by side, each on its share of the device: a launch that owned every
        # CTA slot would make the next one wait for its tail (the longest trajectories of the batch)
        limit = self.cfg.grid_limit
        if self.share > 1:
            self.cfg.grid_limit = max(1, -(-(self.engine.resident_lanes // 128) // self.share))
        try:
            ticket = self.engine.submit(uniq, n_traj, bases, self.cfg, self.seed, job_cost=job_cost)
        finally:
            self.cfg.grid_limit = limit
        return {"ticket": ticket, "uniq": uniq, "inverse": inverse, "n_traj": n_traj, "offsets": offsets,
                "n_leaves": len(handles), "B": B}

    def collect_leaves(self, job: dict) -> np.ndarray:
        """Wait for a batch started by :meth:`submit_leaves`; leaf x of topology u gets the mean of
        its own block of ``trajectories_per_leaf`` trajectories (float64, in leaf order)."""
        out = self.engine.wait(job["ticket"], per_trajectory=True)
        self.total_events += int(out["job_events"].sum())
        self.total_trajectories += int(job["n_traj"].sum())
        self._count_issue_slots(job["uniq"], out["job_events"])
        for h, ev, nt in zip(job["uniq"], out["job_events"], job["n_traj"]):
            self.cost[int(h)] = float(ev) / float(nt)
        B = job["B"]
        block_means = out["reward"].astype(np.float64).reshape(-1, B).mean(axis=1)
        first_block = (job["offsets"] // np.uint64(B)).astype(np.int64)
        inverse = job["inverse"]
        # rank of each leaf among the leaves that share its topology, in selection order
        order = np.argsort(inverse, kind="stable")
        rank = np.empty(len(inverse), dtype=np.int64)
        starts = np.concatenate(([0], np.cumsum(np.bincount(inverse, minlength=len(job["uniq"])))[:-1]))
        rank[order] = np.arange(len(inverse)) - starts[inverse[order]]
        return block_means[first_block[inverse] + rank]


class OscillationTree(CircuiTree):
    """MCTS for oscillating TF networks with the reward computed on the GPU.

    Args (beyond :class:`CircuiTree`):
        trajectories_per_reward: SSA trajectories averaged into one reward (default 32).
        sim_config: ``_capi.SimConfig`` (time grid, parameter ranges); default = SPEC.md.
        device: CUDA device index (default ``LOCAL_RANK`` or 0).
        reward_seed: Philox key of the simulator (default: the tree's ``seed``).
        success_threshold: mean reward above which :meth:`is_success` is true (default 0.4).
        next_traj: first unused trajectory id of the simulator stream.  ``to_file`` saves it next to
            ``reward_seed`` and the numpy RNG state, so a tree restored with ``from_file`` continues
            with fresh Philox streams instead of replaying the ones already used.
    """

    def __init__(self, *args, trajectories_per_reward: int = 32, sim_config=None, device: Optional[int] = None,
                 reward_seed: Optional[int] = None, success_threshold: float = 0.4, next_traj: int = 0, **kwargs):
        super().__init__(*args, **kwargs)
        self.trajectories_per_reward = int(trajectories_per_reward)
        self.success_threshold = float(success_threshold)
        self.reward_seed = int(self.seed if reward_seed is None else reward_seed) & 0xFFFFFFFFFFFFFFFF
        self._reward_engine = RewardEngine(self.grammar, device=device, sim_config=sim_config, seed=self.reward_seed)
        self._reward_engine.next_traj = int(next_traj)
        self._non_serializable_attrs.append("_reward_engine")

    def get_attributes(self, attrs_copy=None) -> dict:
        attrs = super().get_attributes(attrs_copy)
        if attrs_copy is None or "next_traj" in attrs_copy:
            attrs["next_traj"] = int(self._reward_engine.next_traj)
        return attrs

    @property
    def reward_engine(self) -> RewardEngine:
        return self._reward_engine

    def get_reward(self, state: Hashable, **kwargs) -> float:
        return float(self.get_rewards([state], **kwargs)[0])

    def get_rewards(self, states: Sequence[Hashable], **kwargs) -> list[float]:
        out = self._reward_engine.run(list(states), self.trajectories_per_reward)
        return [float(x) for x in out["job_mean"]]

    def search_mcts_batched(self, n_steps: int, batch_size: int, *args, pipeline_depth: int = 1, **kwargs) -> None:
        self._reward_engine.share = max(1, int(pipeline_depth))
        try:
            super().search_mcts_batched(n_steps, batch_size, *args, pipeline_depth=pipeline_depth, **kwargs)
        finally:
            self._reward_engine.share = 1

    def submit_rewards(self, states: Sequence[Hashable], **kwargs):
        """Asynchronous half of :meth:`get_rewards` (used by ``search_mcts_batched(pipeline_depth > 1)``):
        the batch starts on a stream of its own and this returns at once."""
        handles = np.array([self._reward_engine.handle(s) for s in states], dtype=np.uint64)
        return self._reward_engine.submit_leaves(handles, self.trajectories_per_reward)

    def collect_rewards(self, ticket) -> list[float]:
        return [float(x) for x in self._reward_engine.collect_leaves(ticket)]

    def simulate(self, states: Sequence[Hashable], n_traj: int, params: Optional[np.ndarray] = None) -> dict:
        """Per-trajectory rewards and lags for analysis (not part of the reference API)."""
        return self._reward_engine.run(list(states), n_traj, per_trajectory=True, params=params)

    def is_success(self, terminal_state: Hashable) -> bool:
        node = self.graph.nodes[terminal_state]
        return node["visits"] > 0 and node["reward"] / node["visits"] > self.success_threshold


def reward_landscape(engine: RewardEngine, states: Sequence[Hashable], n_per_state: int, chunk: int = 1 << 22,
                     rank: int = 0, world: int = 1) -> dict:
    """Mean reward, its standard error and the success fraction (reward > 0.4) of every state:
    the exhaustive "reward landscape" of BASELINE config 5.  States are dealt round-robin to
    ``world`` ranks (no communication: every (state, trajectory) pair is independent) and simulated
    in launches of at most ``chunk`` trajectories; trajectory ids are ``state_index * n_per_state + k``
    so the result does not depend on the sharding."""
    states = list(states)
    mine = list(range(rank, len(states), world))
    out = {"index": np.array(mine, dtype=np.int64), "mean": np.zeros(len(mine)), "sem": np.zeros(len(mine)),
           "success": np.zeros(len(mine)), "events": np.zeros(len(mine), dtype=np.uint64)}
    per_launch = max(1, chunk // n_per_state)
    for a in range(0, len(mine), per_launch):
        idx = mine[a: a + per_launch]
        bases = [i * n_per_state for i in idx]
        res = engine.run([states[i] for i in idx], n_per_state, per_trajectory=True, traj_base=bases)
        r = res["reward"].reshape(len(idx), n_per_state).astype(np.float64)
        sl = slice(a, a + len(idx))
        out["mean"][sl] = r.mean(axis=1)
        out["sem"][sl] = r.std(axis=1, ddof=1) / np.sqrt(n_per_state) if n_per_state > 1 else 0.0
        out["success"][sl] = (r > 0.4).mean(axis=1)
        out["events"][sl] = res["job_events"]
    return out


def landscape_motifs(grammar, states: Sequence[Hashable], landscape: dict, patterns: Sequence[str], n_per_state: int,
                     confidence: Optional[float] = 0.95):
    """Motif enrichment straight from a reward landscape (:func:`reward_landscape` output): every
    simulated (topology, parameter set) pair is one sample of the null population, the pairs whose
    reward exceeded the success threshold are the successful population; each pattern's 2x2 table is
    tested as in ``CircuiTree.test_pattern_significance`` (reference circuitree.py:1510-1648).
    Returns the same pandas table, sorted by odds ratio."""
    from .motifs import pattern_table

    idx = np.asarray(landscape["index"])
    null = {states[i]: int(n_per_state) for i in idx}
    succ = {states[i]: int(round(float(f) * n_per_state)) for i, f in zip(idx, landscape["success"])}
    succ = {s: c for s, c in succ.items() if c > 0}
    return pattern_table(grammar, patterns, null, succ, confidence=confidence, exclude_self=False)
