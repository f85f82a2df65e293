"""Root-parallel MCTS across the GPUs of one box.

One process per GPU (torchrun), each with its own :class:`NativeTree` (distinct numpy seed and
distinct simulator streams).  Every ``sync_every`` iterations the ranks exchange the *increments*
of (visits, reward) on every node and edge they touched and add the others' increments to their
own tree, so all replicas hold the statistics a single tree would hold had it executed every path
(SURVEY.md section 8e).  The reference has no counterpart: its only parallel search shares one graph
between greenlets (reference circuitree/circuitree.py:786-893).

The exchange is an all-gather of sparse ``(state key, d visits, d reward)`` records -- dense
indexing is impossible for 5 components (~10^9 topologies).  Records travel as one packed
float64/int64 buffer per rank over ``torch.distributed`` (NCCL on the GPUs, gloo in CPU tests);
sizes are exchanged first and buffers padded to the maximum.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

__all__ = ["allgather_deltas", "RootParallelSearch"]


def _pack(d: dict) -> np.ndarray:
    """Deltas -> one int64 buffer: [n_nodes, n_edges, node_key.., node_dv.., node_dr(bits).., edge_*...]."""
    nn, ne = len(d["node_key"]), len(d["edge_pkey"])
    parts = [np.array([nn, ne], dtype=np.int64), d["node_key"].view(np.int64), d["node_dv"],
             d["node_dr"].view(np.int64), d["edge_pkey"].view(np.int64), d["edge_ckey"].view(np.int64), d["edge_dv"],
             d["edge_dr"].view(np.int64)]
    return np.concatenate([np.ascontiguousarray(p, dtype=np.int64) for p in parts])


def _unpack(buf: np.ndarray) -> dict:
    nn, ne = int(buf[0]), int(buf[1])
    o = 2
    out = {}
    for name, n, dt in (("node_key", nn, np.uint64), ("node_dv", nn, np.int64), ("node_dr", nn, np.float64),
                        ("edge_pkey", ne, np.uint64), ("edge_ckey", ne, np.uint64), ("edge_dv", ne, np.int64),
                        ("edge_dr", ne, np.float64)):
        out[name] = buf[o: o + n].copy().view(dt)
        o += n
    return out


def allgather_deltas(deltas: dict, device=None) -> list[dict]:
    """All-gather every rank's delta records; returns the list indexed by rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    mine = torch.from_numpy(_pack(deltas))
    if device is not None:
        mine = mine.to(device)
    size = torch.tensor([mine.numel()], dtype=torch.int64, device=mine.device)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size)
    cap = int(max(int(s.item()) for s in sizes))
    padded = torch.zeros(cap, dtype=torch.int64, device=mine.device)
    padded[: mine.numel()] = mine
    bufs = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    return [_unpack(b[: int(s.item())].cpu().numpy()) for b, s in zip(bufs, sizes)]


class RootParallelSearch:
    """Drives one :class:`NativeTree` per rank with periodic statistic merges.

    Args:
        tree: this rank's ``NativeTree`` (seeded differently on every rank).
        sync_every: iterations between merges.
        batch_size: leaves per reward launch.
        device: torch device for the collective buffers (``cuda:<local_rank>`` with NCCL, None with gloo).
    """

    def __init__(self, tree, sync_every: int, batch_size: int, device=None, pipeline_depth: int = 1):
        import torch.distributed as dist

        self.tree = tree
        self.sync_every = int(sync_every)
        self.batch_size = int(batch_size)
        self.pipeline_depth = int(pipeline_depth)
        self.device = device
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.n_syncs = 0
        self.bytes_exchanged = 0

    def sync(self) -> None:
        deltas = self.tree.collect_deltas()
        if self.world == 1:
            return
        gathered = allgather_deltas(deltas, self.device)
        for r, d in enumerate(gathered):
            self.bytes_exchanged += 8 * (2 + 3 * len(d["node_key"]) + 4 * len(d["edge_pkey"]))
            if r != self.rank:
                self.tree.merge_deltas(d)
        self.n_syncs += 1

    def run(self, n_steps_per_rank: int) -> None:
        done = 0
        while done < n_steps_per_rank:
            k = min(self.sync_every, n_steps_per_rank - done)
            self.tree.search_mcts(k, batch_size=self.batch_size, pipeline_depth=self.pipeline_depth)
            done += k
            self.sync()
