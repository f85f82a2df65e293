"""Root-parallel MCTS across the GPUs of one box.

One process per GPU (torchrun), each with its own :class:`NativeTree` (distinct numpy seed and
distinct simulator streams).  Every ``sync_every`` iterations the ranks exchange what they did since
the previous exchange -- the *increments* of (visits, reward) on every node and edge they touched,
and the keys of the nodes they flagged exhausted -- and add the others' increments to their own
tree, so all replicas hold the statistics a single tree would hold had it executed every path
(SURVEY.md section 8e).  The reference has no counterpart: its only parallel search shares one graph
between greenlets (reference circuitree/circuitree.py:786-893).

The exchange is an all-gather of sparse ``(state key, d visits, d reward)`` records -- dense
indexing is impossible for 5 components (~10^9 topologies).  The native tree packs its records
straight into a pinned host buffer (``ct_tree_pack_deltas``), which travels as one int64 tensor per
rank over ``torch.distributed`` (NCCL on the GPUs, gloo in CPU tests; sizes first, then the buffers
padded to the largest); the received buffers are merged natively (``ct_tree_merge_packed``).  Reward
batches stay in flight on the GPU across an exchange, so the host-side pack/merge overlaps simulation.
"""
from __future__ import annotations

import time
from typing import Optional

import numpy as np

from .native import NativeExhaustionError, SearchPipeline

__all__ = ["allgather_packed", "RootParallelSearch"]


class _Staging:
    """Pinned (when CUDA is in use) host buffers reused across exchanges."""

    def __init__(self, device):
        self.device = device
        self.send = None
        self.recv = None

    def _alloc(self, n: int):
        import torch

        t = torch.empty(n, dtype=torch.int64)
        if self.device is not None and torch.cuda.is_available():
            t = t.pin_memory()
        return t

    def send_buffer(self, n: int):
        if self.send is None or self.send.numel() < n:
            self.send = self._alloc(max(2 * n, 1 << 16))
        return self.send

    def recv_buffer(self, n: int):
        if self.recv is None or self.recv.numel() < n:
            self.recv = self._alloc(max(2 * n, 1 << 16))
        return self.recv


def allgather_packed(buf: np.ndarray, staging: _Staging) -> tuple[np.ndarray, list[int], int]:
    """All-gather one packed delta buffer per rank.  ``buf`` must be a view of
    ``staging.send_buffer``'s numpy array (or any int64 array; it is copied in then).  Returns
    (host array of ``world * cap`` words, per-rank sizes, cap)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    dev = staging.device if staging.device is not None else torch.device("cpu")
    n = int(buf.size)
    size = torch.tensor([n], dtype=torch.int64, device=dev)
    sizes_t = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes_t, size)
    sizes = [int(x) for x in sizes_t.cpu().tolist()]
    cap = max(sizes)
    send = staging.send_buffer(cap)
    send_np = send.numpy()
    if buf.ctypes.data != send_np.ctypes.data:
        send_np[:n] = buf
    mine = send[:cap].to(dev, non_blocking=True) if dev.type == "cuda" else send[:cap]
    gathered = torch.empty(world * cap, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered, mine)
    if dev.type == "cuda":
        host = staging.recv_buffer(world * cap)
        host[: world * cap].copy_(gathered)            # device -> pinned host, synchronous for the caller
        out = host.numpy()[: world * cap]
    else:
        out = gathered.numpy()
    return out, sizes, cap


class RootParallelSearch:
    """Drives one :class:`NativeTree` per rank with periodic statistic merges.

    Args:
        tree: this rank's ``NativeTree`` (seeded differently on every rank).
        sync_every: iterations between merges.
        batch_size: leaves per reward launch.
        device: torch device for the collective buffers (``cuda:<local_rank>`` with NCCL, None with gloo).
        pipeline_depth: reward batches kept in flight (they stay in flight across a merge).
        max_depth: exchange only states with at most this many interactions (-1: everything).
    """

    def __init__(self, tree, sync_every: int, batch_size: int, device=None, pipeline_depth: int = 1, max_depth: int = -1):
        import torch.distributed as dist

        self.tree = tree
        self.sync_every = int(sync_every)
        self.batch_size = int(batch_size)
        self.pipeline_depth = int(pipeline_depth)
        self.max_depth = int(max_depth)
        self.device = device
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.n_syncs = 0
        self.bytes_exchanged = 0          # bytes this rank received (all ranks' buffers) over all syncs
        self.bytes_sent = 0
        self.records_merged = 0
        self.timing = {"pack_s": 0.0, "comm_s": 0.0, "merge_s": 0.0}
        self.exhausted = False            # the whole tree is exhausted (here or on another replica)
        self._staging = _Staging(device)

    def sync(self) -> None:
        """One exchange.  Collective: every rank must call it the same number of times."""
        if self.world == 1:
            return                         # a single tree follows the reference to the letter
        t0 = time.perf_counter()
        send_np = self._staging.send_buffer(1 << 16).numpy()
        buf = self.tree.pack_deltas(out=send_np, max_depth=self.max_depth)
        t1 = time.perf_counter()
        host, sizes, cap = allgather_packed(buf, self._staging)
        t2 = time.perf_counter()
        for r in range(self.world):
            if r == self.rank:
                continue
            part = host[r * cap: r * cap + sizes[r]]
            self.records_merged += int(part[1] + part[2])
            if self.tree.merge_packed(part):
                self.exhausted = True
        t3 = time.perf_counter()
        self.timing["pack_s"] += t1 - t0
        self.timing["comm_s"] += t2 - t1
        self.timing["merge_s"] += t3 - t2
        self.bytes_sent += 8 * int(buf.size)
        self.bytes_exchanged += 8 * int(sum(sizes))
        self.n_syncs += 1

    def run(self, n_steps_per_rank: int) -> int:
        """``n_steps_per_rank`` iterations on this rank, a merge after every ``sync_every``; returns
        the number of iterations done (fewer only when the tree got exhausted).  Every rank runs the
        same number of exchanges whatever happens to its own search, so the collectives always match."""
        pipe = SearchPipeline(self.tree, self.batch_size, self.pipeline_depth)
        done = 0
        try:
            while done < n_steps_per_rank:
                k = min(self.sync_every, n_steps_per_rank - done)
                if not self.exhausted:
                    try:
                        pipe.advance(k)
                    except NativeExhaustionError:
                        self.exhausted = True
                done += k
                if done >= n_steps_per_rank:
                    pipe.drain()               # the last exchange carries every reward
                self.sync()
        finally:
            pipe.drain()
        return pipe.done
