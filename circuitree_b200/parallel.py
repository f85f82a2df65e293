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


def _gather(out, inp, poll: bool) -> None:
    """all_gather_into_tensor; with ``poll`` the calling (helper) thread waits for completion by
    querying, not inside a blocking CUDA call -- a thread parked in a synchronous copy behind a
    collective that waits for a late rank slows the search thread's own launches and copies down."""
    import torch.distributed as dist

    if not poll:
        dist.all_gather_into_tensor(out, inp)
        return
    work = dist.all_gather_into_tensor(out, inp, async_op=True)
    while not work.is_completed():
        time.sleep(0.0005)
    work.wait()


def allgather_packed(buf: np.ndarray, staging: _Staging, poll: bool = False) -> tuple[np.ndarray, list[int], int]:
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
    _gather(sizes_t, size, poll)
    sizes = [int(x) for x in sizes_t.cpu().tolist()]
    cap = max(sizes)
    send = staging.send_buffer(cap)
    send_np = send.numpy()
    if buf.ctypes.data != send_np.ctypes.data:
        send_np[:n] = buf
    mine = send[:cap].to(dev, non_blocking=True) if dev.type == "cuda" else send[:cap]
    gathered = torch.empty(world * cap, dtype=torch.int64, device=dev)
    _gather(gathered, mine, poll)
    if dev.type == "cuda":
        host = staging.recv_buffer(world * cap)
        host[: world * cap].copy_(gathered)            # device -> pinned host, synchronous for the caller
        out = host.numpy()[: world * cap]
    else:
        out = gathered.numpy()
    return out, sizes, cap


class _ExchangeThread:
    """Runs the collectives of the statistic exchanges off the search thread, one after the other, in
    submission order (every rank submits the same number).  A rank that reaches an exchange early keeps
    selecting and simulating instead of waiting for the slowest rank; it merges what arrives when it
    arrives (`RootParallelSearch._poll`)."""

    def __init__(self, device):
        import queue
        import threading

        self.device = device
        self.staging = _Staging(device)
        self.todo: "queue.Queue" = queue.Queue()
        self.done: "queue.Queue" = queue.Queue()
        self.outstanding = 0
        self.comm_s = 0.0
        self.error = None
        self.thread = threading.Thread(target=self._work, daemon=True)
        self.thread.start()

    def _work(self):
        try:
            if self.device is not None:
                import torch

                torch.cuda.set_device(self.device)      # the current device is per thread
            while True:
                buf = self.todo.get()
                if buf is None:
                    return
                t0 = time.perf_counter()
                host, sizes, cap = allgather_packed(buf, self.staging, poll=True)
                parts = [host[r * cap: r * cap + sizes[r]].copy() for r in range(len(sizes))]
                self.comm_s += time.perf_counter() - t0
                self.done.put(parts)
        except BaseException as e:      # surfaced by the search thread
            self.error = e
            self.done.put(None)

    def submit(self, buf: np.ndarray) -> None:
        self.outstanding += 1
        self.todo.put(np.array(buf, dtype=np.int64, copy=True))

    def collect(self, block: bool):
        """Finished exchanges (lists of per-rank buffers), oldest first; all outstanding ones if `block`."""
        import queue

        out = []
        while self.outstanding:
            try:
                parts = self.done.get(block=block)
            except queue.Empty:
                break
            if parts is None:
                raise RuntimeError(f"statistic exchange failed: {self.error!r}")
            self.outstanding -= 1
            out.append(parts)
        return out

    def close(self) -> None:
        self.todo.put(None)
        self.thread.join(timeout=60)


class RootParallelSearch:
    """Drives one :class:`NativeTree` per rank with periodic statistic merges.

    Args:
        tree: this rank's ``NativeTree`` (seeded differently on every rank).
        sync_every: iterations between merges.
        batch_size: leaves per reward launch.
        device: torch device for the collective buffers (``cuda:<local_rank>`` with NCCL, None with gloo).
        pipeline_depth: reward batches kept in flight (they stay in flight across a merge).
        max_depth: exchange only states with at most this many interactions (-1: everything).
        overlap: True (default): the collectives run on a helper thread and a rank merges the others'
            statistics when they arrive (at the next batch boundary) instead of blocking at the exchange --
            measured on 8 GPUs, blocking exchanges spent 7.2 of 18.1 s waiting for the slowest rank
            (profiles/r02_bench_n8_short.json).  False: merge at the exchange itself.  Either way every
            replica ends with every other replica's statistics.
    """

    def __init__(self, tree, sync_every: int, batch_size: int, device=None, pipeline_depth: int = 1, max_depth: int = -1,
                 overlap: bool = True):
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
        self.overlap = bool(overlap)
        self._staging = _Staging(device)
        self._xthread: Optional[_ExchangeThread] = None

    def _merge_parts(self, parts) -> None:
        t0 = time.perf_counter()
        for r, part in enumerate(parts):
            if r == self.rank:
                continue
            self.records_merged += int(part[1] + part[2])
            if self.tree.merge_packed(part):
                self.exhausted = True
        self.bytes_exchanged += 8 * int(sum(len(p) for p in parts))
        self.timing["merge_s"] += time.perf_counter() - t0

    def _poll(self, block: bool = False) -> None:
        """Merge the exchanges that have completed (all outstanding ones if `block`)."""
        if self._xthread is None:
            return
        t0 = time.perf_counter()
        finished = self._xthread.collect(block)
        if block:
            self.timing["wait_s"] = self.timing.get("wait_s", 0.0) + time.perf_counter() - t0
        for parts in finished:
            self._merge_parts(parts)

    def sync(self) -> None:
        """One exchange.  Collective: every rank must call it the same number of times."""
        if self.world == 1:
            return                         # a single tree follows the reference to the letter
        t0 = time.perf_counter()
        send_np = self._staging.send_buffer(1 << 16).numpy()
        buf = self.tree.pack_deltas(out=send_np, max_depth=self.max_depth)
        t1 = time.perf_counter()
        self.timing["pack_s"] += t1 - t0
        self.bytes_sent += 8 * int(buf.size)
        self.n_syncs += 1
        if self.overlap:
            if self._xthread is None:
                self._xthread = _ExchangeThread(self.device)
            self._xthread.submit(buf)
            self._poll()
            return
        host, sizes, cap = allgather_packed(buf, self._staging)
        self.timing["comm_s"] += time.perf_counter() - t1
        self._merge_parts([host[r * cap: r * cap + sizes[r]] for r in range(self.world)])

    def run(self, n_steps_per_rank: int) -> int:
        """``n_steps_per_rank`` iterations on this rank, an exchange after every ``sync_every``; returns
        the number of iterations done (fewer only when the tree got exhausted).  Every rank runs the
        same number of exchanges whatever happens to its own search, so the collectives always match."""
        pipe = SearchPipeline(self.tree, self.batch_size, self.pipeline_depth)
        done = 0
        try:
            while done < n_steps_per_rank:
                k = min(self.sync_every, n_steps_per_rank - done)
                left = k
                while left > 0 and not self.exhausted:
                    step = min(self.batch_size, left)
                    try:
                        pipe.advance(step)
                    except NativeExhaustionError:
                        self.exhausted = True
                    left -= step
                    self._poll()               # statistics that arrived meanwhile
                done += k
                if done >= n_steps_per_rank:
                    pipe.drain()               # the last exchange carries every reward
                self.sync()
        finally:
            pipe.drain()
            if self._xthread is not None:
                self._poll(block=True)
                self.timing["comm_s"] += self._xthread.comm_s
                self._xthread.close()
                self._xthread = None
        return pipe.done
