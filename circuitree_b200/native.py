"""``NativeTree``: the batched MCTS scheduler on the native (C++) tree.

Same search as :class:`circuitree_b200.CircuiTree` -- identical RNG consumption, UCB arithmetic,
virtual-loss visit back-propagation and exhaustion bookkeeping (reference
circuitree/circuitree.py:315-537) -- but on bit-packed states inside ``libcircuitree_b200.so``
(csrc/native_tree.cpp), so that selection is not the ceiling once rewards come from the GPU
(the reference spends 59-95 % of a step canonicalising strings).

``to_networkx()`` materialises the reference's ``nx.DiGraph`` (string node ids, ``visits`` /
``reward`` / ``is_exhausted`` attributes) for inspection, GML export and parity tests.
"""
from __future__ import annotations

import ctypes
from typing import Callable, Optional, Sequence

import numpy as np

from . import _capi
from .models import DimersGrammar, SimpleNetworkGrammar

__all__ = ["NativeTree", "NativeExhaustionError", "SearchPipeline"]

_bound = False


class NativeExhaustionError(RuntimeError):
    """The whole tree has been visited to exhaustion (reference ``ExhaustionError``)."""


def _lib():
    global _bound
    L = _capi.lib()
    if not _bound:
        vp, i32, i64, u64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double
        L.ct_tree_last_error.restype = ctypes.c_char_p
        L.ct_tree_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p, i32, ctypes.c_char_p, dbl, dbl, ctypes.POINTER(vp)]
        L.ct_tree_create_dimers.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, i32, i32, ctypes.c_char_p, dbl, dbl,
                                            ctypes.POINTER(vp)]
        L.ct_tree_destroy.argtypes = [vp]
        L.ct_tree_set_rng.argtypes = [vp, vp, vp, i32, ctypes.c_uint32]
        L.ct_tree_get_rng.argtypes = [vp, vp, vp, ctypes.POINTER(i32), ctypes.POINTER(ctypes.c_uint32)]
        L.ct_tree_set_log_table.argtypes = [vp, vp, i64]
        L.ct_tree_rng_draw.argtypes = [vp, i32, i64, vp]
        L.ct_tree_select.argtypes = [vp, i32, vp, vp, ctypes.POINTER(i32)]
        L.ct_tree_backprop.argtypes = [vp, i32, vp]
        L.ct_tree_pending.argtypes = [vp, ctypes.POINTER(i32), vp, vp]
        L.ct_tree_size.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64)]
        L.ct_tree_export.argtypes = [vp] + [vp] * 8
        L.ct_tree_key_to_string.argtypes = [vp, u64, ctypes.c_char_p, i32]
        L.ct_tree_string_to_key.argtypes = [vp, ctypes.c_char_p, i32, ctypes.POINTER(u64)]
        L.ct_tree_actions.argtypes = [vp, u64, vp, i32, ctypes.POINTER(i32)]
        L.ct_tree_do_action.argtypes = [vp, u64, i32, ctypes.POINTER(u64)]
        L.ct_tree_key_to_handle.argtypes = [vp, u64, ctypes.POINTER(u64)]
        L.ct_tree_grow.argtypes = [vp]
        L.ct_tree_pack_deltas.argtypes = [vp, i32, vp, i64, ctypes.POINTER(i64)]
        L.ct_tree_merge_packed.argtypes = [vp, vp, i64, ctypes.POINTER(i32)]
        for name in ("ct_tree_create", "ct_tree_create_dimers", "ct_tree_destroy", "ct_tree_set_rng", "ct_tree_get_rng", "ct_tree_set_log_table",
                     "ct_tree_rng_draw", "ct_tree_select", "ct_tree_backprop", "ct_tree_pending", "ct_tree_size",
                     "ct_tree_export", "ct_tree_key_to_string", "ct_tree_string_to_key", "ct_tree_actions",
                     "ct_tree_do_action", "ct_tree_key_to_handle", "ct_tree_grow", "ct_tree_pack_deltas",
                     "ct_tree_merge_packed"):
            getattr(L, name).restype = ctypes.c_int
        _bound = True
    return L


def _check(rc: int) -> None:
    if rc == 100:
        raise NativeExhaustionError(_lib().ct_tree_last_error().decode())
    if rc != 0:
        raise _capi.EngineError(f"native tree error {rc}: {_lib().ct_tree_last_error().decode()}")


_MASK64 = (1 << 64) - 1


class NativeTree:
    """MCTS on the native tree.  Rewards come from ``reward_fn(states) -> list[float]`` (any
    callable, e.g. a test stub) or, by default, from the GPU :class:`RewardEngine`.

    Args mirror ``CircuiTree`` (reference circuitree.py:76-127): ``grammar`` is a
    :class:`SimpleNetworkGrammar` or a :class:`DimersGrammar` whose root holds every component (and
    regulator), e.g. ``"ABC::"`` / ``"ABC+DE::"``.
    """

    def __init__(self, grammar: SimpleNetworkGrammar, root: str, exploration_constant: Optional[float] = None,
                 seed: Optional[int] = None, n_exhausted: Optional[float] = None, reward_fn: Optional[Callable] = None,
                 trajectories_per_reward: int = 32, device: Optional[int] = None, sim_config=None,
                 reward_seed: Optional[int] = None, max_steps_hint: int = 1 << 20):
        if getattr(grammar, "fixed_components", None):
            raise _capi.EngineError("NativeTree does not support fixed_components; use CircuiTree")
        self.grammar = grammar
        self.root = root
        self.rg = np.random.default_rng(seed)
        self.seed = self.rg.bit_generator._seed_seq.entropy
        self.exploration_constant = float(np.sqrt(2) if exploration_constant is None else exploration_constant)
        self.n_exhausted = n_exhausted
        comps = "".join(c[0].upper() for c in grammar.components)
        types = "".join(x[0].lower() for x in grammar.interactions)
        self._h = ctypes.c_void_p(None)
        n_exh = float("nan") if n_exhausted is None else float(n_exhausted)
        self.is_dimers = isinstance(grammar, DimersGrammar)
        if self.is_dimers:
            regs = "".join(r[0].upper() for r in grammar.regulators)
            _check(_lib().ct_tree_create_dimers(comps.encode(), regs.encode(), types.encode(), int(grammar.max_interactions),
                                                int(grammar.max_interactions_per_promoter), root.encode(),
                                                self.exploration_constant, n_exh, ctypes.byref(self._h)))
        else:
            _check(_lib().ct_tree_create(comps.encode(), types.encode(), int(grammar.max_interactions), root.encode(),
                                         self.exploration_constant, n_exh, ctypes.byref(self._h)))
        with np.errstate(divide="ignore"):
            table = np.log(np.arange(max_steps_hint + 1))     # numpy's own log, as the reference uses
        _check(_lib().ct_tree_set_log_table(self._h, table.ctypes.data, len(table)))
        self._push_rng()
        self.reward_fn = reward_fn
        self.trajectories_per_reward = int(trajectories_per_reward)
        self._reward_engine = None
        self._engine_args = dict(device=device, sim_config=sim_config,
                                 seed=int(self.seed if reward_seed is None else reward_seed) & _MASK64)
        self._strings: dict[int, str] = {}
        self.timing = {"select_s": 0.0, "submit_s": 0.0, "reward_s": 0.0, "backprop_s": 0.0}

    def __del__(self):
        try:
            if self._h.value:
                _lib().ct_tree_destroy(self._h)
                self._h = ctypes.c_void_p(None)
        except Exception:
            pass

    # -- RNG hand-over: the numpy Generator is the single source of truth between calls --------------
    def _push_rng(self) -> None:
        st = self.rg.bit_generator.state
        if st["bit_generator"] != "PCG64":
            raise _capi.EngineError("NativeTree needs numpy's PCG64 bit generator")
        s, inc = st["state"]["state"], st["state"]["inc"]
        a = np.array([s & _MASK64, s >> 64], dtype=np.uint64)
        b = np.array([inc & _MASK64, inc >> 64], dtype=np.uint64)
        _check(_lib().ct_tree_set_rng(self._h, a.ctypes.data, b.ctypes.data, int(st["has_uint32"]), int(st["uinteger"])))

    def _pull_rng(self) -> None:
        a = np.zeros(2, dtype=np.uint64)
        b = np.zeros(2, dtype=np.uint64)
        has, uint = ctypes.c_int32(0), ctypes.c_uint32(0)
        _check(_lib().ct_tree_get_rng(self._h, a.ctypes.data, b.ctypes.data, ctypes.byref(has), ctypes.byref(uint)))
        st = self.rg.bit_generator.state
        st["state"]["state"] = int(a[0]) | (int(a[1]) << 64)
        st["state"]["inc"] = int(b[0]) | (int(b[1]) << 64)
        st["has_uint32"] = int(has.value)
        st["uinteger"] = int(uint.value)
        self.rg.bit_generator.state = st

    # -- packed states <-> strings ---------------------------------------------------------------------
    def key_to_string(self, key: int) -> str:
        s = self._strings.get(key)
        if s is None:
            buf = ctypes.create_string_buffer(256)
            _check(_lib().ct_tree_key_to_string(self._h, key, buf, 256))
            s = buf.value.decode()
            if len(self._strings) < 2_000_000:
                self._strings[key] = s
        return s

    def string_to_key(self, state: str, canonical: bool = True) -> int:
        k = ctypes.c_uint64(0)
        _check(_lib().ct_tree_string_to_key(self._h, state.encode(), int(canonical), ctypes.byref(k)))
        return int(k.value)

    def get_unique_state(self, state: str) -> str:
        return self.key_to_string(self.string_to_key(state, canonical=True))

    def get_actions(self, state: str) -> list[str]:
        key = self.string_to_key(state, canonical=False)
        n = ctypes.c_int32(0)
        ids = np.zeros(512, dtype=np.int32)
        _check(_lib().ct_tree_actions(self._h, key, ids.ctypes.data, 512, ctypes.byref(n)))
        comps = [c[0].upper() for c in self.grammar.components]
        types = [x[0].lower() for x in self.grammar.interactions]
        dimers = self.grammar.dimer_options if self.is_dimers else None
        out = []
        for a in ids[: n.value]:
            if a == 0:
                out.append("*terminate*")
            elif self.is_dimers:        # 1 + ((dimer * C + declared component) * T + declared type)
                a -= 1
                t, dc, d = a % len(types), (a // len(types)) % len(comps), a // (len(types) * len(comps))
                out.append(dimers[d] + types[t] + comps[dc])
            else:
                a -= 1
                t, pair = a % len(types), a // len(types)
                out.append(comps[pair // len(comps)] + comps[pair % len(comps)] + types[t])
        return out

    def handle_of_key(self, key: int) -> int:
        h = ctypes.c_uint64(0)
        _check(_lib().ct_tree_key_to_handle(self._h, key, ctypes.byref(h)))
        return int(h.value)

    # -- search -------------------------------------------------------------------------------------------
    @property
    def reward_engine(self):
        if self._reward_engine is None:
            from .oscillation import RewardEngine

            self._reward_engine = RewardEngine(self.grammar, **self._engine_args)
        return self._reward_engine

    def select(self, k: int):
        """k selections under virtual loss -> (packed sim states, GPU topology handles)."""
        keys = np.zeros(k, dtype=np.uint64)
        handles = np.zeros(k, dtype=np.uint64)
        done = ctypes.c_int32(0)
        self._push_rng()
        rc = _lib().ct_tree_select(self._h, k, keys.ctypes.data, handles.ctypes.data, ctypes.byref(done))
        self._pull_rng()
        _check(rc)
        return keys, handles

    def backprop(self, rewards: Sequence[float]) -> None:
        r = np.ascontiguousarray(rewards, dtype=np.float64)
        _check(_lib().ct_tree_backprop(self._h, len(r), r.ctypes.data))

    def pending_paths(self) -> list[list[str]]:
        n = ctypes.c_int32(0)
        _check(_lib().ct_tree_pending(self._h, ctypes.byref(n), None, None))
        lengths = np.zeros(max(1, n.value), dtype=np.int32)
        _check(_lib().ct_tree_pending(self._h, ctypes.byref(n), lengths.ctypes.data, None))
        ids = np.zeros(max(1, int(lengths[: n.value].sum())), dtype=np.int32)
        _check(_lib().ct_tree_pending(self._h, ctypes.byref(n), lengths.ctypes.data, ids.ctypes.data))
        keys = self.export()["node_keys"]
        out, off = [], 0
        for ln in lengths[: n.value]:
            out.append([self.key_to_string(int(keys[i])) for i in ids[off: off + ln]])
            off += ln
        return out

    def _submit(self, keys: np.ndarray, handles: np.ndarray) -> dict:
        """Start the reward batch of a set of selected leaves (asynchronous on the GPU path)."""
        if self.reward_fn is not None:
            return {"rewards": np.asarray(self.reward_fn([self.key_to_string(int(k)) for k in keys]), dtype=np.float64)}
        return self.reward_engine.submit_leaves(handles, self.trajectories_per_reward)

    def _collect(self, job: dict) -> np.ndarray:
        """Wait for a reward batch; leaf x of topology u gets the mean of its own block of
        ``trajectories_per_reward`` trajectories."""
        if "rewards" in job:
            return job["rewards"]
        return self.reward_engine.collect_leaves(job)

    def _rewards(self, keys: np.ndarray, handles: np.ndarray) -> np.ndarray:
        return self._collect(self._submit(keys, handles))

    def search_mcts(self, n_steps: int, batch_size: int = 1, callback: Optional[Callable] = None,
                    pipeline_depth: int = 1) -> None:
        """``n_steps`` iterations with ``batch_size`` leaves per reward launch; ``batch_size=1``
        is the reference's sequential search (circuitree.py:693-771).

        ``pipeline_depth > 1`` keeps that many reward batches in flight: the leaves of the next
        batch are selected -- with the visits of all in-flight leaves already counted, i.e. under
        virtual loss exactly as in the reference's parallel mode (circuitree.py:532-535) -- while
        earlier batches are still simulating, so their kernels overlap on the GPU.  Rewards are
        back-propagated in selection order.

        Exhaustion (finite ``n_exhausted``) with leaves in flight follows the reference's greenlet
        semantics (circuitree.py:377-420 under :786-893): flagging a node un-counts the edge totals
        of that moment -- including the virtual visits of in-flight leaves -- from its parents, and a
        reward that arrives later is still added along its whole path."""
        pipe = SearchPipeline(self, batch_size, pipeline_depth, callback)
        try:
            pipe.advance(n_steps)
        finally:
            pipe.drain()

    def grow_tree(self) -> None:
        _check(_lib().ct_tree_grow(self._h))

    # -- export ---------------------------------------------------------------------------------------------
    def size(self) -> tuple[int, int]:
        a, b = ctypes.c_int64(0), ctypes.c_int64(0)
        _check(_lib().ct_tree_size(self._h, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value), int(b.value)

    def export(self) -> dict:
        nn, ne = self.size()
        out = {
            "node_keys": np.zeros(nn, np.uint64), "node_visits": np.zeros(nn, np.int64), "node_reward": np.zeros(nn, np.float64),
            "node_exhausted": np.zeros(nn, np.uint8), "edge_parent": np.zeros(ne, np.int32), "edge_child": np.zeros(ne, np.int32),
            "edge_visits": np.zeros(ne, np.int64), "edge_reward": np.zeros(ne, np.float64),
        }
        _check(_lib().ct_tree_export(self._h, *[out[k].ctypes.data for k in
                                               ("node_keys", "node_visits", "node_reward", "node_exhausted", "edge_parent",
                                                "edge_child", "edge_visits", "edge_reward")]))
        return out

    def to_networkx(self):
        import networkx as nx

        ex = self.export()
        g = nx.DiGraph()
        names = [self.key_to_string(int(k)) for k in ex["node_keys"]]
        for name, v, r, x in zip(names, ex["node_visits"], ex["node_reward"], ex["node_exhausted"]):
            attrs = dict(visits=int(v), reward=float(r) if (r != 0 or v != 0) else 0)
            if x:
                attrs["is_exhausted"] = True
            g.add_node(name, **attrs)
        for p, c, v, r in zip(ex["edge_parent"], ex["edge_child"], ex["edge_visits"], ex["edge_reward"]):
            g.add_edge(names[p], names[c], visits=int(v), reward=float(r) if (r != 0 or v != 0) else 0)
        g.root = self.root
        return g

    # -- multi-GPU merge --------------------------------------------------------------------------------------
    PACK_HEADER = 5

    def pack_deltas(self, out: Optional[np.ndarray] = None, max_depth: int = -1) -> np.ndarray:
        """What this replica did since the previous call (sparse statistic increments + exhaustion
        events) as one int64 buffer (``ct_tree_pack_deltas``).  ``out``: a caller-owned int64 array
        (e.g. the numpy view of a pinned torch tensor) used when it is large enough."""
        n = ctypes.c_int64(0)
        _check(_lib().ct_tree_pack_deltas(self._h, int(max_depth), None, 0, ctypes.byref(n)))
        if out is None or out.size < n.value:
            out = np.empty(max(int(n.value), self.PACK_HEADER), dtype=np.int64)
        _check(_lib().ct_tree_pack_deltas(self._h, int(max_depth), out.ctypes.data, out.size, ctypes.byref(n)))
        return out[: n.value]

    def merge_packed(self, buf: np.ndarray) -> bool:
        """Add another replica's :meth:`pack_deltas` buffer; True when the whole tree is exhausted."""
        buf = np.ascontiguousarray(buf, dtype=np.int64)
        flag = ctypes.c_int32(0)
        _check(_lib().ct_tree_merge_packed(self._h, buf.ctypes.data, buf.size, ctypes.byref(flag)))
        return bool(flag.value)

    @staticmethod
    def unpack_deltas(buf: np.ndarray) -> dict:
        """Readable view of a packed buffer (tests, diagnostics).  Keys are one word per state for
        SimpleNetworkGrammar trees (the packed state) and ``key_words`` words for DimersGrammar trees (the
        interaction list itself: their keys are process-local); key arrays then have shape (n, key_words)."""
        nn, ne, nx = int(buf[1]), int(buf[2]), int(buf[3])
        w = int(buf[4]) >> 8
        o, out = NativeTree.PACK_HEADER, {"tree_exhausted": bool(buf[4] & 1), "key_words": w}
        for name, n, dt, width in (("node_key", nn, np.uint64, w), ("node_dv", nn, np.int64, 1), ("node_dr", nn, np.float64, 1),
                                   ("edge_pkey", ne, np.uint64, w), ("edge_ckey", ne, np.uint64, w), ("edge_dv", ne, np.int64, 1),
                                   ("edge_dr", ne, np.float64, 1), ("exhausted", nx, np.uint64, w)):
            a = buf[o: o + n * width].view(dt)
            out[name] = a if width == 1 else a.reshape(n, width)
            o += n * width
        return out

    # dict-style aliases of the two calls above
    def collect_deltas(self, max_depth: int = -1) -> dict:
        buf = self.pack_deltas(max_depth=max_depth)
        d = self.unpack_deltas(buf)
        d["packed"] = buf
        return d

    def merge_deltas(self, d: dict) -> bool:
        return self.merge_packed(d["packed"])


class SearchPipeline:
    """The select -> submit -> collect -> back-propagate loop of :meth:`NativeTree.search_mcts` as
    an object, so that a caller (``RootParallelSearch``) can do other work -- a statistics merge --
    between :meth:`advance` calls while reward batches stay in flight on the GPU."""

    def __init__(self, tree: NativeTree, batch_size: int = 1, pipeline_depth: int = 1, callback: Optional[Callable] = None):
        from collections import deque

        self.tree = tree
        self.batch_size = max(1, int(batch_size))
        self.depth = max(1, int(pipeline_depth))
        self.callback = callback
        self.inflight = deque()
        self.selected = 0      # leaves selected
        self.done = 0          # leaves back-propagated
        if tree.reward_fn is None:
            tree.reward_engine.share = self.depth      # batches in flight share the device

    def _retire(self) -> None:
        import time

        tree = self.tree
        k, keys, paths, job = self.inflight.popleft()
        t0 = time.perf_counter()
        rewards = tree._collect(job)
        t1 = time.perf_counter()
        tree.backprop(rewards)
        tree.timing["reward_s"] += t1 - t0
        tree.timing["backprop_s"] += time.perf_counter() - t1
        if self.callback is not None:
            for x in range(k):
                self.callback(tree, self.done + x, paths[x], tree.key_to_string(int(keys[x])), float(rewards[x]))
        self.done += k

    def advance(self, n_steps: int) -> None:
        """Select and submit ``n_steps`` more leaves; retire batches only as the depth requires."""
        import time

        tree = self.tree
        left = int(n_steps)
        while left > 0:
            k = min(self.batch_size, left)
            t0 = time.perf_counter()
            keys, handles = tree.select(k)
            t1 = time.perf_counter()
            paths = tree.pending_paths()[-k:] if self.callback is not None else None
            job = tree._submit(keys, handles)
            tree.timing["select_s"] += t1 - t0
            tree.timing["submit_s"] += time.perf_counter() - t1
            self.inflight.append((k, keys, paths, job))
            self.selected += k
            left -= k
            while len(self.inflight) >= self.depth:
                self._retire()

    def drain(self) -> None:
        """Wait for and back-propagate every batch still in flight."""
        while self.inflight:
            self._retire()
        if self.tree.reward_fn is None:
            self.tree.reward_engine.share = 1
