"""ctypes binding of ``libcircuitree_b200.so`` (the C ABI declared in include/circuitree_b200.h).

This is the only module that touches the shared library.  There is no CPU fallback: if the
library is missing or no CUDA device is present, the calls raise.
"""
from __future__ import annotations

import ctypes
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

N_PARAMS = 10          # pairwise model (SPEC.md section 4)
N_PARAMS_DIMER = 12    # dimer model (SPEC.md section 9)
MAX_PARAMS = 12
MAX_INTERACTIONS = 10
DIMER_TAG = 1 << 63    # set in handles of k = 3 topologies
MAX_SPECIES = 5
MAX_NDEC = 126

import os as _os

# CIRCUITREE_B200_LIB selects another build of the same library (kernel-variant experiments only)
_LIB_PATH = Path(_os.environ.get("CIRCUITREE_B200_LIB") or Path(__file__).resolve().parent / "libcircuitree_b200.so")
_lib: Optional[ctypes.CDLL] = None


class EngineError(RuntimeError):
    """Raised when a C-ABI call returns a non-zero status."""


class SimConfig(ctypes.Structure):
    """Mirror of ``ct_sim_config`` (SPEC.md sections 4-7)."""

    _fields_ = [
        ("nt", ctypes.c_int32),
        ("decim", ctypes.c_int32),
        ("dt", ctypes.c_float),
        ("init_mean", ctypes.c_float),
        ("schedule", ctypes.c_int32),
        ("lpt", ctypes.c_int32),
        ("grid_limit", ctypes.c_int32),
        ("fast_path", ctypes.c_int32),
        ("ranges", (ctypes.c_float * 3) * MAX_PARAMS),
    ]

    def ranges_array(self, n: int = N_PARAMS) -> np.ndarray:
        """First ``n`` (lo, hi, is_log) rows: 10 for the pairwise model, 12 for the dimer model."""
        return np.array([[self.ranges[i][j] for j in range(3)] for i in range(n)], dtype=np.float64)

    def set_ranges(self, ranges) -> None:
        """Overwrite the first ``len(ranges)`` rows (10 or 12)."""
        r = np.asarray(ranges, dtype=np.float32).reshape(-1, 3)
        if r.shape[0] > MAX_PARAMS:
            raise ValueError(f"at most {MAX_PARAMS} parameter ranges")
        for i in range(r.shape[0]):
            for j in range(3):
                self.ranges[i][j] = float(r[i, j])

    @property
    def n_dec(self) -> int:
        return self.nt // self.decim


def library_path() -> Path:
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise EngineError(
            f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  circuitree_b200 has no CPU fallback for the reward path."
        )
    L = ctypes.CDLL(str(_LIB_PATH))
    vp, i32, u64, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint64, ctypes.c_int64
    L.ct_last_error.restype = ctypes.c_char_p
    L.ct_last_error.argtypes = []
    L.ct_version.restype = ctypes.c_int
    L.ct_default_config.argtypes = [ctypes.POINTER(SimConfig)]
    L.ct_engine_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    L.ct_engine_destroy.argtypes = [vp]
    L.ct_engine_info.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32)]
    L.ct_compile_topology.argtypes = [i32, vp, i32, vp, i32, i32, ctypes.POINTER(u64)]
    L.ct_topology_info.argtypes = [u64, ctypes.POINTER(i32), ctypes.POINTER(i32)]
    L.ct_check_params.argtypes = [vp, i64, i32]
    L.ct_launch_rewards.argtypes = [vp, vp, vp, vp, vp, i32, ctypes.POINTER(SimConfig), u64, vp, vp, vp, vp, vp, vp, i32]
    L.ct_reduce_jobs.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
    L.ct_rewards_host.argtypes = [vp, vp, vp, vp, vp, i32, ctypes.POINTER(SimConfig), u64, vp, vp, vp, vp, vp]
    L.ct_reward_from_series_host.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    L.ct_rewards_submit.argtypes = [vp, vp, vp, vp, vp, i32, ctypes.POINTER(SimConfig), u64, vp, ctypes.POINTER(i64)]
    L.ct_rewards_wait.argtypes = [vp, i64, vp, vp, vp, vp]
    L.ct_engine_launch_count.restype = i64
    L.ct_engine_launch_count.argtypes = [vp]
    L.ct_engine_warp_iterations.restype = i64
    L.ct_engine_warp_iterations.argtypes = [vp]
    for name in ("ct_topology_info", "ct_check_params", "ct_default_config", "ct_engine_create", "ct_engine_destroy", "ct_engine_info",
                 "ct_compile_topology", "ct_launch_rewards", "ct_reduce_jobs", "ct_rewards_host",
                 "ct_reward_from_series_host", "ct_rewards_submit", "ct_rewards_wait"):
        getattr(L, name).restype = ctypes.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().ct_last_error().decode("utf-8", "replace")
        raise EngineError(f"circuitree_b200 C ABI error {rc}: {msg}")


def default_config() -> SimConfig:
    cfg = SimConfig()
    check(lib().ct_default_config(ctypes.byref(cfg)))
    return cfg


def is_dimer_handle(handle: int) -> bool:
    return bool(int(handle) & DIMER_TAG)


def _params_array(params, handles, total: int) -> np.ndarray:
    width = N_PARAMS_DIMER if len(handles) and is_dimer_handle(handles[0]) else N_PARAMS
    return np.ascontiguousarray(params, dtype=np.float32).reshape(total, width)


def compile_topology(n_species: int, activations, inhibitions, k: int = 2) -> int:
    """``parse_genotype`` arrays -> 64-bit topology handle: k = 2 rows (regulator, target) of
    SimpleNetworkGrammar (reference models.py:441-480); k = 3 rows (monomer1, monomer2, target) of
    DimersGrammar (reference models.py:844-884), ``n_species`` = components + regulators."""
    act = np.ascontiguousarray(np.asarray(activations, dtype=np.int64).reshape(-1, k))
    inh = np.ascontiguousarray(np.asarray(inhibitions, dtype=np.int64).reshape(-1, k))
    h = ctypes.c_uint64(0)
    check(lib().ct_compile_topology(n_species, act.ctypes.data, act.shape[0], inh.ctypes.data, inh.shape[0], k,
                                    ctypes.byref(h)))
    return int(h.value)


def topology_info(handle: int) -> tuple[int, int]:
    """(species, reactions) of a compiled topology (``ct_topology_info``)."""
    nsp, nr = ctypes.c_int32(0), ctypes.c_int32(0)
    check(lib().ct_topology_info(int(handle), ctypes.byref(nsp), ctypes.byref(nr)))
    return int(nsp.value), int(nr.value)


def check_params(params: np.ndarray) -> None:
    """Raise unless every explicit rate is finite and non-negative (``ct_check_params``; the host-buffer entry
    points apply the same check themselves)."""
    a = np.ascontiguousarray(params, dtype=np.float32)
    if a.ndim != 2:
        raise EngineError("explicit parameters must have shape (trajectories, parameters per trajectory)")
    check(lib().ct_check_params(a.ctypes.data, a.shape[0], a.shape[1]))


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a)


class Engine:
    """One engine per CUDA device (``ct_engine``)."""

    def __init__(self, device: int = 0):
        self._h = ctypes.c_void_p(None)
        check(lib().ct_engine_create(device, ctypes.byref(self._h)))
        self.device = device
        n_sm, lanes = ctypes.c_int32(0), ctypes.c_int32(0)
        check(lib().ct_engine_info(self._h, ctypes.byref(n_sm), ctypes.byref(lanes)))
        self.n_sm = int(n_sm.value)
        self.resident_lanes = int(lanes.value)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().ct_engine_destroy(self._h)
            self._h = ctypes.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def warp_iterations(self) -> int:
        """Debug counter, see ``ct_engine_warp_iterations``."""
        return int(lib().ct_engine_warp_iterations(self._h))

    @property
    def launch_count(self) -> int:
        return int(lib().ct_engine_launch_count(self._h))

    # -- device-buffer path (torch tensors are only buffers: pass tensor.data_ptr()) --------------
    def launch_rewards(self, handles: np.ndarray, n_traj: np.ndarray, traj_base: np.ndarray, cfg: SimConfig,
                       seed: int, stream: int, d_reward: int, d_lag: int = 0, d_events: int = 0,
                       d_params: int = 0, d_series: int = 0, series_stride: int = 0, job_cost=None) -> None:
        handles = np.ascontiguousarray(handles, dtype=np.uint64)
        n_traj = np.ascontiguousarray(n_traj, dtype=np.uint32)
        traj_base = np.ascontiguousarray(traj_base, dtype=np.uint64)
        if job_cost is not None:
            job_cost = np.ascontiguousarray(job_cost, dtype=np.float32)
        check(lib().ct_launch_rewards(self._h, handles.ctypes.data, n_traj.ctypes.data, traj_base.ctypes.data,
                                      _ptr(job_cost), len(handles), ctypes.byref(cfg), seed, stream or None, d_params or None,
                                      d_reward or None, d_lag or None, d_events or None, d_series or None,
                                      series_stride))

    def reduce_jobs(self, n_traj: np.ndarray, stream: int, d_reward: int, d_events: int, d_job_mean: int,
                    d_job_events: int) -> None:
        n_traj = np.ascontiguousarray(n_traj, dtype=np.uint32)
        check(lib().ct_reduce_jobs(self._h, n_traj.ctypes.data, len(n_traj), stream or None, d_reward or None,
                                   d_events or None, d_job_mean or None, d_job_events or None))

    # -- host-buffer path ---------------------------------------------------------------------------
    def rewards_host(self, handles: Sequence[int], n_traj: Sequence[int], traj_base: Sequence[int], cfg: SimConfig,
                     seed: int, params: Optional[np.ndarray] = None, per_trajectory: bool = False, job_cost=None):
        handles = np.ascontiguousarray(handles, dtype=np.uint64)
        n_traj = np.ascontiguousarray(n_traj, dtype=np.uint32)
        traj_base = np.ascontiguousarray(traj_base, dtype=np.uint64)
        n_jobs = len(handles)
        total = int(n_traj.sum())
        job_mean = np.zeros(n_jobs, dtype=np.float64)
        job_events = np.zeros(n_jobs, dtype=np.uint64)
        reward = np.zeros(total, dtype=np.float32) if per_trajectory else None
        lag = np.zeros(total, dtype=np.int32) if per_trajectory else None
        if params is not None:
            params = _params_array(params, handles, total)
        if job_cost is not None:
            job_cost = np.ascontiguousarray(job_cost, dtype=np.float32)
        check(lib().ct_rewards_host(self._h, handles.ctypes.data, n_traj.ctypes.data, traj_base.ctypes.data, _ptr(job_cost), n_jobs,
                                    ctypes.byref(cfg), seed, _ptr(params), job_mean.ctypes.data,
                                    job_events.ctypes.data, _ptr(reward), _ptr(lag)))
        out = {"job_mean": job_mean, "job_events": job_events}
        if per_trajectory:
            out["reward"] = reward
            out["lag"] = lag
        return out

    def submit(self, handles, n_traj, traj_base, cfg: SimConfig, seed: int, params: Optional[np.ndarray] = None,
               job_cost=None) -> dict:
        """Asynchronous ``rewards_host``: returns a ticket dict to pass to :meth:`wait`."""
        handles = np.ascontiguousarray(handles, dtype=np.uint64)
        n_traj = np.ascontiguousarray(n_traj, dtype=np.uint32)
        traj_base = np.ascontiguousarray(traj_base, dtype=np.uint64)
        total = int(n_traj.sum())
        if params is not None:
            params = _params_array(params, handles, total)
        if job_cost is not None:
            job_cost = np.ascontiguousarray(job_cost, dtype=np.float32)
        ticket = ctypes.c_int64(0)
        check(lib().ct_rewards_submit(self._h, handles.ctypes.data, n_traj.ctypes.data, traj_base.ctypes.data,
                                      _ptr(job_cost), len(handles), ctypes.byref(cfg), seed, _ptr(params),
                                      ctypes.byref(ticket)))
        return {"ticket": int(ticket.value), "n_jobs": len(handles), "total": total, "params": params}

    def wait(self, ticket: dict, per_trajectory: bool = False) -> dict:
        n_jobs, total = ticket["n_jobs"], ticket["total"]
        job_mean = np.zeros(n_jobs, dtype=np.float64)
        job_events = np.zeros(n_jobs, dtype=np.uint64)
        reward = np.zeros(total, dtype=np.float32) if per_trajectory else None
        lag = np.zeros(total, dtype=np.int32) if per_trajectory else None
        check(lib().ct_rewards_wait(self._h, ticket["ticket"], job_mean.ctypes.data, job_events.ctypes.data,
                                    _ptr(reward), _ptr(lag)))
        out = {"job_mean": job_mean, "job_events": job_events}
        if per_trajectory:
            out["reward"] = reward
            out["lag"] = lag
        return out

    def reward_from_series(self, series: np.ndarray):
        series = np.ascontiguousarray(series, dtype=np.uint8)
        n, nsp, n_dec = series.shape
        reward = np.zeros(n, dtype=np.float32)
        lag = np.zeros(n, dtype=np.int32)
        check(lib().ct_reward_from_series_host(self._h, series.ctypes.data, n, nsp, n_dec, reward.ctypes.data,
                                               lag.ctypes.data))
        return reward, lag
