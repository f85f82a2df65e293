"""ctypes loader for oracle/ssa_oracle.c -- the CPU checker of the reward path.

TEST INFRASTRUCTURE ONLY (see the header of ssa_oracle.c): imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by circuitree_b200/.  PARITY UNPINNED: upstream has no simulator; the
model is this repo's SPEC.md.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libct_oracle.so"
_lib = None

N_PARAMS = 10


def build(force: bool = False) -> Path:
    """Compile the C restatement with the system gcc (seconds)."""
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < (_HERE / "ssa_oracle.c").stat().st_mtime:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True, env=env)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_LIB_PATH))
        L.ct_oracle_run_pairwise.restype = ctypes.c_int
        L.ct_oracle_run_pairwise.argtypes = [
            ctypes.c_int,
            ctypes.c_void_p, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64,
            ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
            ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ]
        L.ct_oracle_run_pairwise_ex.restype = ctypes.c_int
        L.ct_oracle_run_pairwise_ex.argtypes = L.ct_oracle_run_pairwise.argtypes + [ctypes.c_void_p]
        L.ct_oracle_run_dimers.restype = ctypes.c_int
        L.ct_oracle_run_dimers.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64,
            ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ]
        L.ct_oracle_reward_from_series.restype = ctypes.c_double
        L.ct_oracle_reward_from_series.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.ct_oracle_philox.restype = None
        L.ct_oracle_philox.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.ct_oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().ct_oracle_max_threads())


def philox(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().ct_oracle_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def _ptr(a):
    return None if a is None else a.ctypes.data


def run_pairwise(n_species, activations, inhibitions, ranges, n_traj, *, seed=0, traj_base=0,
                 explicit_params=None, nt=2000, decim=16, dt=20.0, init_mean=10.0,
                 n_threads=1, want_series=False, want_params=False, want_final=False):
    """Simulate ``n_traj`` trajectories of one pairwise topology.

    ``activations`` / ``inhibitions`` are the int[n,2] arrays returned by
    ``SimpleNetworkGrammar.parse_genotype`` (reference circuitree/models.py:441-480).
    Returns a dict with reward (f32), lag (i32, decimated samples), n_events (u64)
    and optionally series (u8 [n_traj, n_species, nt//decim]), params (f64) and final
    (f64 [n_traj, 3, n_species]: mRNA, free protein, molecules bound at promoters at the end).
    """
    act = np.ascontiguousarray(np.asarray(activations, dtype=np.int32).reshape(-1, 2))
    inh = np.ascontiguousarray(np.asarray(inhibitions, dtype=np.int32).reshape(-1, 2))
    rng = np.ascontiguousarray(np.asarray(ranges, dtype=np.float64).reshape(N_PARAMS, 3))
    exp = None
    if explicit_params is not None:
        exp = np.ascontiguousarray(np.asarray(explicit_params, dtype=np.float64).reshape(n_traj, N_PARAMS))
    n_dec = nt // decim
    reward = np.zeros(n_traj, dtype=np.float32)
    lag = np.zeros(n_traj, dtype=np.int32)
    n_events = np.zeros(n_traj, dtype=np.uint64)
    series = np.zeros((n_traj, n_species, n_dec), dtype=np.uint8) if want_series else None
    params = np.zeros((n_traj, N_PARAMS), dtype=np.float64) if want_params else None
    final = np.zeros((n_traj, 3, n_species), dtype=np.float64) if want_final else None
    rc = lib().ct_oracle_run_pairwise_ex(
        n_species, _ptr(act), act.shape[0], _ptr(inh), inh.shape[0],
        _ptr(rng), _ptr(exp), seed, traj_base, n_traj, nt, decim, dt, init_mean, n_threads,
        _ptr(reward), _ptr(lag), _ptr(n_events), _ptr(series), _ptr(params), _ptr(final))
    if rc != 0:
        raise ValueError(f"ct_oracle_run_pairwise failed with code {rc}")
    out = {"reward": reward, "lag": lag, "n_events": n_events}
    if want_series:
        out["series"] = series
    if want_params:
        out["params"] = params
    if want_final:
        out["final"] = final
    return out


def reward_from_series(q: np.ndarray):
    """Reward and lag from a companded series u8[n_species, n_dec] (SPEC.md §7)."""
    q = np.ascontiguousarray(q, dtype=np.uint8)
    lag = ctypes.c_int(0)
    r = lib().ct_oracle_reward_from_series(q.ctypes.data, q.shape[0], q.shape[1], ctypes.byref(lag))
    return float(r), int(lag.value)


N_PARAMS_DIMER = 12


def run_dimers(n_tf, n_comp, activations, inhibitions, ranges, n_traj, *, seed=0, traj_base=0, explicit_params=None,
               nt=2000, decim=16, dt=20.0, init_mean=10.0, want_series=False, n_threads=1):
    """Dimer model (SPEC.md section 9).  ``activations`` / ``inhibitions`` are the int[n,3] arrays of
    ``DimersGrammar.parse_genotype`` (reference circuitree/models.py:844-884).  ``n_threads`` > 1 splits
    the trajectory range over Python threads (the C call releases the GIL)."""
    act = np.ascontiguousarray(np.asarray(activations, dtype=np.int32).reshape(-1, 3))
    inh = np.ascontiguousarray(np.asarray(inhibitions, dtype=np.int32).reshape(-1, 3))
    rng = np.ascontiguousarray(np.asarray(ranges, dtype=np.float64).reshape(N_PARAMS_DIMER, 3))
    exp = None
    if explicit_params is not None:
        exp = np.ascontiguousarray(np.asarray(explicit_params, dtype=np.float64).reshape(n_traj, N_PARAMS_DIMER))
    n_dec = nt // decim
    reward = np.zeros(n_traj, dtype=np.float32)
    lag = np.zeros(n_traj, dtype=np.int32)
    n_events = np.zeros(n_traj, dtype=np.uint64)
    series = np.zeros((n_traj, n_tf, n_dec), dtype=np.uint8) if want_series else None

    def work(a, b):
        rc = lib().ct_oracle_run_dimers(
            n_tf, n_comp, _ptr(act), act.shape[0], _ptr(inh), inh.shape[0], _ptr(rng),
            None if exp is None else exp[a:b].ctypes.data, seed, traj_base + a, b - a, nt, decim, dt, init_mean,
            reward[a:b].ctypes.data, lag[a:b].ctypes.data, n_events[a:b].ctypes.data,
            None if series is None else series[a:b].ctypes.data)
        if rc != 0:
            raise ValueError(f"ct_oracle_run_dimers failed with code {rc}")

    if n_threads <= 1 or n_traj < 2 * n_threads:
        work(0, n_traj)
    else:
        from concurrent.futures import ThreadPoolExecutor

        cuts = np.linspace(0, n_traj, n_threads * 4 + 1).astype(int)
        with ThreadPoolExecutor(n_threads) as ex:
            list(ex.map(lambda ab: work(*ab), zip(cuts[:-1], cuts[1:])))
    out = {"reward": reward, "lag": lag, "n_events": n_events}
    if want_series:
        out["series"] = series
    return out
