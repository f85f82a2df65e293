"""Records the CPU checker's results for the randomised parity sweep (tests/test_gpu_ssa.py::
test_random_topology_sweep): 200 random terminal topologies (100 with three, 100 with four
components) x 1000 trajectories on a 640-point grid.  Run from the repo root:

    python tests/golden/make_sweep_oracle.py

Stored per topology: the state string, mean and variance of reward and of events per trajectory
(float64), the lag histogram, and the first 16 trajectories' raw (reward, lag, n_events) so that
tests/test_oracle.py can re-run a slice and compare bit for bit.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from circuitree_b200 import CircuiTree, SimpleNetworkGrammar, _capi_defaults  # noqa: E402
from oracle import ssa_oracle as so  # noqa: E402

N_TRAJ, NT, DECIM, SEED, N_RAW = 1000, 640, 16, 4242, 16
RANGES = _capi_defaults.SPEC_DEFAULTS["ranges"]


class _T(CircuiTree):
    def get_reward(self, state, **kw):
        return 0.0


def random_terminals(n_comp: int, count: int, seed: int) -> list[str]:
    comps = list("ABCDE"[:n_comp])
    g = SimpleNetworkGrammar(comps, ["activates", "inhibits"], root="".join(comps) + "::")
    t = _T(grammar=g, root=g.root, seed=seed, n_exhausted=np.inf)
    out: list[str] = []
    while len(out) < count:
        s = t.get_random_terminal_descendant(g.root)
        if s not in out:
            out.append(s)
    return out


def states() -> list[str]:
    return random_terminals(3, 100, 11) + random_terminals(4, 100, 12)


def run_state(index: int, state: str, n: int = N_TRAJ, n_threads=None):
    comps, act, inh = SimpleNetworkGrammar.parse_genotype(state)
    return so.run_pairwise(len(comps), act, inh, RANGES, n, seed=SEED, traj_base=index * 100_000, nt=NT, decim=DECIM, dt=20.0,
                           init_mean=10.0, n_threads=n_threads or so.max_threads())


def main():
    sts = states()
    cols = {k: [] for k in ("reward_mean", "reward_var", "events_mean", "events_var")}
    raw_r, raw_l, raw_e, hist = [], [], [], []
    for i, s in enumerate(sts):
        r = run_state(i, s)
        rw, ev = r["reward"].astype(np.float64), r["n_events"].astype(np.float64)
        cols["reward_mean"].append(rw.mean()); cols["reward_var"].append(rw.var(ddof=1))
        cols["events_mean"].append(ev.mean()); cols["events_var"].append(ev.var(ddof=1))
        hist.append(np.bincount(r["lag"], minlength=NT // DECIM // 2 + 1))
        raw_r.append(r["reward"][:N_RAW]); raw_l.append(r["lag"][:N_RAW].astype(np.int16)); raw_e.append(r["n_events"][:N_RAW].astype(np.uint32))
        print(i, s, round(rw.mean(), 4), int(ev.mean()), flush=True)
    np.savez_compressed(Path(__file__).with_name("sweep_oracle.npz"), states=np.array(sts), lag_hist=np.array(hist, dtype=np.int32),
                        raw_reward=np.array(raw_r), raw_lag=np.array(raw_l), raw_events=np.array(raw_e),
                        **{k: np.array(v) for k, v in cols.items()})


if __name__ == "__main__":
    main()
