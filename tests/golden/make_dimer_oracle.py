"""Records the CPU checker's per-trajectory results for the dimer-model parity cases
(tests/test_gpu_dimers.py) so that the GPU run does not spend minutes of box time inside the
single-precision-independent, double-precision C checker (1.7e6 events per trajectory for some
of these topologies).  Run from the repo root:  python tests/golden/make_dimer_oracle.py

The stored arrays are exactly what ``oracle.ssa_oracle.run_dimers`` returns for the listed
(case, seed, n, grid); tests/test_oracle.py re-runs a slice of every case and compares bit for bit.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from circuitree_b200 import DimersGrammar, _capi_defaults  # noqa: E402
from oracle import ssa_oracle as so  # noqa: E402

AI = ["activates", "inhibits"]
RANGES = _capi_defaults.SPEC_DEFAULTS["ranges"] + _capi_defaults.SPEC_DEFAULTS["dimer_ranges"]
G4 = dict(components=["A", "B", "C"], regulators=["R"], interactions=AI)
G5 = dict(components=list("ABCDE"), regulators=[], interactions=AI)

# name -> (grammar kwargs, state, n trajectories, seed, nt)
CASES = {
    "homodimer_repressilator": (G4, "*ABC+R::AAiB_BBiC_CCiA", 10_000, 202, 2000),
    "heterodimer_regulator": (G4, "*ABC+R::ARaB_BBiC_CCiA", 2500, 202, 800),
    "no_interactions": (G4, "*ABC+R::", 2500, 202, 800),
    "shared_heterodimer": (G4, "*ABC+R::ABiA_ABiB_CRaC_CCiA", 2500, 202, 800),
    "five_tf_ten_interactions": (G5, "*ABCDE+::AAiB_BBiC_CCiD_DDiE_EEiA_ABaA_BCaB_CDaC_DEaD_AEaE", 1000, 32, 800),
    "same_stream": (G4, "*ABC+R::AAiB_BBiC_CCiA", 1000, 9, 2000),
    # a promoter with an activating AND an inhibiting dimer (km_act_rep, the A bookkeeping), a heterodimer
    # with the regulator, and a heterodimer of two components on a third promoter
    "mixed_promoter": (G4, "*ABC+R::AAaA_BRiA_ABiC", 2500, 202, 800),
}


def run_case(name, n=None, n_threads=None):
    gk, state, n_full, seed, nt = CASES[name]
    comps, regs, act, inh = DimersGrammar(**gk).parse_genotype(state)
    return so.run_dimers(len(comps) + len(regs), len(comps), act, inh, RANGES, n_full if n is None else n, seed=seed, nt=nt,
                         decim=16, dt=20.0, init_mean=10.0, n_threads=n_threads or so.max_threads())


def main():
    out = {}
    target = Path(__file__).with_name("dimer_oracle.npz")
    only = sys.argv[1:]          # names to (re)record; the other cases are kept from the existing file
    if only and target.exists():
        out = dict(np.load(target))
    for name in (only or CASES):
        r = run_case(name)
        out[name + "/reward"] = r["reward"]
        out[name + "/lag"] = r["lag"].astype(np.int16)
        out[name + "/n_events"] = r["n_events"].astype(np.uint32)
        print(name, r["reward"].mean(), r["n_events"].mean(), flush=True)
    np.savez_compressed(target, **out)


if __name__ == "__main__":
    main()
