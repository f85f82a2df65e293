"""Tiny launches of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py

Short grids and a few hundred trajectories per launch: the tools slow kernels down 10-100x."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from circuitree_b200 import DimersGrammar, SimpleNetworkGrammar, _capi

eng = _capi.Engine(0)
AI = ["activates", "inhibits"]


def pair(state):
    c, a, i = SimpleNetworkGrammar.parse_genotype(state)
    return _capi.compile_topology(len(c), a, i)


def dim(g, state):
    c, r, a, i = g.parse_genotype(state)
    return _capi.compile_topology(len(c) + len(r), a, i, k=3)


def cfg(schedule=0, fast=1, nt=96):
    c = _capi.default_config()
    c.nt, c.decim, c.schedule, c.fast_path = nt, 16, schedule, fast
    return c


total = 0
# pairwise: single-regulator loop, general loop, mixed lanes, 3 / 4 / 5 components, ragged jobs
for states, sched, fast in (
        (["*ABC::ABi_BCi_CAi"], 1, 1), (["*ABC::ABi_BCi_CAi"], 1, 0), (["*ABC::AAa_ABi_BBa_BCi_CAi_CCa", "*AB::ABa_BAi", "*ABC::"], 2, 1),
        (["*ABCD::ABi_BCi_CDi_DAi", "*ABCD::ABa_CBi_BCa_DCi_ADi_CAa_DDa"], 2, 1), (["*ABCDE::ABi_BCi_CDi_DEi_EAi"], 1, 1),
        (["*ABCDE::ABi_BCa_CDi_DEi_EAi_AAa_EBi", "*ABCDE::", "*ABC::CAa"], 2, 1)):
    hs = [pair(s) for s in states]
    counts = [200 + 37 * k for k in range(len(hs))]
    out = eng.rewards_host(hs, counts, [1000 * k for k in range(len(hs))], cfg(sched, fast), 3, per_trajectory=True)
    total += int(out["job_events"].sum())
# dimers: one topology per warp, per-lane programs, cooperative sub-groups of 8 and of 4 lanes
G4 = DimersGrammar(components=list("ABC"), regulators=["R"], interactions=AI)
G3 = DimersGrammar(components=list("ABC"), regulators=[], interactions=AI)
G5 = DimersGrammar(components=list("ABC"), regulators=list("DE"), interactions=AI)
for g, states in ((G4, ["*ABC+R::AAiB_BBiC_CCiA", "*ABC+R::AAaA_BRiA_ABiC", "*ABC+R::"]), (G3, ["*ABC+::AAiB_BBiC_CCiA", "*ABC+::ABaA_ABiB_BCaC_ACiC"]),
                  (G5, ["*ABC+DE::ADiB_BEiC_CCiA_ABaA_DEaB_BCaC", "*ABC+DE::DDaA"])):
    hs = [dim(g, s) for s in states]
    for sched in (1, 2, 3):
        out = eng.rewards_host(hs, [40 + 9 * k for k in range(len(hs))], [1000 * k for k in range(len(hs))], cfg(sched, nt=64), 5,
                               per_trajectory=True)
        total += int(out["job_events"].sum())
# asynchronous pair + the stand-alone reward kernel
t = eng.submit([pair("*ABC::ABi_BCi_CAi")], [128], [0], cfg(), 9)
out = eng.wait(t, per_trajectory=True)
rw, lg = eng.reward_from_series(np.random.default_rng(0).integers(0, 255, size=(64, 3, 125), dtype=np.uint8))
print("sanitize smoke done:", total, "events", float(out["reward"].mean()), float(rw.mean()))
