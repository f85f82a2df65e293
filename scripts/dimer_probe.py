"""GPU probe for the dimer kernel: timings per topology and batch size (progress flushed line by line)."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from circuitree_b200 import DimersGrammar, _capi

AI = ["activates", "inhibits"]
G = DimersGrammar(components=["A", "B", "C"], regulators=["R"], interactions=AI)
G5 = DimersGrammar(components=list("ABCDE"), regulators=[], interactions=AI)
eng = _capi.Engine(0)
print("engine up", flush=True)


def handle(g, s):
    comps, regs, act, inh = g.parse_genotype(s)
    return _capi.compile_topology(len(comps) + len(regs), act, inh, k=3)


cases = [(G, "*ABC+R::AAiB_BBiC_CCiA"), (G, "*ABC+R::"), (G, "*ABC+R::ABiA_ABiB_CRaC_CCiA"),
         (G5, "*ABCDE+::AAiB_BBiC_CCiD_DDiE_EEiA_ABaA_BCaB_CDaC_DEaD_AEaE")]
sizes = [int(x) for x in sys.argv[1:]] or [256, 4096]
for nt in (64, 2000):
    cfg = _capi.default_config()
    cfg.nt = nt
    cfg.decim = 1 if nt == 64 else 16
    for g, s in cases:
        for n in sizes:
            t0 = time.perf_counter()
            out = eng.rewards_host([handle(g, s)], [n], [0], cfg, 1, per_trajectory=True)
            dt = time.perf_counter() - t0
            ev = int(out["job_events"][0])
            print(f"nt={nt} {s} n={n}: {dt*1e3:.1f} ms, {ev} events, {ev/dt:.3e} steps/s, reward {out['job_mean'][0]:.4f}", flush=True)
