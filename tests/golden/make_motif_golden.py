"""Records the reference's motif-significance path for parity tests:
sampling of terminal / successful states, grow_tree_from_leaves, the contingency tests and odds
ratios of test_pattern_significance (reference circuitree/circuitree.py:1200-1648, 1680-1800).
Run in the build container, where /root/reference is mounted:

    python tests/golden/make_motif_golden.py

Writes tests/golden/motif_vectors.json.gz.  The toy reward / success rule below is duplicated in
tests/test_motifs.py so that both sides search the same problem.
"""
import gzip
import json
import sys
import zlib
from pathlib import Path

import numpy as np

sys.path.insert(0, "/root/reference")
import circuitree as ref  # noqa: E402
from circuitree.models import SimpleNetworkGrammar  # noqa: E402

PATTERNS = ["ABi_BAi", "ABa_BAi", "AAa", "ABi_BCi_CAi", "AAi_BBi_CCi", "ABa", "ABa_BCa_CAa", "AAa_ABi_BAi_BBa"]


def toy_reward(state: str) -> float:
    return (zlib.crc32(state.encode()) % 1000) / 1000.0


def toy_success(state: str) -> bool:
    return state.startswith("*") and zlib.crc32(("s" + state).encode()) % 4 == 0


class Tree(ref.CircuiTree):
    def get_reward(self, state, **kw):
        return toy_reward(state)

    def is_success(self, state):
        return toy_success(state)


def df_to_json(df):
    out = {"columns": list(df.columns), "rows": []}
    for _, row in df.iterrows():
        vals = []
        for v in row.values:
            if v is None or (not isinstance(v, str) and v is not True and v is not False and np.ndim(v) == 0 and (
                    v is ref.circuitree.pd.NA or (isinstance(v, float) and np.isnan(v)))):
                vals.append(None)
            elif isinstance(v, (np.integer, int)) and not isinstance(v, bool):
                vals.append(int(v))
            elif isinstance(v, (np.floating, float)):
                vals.append("inf" if np.isinf(v) else float(v))
            else:
                vals.append(str(v))
        out["rows"].append(vals)
    return out


def main():
    grammar = SimpleNetworkGrammar(components=["A", "B", "C"], interactions=["activates", "inhibits"], root="ABC::")
    out = {"patterns": PATTERNS}
    t = Tree(grammar=grammar, root="ABC::", seed=7, n_exhausted=np.inf)
    t.grow_tree()
    out["n_terminal"] = sum(1 for _ in t.terminal_states)
    out["n_success"] = sum(1 for s in t.terminal_states if t.is_success(s))
    out["sample_terminal_states"] = [str(s) for s in t.sample_terminal_states(150)]
    out["by_rejection"] = [str(s) for s in t.sample_successful_circuits_by_rejection(60)]
    out["by_enumeration"] = [str(s) for s in t.sample_successful_circuits_by_enumeration(60)]
    out["rng_after_sampling"] = int(t.rg.integers(0, 2**62))
    leaves = sorted(s for s in t.terminal_states if t.is_success(s))[:25]
    g = t.grow_tree_from_leaves(leaves)
    out["from_leaves"] = {"leaves": leaves, "nodes": sorted(map(str, g.nodes)), "edges": sorted([str(a), str(b)] for a, b in g.edges)}

    t2 = Tree(grammar=grammar, root="ABC::", seed=11, n_exhausted=np.inf)
    t2.grow_tree()
    df = t2.test_pattern_significance(PATTERNS, n_samples=400, confidence=0.95)
    out["significance"] = df_to_json(df)
    out["rng_after_significance"] = int(t2.rg.integers(0, 2**62))
    df = t2.test_pattern_significance(PATTERNS[:4], n_samples=120, confidence=None, sampling_method="enumeration",
                                      exclude_self=False, correction=False)
    out["significance_enumeration"] = df_to_json(df)

    tables = [[[30, 50], [70, 150]], [[3, 40], [97, 160]], [[0, 10], [50, 40]], [[2, 3], [4, 60]], [[12, 12], [12, 12]]]
    out["contingency"] = []
    for tab in tables:
        for corr, barnard in ((True, True), (False, False)):
            d = ref.circuitree.contingency_test(np.array(tab), correction=corr, barnard_ok=barnard)
            out["contingency"].append({"table": tab, "correction": corr, "barnard_ok": barnard, "df": df_to_json(d.reset_index())})
    abcd = np.array([[30, 50, 70, 150], [3, 40, 97, 160], [0, 10, 50, 40], [5, 0, 4, 60], [12, 12, 12, 12]])
    orr = ref.circuitree.compute_odds_ratios(abcd)
    o2, lo, hi = ref.circuitree.compute_odds_ratios_with_ci(abcd, 0.9)
    js = lambda a: [None if np.isnan(x) else ("inf" if np.isinf(x) else float(x)) for x in a]
    out["odds"] = {"abcd": abcd.tolist(), "ratios": js(orr), "ratios_ci": js(o2), "lo": js(lo), "hi": js(hi), "confidence": 0.9}

    path = Path(__file__).with_name("motif_vectors.json.gz")
    with gzip.open(path, "wt") as f:
        json.dump(out, f)
    print("wrote", path, {k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})
    print(json.dumps(out["significance"], indent=0)[:1500])


if __name__ == "__main__":
    main()
