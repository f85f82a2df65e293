"""Motif significance on search / landscape results.

Host-side mirror of the reference's sampling and contingency-test path
(reference circuitree/circuitree.py:1200-1648 and the helpers at :1680-1800): draw terminal states
from the whole design space ("null") and from the successful part of it, count in how many of
each a pattern occurs, test the 2x2 table and report odds ratios.  Same names, argument meaning,
RNG consumption and output tables as the reference (pinned by tests/golden/motif_vectors.json.gz,
recorded from the reference), but counted over *distinct* samples (a pattern test per distinct
state, weighted by multiplicity) and with vectorised odds ratios, so that the 10^4-10^6 samples a
GPU-scale landscape produces are cheap to analyse.

``MotifMixin`` is mixed into :class:`circuitree_b200.CircuiTree`.
"""
from __future__ import annotations

from collections import Counter
from typing import Any, Hashable, Iterable, Literal, Optional, Sequence

import networkx as nx
import numpy as np

__all__ = ["contingency_test", "compute_odds_ratios", "compute_odds_ratios_with_ci", "compute_odds_ratio_and_ci",
           "pattern_table", "MotifMixin"]


# ---- 2x2 tables -------------------------------------------------------------------------------------
def _odds(abcd: np.ndarray) -> np.ndarray:
    """a d / (b c) per row of ``[a, b, c, d]``; ``inf`` where b c = 0 (reference circuitree.py:1704-1727)."""
    t = np.asarray(abcd, dtype=np.float64).reshape(-1, 4)
    ad, bc = t[:, 0] * t[:, 3], t[:, 1] * t[:, 2]
    out = np.full(len(t), np.inf)
    np.divide(ad, bc, out=out, where=bc != 0)
    return out


def compute_odds_ratios(abcd: np.ndarray, progress: bool = False) -> np.ndarray:
    """Odds ratios of many 2x2 tables ``[[a, b], [c, d]]`` given as rows ``[a, b, c, d]``."""
    return _odds(abcd)


def compute_odds_ratios_with_ci(abcd: np.ndarray, confidence_level: float, progress: bool = False):
    """Odds ratios with the normal-approximation confidence interval of the log odds ratio
    (Woolf); the interval is NaN for a table with an empty cell (reference circuitree.py:1680-1700, :1730-1752)."""
    from scipy import stats

    t = np.asarray(abcd, dtype=np.float64).reshape(-1, 4)
    ratios = _odds(t)
    ok = (t != 0).all(axis=1)
    lo = np.full(len(t), np.nan)
    hi = np.full(len(t), np.nan)
    if ok.any():
        z = stats.norm.ppf((1.0 + confidence_level) / 2.0)
        se = np.sqrt((1.0 / t[ok]).sum(axis=1))
        log_or = np.log(ratios[ok])
        lo[ok] = np.exp(log_or - z * se)
        hi[ok] = np.exp(log_or + z * se)
    return ratios, lo, hi


def compute_odds_ratio_and_ci(table: np.ndarray, confidence_level: float):
    r, lo, hi = compute_odds_ratios_with_ci(np.asarray(table).reshape(1, 4), confidence_level)
    return float(r[0]), (float(lo[0]), float(hi[0]))


def _test_table(table: np.ndarray, correction: bool, barnard_ok: bool):
    """(test name or None, statistic, p) of one 2x2 table: chi-square when every cell holds at
    least 5, Barnard's exact test for smaller non-empty cells if allowed, nothing for an empty cell
    (reference circuitree.py:1755-1790)."""
    from scipy import stats

    smallest = int(np.min(table))
    if smallest == 0 or (smallest < 5 and not barnard_ok):
        return None, np.nan, np.nan
    if smallest < 5:
        try:
            res = stats.barnard_exact(table, alternative="two-sided")
        except MemoryError:
            return None, np.nan, np.nan
        return "barnard", float(res.statistic), float(res.pvalue)
    res = stats.chi2_contingency(table, correction=correction)
    return "chi2", float(res.statistic), float(res.pvalue)


def contingency_test(table: np.ndarray, correction: bool = True, barnard_ok: bool = True):
    """Two-tailed test for P(has_pattern | successful) != P(has_pattern); rows = pattern present /
    absent, columns = successful / overall paths.  Returns the reference's DataFrame layout."""
    import pandas as pd

    table = np.asarray(table)
    name, stat, p = _test_table(table, correction, barnard_ok)
    df = pd.DataFrame(table, index=["has_pattern", "lacks_pattern"], columns=["successful_paths", "overall_paths"])
    df.index.name = "pattern_present"
    df["test"] = pd.NA if name is None else name
    df["statistic"] = stat
    df["pvalue"] = p
    return df


def pattern_table(grammar, patterns: Sequence[Any], null_samples: Iterable[Hashable], succ_samples: Iterable[Hashable],
                  confidence: Optional[float] = 0.95, correction: bool = True, barnard_ok: bool = True,
                  exclude_self: bool = True):
    """The table ``test_pattern_significance`` returns, from given samples: per pattern the test,
    the four counts, the Bonferroni-corrected p-value and the odds ratio (with confidence interval),
    sorted by odds ratio, largest first."""
    import pandas as pd

    patterns = list(patterns)
    null_count, succ_count = Counter(null_samples), Counter(succ_samples)
    distinct = list(set(null_count) | set(succ_count))
    rows = []
    for pat in patterns:
        hit = {s for s in distinct if s != pat or not exclude_self if grammar.has_pattern(s, pat)}
        skip = pat if exclude_self else None       # a sample equal to the pattern itself is left out of both sides
        n_null = sum(c for s, c in null_count.items() if s != skip)
        n_succ = sum(c for s, c in succ_count.items() if s != skip)
        in_null = sum(c for s, c in null_count.items() if s in hit)
        in_succ = sum(c for s, c in succ_count.items() if s in hit)
        name, stat, p = _test_table(np.array([[in_succ, in_null], [n_succ - in_succ, n_null - in_null]]), correction, barnard_ok)
        rows.append((pat, pd.NA if name is None else name, stat, p, n_null - in_null, in_null, n_succ - in_succ, in_succ))
    df = pd.DataFrame(rows, columns=["pattern", "test", "statistic", "pvalue", "others_in_null", "pattern_in_null",
                                     "others_in_succ", "pattern_in_succ"])
    df["p_corrected"] = df["pvalue"] * len(df)
    abcd = df[["pattern_in_succ", "pattern_in_null", "others_in_succ", "others_in_null"]].to_numpy()
    if confidence is None:
        df["odds_ratio"] = compute_odds_ratios(abcd)
    else:
        ratios, lo, hi = compute_odds_ratios_with_ci(abcd, confidence)
        level = f"{int(confidence * 100)}%"
        df["odds_ratio"] = ratios
        df[f"ci_{level}_low"] = lo
        df[f"ci_{level}_high"] = hi
    return df.sort_values("odds_ratio", ascending=False)


# ---- sampling ---------------------------------------------------------------------------------------
class MotifMixin:
    """Sampling and pattern-significance methods of ``CircuiTree`` (reference circuitree.py:1200-1648)."""

    def _spawned_generators(self, n: int):
        """One fresh generator per sample, as the reference's multi-process paths create them
        (``SeedSequence(seed).spawn(1)[0]`` per draw); consumed here in order, in this process."""
        seq = np.random.SeedSequence(self.seed)
        return (np.random.default_rng(seq.spawn(1)[0]) for _ in range(n))

    def sample_terminal_states(self, n_samples: int, progress: bool = False, nprocs: int = 1, chunksize: int = 100) -> list:
        """``n_samples`` random terminal states of the grammar: random walks from the root."""
        if nprocs == 1:
            return [self.get_random_terminal_descendant(self.root) for _ in range(n_samples)]
        return [self._get_random_terminal_descendant(self.grammar, self.root, rg) for rg in self._spawned_generators(n_samples)]

    @staticmethod
    def _sample_leaf(graph: nx.DiGraph, root: Hashable, rg: np.random.Generator):
        """Random walk along the edges of ``graph`` until a node without successors."""
        node = root
        while True:
            nxt = list(graph.successors(node))
            if not nxt:
                return node
            node = rg.choice(nxt)

    def _sample_leaves(self, n_leaves: int, graph: Optional[nx.DiGraph] = None, root: Optional[Hashable] = True,
                       rg: Optional[np.random.Generator] = None, terminal: bool = True, progress: bool = False,
                       max_iter: int = 10_000_000) -> list:
        """Leaves of ``graph`` by random traversal from the root; with ``terminal`` non-terminal
        leaves are rejected."""
        graph = self.graph if graph is None else graph
        root = self.root if root is True else root
        rg = self.rg if rg is None else rg
        if terminal and not any(self.grammar.is_terminal(n) for n in graph.nodes):
            raise ValueError("No terminal states in the graph.")
        samples: list = []
        for _ in range(max_iter):
            leaf = self._sample_leaf(graph, root, rg)
            if terminal and not self.grammar.is_terminal(leaf):
                continue
            samples.append(leaf)
            if len(samples) == n_leaves:
                return samples
        raise RuntimeError(f"Maximum number of iterations reached: {max_iter}")

    def _successful_terminals(self) -> list:
        try:
            self.is_success(self.root)
        except NotImplementedError:
            raise NotImplementedError(
                "The CircuiTree subclass must implement the is_success() method to use this function.") from None
        return [s for s in self.terminal_states if self.is_success(s)]

    def sample_successful_circuits_by_enumeration(self, n_samples: int, progress: bool = False, nprocs: int = 1,
                                                  chunksize: int = 100) -> list:
        """Successful terminal states by random traversal of the graph of *all* paths from the root
        to a successful terminal state.  (The reference builds that graph from a ``set``, so its
        traversal order -- and with it the sample -- changes with ``PYTHONHASHSEED``; here the
        successful states enter in graph order, which makes the draw reproducible.)"""
        paths = self.grow_tree_from_leaves(self._successful_terminals())
        if nprocs == 1:
            return self._sample_leaves(n_samples, graph=paths, terminal=True)
        return [self._sample_leaf(paths, self.root, rg) for rg in self._spawned_generators(n_samples)]

    def sample_successful_circuits_by_rejection(self, n_samples: int, max_iter: int = 10_000_000, progress: bool = False,
                                                nprocs: int = 1, chunksize: int = 100) -> list:
        """Successful terminal states by rejection: random walks from the root, kept when they end
        in a successful terminal state of the search graph."""
        wanted = set(self._successful_terminals())
        samples: list = []
        if nprocs == 1:
            for _ in range(max_iter):
                state = self.get_random_terminal_descendant(self.root)
                if state in wanted:
                    samples.append(state)
                    if len(samples) == n_samples:
                        return samples
            raise RuntimeError(f"Maximum number of iterations reached: {max_iter}")
        # one seeded stream per requested sample, each walking until it hits a successful state
        for seed in np.random.SeedSequence(self.seed).generate_state(n_samples):
            rg = np.random.default_rng(seed)
            for _ in range(max_iter):
                state = self._get_random_terminal_descendant(self.grammar, self.root, rg)
                if state in wanted:
                    samples.append(state)
                    break
        if len(samples) < n_samples:
            raise RuntimeError(f"Maximum number of iterations reached: {max_iter}")
        return samples

    def grow_tree_from_leaves(self, leaves: Iterable[Hashable]) -> nx.DiGraph:
        """The graph of all paths from the root to a node in ``leaves``, grown backwards with the
        grammar's undo actions (same insertion order as the reference, circuitree.py:1650-1678,
        because random traversal follows networkx's successor order)."""
        leaves = list(leaves)
        graph = nx.DiGraph()
        graph.add_nodes_from(leaves)
        # LIFO work list of states to open (their undo actions go on top, in order) and of
        # (child, undo action) links to resolve
        todo: list = [(leaf, None) for leaf in leaves]
        while todo:
            state, action = todo.pop()
            if action is None:
                graph.add_node(state, **self.default_attrs)
                todo.extend((state, a) for a in self.grammar.get_undo_actions(state))
                continue
            parent = self._undo_action(state, action)
            if graph.has_edge(parent, state):
                continue
            if parent not in graph:
                graph.add_node(parent, **self.default_attrs)
                todo.extend((parent, a) for a in self.grammar.get_undo_actions(parent))
            graph.add_edge(parent, state, **self.default_attrs)
        return graph

    @staticmethod
    def _contingency_test(pattern: Hashable, grammar, null_samples: list, succ_samples: list, correction: bool = True,
                          barnard_ok: bool = True, exclude_self: bool = True):
        """The 2x2 table of one pattern with its test, indexed by (pattern, has_pattern)."""
        import pandas as pd

        if exclude_self:
            null_samples = [s for s in null_samples if s != pattern]
            succ_samples = [s for s in succ_samples if s != pattern]
        in_null = sum(grammar.has_pattern(s, pattern) for s in null_samples)
        in_succ = sum(grammar.has_pattern(s, pattern) for s in succ_samples)
        df = contingency_test(np.array([[in_succ, in_null], [len(succ_samples) - in_succ, len(null_samples) - in_null]]),
                              correction=correction, barnard_ok=barnard_ok)
        df.index = pd.MultiIndex.from_tuples([(pattern, True), (pattern, False)], names=["pattern", "has_pattern"])
        return df

    def test_pattern_significance(self, patterns: Iterable[Any], n_samples: int, confidence: Optional[float] = 0.95,
                                  correction: bool = True, progress: bool = False, null_samples: Optional[list] = None,
                                  succ_samples: Optional[list] = None,
                                  sampling_method: Literal["rejection", "enumeration"] = "rejection", nprocs_sampling: int = 1,
                                  nprocs_testing: int = 1, max_iter: int = 10_000_000, null_kwargs: Optional[dict] = None,
                                  succ_kwargs: Optional[dict] = None, barnard_ok: bool = True, exclude_self: bool = True):
        """Is a pattern enriched among successful circuits?  Draws ``n_samples`` terminal states
        from the whole design space and ``n_samples`` successful ones (``is_success``), and tests
        each pattern's 2x2 table; returns a pandas DataFrame sorted by odds ratio (columns as in
        the reference, circuitree.py:1510-1648).  Samples can be supplied instead, e.g. the states
        of a GPU reward landscape weighted by their success fraction."""
        if null_samples is None:
            null_samples = self.sample_terminal_states(n_samples, nprocs=nprocs_sampling, **(null_kwargs or {}))
        if succ_samples is None:
            if sampling_method == "enumeration":
                succ_samples = self.sample_successful_circuits_by_enumeration(n_samples, nprocs=nprocs_sampling, **(succ_kwargs or {}))
            elif sampling_method == "rejection":
                succ_samples = self.sample_successful_circuits_by_rejection(n_samples, max_iter=max_iter, nprocs=nprocs_sampling,
                                                                            **(succ_kwargs or {}))
            else:
                raise ValueError(f"Invalid sampling method: {sampling_method}. Must be one of ['rejection', 'enumeration'].")
        return pattern_table(self.grammar, patterns, null_samples, succ_samples, confidence=confidence, correction=correction,
                             barnard_ok=barnard_ok, exclude_self=exclude_self)
