"""Built-in grammars: pairwise networks and dimerising TF networks.

Host-side mirrors of the reference's ``SimpleNetworkGrammar`` (reference
circuitree/models.py:23-480) and ``DimersGrammar`` (models.py:524-939).  State strings, action
strings, action order and canonical forms are bit-identical to the reference wherever the
reference is deterministic and runnable (pinned by tests/golden); deviations are listed in each
docstring.  The implementation is independent: states are parsed once into index tuples and
canonicalised over precomputed permutation tables instead of recolouring strings.
"""
from __future__ import annotations

from collections import Counter, OrderedDict
from itertools import combinations, permutations, product
from typing import Iterable, Optional

import numpy as np

from .grammar import CircuitGrammar

__all__ = ["SimpleNetworkGrammar", "DimersGrammar", "SimpleNetworkTree", "DimerNetworkTree"]

TERMINATE = "*terminate*"


class _LRU:
    """Tiny LRU map (``maxsize=None`` = unbounded) used for canonical-form memoisation."""

    def __init__(self, maxsize: Optional[int]):
        self.maxsize = maxsize
        self.data: OrderedDict = OrderedDict()

    def get(self, key):
        try:
            val = self.data[key]
        except KeyError:
            return None
        if self.maxsize is not None:
            self.data.move_to_end(key)
        return val

    def put(self, key, val):
        self.data[key] = val
        if self.maxsize is not None and len(self.data) > self.maxsize:
            self.data.popitem(last=False)


class SimpleNetworkGrammar(CircuitGrammar):
    """Pairwise interaction networks, e.g. ``*MP::MPi_PMa`` (reference models.py:23-480).

    * components: one uppercase letter each; interaction types: one lowercase letter each
    * state = ``<components>::<interactions joined by _>``; a leading ``*`` marks a terminal state
    * actions: a 3-character interaction ``ABa``, a 1-character component, or ``*terminate*``

    Same constructor arguments and serialised attributes as the reference.  Deviation: the
    reference lists "add component" actions in Python ``set`` order (hash-seed dependent,
    models.py:195); here they come in declaration order.
    """

    def __init__(
        self,
        components: str | Iterable[str],
        interactions: str | Iterable[str],
        max_interactions: Optional[int] = None,
        root: Optional[str] = None,
        fixed_components: Optional[list[str]] = None,
        cache_maxsize: int | None = 128,
        *args,
        **kwargs,
    ):
        super().__init__(*args, **kwargs)
        components = list(components)
        interactions = list(interactions)
        if len({c[0].upper() for c in components}) < len(components):
            raise ValueError("First character of each component must be unique")
        if len({x[0].lower() for x in interactions}) < len(interactions):
            raise ValueError("First character of each interaction must be unique")

        self.root = root
        self.components = components
        self.component_map = {c[0].upper(): c for c in components}
        self.interactions = interactions
        self.interaction_map = {x[0].lower(): x for x in interactions}
        self.max_interactions = len(components) ** 2 if max_interactions is None else max_interactions
        self.fixed_components = list(fixed_components or [])
        self.cache_maxsize = cache_maxsize
        self._non_serializable_attrs.extend(
            ["component_map", "interaction_map", "edge_options", "component_codes", "_recolor",
             "get_interaction_recolorings", "_codes", "_types", "_edge_groups", "_perm_maps", "_unique_cache",
             "_ixn_cache"]
        )
        self._build_tables()

    # -- derived tables (never serialised) -----------------------------------------------------------
    def _build_tables(self) -> None:
        self._codes = [c[0].upper() for c in self.components]
        self._types = [x[0].lower() for x in self.interactions]
        # one group per ordered pair, in product(components, components) order (models.py:149-158)
        self._edge_groups = [(a + b, [a + b + t for t in self._types]) for a, b in product(self._codes, self._codes)]
        movable = [c for c, full in zip(self._codes, self.components)
                   if full not in self.fixed_components and c not in self.fixed_components]
        self._perm_maps = [str.maketrans("".join(movable), "".join(p)) for p in permutations(movable)]
        self._unique_cache = _LRU(self.cache_maxsize)
        self._ixn_cache = _LRU(self.cache_maxsize)

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in ("_perm_maps", "_unique_cache", "_ixn_cache"):
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        self.__dict__ = state
        self._build_tables()

    @property
    def component_codes(self) -> set[str]:
        return set(self._codes)

    @property
    def _recolorable_components(self) -> list[str]:
        return [c for c, full in zip(self._codes, self.components)
                if full not in self.fixed_components and c not in self.fixed_components]

    # -- game rules ------------------------------------------------------------------------------------
    @staticmethod
    def is_terminal(genotype: str) -> bool:
        return genotype.startswith("*")

    def get_actions(self, genotype: str) -> list[str]:
        """Legal actions (reference models.py:164-226), including the reference's quirk that a
        state with zero interactions offers only interactions: no ``*terminate*`` and no
        component additions (models.py:198-204)."""
        if genotype.startswith("*"):
            return []
        comp_part, ixn_part = genotype.strip("*").split("::")
        present = set(comp_part)
        pairs = {x[:2] for x in ixn_part.split("_") if x}
        if len(pairs) >= self.max_interactions:
            return [TERMINATE]
        if not pairs:
            out: list[str] = []
            for pair, group in self._edge_groups:
                if group and pair[0] in present and pair[1] in present:
                    out.extend(group)
            return out
        actions = [TERMINATE]
        actions.extend(c for c in self._codes if c not in present)
        for pair, group in self._edge_groups:
            if not group or pair in pairs:
                continue
            a, b = pair
            if a in present and b in present and (a in ixn_part or b in ixn_part):
                actions.extend(group)
        return actions

    def do_action(self, genotype: str, action: str) -> str:
        """Apply an action (reference models.py:228-270)."""
        if action == TERMINATE:
            return "*" + genotype
        comp_part, ixn_part = genotype.split("::")
        if len(action) == 1:
            return f"{comp_part}{action}::{ixn_part}"
        if len(action) == 3:
            return f"{comp_part}::{ixn_part}_{action}" if ixn_part else f"{comp_part}::{action}"
        raise ValueError(f"Invalid action: {action!r}")

    # -- canonical form ----------------------------------------------------------------------------------
    def _interaction_recolorings(self, ixn_part: str) -> list[str]:
        """Sorted-and-joined interactions under every permutation of the movable components, in
        ``itertools.permutations`` order (what reference models.py:350-359 returns)."""
        hit = self._ixn_cache.get(ixn_part)
        if hit is not None:
            return hit
        items = ixn_part.split("_")
        out = ["_".join(sorted(x.translate(m) for x in items)).strip("_") for m in self._perm_maps]
        self._ixn_cache.put(ixn_part, out)
        return out

    # public name kept for API parity with the reference attribute of the same name
    def get_interaction_recolorings(self, interactions: str) -> list[str]:
        return list(self._interaction_recolorings(interactions))

    def get_component_recolorings(self, components: str) -> list[str]:
        return ["".join(sorted(components.translate(m))) for m in self._perm_maps]

    def get_recolorings(self, genotype: str) -> list[str]:
        prefix = "*" if genotype.startswith("*") else ""
        comp_part, ixn_part = genotype.strip("*").split("::")
        return [f"{prefix}{c}::{i}" for c, i in
                zip(self.get_component_recolorings(comp_part), self._interaction_recolorings(ixn_part))]

    def get_unique_state(self, genotype: str) -> str:
        """Lexicographically smallest relabelling (reference models.py:384-400)."""
        hit = self._unique_cache.get(genotype)
        if hit is not None:
            return hit
        best = min(self.get_recolorings(genotype))
        self._unique_cache.put(genotype, best)
        return best

    def has_pattern(self, state: str, pattern: str) -> bool:
        """Whether some relabelling of ``pattern`` is a subset of the state's interactions
        (reference models.py:402-439)."""
        if "::" in pattern or "*" in pattern:
            raise ValueError("Pattern code should only contain interactions, no components")
        if "::" not in state:
            raise ValueError("State code should contain both components and interactions")
        ixn_part = state.split("::")[1]
        if not ixn_part:
            return False
        have = set(ixn_part.split("_"))
        return any(set(r.split("_")) <= have for r in self._interaction_recolorings(pattern))

    # -- undo (experimental upstream, models.py:272-320) -------------------------------------------------
    def get_undo_actions(self, genotype: str) -> list[str]:
        if genotype == self.root:
            return []
        if genotype.startswith("*"):
            return ["*undo_terminate*"]
        comp_part, ixn_part = genotype.split("::")
        out: list[str] = []
        if ixn_part:
            ixns = ixn_part.split("_")
            for k, ixn in enumerate(ixns):
                rest = ixns[:k] + ixns[k + 1:]
                if _n_weak_groups_single_pass([set(x[:2]) for x in rest]) < 2:
                    out.append(ixn)
        if len(comp_part) - len(self.root.split("::")[0]):
            out.extend(set(comp_part) - set(ixn_part))
        return out

    def undo_action(self, genotype: str, action: str) -> str:
        if action == "*undo_terminate*":
            if genotype.startswith("*"):
                return genotype[1:]
            raise ValueError(f"Cannot undo termination on a non-terminal genotype: {genotype}")
        comp_part, ixn_part = genotype.split("::")
        if len(action) == 1:
            comp_part = comp_part.replace(action, "")
        elif len(action) == 3:
            ixns = ixn_part.split("_")
            ixns.remove(action)
            ixn_part = "_".join(ixns)
        return f"{comp_part}::{ixn_part}"

    # -- bridge to the simulator -------------------------------------------------------------------------
    @staticmethod
    def parse_genotype(genotype: str, nonterminal_ok: bool = False):
        """``(components, activations int[n,2], inhibitions int[n,2])`` with rows
        ``(regulator, target)`` indexing the components string (reference models.py:441-480).
        This is the topology input of the CUDA engine (``ct_compile_topology``)."""
        if not genotype.startswith("*") and not nonterminal_ok:
            raise ValueError(f"Assembly incomplete. Genotype {genotype} is not a terminal genotype.")
        comp_part, ixn_part = genotype.strip("*").split("::")
        where = {c: k for k, c in enumerate(comp_part)}
        act, inh = [], []
        for ixn in ixn_part.split("_"):
            if not ixn:
                continue
            src, dst, kind = ixn
            row = (where[src], where[dst])
            if kind.lower() == "a":
                act.append(row)
            elif kind.lower() == "i":
                inh.append(row)
        return comp_part, np.array(act, dtype=np.int_), np.array(inh, dtype=np.int_)


def _n_weak_groups_single_pass(sets: list[set]) -> int:
    """Number of groups left by ONE merge of the first overlapping pair -- the behaviour of the
    reference helper ``merge_overlapping_sets`` (utils.py:27-45), whose loop runs a single pass."""
    sets = [set(s) for s in sets]
    for i, j in combinations(range(len(sets)), 2):
        if sets[i] & sets[j]:
            sets[i] |= sets.pop(j)
            break
    return len(sets)


class DimersGrammar(CircuitGrammar):
    """(Experimental upstream) dimerising TF networks, e.g. ``*AB+R::ARiB_BBaB``
    (reference models.py:524-939).

    state = ``<components>+<regulators>::<interactions>``; an interaction is 4 characters: the two
    monomers of the dimer, the regulation type, the target component.

    Bit-exact with the reference (pinned by tests/golden): ``do_action``, ``get_genotype_parts``,
    ``parse_genotype``, ``dimer_options``, and ``get_actions`` for states with at least one
    interaction (as a set; upstream order depends on the hash seed).  The following reference
    methods raise or mis-sort on every input, so they are implemented from
    their documented intent (models.py:759-778) and labelled as restatements:

    * ``get_unique_state``: smallest relabelling over permutations of components and of
      regulators, monomers sorted within each dimer, interactions sorted as whole strings.
    * ``has_pattern``: subset test under the same relabellings.
    * ``get_actions`` on a state without interactions: every interaction whose three species are
      present (the analogue of SimpleNetworkGrammar's rule).
    """

    def __init__(
        self,
        components: Iterable[str],
        regulators: Iterable[str],
        interactions: Iterable[str],
        max_interactions: Optional[int] = None,
        max_interactions_per_promoter: int = 2,
        *args,
        **kwargs,
    ):
        super().__init__(*args, **kwargs)
        components, regulators, interactions = list(components), list(regulators), list(interactions)
        if len({c[0].upper() for c in components}) < len(components):
            raise ValueError("First character of each component must be unique")
        if len({r[0].upper() for r in regulators}) < len(regulators):
            raise ValueError("First character of each regulator must be unique")
        if len({x[0].lower() for x in interactions}) < len(interactions):
            raise ValueError("First character of each interaction must be unique")
        self.components = components
        self.regulators = regulators
        self.interactions = interactions
        if max_interactions is None:
            self.max_interactions = len(components) * (len(components) + len(regulators)) ** 2
        else:
            self.max_interactions = max_interactions
        self.max_interactions_per_promoter = max_interactions_per_promoter
        self._non_serializable_attrs.extend(
            ["_component_codes", "_regulator_codes", "edge_options", "_ccodes", "_rcodes", "_tcodes", "_groups",
             "_unique_cache"]
        )
        self._ccodes = [c[0].upper() for c in components]
        self._rcodes = [r[0].upper() for r in regulators]
        self._tcodes = [x[0].lower() for x in interactions]
        self._groups = [(d, t, [d + k + t for k in self._tcodes]) for d, t in product(self.dimer_options, self._ccodes)]
        self._unique_cache: dict[str, str] = {}

    @property
    def dimer_options(self) -> list[str]:
        """Sorted two-letter dimers: homodimers and heterodimers of components, plus
        component-regulator heterodimers (reference models.py:596-612)."""
        dimers = [c + c for c in self._ccodes]
        dimers += [a + b for a, b in combinations(sorted(self._ccodes), 2)]
        dimers += ["".join(sorted(r + c)) for r, c in product(self._rcodes, self._ccodes)]
        return sorted(dimers)

    @property
    def edge_options(self) -> list[list[str]]:
        return [list(g) for _, _, g in self._groups]

    @staticmethod
    def is_terminal(genotype: str) -> bool:
        return genotype.startswith("*")

    @staticmethod
    def get_genotype_parts(genotype: str) -> tuple[str, str, str]:
        """``(components, regulators, interactions)`` (reference models.py:828-842)."""
        head, ixn_part = genotype.strip("*").split("::")
        comps, regs = head.split("+")
        return comps, regs, ixn_part

    def get_actions(self, genotype: str) -> list[str]:
        """Reference models.py:646-739 for states with interactions; see the class docstring for
        the zero-interaction case."""
        if genotype.startswith("*"):
            return []
        comps, regs, ixn_part = self.get_genotype_parts(genotype)
        present = set(comps + regs)
        ixns = [x for x in ixn_part.split("_") if x]
        triplets = {x[:2] + x[3] for x in ixns}
        if len(triplets) >= self.max_interactions:
            return [TERMINATE]
        adds = [c for c in self._ccodes if c not in comps] + ["+" + r for r in self._rcodes if r not in regs]
        if not triplets:
            out: list[str] = []
            for dimer, target, group in self._groups:
                if set(dimer + target) <= present:
                    out.extend(group)
            return out
        actions = [TERMINATE] + adds
        in_use = set("".join(triplets))
        load = Counter(x[3] for x in ixns)
        full = {p for p in comps if load[p] >= self.max_interactions_per_promoter}
        for dimer, target, group in self._groups:
            trip = dimer + target
            if target in full or trip in triplets:
                continue
            if not set(trip) <= present or not (set(trip) & in_use):
                continue
            actions.extend(group)
        return actions

    def do_action(self, genotype: str, action: str) -> str:
        """Reference models.py:741-757."""
        comps, regs, ixn_part = self.get_genotype_parts(genotype)
        # the reference strips the terminal marker here as well
        n = len(action)
        if action == TERMINATE:
            comps = "*" + comps
        elif n == 1:
            comps += action
        elif n == 2 and action[0] == "+":
            regs += action[1]
        elif n == 4:
            ixn_part = f"{ixn_part}_{action}" if ixn_part else action
        else:
            raise ValueError(f"Invalid action: {action}")
        return f"{comps}+{regs}::{ixn_part}"

    @staticmethod
    def _normalise(ixns: list[str]) -> str:
        return "_".join(sorted("".join(sorted(x[:2])) + x[2:] for x in ixns))

    def _relabelings(self, comps: str, regs: str, ixn_part: str):
        ixns = [x for x in ixn_part.split("_") if x]
        src = comps + regs
        for cp in permutations(comps):
            for rp in permutations(regs):
                table = str.maketrans(src, "".join(cp) + "".join(rp))
                yield self._normalise([x.translate(table) for x in ixns])

    def get_unique_state(self, genotype: str) -> str:
        """Canonical form from the documented intent (models.py:759-778); the reference
        implementation raises on every input."""
        hit = self._unique_cache.get(genotype)
        if hit is not None:
            return hit
        prefix = "*" if genotype.startswith("*") else ""
        comps, regs, ixn_part = self.get_genotype_parts(genotype)
        best = min(self._relabelings(comps, regs, ixn_part))
        out = f"{prefix}{''.join(sorted(comps))}+{''.join(sorted(regs))}::{best}"
        if len(self._unique_cache) < 1_000_000:
            self._unique_cache[genotype] = out
        return out

    @staticmethod
    def parse_genotype(genotype: str):
        """``(components, regulators, activations int[n,3], inhibitions int[n,3])`` with rows
        ``(monomer1, monomer2, target)``, monomers sorted, indexing ``components + regulators``
        (reference models.py:844-884)."""
        head, ixn_part = genotype.strip("*").split("::")
        comps, regs = head.split("+")
        where = {c: k for k, c in enumerate(comps + regs)}
        act, inh = [], []
        for ixn in ixn_part.split("_") if ixn_part else []:
            *monomers, kind, target = ixn
            m1, m2 = sorted(monomers)
            row = (where[m1], where[m2], where[target])
            if kind.lower() == "a":
                act.append(row)
            elif kind.lower() == "i":
                inh.append(row)
            else:
                raise ValueError(f"Unknown regulation type {kind} in {genotype}")
        return comps, regs, np.array(act, dtype=np.int_), np.array(inh, dtype=np.int_)

    def has_pattern(self, state: str, pattern: str) -> bool:
        """Subset test under relabelling (documented intent of reference models.py:886-939)."""
        if "::" in pattern or "*" in pattern:
            raise ValueError("pattern should only contain interactions, no components or regulators")
        if "::" not in state or "+" not in state:
            raise ValueError("state should contain components, regulators and interactions")
        if pattern == "":
            raise ValueError("pattern should be non-empty")
        comps, regs, ixn_part = self.get_genotype_parts(state)
        for ixn in pattern.split("_"):
            if any(ch not in comps + regs for ch in ixn[:2] + ixn[3]):
                raise ValueError("All components and regulators in the pattern should also be present in the state.")
        if not ixn_part:
            return False
        have = {"".join(sorted(x[:2])) + x[3] for x in ixn_part.split("_")}
        for relabeled in self._relabelings(comps, regs, pattern):
            if {x[:2] + x[3] for x in relabeled.split("_")} <= have:
                return True
        return False


def _tree_base():
    from .circuitree import CircuiTree      # (circuitree.py looks grammars up in this module lazily)

    return CircuiTree


class SimpleNetworkTree(_tree_base()):
    """``CircuiTree`` that builds its :class:`SimpleNetworkGrammar` from keyword arguments
    (reference circuitree/models.py:483-521; default ``seed=2023`` as upstream).  ``get_reward`` is
    still the subclass's to supply."""

    def __init__(self, grammar: Optional[SimpleNetworkGrammar] = None, components=None, interactions=None,
                 max_interactions: Optional[int] = None, root: Optional[str] = None, fixed_components=None, **tree_kwargs):
        if grammar is None:
            grammar = SimpleNetworkGrammar(components=components, interactions=interactions, max_interactions=max_interactions,
                                           root=root, fixed_components=fixed_components)
        tree_kwargs.setdefault("seed", 2023)
        super().__init__(grammar=grammar, root=root, **tree_kwargs)


class DimerNetworkTree(_tree_base()):
    """``CircuiTree`` over a :class:`DimersGrammar` built from keyword arguments (reference
    circuitree/models.py:942-980; default ``seed=2023`` as upstream)."""

    def __init__(self, components, regulators, interactions, max_interactions: Optional[int] = None,
                 max_interactions_per_promoter: int = 2, root: Optional[str] = None, **tree_kwargs):
        grammar = DimersGrammar(components=components, regulators=regulators, interactions=interactions,
                                max_interactions=max_interactions, max_interactions_per_promoter=max_interactions_per_promoter,
                                root=root)
        tree_kwargs.setdefault("seed", 2023)
        super().__init__(grammar=grammar, root=root, **tree_kwargs)
