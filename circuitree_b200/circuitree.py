"""MCTS over circuit topologies: host-side mirror of the reference's ``CircuiTree``.

Same public names, argument meaning, RNG consumption and graph bookkeeping as the reference
(reference circuitree/circuitree.py:64-1122, :1432-1469), so that with identical rewards the
search graph after N iterations is identical node for node (pinned by tests/golden).  On top of
that it adds the batched scheduler the GPU reward path needs (``get_rewards``,
``search_mcts_batched``): K leaves are selected under virtual loss, simulated in one launch,
then back-propagated -- K = 1 is exactly ``search_mcts``.

Sampling of terminal / successful states and the pattern-significance tests (reference
circuitree.py:1200-1648) live in ``motifs.py`` (``MotifMixin``).  Out of scope here (SURVEY.md
section 8): complexity graphs, plotting.
"""
from __future__ import annotations

import json
from abc import ABC, abstractmethod
from itertools import chain, cycle, islice, repeat
from pathlib import Path
from typing import Any, Callable, Hashable, Iterable, Iterator, Literal, Optional, Sequence

import networkx as nx
import numpy as np

from .grammar import CircuitGrammar
from .motifs import MotifMixin

__all__ = ["ucb_score", "ExhaustionError", "CircuiTree"]


def ucb_score(graph: nx.DiGraph, parent: Hashable, node: Hashable,
              exploration_constant: Optional[float] = np.sqrt(2.0), **kw) -> float:
    """UCB1 of the edge parent->node: ``Q/n + c*sqrt(ln N_parent / n)``; infinite for an
    unvisited edge (reference circuitree.py:21-57).  Evaluated with the same numpy scalar
    functions as the reference so scores agree to the last bit."""
    edge = graph.edges[parent, node]
    n = edge["visits"]
    if n == 0:
        return np.inf
    return edge["reward"] / n + exploration_constant * np.sqrt(np.log(graph.nodes[parent]["visits"]) / n)


class ExhaustionError(RuntimeError):
    """Every node of the tree has been visited to exhaustion (reference circuitree.py:60)."""


class CircuiTree(MotifMixin, ABC):
    """Monte Carlo tree search over a :class:`CircuitGrammar`; subclass and implement
    :meth:`get_reward` (reference circuitree.py:64-152)."""

    def __init__(
        self,
        grammar: CircuitGrammar,
        root: str,
        exploration_constant: Optional[float] = None,
        seed: Optional[int] = None,
        n_exhausted: Optional[int | float] = None,
        graph: Optional[nx.DiGraph] = None,
        tree_shape: Optional[Literal["tree", "dag"]] = None,
        compute_unique: bool = True,
        **kwargs,
    ):
        self.rg = np.random.default_rng(seed)
        self.seed: int = self.rg.bit_generator._seed_seq.entropy
        self.root = root
        if graph is None:
            self.graph = nx.DiGraph()
            self.graph.add_node(root, visits=0, reward=0)
        else:
            if root not in graph:
                raise ValueError(f"Supplied graph does not contain the root node: {root}")
            self.graph = graph
        self.graph.root = root
        self.compute_unique = compute_unique
        if tree_shape is not None:
            if tree_shape not in ("tree", "dag"):
                raise ValueError("Argument `tree_shape` must be `tree` or `dag`.")
            import warnings

            warnings.warn("The `tree_shape` argument is deprecated and will be removed in a future version. "
                          "Please use `compute_symmetries` instead.")
        self.grammar = grammar
        self.exploration_constant = np.sqrt(2) if exploration_constant is None else exploration_constant
        self.n_exhausted = n_exhausted
        self._non_serializable_attrs = ["_non_serializable_attrs", "rg", "graph", "_child_memo"]
        self._child_memo: dict = {}        # (state, action) -> child state: the tree revisits the same edges all the time

    # ---------------------------------------------------------------------------------------------
    # reward plug-in boundary
    # ---------------------------------------------------------------------------------------------
    @abstractmethod
    def get_reward(self, state: Hashable, **kwargs) -> float | int:
        """Reward of a terminal state (reference circuitree.py:135-152)."""

    def get_rewards(self, states: Sequence[Hashable], **kwargs) -> list[float]:
        """Batched form of :meth:`get_reward` used by :meth:`search_mcts_batched` and
        :meth:`search_bfs`; the default loops, the GPU subclass overrides it with one launch."""
        return [self.get_reward(s, **kwargs) for s in states]

    @property
    def default_attrs(self) -> dict:
        return dict(visits=0, reward=0)

    @property
    def terminal_states(self) -> Iterator[Hashable]:
        return (n for n in self.graph.nodes if self.grammar.is_terminal(n))

    # ---------------------------------------------------------------------------------------------
    # state transitions
    # ---------------------------------------------------------------------------------------------
    def _do_action(self, state: Hashable, action: Hashable) -> Hashable:
        nxt = self.grammar.do_action(state, action)
        return self.grammar.get_unique_state(nxt) if self.compute_unique else nxt

    _CHILD_MEMO_MAX = 1 << 20

    def _child(self, state: Hashable, action: Hashable) -> Hashable:
        """``_do_action`` memoised per (state, action): a pure function of the grammar, evaluated
        for every child of every node on every selection path."""
        key = (state, action)
        try:
            return self._child_memo[key]
        except KeyError:
            pass
        except TypeError:          # unhashable state or action
            return self._do_action(state, action)
        if len(self._child_memo) >= self._CHILD_MEMO_MAX:
            self._child_memo.clear()
        child = self._child_memo[key] = self._do_action(state, action)
        return child

    def _undo_action(self, state: Hashable, action: Hashable) -> Hashable:
        if state == self.root:
            return None
        prev = self.grammar.undo_action(state, action)
        return self.grammar.get_unique_state(prev) if self.compute_unique else prev

    @staticmethod
    def _get_random_terminal_descendant(grammar: CircuitGrammar, start: Hashable, rg: np.random.Generator) -> Hashable:
        """Uniform random walk to a terminal state, canonicalising every step; one
        ``rg.choice`` per step like the reference (circuitree.py:218-246)."""
        state = start
        while not grammar.is_terminal(state):
            actions = grammar.get_actions(state)
            state = grammar.get_unique_state(grammar.do_action(state, rg.choice(actions)))
        return state

    def get_random_terminal_descendant(self, start: Hashable, rg: Optional[np.random.Generator] = None) -> Hashable:
        return self._get_random_terminal_descendant(self.grammar, start, self.rg if rg is None else rg)

    # ---------------------------------------------------------------------------------------------
    # one MCTS iteration
    # ---------------------------------------------------------------------------------------------
    def _is_exhausted(self, node: Hashable) -> bool:
        return node in self.graph.nodes and self.graph.nodes[node].get("is_exhausted", False)

    def select_and_expand(self, rg: Optional[np.random.Generator] = None) -> list[Hashable]:
        """UCB descent from the root; the first unexpanded edge met (after the per-level
        shuffle) is expanded and ends the path (reference circuitree.py:315-384).

        Deviation: ``n_exhausted=None`` means "never exhaust"; the reference checkout raises
        ``TypeError`` on reaching a terminal node in that case."""
        rg = self.rg if rg is None else rg
        track_exhaustion = self.n_exhausted is not None
        # fast path: the stock UCB read straight off networkx's adjacency dicts (same arithmetic, same
        # numpy scalar functions as ucb_score; skipped when a subclass overrides the scoring or _do_action)
        stock = type(self).get_ucb_score is CircuiTree.get_ucb_score and type(self)._do_action is CircuiTree._do_action
        adj, nodes, c_explore = self.graph._adj, self.graph._node, self.exploration_constant
        child_of = self._child if stock else self._do_action
        node = self.root
        path = [node]
        actions = self.grammar.get_actions(node)
        while actions:
            rg.shuffle(actions)
            best, best_score = None, -np.inf
            out_edges = adj.get(node, {}) if stock else None
            log_n = None
            for action in actions:
                child = child_of(node, action)
                if track_exhaustion and self._is_exhausted(child):
                    continue
                if stock:
                    edge = out_edges.get(child)
                    if edge is None or edge["visits"] == 0:
                        score = np.inf
                    else:
                        if log_n is None:
                            log_n = np.log(nodes[node]["visits"])
                        n = edge["visits"]
                        score = edge["reward"] / n + c_explore * np.sqrt(log_n / n)
                else:
                    score = self.get_ucb_score(node, child)
                if score == np.inf:
                    self.expand_edge(node, child)
                    path.append(child)
                    return path
                if score > best_score:
                    best, best_score = child, score
            node = best
            path.append(node)
            actions = self.grammar.get_actions(node)
        if track_exhaustion and self.graph.nodes[node].get("visits", 0) >= self.n_exhausted - 1:
            self.mark_as_exhausted(node)
        return path

    def mark_as_exhausted(self, node: Hashable) -> None:
        """Flag ``node``, make its parents forget what they learnt through it, and recurse into
        parents whose children are all flagged (reference circuitree.py:386-420)."""
        self.graph.nodes[node]["is_exhausted"] = True
        if node == self.root:
            raise ExhaustionError("Every node in the tree has been visited to exhaustion.")
        for parent in self.graph.predecessors(node):
            edge = self.graph.edges[parent, node]
            self.graph.nodes[parent]["visits"] -= edge["visits"]
            self.graph.nodes[parent]["reward"] -= edge["reward"]
        for parent in self.graph.predecessors(node):
            if all(self.graph.nodes[c].get("is_exhausted", False) for c in self.graph.successors(parent)):
                self.mark_as_exhausted(parent)

    def expand_edge(self, parent: Hashable, child: Hashable) -> None:
        if not self.graph.has_node(child):
            self.graph.add_node(child, **self.default_attrs)
        self.graph.add_edge(parent, child, **self.default_attrs)

    def get_ucb_score(self, parent: Hashable, child: Hashable) -> float:
        if self.graph.has_edge(parent, child):
            return ucb_score(self.graph, parent, child, self.exploration_constant)
        return np.inf

    def _backpropagate(self, path: list, attr: str, value: float | int) -> None:
        """``attr += value`` on every node and edge of ``path`` (reference circuitree.py:451-466)."""
        nodes, edges = self.graph.nodes, self.graph.edges
        child = path[-1]
        nodes[child][attr] += value
        for parent in reversed(path[:-1]):
            edges[parent, child][attr] += value
            nodes[parent][attr] += value
            child = parent

    def backpropagate_visit(self, selection_path: list) -> None:
        self._backpropagate(selection_path, "visits", 1)

    def backpropagate_reward(self, selection_path: list, reward: float | int) -> None:
        self._backpropagate(selection_path, "reward", reward)

    def traverse(self, **kwargs) -> tuple[list[Hashable], float | int, Hashable]:
        """Select/expand, roll out, visit++ (virtual loss), reward, reward += -- in the
        reference's order (circuitree.py:501-537)."""
        path = self.select_and_expand()
        sim_node = self.get_random_terminal_descendant(path[-1])
        self.backpropagate_visit(path)
        reward = self.get_reward(sim_node, **kwargs)
        self.backpropagate_reward(path, reward)
        return path, reward, sim_node

    # ---------------------------------------------------------------------------------------------
    # exhaustive expansion
    # ---------------------------------------------------------------------------------------------
    def grow_tree(self, root: Hashable = None, n_visits: int = 0, print_updates: bool = False,
                  print_every: int = 1000) -> None:
        """Add every reachable state and edge to the graph (reference circuitree.py:539-586)."""
        if root is None:
            root = self.root
            if print_updates:
                print(f"Adding root: {root}")
            self.graph.add_node(root, visits=n_visits, reward=0)
        graph, grammar = self.graph, self.grammar
        pending = [(root, a) for a in grammar.get_actions(root)]
        n_added = 1
        while pending:
            node, action = pending.pop()
            if grammar.is_terminal(node):
                continue
            child = self._do_action(node, action)
            if child not in graph.nodes:
                n_added += 1
                graph.add_node(child, visits=n_visits, reward=0)
                pending.extend((child, a) for a in grammar.get_actions(child))
                if print_updates and n_added % print_every == 0:
                    print(f"Graph size: {n_added} nodes.")
            if not graph.has_edge(node, child):
                graph.add_edge(node, child, visits=n_visits, reward=0)

    def enumerate_terminal_states(self, root: Optional[Hashable] = None, progress: bool = False,
                                  max_iter: int = None) -> set[Hashable]:
        """All terminal states reachable from ``root`` (reference circuitree.py:1432-1469)."""
        root = self.root if root is None else root
        limit = np.inf if max_iter is None else max_iter
        seen, terminals, stack, k = set(), set(), [root], 0
        while stack and k < limit:
            state = stack.pop()
            seen.add(state)
            if self.grammar.is_terminal(state):
                terminals.add(state)
            else:
                stack.extend({self._do_action(state, a) for a in self.grammar.get_actions(state)} - seen)
            k += 1
        print(f"Found {len(terminals)} terminal states.")
        return terminals

    def bfs_iterator(self, root=None, shuffle=False) -> Iterator[Hashable]:
        """Terminal nodes in breadth-first layer order (reference circuitree.py:588-599)."""
        root = self.root if root is None else root
        layers: Iterable = nx.bfs_layers(self.graph, root)
        if shuffle:
            layers = list(layers)
            for layer in layers:
                self.rg.shuffle(layer)
        return filter(self.grammar.is_terminal, chain.from_iterable(layers))

    # ---------------------------------------------------------------------------------------------
    # search drivers
    # ---------------------------------------------------------------------------------------------
    def search_bfs(self, n_steps: Optional[int] = None, n_repeats: Optional[int] = None,
                   n_cycles: Optional[int] = None, callback: Optional[Callable] = None, callback_every: int = 1,
                   shuffle: bool = False, progress: bool = False, run_kwargs: Optional[dict] = None,
                   batch_size: int = 1) -> None:
        """Visit every terminal in BFS order, ``visits += 1; reward += get_reward`` on the node
        only (reference circuitree.py:601-691).  ``batch_size > 1`` groups consecutive visits into
        one :meth:`get_rewards` call; the resulting node statistics are the same."""
        if self.graph.number_of_nodes() < 2:
            self.graph.add_node(self.root, **self.default_attrs)
            self.grow_tree(root=self.root, n_visits=0)
        run_kwargs = {} if run_kwargs is None else run_kwargs
        it: Iterable = self.bfs_iterator(root=self.root, shuffle=shuffle)
        if not ((n_steps is None) ^ (n_cycles is None)):
            raise ValueError("Must specify exactly one of n_steps or n_cycles.")
        if n_repeats is not None:
            it = chain.from_iterable(repeat(elem, n_repeats) for elem in it)
        if n_cycles is not None:
            # the reference repeats one exhausted iterator object (circuitree.py:672-673), so only
            # the first cycle yields anything; materialise to keep that observable behaviour
            it = chain.from_iterable(repeat(iter(it), n_cycles))
        elif n_steps is not None:
            it = islice(cycle(it), n_steps)
        if progress:
            from tqdm import tqdm

            it = tqdm(it, desc="BFS search")
        if callback is not None:
            callback(self, None, None)
        nodes = self.graph.nodes
        i = 0
        it = iter(it)
        while True:
            batch = list(islice(it, max(1, batch_size)))
            if not batch:
                break
            for node in batch:
                nodes[node]["visits"] += 1
            rewards = [self.get_reward(batch[0], **run_kwargs)] if batch_size <= 1 else self.get_rewards(batch, **run_kwargs)
            for node, reward in zip(batch, rewards):
                nodes[node]["reward"] += reward
                if callback is not None and i % callback_every == 0:
                    callback(self.graph, node, reward)
                i += 1

    def search_mcts(self, n_steps: int, callback: Optional[Callable] = None, callback_every: int = 1,
                    progress_bar: bool = False, run_kwargs: Optional[dict] = None,
                    callback_before_start: bool = True) -> None:
        """``n_steps`` sequential MCTS iterations (reference circuitree.py:693-784)."""
        if progress_bar:
            from tqdm import trange

            iterator: Iterable[int] = trange(n_steps, desc="MCTS search")
        else:
            iterator = range(n_steps)
        if callback is not None and callback_before_start:
            callback(self, -1, [None], None, None)
        run_kwargs = {} if run_kwargs is None else run_kwargs
        print(f"Starting MCTS search with {n_steps} iterations.")
        if callback is None:
            self._run_mcts(self, iterator, **run_kwargs)
        else:
            self._run_mcts_with_callback(self, iterator, callback, callback_every, **run_kwargs)

    @staticmethod
    def _run_mcts(tree: "CircuiTree", iterator: Iterable[int], **kwargs) -> None:
        for _ in iterator:
            tree.traverse(**kwargs)

    @staticmethod
    def _run_mcts_with_callback(tree: "CircuiTree", iterator: Iterable[int], callback: Optional[Callable],
                                callback_every: int, **kwargs) -> None:
        for i in iterator:
            path, reward, sim_node = tree.traverse(**kwargs)
            if callback is not None and i % callback_every == 0:
                callback(tree, i, path, sim_node, reward)

    def submit_rewards(self, states: Sequence[Hashable], **kwargs):
        """Start computing the rewards of a batch of states and return a ticket for
        :meth:`collect_rewards`.  The default computes them on the spot; subclasses with an
        asynchronous backend (``OscillationTree``) return at once."""
        return {"rewards": self.get_rewards(states, **kwargs)}

    def collect_rewards(self, ticket) -> list[float]:
        return ticket["rewards"]

    def search_mcts_batched(self, n_steps: int, batch_size: int, callback: Optional[Callable] = None,
                            callback_every: int = 1, callback_before_start: bool = True,
                            run_kwargs: Optional[dict] = None, pipeline_depth: int = 1) -> None:
        """MCTS with ``batch_size`` leaves in flight: each leaf is selected, rolled out and
        visit-back-propagated (virtual loss, the semantics of reference circuitree.py:529-535 under
        ``search_mcts_parallel``), then all rewards come from ONE :meth:`get_rewards` call and are
        back-propagated in selection order.  ``batch_size=1`` reproduces :meth:`search_mcts`.

        ``pipeline_depth > 1`` keeps that many reward batches in flight (:meth:`submit_rewards` /
        :meth:`collect_rewards`): the next batch is selected, with the visits of every in-flight leaf
        already counted, while earlier batches are still being simulated.

        Finite ``n_exhausted`` with leaves in flight follows what the reference's greenlets do when
        they all block in ``get_reward`` (circuitree.py:377-420 under :786-893): flagging a node
        un-counts the edge totals of that moment -- virtual visits of in-flight leaves included --
        from its parents, and a reward that arrives later is still added along its whole path.  When
        the tree is exhausted (``ExhaustionError``) the complete batches in flight are finished first;
        the leaves of the batch being selected keep their visits and never get rewards.  Pinned by
        tests/golden/batched_vectors.json.gz (reference methods driven in this schedule)."""
        if batch_size < 1:
            raise ValueError("batch_size must be at least 1.")
        if pipeline_depth < 1:
            raise ValueError("pipeline_depth must be at least 1.")
        if callback is not None and callback_before_start:
            callback(self, -1, [None], None, None)
        run_kwargs = {} if run_kwargs is None else run_kwargs
        inflight: list = []
        step = 0          # leaves back-propagated
        selected = 0      # leaves selected

        def retire() -> None:
            nonlocal step
            paths, sims, ticket = inflight.pop(0)
            rewards = ticket if isinstance(ticket, list) else self.collect_rewards(ticket)
            for path, sim, reward in zip(paths, sims, rewards):
                self.backpropagate_reward(path, reward)
                if callback is not None and step % callback_every == 0:
                    callback(self, step, path, sim, reward)
                step += 1

        try:
            while selected < n_steps:
                k = min(batch_size, n_steps - selected)
                paths, sims = [], []
                for _ in range(k):
                    path = self.select_and_expand()
                    sims.append(self.get_random_terminal_descendant(path[-1]))
                    self.backpropagate_visit(path)
                    paths.append(path)
                selected += k
                if pipeline_depth == 1:
                    inflight.append((paths, sims, list(self.get_rewards(sims, **run_kwargs))))
                else:
                    inflight.append((paths, sims, self.submit_rewards(sims, **run_kwargs)))
                while len(inflight) >= pipeline_depth:
                    retire()
        finally:
            while inflight:
                retire()

    def search_mcts_parallel(self, n_steps: int, n_threads: int, callback: Optional[Callable] = None,
                             callback_every: int = 1, callback_before_start: bool = True,
                             run_kwargs: Optional[dict] = None, logger: Optional[Any] = None) -> None:
        """Parallel MCTS.  The reference spawns ``n_threads`` gevent greenlets on one graph and
        relies on ``get_reward`` blocking on I/O (circuitree.py:786-893); here the same virtual-loss
        semantics are realised as ``n_threads`` leaves per batched reward launch."""
        if n_threads < 1:
            raise ValueError("Number of threads must be at least 1.")
        if n_threads == 1:
            print("Detected n_threads==1. Running MCTS with a single thread.")
            self.search_mcts(n_steps=n_steps, callback_every=callback_every, callback=callback, run_kwargs=run_kwargs)
            return
        print(f"Starting MCTS search with {n_steps} iters, {n_threads} leaves per batch.")
        if logger is not None:
            run_kwargs = dict(run_kwargs or {}, logger=logger)
        self.search_mcts_batched(n_steps, n_threads, callback=callback, callback_every=callback_every,
                                 callback_before_start=callback_before_start, run_kwargs=run_kwargs)

    def is_success(self, terminal_state: Hashable) -> bool:
        raise NotImplementedError

    # ---------------------------------------------------------------------------------------------
    # (de)serialisation: GML graph + JSON attributes, reference-compatible
    # ---------------------------------------------------------------------------------------------
    def copy_graph(self) -> nx.DiGraph:
        return self.graph.copy()

    def get_rng_state(self) -> dict:
        """JSON-serialisable state of ``self.rg`` (the reference does not save it -- ``rg`` is excluded
        at circuitree.py:129-133 -- so its resumed searches are not reproducible)."""
        st = self.rg.bit_generator.state
        return json.loads(json.dumps(st, default=_json_default))

    def set_rng_state(self, state: dict) -> None:
        self.rg.bit_generator.state = state

    def get_attributes(self, attrs_copy: Optional[Iterable[str]]) -> dict:
        """Serialisable attributes; objects with ``to_dict`` are converted (circuitree.py:922-966)."""
        blocked = set(self._non_serializable_attrs)
        if attrs_copy is None:
            keys = set(self.__dict__) - blocked
        else:
            keys = set(attrs_copy) & set(self.__dict__)
            bad = keys & blocked
            if bad:
                raise ValueError(f"Attempting to save non-serializable attributes: {', '.join(bad)}.")
        return {k: (v.to_dict() if hasattr(v, "to_dict") else v) for k, v in self.__dict__.items() if k in keys}

    def generate_gml(self) -> str:
        return nx.generate_gml(self.graph)

    def to_string(self) -> tuple[str, str]:
        return "\n".join(self.generate_gml()), json.dumps(self.get_attributes(None), indent=4, default=_json_default)

    def to_file(self, gml_file: str | Path, json_file: Optional[str | Path] = None,
                save_attrs: Optional[Iterable[str]] = None, compress: bool = False, **kwargs):
        """Write the graph as GML (optionally gzipped) and the attributes as JSON; loadable by the
        reference's ``from_file`` (circuitree.py:983-1033)."""
        gml_target = Path(gml_file).with_suffix(".gml.gz" if compress else ".gml")
        nx.write_gml(self.graph, gml_target, **kwargs)
        if json_file is None:
            return gml_target
        json_target = Path(json_file).with_suffix(".json")
        attrs = self.get_attributes(save_attrs)
        if save_attrs is None:
            attrs["rng_state"] = self.get_rng_state()      # extra key; the reference's from_file passes it on as a kwarg
        with json_target.open("w") as f:
            json.dump(attrs, f, indent=4, default=_json_default)
        return gml_target, json_target

    @staticmethod
    def _generate_grammar(grammar_cls, grammar_kwargs: dict) -> CircuitGrammar:
        name = grammar_kwargs.pop("__grammar_cls__", None)
        if grammar_cls is None:
            if name is None:
                raise ValueError("Must specify grammar class as a keyword argument (from_file(grammar_cls=...)) or "
                                 "in the attributes json file (grammar = {'__grammar_cls__': ..., **grammar_kws}).")
            from . import models

            grammar_cls = getattr(models, name, None)
            if grammar_cls is None:
                raise ValueError(f"Grammar class {name} not found. Pass it as grammar_cls=.")
        grammar_kwargs.pop("_non_serializable_attrs", None)
        known = {k: v for k, v in grammar_kwargs.items() if not k.startswith("_")}
        return grammar_cls(**known)

    @classmethod
    def from_file(cls, graph_gml: str | Path | None, attrs_json: str | Path, grammar_cls=None,
                  grammar_kwargs: Optional[dict] = None, **kwargs):
        """Load a tree written by :meth:`to_file` or by the reference (circuitree.py:1068-1122).
        Unlike the reference, built-in grammar classes are found by name without being passed."""
        if _is_file(attrs_json):
            with open(attrs_json, "r") as f:
                kwargs.update(json.load(f))
        else:
            kwargs.update(json.loads(attrs_json))
        if graph_gml is None:
            graph = None
        elif _is_file(graph_gml):
            graph = nx.read_gml(graph_gml)
        else:
            graph = nx.parse_gml(graph_gml)
        grammar = cls._generate_grammar(grammar_cls, kwargs.pop("grammar", {}) | (grammar_kwargs or {}))
        rng_state = kwargs.pop("rng_state", None)
        tree = cls(grammar=grammar, graph=graph, **kwargs)
        if rng_state is not None:
            tree.set_rng_state(rng_state)       # resume is bit-reproducible
        return tree


def _is_file(p) -> bool:
    try:
        return Path(p).is_file()
    except (OSError, ValueError):
        return False


def _json_default(o):
    if isinstance(o, np.integer):
        return int(o)
    if isinstance(o, np.floating):
        return float(o)
    raise TypeError(f"Object of type {type(o).__name__} is not JSON serializable")
