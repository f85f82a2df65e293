"""Grammar interface of the circuit-assembly game.

Host-side mirror of the reference's ``CircuitGrammar`` (reference circuitree/grammar.py:7-189):
same method names, argument meaning and serialisation keys, so user grammars written for the
reference work unchanged with :class:`circuitree_b200.CircuiTree`.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Hashable, Iterable, Iterator

__all__ = ["CircuitGrammar"]


class CircuitGrammar(ABC):
    """Rules of an assembly game: states, actions, terminal test, canonical form.

    Subclasses implement :meth:`get_actions`, :meth:`do_action`, :meth:`is_terminal` and
    :meth:`get_unique_state` (reference grammar.py:34-95).  Actions must never form a cycle.
    """

    def __init__(self, *args, **kwargs):
        # names of attributes that to_dict() leaves out (reference grammar.py:30-32)
        self._non_serializable_attrs = ["_non_serializable_attrs"]

    @abstractmethod
    def get_actions(self, state: Hashable) -> Iterable[Any]:
        """Actions available in ``state``; empty for a terminal state."""

    @abstractmethod
    def do_action(self, state: Hashable, action: Any) -> Hashable:
        """The state reached by taking ``action`` in ``state`` (not necessarily canonical)."""

    @abstractmethod
    def is_terminal(self, state: Hashable) -> bool:
        """Whether the game has ended in ``state``."""

    @abstractmethod
    def get_unique_state(self, state: Hashable) -> Hashable:
        """A canonical ID shared by all states that describe the same topology."""

    def has_pattern(self, state: Hashable, pattern: Hashable) -> bool:
        """Optional motif test (reference grammar.py:97-111)."""
        raise NotImplementedError

    def get_undo_actions(self, state: Hashable) -> Iterable[Any]:
        """(Experimental upstream) actions that can be undone from ``state``."""
        raise NotImplementedError

    def undo_action(self, state: Hashable, action: Any) -> Hashable:
        """(Experimental upstream) undo one action."""
        raise NotImplementedError

    def iter_states(self, root: Hashable, terminal_only: bool = True) -> Iterator[Hashable]:
        """Depth-first iteration over the states reachable from ``root``.

        Yields the same *set* of states as the reference (grammar.py:113-153).  Unlike the
        reference, every state is expanded once: upstream only marks *yielded* states as visited,
        so with ``terminal_only=True`` it re-expands non-terminals along every path and does not
        finish on the 3-component space.
        """
        seen = {root}
        stack = [root]
        while stack:
            state = stack.pop()
            if not terminal_only or self.is_terminal(state):
                yield state
            for action in self.get_actions(state):
                nxt = self.get_unique_state(self.do_action(state, action))
                if nxt not in seen:
                    seen.add(nxt)
                    stack.append(nxt)

    def _get_dict_serializable(self) -> dict:
        skip = set(self._non_serializable_attrs)
        return {k: v for k, v in self.__dict__.items() if k not in skip}

    def to_dict(self) -> dict:
        """JSON-serialisable description; ``__grammar_cls__`` names the class (grammar.py:176-189)."""
        return {"__grammar_cls__": type(self).__name__, **self._get_dict_serializable()}
