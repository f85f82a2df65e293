"""circuitree_b200 -- B200-native rollout engine behind the CircuiTree reward path.

Host-side API mirrors the reference package (``CircuiTree``, ``CircuitGrammar``,
``SimpleNetworkGrammar``, ``DimersGrammar`` ...); the reward path itself is CUDA (sm_100a)
reached through the C ABI in ``include/circuitree_b200.h``.
"""
from . import _capi, _capi_defaults
from .circuitree import CircuiTree, ExhaustionError, ucb_score
from .grammar import CircuitGrammar
from .models import DimerNetworkTree, DimersGrammar, SimpleNetworkGrammar, SimpleNetworkTree
from .motifs import compute_odds_ratio_and_ci, compute_odds_ratios, compute_odds_ratios_with_ci, contingency_test, pattern_table
from .oscillation import OscillationTree, RewardEngine, landscape_motifs, reward_landscape

__version__ = "0.1.0"
__all__ = ["CircuiTree", "CircuitGrammar", "DimersGrammar", "ExhaustionError", "OscillationTree", "RewardEngine", "reward_landscape",
           "SimpleNetworkGrammar", "ucb_score", "compute_odds_ratio_and_ci", "compute_odds_ratios", "compute_odds_ratios_with_ci", "contingency_test",
           "pattern_table", "landscape_motifs", "SimpleNetworkTree", "DimerNetworkTree"]
