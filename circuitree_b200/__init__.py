"""circuitree_b200 -- B200-native rollout engine behind the CircuiTree reward path."""
__version__ = "0.1.0"
