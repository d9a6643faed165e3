"""desirna_b200 -- B200-native batched RNA folding engine behind DesiRNA's scoring entry points."""
__all__ = ["engine"]
