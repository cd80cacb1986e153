"""cvxpygen_b200 -- B200-native batched-solve backend for CVXPYgen's hot path."""
__version__ = '0.1.0'
