"""Mirror of torchreid.metrics (reference torchreid/metrics/__init__.py:3-5) for the hot path."""
from .rank import evaluate_rank
from .distance import compute_distance_matrix

__all__ = ['evaluate_rank', 'compute_distance_matrix']
