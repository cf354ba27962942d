"""pytorch_bayesian.prune: PruneNormal (mirror of pytorch_bayesian/prune/prune.py:5-22)."""
from .prune import PruneNormal, set_in_place

__all__ = ['PruneNormal']
