"""Host-side traversal helpers; same names and semantics as pytorch_bayesian/utils/utils.py
(they define WHICH tensors the KL and prune kernels see, and in what order — SURVEY §8 a8)."""
from .traversal import (_item_or_list, _ntuple, _single, _pair, _triple, apply_wb, traverse,
                        variational_tensors)

__all__ = ['_item_or_list', '_single', '_pair', '_triple', 'apply_wb', 'traverse']
