"""Traversal of a model's Bayesian modules (mirror of pytorch_bayesian/utils/utils.py:10-67).

Semantics kept on purpose: only Sequential / ModuleList /
ModuleDict / BayesianNetworkModule containers are descended into, so Bayesian layers nested in any
other nn.Module are invisible to KLDivergence and PruneNormal; results come back in registration
order, weight before bias; "nothing found" is None, not an empty list.
"""
from collections.abc import Iterable
from itertools import repeat

from torch.nn.modules.container import ModuleDict, ModuleList, Sequential


def _item_or_list(n):
    """utils.py:10-11 — a one-element list collapses to its element."""
    return n[0] if len(n) == 1 else n


def _ntuple(n):
    """utils.py:14-19 — iterables pass through unchanged (no length check), scalars repeat n times."""
    def parse(x):
        return x if isinstance(x, Iterable) else tuple(repeat(x, n))
    return parse


_single = _ntuple(1)
_pair = _ntuple(2)
_triple = _ntuple(3)


def apply_wb(module, fn, *args, pass_module=False, pass_type=False, **kwargs):
    """utils.py:30-47 — fn(param, ...) for module.weight then module.bias (skipping None);
    non-None results are collected; returns the list, or None when it is empty."""
    if pass_module:
        kwargs['module'] = module
    collected = []
    for kind, param in (('w', module.weight), ('b', module.bias)):
        if param is None:
            continue
        if pass_type:
            kwargs['type'] = kind
        out = fn(param, *args, **kwargs)
        if out is not None:
            collected.append(out)
    return collected or None


def _containers():
    from ..nn.container import BayesianNetworkModule
    return (ModuleList, ModuleDict, Sequential, BayesianNetworkModule)


def traverse(module, fn, *args, **kwargs):
    """utils.py:50-67 — depth-first over container children; fn is applied to BayesianModules; list
    results are concatenated; anything else (including a non-list fn result) yields None."""
    from ..nn.container import BayesianModule
    if isinstance(module, _containers()):
        found = []
        for child in module.children():
            sub = traverse(child, fn, *args, **kwargs)
            if sub is not None:
                found += sub
        return found or None
    if isinstance(module, BayesianModule):
        out = fn(module, *args, **kwargs)
        if isinstance(out, list) and out:
            return out
    return None


def variational_tensors(model):
    """[(param, module, 'w'|'b')] in the order KLDivergence / PruneNormal visit them."""
    return model.traverse(lambda m: apply_wb(m, lambda p, module, type: (p, module, type),
                                             pass_module=True, pass_type=True))
