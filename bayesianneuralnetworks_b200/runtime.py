"""Process-wide knobs of the variational hot path that the reference does not have (additive,
module-level; no constructor signature changes — SURVEY §8b):

  manual_seed(seed)        key of the in-kernel Philox stream (default 0x5EED)
  set_precision(mode)      'fp32' (three-term TF32 split, 1e-5 parity class; default) or 'tf32'
  set_mc_batching(mode)    'auto' | 'always' | 'never': fold the S Monte-Carlo passes of
                           BayesianNetworkModule.forward (container.py:32-37) into one launch per layer
  set_sample_partition(rank, world)   this process evaluates MC samples s with s % world == rank ... see below
  injected_eps({weight: eps})         test-only: feed recorded torch.randn_like draws to the kernels

Every variational tensor (WeightNormal) owns a Philox stream (tensor_id) and a draw counter; eps
is a pure function of (seed, tensor_id, draw index, element), so the backward pass, `.sampled`
and any partition of the draws over GPUs regenerate identical numbers.
"""
import contextlib
import itertools
import threading

from . import _C

_state = {
    "seed": 0x5EED,
    "precision": _C.PREC_FP32X3,
    "mc_batching": "auto",
}
_tensor_ids = itertools.count(1)
_tls = threading.local()


def manual_seed(seed):
    """Seed of the Philox stream shared by every variational tensor of this process."""
    _state["seed"] = int(seed) & 0xFFFFFFFFFFFFFFFF


def seed():
    return _state["seed"]


def set_precision(mode):
    modes = {"fp32": _C.PREC_FP32X3, "tf32": _C.PREC_TF32}
    if mode not in modes:
        raise ValueError(f"precision must be one of {sorted(modes)}, got {mode!r}")
    _state["precision"] = modes[mode]


def precision():
    return _state["precision"]


def precision_name():
    return "fp32" if _state["precision"] == _C.PREC_FP32X3 else "tf32"


def set_mc_batching(mode):
    if mode not in ("auto", "always", "never"):
        raise ValueError("mc batching mode must be 'auto', 'always' or 'never'")
    _state["mc_batching"] = mode


def mc_batching():
    return _state["mc_batching"]


def next_tensor_id():
    return next(_tensor_ids)


# ---------------------------------------------------------------------------------------------
# Device-side step counter: lets a captured CUDA graph draw fresh eps on every replay.  The draw
# indices a launch uses are host integers baked into the captured kernel arguments; with the device
# counter enabled every kernel ADDS the value of one uint64 in device memory to the Philox `step`
# word (bnn_rng.step_dev), and `advance_rng_step()` — one tiny in-graph add — moves all streams on.
# Call it at the START of each training step so forward, backward and `.sampled` agree.
# ---------------------------------------------------------------------------------------------
_device_step = {"enabled": False, "counters": {}}


def graph_safe_rng(enabled=True):
    _device_step["enabled"] = bool(enabled)


def step_counter(device):
    """The device step counter of `device` (int64[1] tensor) or None when graph-safe RNG is off."""
    if not _device_step["enabled"]:
        return None
    import torch
    key = torch.device(device)
    if key.index is None:
        key = torch.device(key.type, torch.cuda.current_device())
    c = _device_step["counters"].get(key)
    if c is None:
        c = torch.zeros(1, dtype=torch.int64, device=key)
        _device_step["counters"][key] = c
    return c


def advance_rng_step(device=None):
    """Advance every Philox stream on `device` by one step (graph-capturable: a single in-place add)."""
    import torch
    c = step_counter(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    if c is None:
        raise RuntimeError("advance_rng_step() needs graph_safe_rng(True)")
    c.add_(1)


# ---------------------------------------------------------------------------------------------
# Monte-Carlo batch context: set by BayesianNetworkModule.forward while it runs `_forward` ONCE for
# all S samples.  Activations enter with B rows (shared by all samples); the first Bayesian layer
# expands them to S*B rows (sample-major), later layers see S independent row blocks.
# ---------------------------------------------------------------------------------------------
class MCContext:
    def __init__(self, samples, rows, sample_offset=0, total_samples=None):
        self.samples = samples            # samples evaluated by this process in this pass
        self.rows = rows                  # leading dimension of the un-expanded input
        self.expanded = False
        self.sample_offset = sample_offset            # first global sample index of this process
        self.total_samples = total_samples or samples  # draws consumed per pass by every process


def current_mc():
    return getattr(_tls, "mc", None)


def mc_expand_rows(x):
    """For the torch-composite Bayesian layers (Flipout, full covariance) inside a batched Monte-Carlo pass:
    (S, x with S*B sample-major rows, context).  The first Bayesian layer of the network receives the B un-expanded
    rows and repeats them; later layers already see S*B rows.  Outside a batched pass: (1, x, None)."""
    ctx = current_mc()
    if ctx is None:
        return 1, x, None
    if not ctx.expanded:
        if x.shape[0] != ctx.rows:
            raise RuntimeError("batched Monte-Carlo forward: the first Bayesian layer must see the network "
                               f"input rows ({ctx.rows}), got {x.shape[0]}")
        x = x.unsqueeze(0).expand((ctx.samples,) + tuple(x.shape)).reshape((ctx.samples * x.shape[0],) + tuple(x.shape[1:]))
        ctx.expanded = True
    return ctx.samples, x, ctx


@contextlib.contextmanager
def mc_batch(ctx):
    prev = getattr(_tls, "mc", None)
    _tls.mc = ctx
    try:
        yield ctx
    finally:
        _tls.mc = prev


# ---------------------------------------------------------------------------------------------
# sample partition across processes: rank r of R evaluates the contiguous block of
# global sample indices [r*S/R, (r+1)*S/R) of every pass; all ranks advance every draw counter by
# the global S, so the union over ranks reproduces the single-process stream.
# ---------------------------------------------------------------------------------------------
_partition = {"rank": 0, "world": 1}


def set_sample_partition(rank, world):
    if not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    _partition["rank"], _partition["world"] = int(rank), int(world)


def sample_partition():
    return _partition["rank"], _partition["world"]


# ---------------------------------------------------------------------------------------------
# eps injection (parity tests against the reference's recorded torch.randn_like draws)
# ---------------------------------------------------------------------------------------------
def injected_for(weight):
    table = getattr(_tls, "eps", None)
    if table is None:
        return None
    return table.get(id(weight))


@contextlib.contextmanager
def injected_eps(mapping):
    """mapping: {WeightNormal instance: eps tensor [S, *shape]} consumed by the next forward (and its
    backward).  Test-only: the kernels read eps from memory instead of generating it."""
    prev = getattr(_tls, "eps", None)
    _tls.eps = {id(k): v for k, v in mapping.items()}
    try:
        yield
    finally:
        _tls.eps = prev
