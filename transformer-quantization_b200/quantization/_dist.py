"""The single collective call site of the package: cross-rank reduction of calibration statistics.

Inference is replica data-parallel (one process per GPU, no communication).  During calibration
each rank sees its own shard of the calibration batches; reducing the batch statistics BEFORE the
estimator update (EMA / running min-max / loss accumulation) makes N-rank calibration equal to
single-GPU calibration on the concatenated batch (exact for min/max; fp64 sums for MSE losses).
Payloads are 8 B .. ~100 KB, i.e. latency-bound: one packed all-reduce per estimator update over
NCCL (NVLink 5 / NVSwitch) -- or gloo in the CPU tests.

OPT-IN: the reduction only happens inside ``calibration_sync()``.  ``utils.pass_data_for_range_estimation``
(the calibration loop) enters it when ``torch.distributed`` is initialised with more than one rank and
``TQ_DIST_CALIBRATION`` != "0"; nothing else does.  Outside of it -- plain forwards, quantization-aware
training under DDP (``estimate_ranges_train``), a rank-0-only evaluation pass while a quantizer is still
estimating -- estimators are rank-local exactly like the reference's, and no collective can hang on a
rank-asymmetric forward.  Weight estimators never reduce (``local()``: the weights are replicated).
Not reduced even inside the context (documented, rank-local): the cross-entropy estimator's host candidate
loop and the percentile path of the current-min-max estimator.
"""
import contextlib
import os

import torch
import torch.distributed as dist

_group = None
_depth = 0          # > 0 inside calibration_sync()
_local = 0          # > 0 inside local(): reductions suspended (weight quantizers)
_stats = {'calls': 0, 'bytes': 0}


def set_group(group):
    """Use a specific process group for calibration reductions (default: the world group)."""
    global _group
    _group = group


def distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(_group) > 1


@contextlib.contextmanager
def calibration_sync(enable=True):
    """Estimator statistics are all-reduced over the ranks inside this context (if torch.distributed
    is initialised with world size > 1 and TQ_DIST_CALIBRATION != "0").  EVERY rank must run the same
    sequence of estimator updates inside it."""
    global _depth
    on = bool(enable) and os.environ.get('TQ_DIST_CALIBRATION', '1') != '0'
    _depth += int(on)
    try:
        yield
    finally:
        _depth -= int(on)


@contextlib.contextmanager
def local():
    """Suspend the reduction (used around weight quantizers: weights are identical on every rank)."""
    global _local
    _local += 1
    try:
        yield
    finally:
        _local -= 1


def enabled():
    return _depth > 0 and _local == 0 and distributed()


def stats(reset=False):
    """collective calls / payload bytes issued so far (bench.py's calibration leg reports them)"""
    out = dict(_stats)
    if reset:
        _stats['calls'] = _stats['bytes'] = 0
    return out


def _count(t):
    _stats['calls'] += 1
    _stats['bytes'] += t.numel() * t.element_size()


def allreduce_minmax(mn, mx):
    """(min, max) over all ranks.  Packs [-mn, mx] so ONE all-reduce(MAX) serves both."""
    if not enabled():
        return mn, mx
    packed = torch.stack([-mn.reshape(-1), mx.reshape(-1)])
    _count(packed)
    dist.all_reduce(packed, op=dist.ReduceOp.MAX, group=_group)
    return (-packed[0]).reshape(mn.shape), packed[1].reshape(mx.shape)


def allreduce_sum(t):
    if not enabled():
        return t
    _count(t)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_group)
    return t
