"""The single collective call site of the package: cross-rank reduction of calibration statistics.

Inference is replica data-parallel (one process per GPU, no communication).  During calibration
each rank sees its own shard of the calibration batches; reducing the batch statistics BEFORE the
estimator update (EMA / running min-max / loss accumulation) makes N-rank calibration equal to
single-GPU calibration on the concatenated batch (exact for min/max; fp64 sums for MSE losses).
Payloads are 8 B .. ~100 KB, i.e. latency-bound: one packed all-reduce per estimator update over
NCCL (NVLink 5 / NVSwitch) -- or gloo in the CPU tests.

Enabled when ``torch.distributed`` is initialised and ``TQ_DIST_CALIBRATION`` != "0".
"""
import os

import torch
import torch.distributed as dist

_group = None


def set_group(group):
    """Use a specific process group for calibration reductions (default: the world group)."""
    global _group
    _group = group


def enabled():
    return (dist.is_available() and dist.is_initialized() and dist.get_world_size(_group) > 1
            and os.environ.get('TQ_DIST_CALIBRATION', '1') != '0')


def allreduce_minmax(mn, mx):
    """(min, max) over all ranks.  Packs [-mn, mx] so ONE all-reduce(MAX) serves both."""
    if not enabled():
        return mn, mx
    packed = torch.stack([-mn.reshape(-1), mx.reshape(-1)])
    dist.all_reduce(packed, op=dist.ReduceOp.MAX, group=_group)
    return (-packed[0]).reshape(mn.shape), packed[1].reshape(mx.shape)


def allreduce_sum(t):
    if not enabled():
        return t
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_group)
    return t
