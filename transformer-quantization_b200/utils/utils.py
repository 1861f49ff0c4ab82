"""Calibration loop and small helpers (the hot-path-adjacent part of the reference's
utils/utils.py: ``pass_data_for_range_estimation`` :47-79, ``StopForwardException``, ``DotDict``,
``seed_all``, ``Stopwatch``)."""
import os
import random
import time

import numpy as np
import torch

from quantization import _dist
from quantization.range_estimators import RangeEstimators


class StopForwardException(Exception):
    """Thrown by hooks to stop a forward pass early."""


def seed_all(seed=1029):
    """seed every generator and force deterministic cuDNN (same defaults as the reference, utils/utils.py:16-24)"""
    random.seed(seed)
    os.environ['PYTHONHASHSEED'] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True


def count_params(module):
    return len(torch.nn.utils.parameters_to_vector(module.parameters()))


def count_embedding_params(model):
    return sum(count_params(m) for m in model.modules() if isinstance(m, torch.nn.Embedding))


def get_layer_by_name(model, layer_name):
    for name, module in model.named_modules():
        if name == layer_name:
            return module
    return None


def pass_data_for_range_estimation(loader, model, act_quant, weight_quant, max_num_batches=20,
                                   cross_entropy_layer=None, inp_idx=0):
    """Calibration: run up to ``max_num_batches`` batches through the model in eval mode with the
    requested quantizers active, so every estimator in ``estimate_ranges`` state sees data.

    Data-parallel calibration: launch one process per GPU, give each rank its own shard of the
    loader (the same number of batches on every rank); with torch.distributed initialised this loop
    -- and only this loop -- runs inside ``_dist.calibration_sync()``: the activation estimators all-reduce
    their statistics, so all ranks end with identical ranges.
    """
    model.set_quant_state(weight_quant, act_quant)
    model.eval()

    if cross_entropy_layer is not None:
        layer_xent = get_layer_by_name(model, cross_entropy_layer)
        if not layer_xent:
            raise ValueError('Cross-entropy layer not found')
        print(f'Set cross entropy estimator for layer "{cross_entropy_layer}"')
        mgr = layer_xent.activation_quantizer
        mgr.range_estimator = RangeEstimators.cross_entropy.cls(
            per_channel=mgr.per_channel, quantizer=mgr.quantizer, **mgr.init_params)

    device = next(model.parameters()).device
    # the ONE place that opts into the cross-rank reduction of estimator statistics: every rank runs this same
    # loop over its own shard (same number of batches), nothing else in the package issues a collective
    with _dist.calibration_sync(enable=_dist.distributed()):
        for i, data in enumerate(loader):
            try:
                if isinstance(data, (tuple, list)):
                    model(data[inp_idx].to(device=device))
                else:
                    model(**{k: v.to(device=device) for k, v in data.items()})
            except StopForwardException:
                pass
            if i >= max_num_batches - 1 or not act_quant:
                break


class DotDict(dict):
    """dict with attribute access; a missing key raises AttributeError like the reference's (utils/utils.py:100-103),
    so ``hasattr(config, 'typo')`` is False instead of silently reading None."""

    def __setattr__(self, key, value):
        self[key] = value

    def __delattr__(self, key):
        del self[key]

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(f"DotDict instance has no key '{key}' ({self.keys()})") from None

    def __getstate__(self):
        return self.__dict__

    def __setstate__(self, d):
        self.__dict__.update(d)


class Stopwatch:
    """Wall-clock timer; ``start`` / ``stop`` / ``reset`` return self so calls chain (``Stopwatch().start()``),
    usable as a context manager (prints on exit when ``verbose``).  Method set of the reference's
    (utils/utils.py:106-179)."""

    def __init__(self, name=None, verbose=False):
        self._name, self._verbose = name, verbose
        self._t0 = None
        self._total = 0.0

    def start(self):
        self._t0 = time.perf_counter()
        return self

    def stop(self):
        self._accumulate()
        self._t0 = None
        return self

    def reset(self):
        self._t0, self._total = None, 0.0
        return self

    def _accumulate(self):
        if self._t0 is not None:
            now = time.perf_counter()
            self._total += now - self._t0
            self._t0 = now

    def format(self):
        self._accumulate()
        prefix = f'[{self._name}]' if self._name is not None else 'Elapsed time'
        return f'{prefix}: {self._total:.3f} sec'

    def print(self):
        print(self.format())

    def get_total_duration(self):
        self._accumulate()
        return self._total

    def __enter__(self):
        return self.start()

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.stop()
        if self._verbose:
            self.print()
