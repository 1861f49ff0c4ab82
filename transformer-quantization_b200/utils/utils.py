"""Calibration loop and small helpers (the hot-path-adjacent part of the reference's
utils/utils.py: ``pass_data_for_range_estimation`` :47-79, ``StopForwardException``, ``DotDict``,
``seed_all``, ``Stopwatch``)."""
import random
import time

import numpy as np
import torch

from quantization.range_estimators import RangeEstimators


class StopForwardException(Exception):
    """Thrown by hooks to stop a forward pass early."""


def seed_all(seed=1000, deterministic=False):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    if deterministic:
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False


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
    loader; with torch.distributed initialised the estimators all-reduce their statistics
    (quantization/_dist.py), so all ranks end with identical ranges.
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
    """dict with attribute access; missing keys read as None."""

    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__

    def __getattr__(self, key):
        return self.get(key)

    def __getstate__(self):
        return self.__dict__

    def __setstate__(self, d):
        self.__dict__.update(d)


class Stopwatch:
    """Wall-clock timer usable as a context manager."""

    def __init__(self, name='', verbose=True):
        self.name, self.verbose = name, verbose
        self._t0 = None
        self.total = 0.0

    def start(self):
        self._t0 = time.perf_counter()

    def stop(self):
        if self._t0 is not None:
            self.total += time.perf_counter() - self._t0
            self._t0 = None

    def reset(self):
        self._t0, self.total = None, 0.0

    def format(self):
        return f'{self.name}: {self.total:.3f}s'

    def __enter__(self):
        self.start()
        return self

    def __exit__(self, *exc):
        self.stop()
        if self.verbose:
            print(self.format())
