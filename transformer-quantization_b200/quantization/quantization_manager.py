"""QuantizationManager: one quantizer + one range estimator + the estimate / fix / learn state
machine of a quantizer site.

Host-side mirror of the reference's quantization/quantization_manager.py -- constructor signature,
attribute names, states and error behaviour are the reference's (models/quantized_*.py and main.py
poke at all of them); the data path is this library's:

    estimate state : min/max reduction -> estimator update -> set_quant_range -> QDQ   (4 kernels)
    fixed state    : QDQ only                                                           (1 kernel)

None of it synchronises the host, so a fixed-state forward is CUDA-graph capturable.
"""
from enum import Enum

from torch import nn

from quantization.quantizers import QMethods, QuantizerNotInitializedError
from quantization.range_estimators import RangeEstimators


class Qstates(Enum):
    estimate_ranges = 0        # ranges are updated in eval and train mode
    fix_ranges = 1             # quantization ranges are fixed for train and eval
    learn_ranges = 2           # quantization params are nn.Parameters
    estimate_ranges_train = 3  # ranges are updated during train and fixed for eval


def _state_switch(target, precondition=None):
    """method that moves the manager to ``target`` (after an optional hook)"""

    def switch(self):
        if precondition is not None:
            precondition(self)
        self.state = target

    switch.__name__ = target.name
    return switch


def _require_initialized(mgr):
    if not mgr.quantizer.is_initialized:
        raise QuantizerNotInitializedError()


class QuantizationManager(nn.Module):
    """Quantization and range estimation of one tensor site.

    qmethod     QMethods member: which quantizer class
    init        RangeEstimators member: how the range is found
    per_channel one grid per output channel (weights)
    axis, n_groups  per-embedding / per-embedding-group activation quantization
    x_min, x_max    optional fixed range: the manager starts in ``fix_ranges`` and owns no estimator
    qparams     kwargs of the quantizer (n_bits, scale_domain, ...)
    init_params kwargs of the estimator (momentum, num_candidates, opt_method, ...)
    """

    def __init__(self, qmethod=QMethods.symmetric_uniform, init=RangeEstimators.current_minmax,
                 per_channel=False, axis=None, n_groups=None, x_min=None, x_max=None, qparams=None,
                 init_params=None):
        super().__init__()
        self.state = Qstates.estimate_ranges
        self.qmethod, self.init = qmethod, init
        self.per_channel, self.axis, self.n_groups = per_channel, axis, n_groups
        self.qparams = qparams or {}
        self.init_params = init_params or {}
        self.range_estimator = None
        self.quantizer = qmethod.cls(per_channel=per_channel, axis=axis, **qparams)

        fixed_range = x_min is not None and x_max is not None
        if fixed_range:
            self.set_quant_range(x_min, x_max)
            self.state = Qstates.fix_ranges
        else:
            self.range_estimator = init.cls(per_channel=per_channel, quantizer=self.quantizer, axis=axis,
                                            n_groups=n_groups, **self.init_params)

    n_bits = property(lambda self: self.quantizer.n_bits)

    estimate_ranges = _state_switch(Qstates.estimate_ranges)
    estimate_ranges_train = _state_switch(Qstates.estimate_ranges_train)
    fix_ranges = _state_switch(Qstates.fix_ranges, _require_initialized)
    learn_ranges = _state_switch(Qstates.learn_ranges, lambda self: self.quantizer.make_range_trainable())

    def reset_ranges(self):
        self.range_estimator.reset()
        self.quantizer.reset()
        self.estimate_ranges()

    def _updates_ranges(self):
        if self.state is Qstates.estimate_ranges:
            return True
        return self.state is Qstates.estimate_ranges_train and self.training

    def forward(self, x):
        est = self.range_estimator
        if est.per_group_range_estimation:
            est(x)                      # FP32 pass: only record per-dim ranges for the PEG permutation
            return x
        if self._updates_ranges():
            self.set_quant_range(*est(x))
        return self.quantizer(x)

    def set_quant_range(self, x_min, x_max):
        self.quantizer.set_quant_range(x_min, x_max)

    def extra_repr(self):
        return f'state={self.state.name}'
