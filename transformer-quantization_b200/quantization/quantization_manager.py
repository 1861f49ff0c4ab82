"""QuantizationManager: one quantizer + one range estimator + the estimate/fix/learn state machine.

Mirror of the reference's quantization/quantization_manager.py (names, constructor signature,
states and error behaviour identical).  ``forward`` enqueues at most four kernels and never
synchronises the host:

    estimate state : min/max reduction -> estimator update -> set_quant_range -> QDQ
    fixed state    : QDQ only (one kernel; CUDA-graph capturable)
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


class QuantizationManager(nn.Module):
    """Quantization + range estimation for one tensor site.

    Parameters (as in the reference, quantization_manager.py:19-40): ``qmethod`` (QMethods member),
    ``init`` (RangeEstimators member), ``per_channel``, ``axis``, ``n_groups``, optional fixed
    ``x_min`` / ``x_max``, ``qparams`` (kwargs of the quantizer, e.g. n_bits) and ``init_params``
    (kwargs of the estimator).
    """

    def __init__(self, qmethod=QMethods.symmetric_uniform, init=RangeEstimators.current_minmax,
                 per_channel=False, axis=None, n_groups=None, x_min=None, x_max=None, qparams=None,
                 init_params=None):
        super().__init__()
        self.state = Qstates.estimate_ranges
        self.qmethod = qmethod
        self.init = init
        self.per_channel = per_channel
        self.axis = axis
        self.n_groups = n_groups
        self.qparams = qparams if qparams else {}
        self.init_params = init_params if init_params else {}
        self.range_estimator = None

        self.quantizer = self.qmethod.cls(per_channel=per_channel, axis=axis, **qparams)

        if x_min is not None and x_max is not None:
            # fixed, user-supplied range: no estimator is created (as in the reference)
            self.set_quant_range(x_min, x_max)
            self.state = Qstates.fix_ranges
        else:
            self.range_estimator = self.init.cls(per_channel=self.per_channel, quantizer=self.quantizer,
                                                 axis=self.axis, n_groups=self.n_groups,
                                                 **self.init_params)

    @property
    def n_bits(self):
        return self.quantizer.n_bits

    # ---- state switches ------------------------------------------------------------------------
    def estimate_ranges(self):
        self.state = Qstates.estimate_ranges

    def fix_ranges(self):
        if not self.quantizer.is_initialized:
            raise QuantizerNotInitializedError()
        self.state = Qstates.fix_ranges

    def learn_ranges(self):
        self.quantizer.make_range_trainable()
        self.state = Qstates.learn_ranges

    def estimate_ranges_train(self):
        self.state = Qstates.estimate_ranges_train

    def reset_ranges(self):
        self.range_estimator.reset()
        self.quantizer.reset()
        self.estimate_ranges()

    def _updates_ranges(self):
        return self.state == Qstates.estimate_ranges or (
            self.state == Qstates.estimate_ranges_train and self.training)

    def forward(self, x):
        if self.range_estimator.per_group_range_estimation:
            # FP32 pass that only records per-dim ranges for the PEG permutation
            self.range_estimator(x)
            return x
        if self._updates_ranges():
            cur_xmin, cur_xmax = self.range_estimator(x)     # per tensor, per axis or per channel
            self.set_quant_range(cur_xmin, cur_xmax)
        return self.quantizer(x)

    def set_quant_range(self, x_min, x_max):
        self.quantizer.set_quant_range(x_min, x_max)

    def extra_repr(self):
        return 'state={}'.format(self.state.name)
