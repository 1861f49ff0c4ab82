"""QuantizedModule / QuantizedActivation / FP32Acts: per-module quantization configuration and
switches.

Host-side mirror of the reference's quantization/base_quantized_classes.py: the constructor takes
the reference's keywords (:41-45), the flags are the reference's (``_quant_w``, ``_quant_a``,
``cached_params``, ``caching``) and the quantized-parameter cache is dropped at the same points
(``train(True)``, ``_apply`` i.e. .to()/.cuda(), the weight switches).  Pure control logic.
"""
from torch import nn

from quantization.quantization_manager import QuantizationManager
from quantization.quantizers import QMethods
from quantization.range_estimators import RangeEstimators


def _switch_initialized(method_name):
    """module.apply() callback calling ``method_name`` on every INITIALISED QuantizationManager
    (un-initialised ones are skipped like in the reference, :11-32)."""

    def visit(layer):
        if isinstance(layer, QuantizationManager) and layer.quantizer.is_initialized:
            getattr(layer, method_name)()

    visit.__name__ = '_set_layer_' + method_name
    return visit


_set_layer_learn_ranges = _switch_initialized('learn_ranges')
_set_layer_fix_ranges = _switch_initialized('fix_ranges')
_set_layer_estimate_ranges = _switch_initialized('estimate_ranges')
_set_layer_estimate_ranges_train = _switch_initialized('estimate_ranges_train')


def _apply_to_managers(visitor):
    def method(self):
        self.apply(visitor)

    return method


class QuantizedModule(nn.Module):
    """Base of every module that owns quantizers: quantization config, weight / activation on-off
    flags and the eval-time cache of fake-quantized parameters."""

    def __init__(self, *args, method=QMethods.asymmetric_uniform, act_method=None, n_bits=8,
                 n_bits_act=None, per_channel_weights=False, per_channel_acts=False, percentile=None,
                 weight_range_method=RangeEstimators.current_minmax, weight_range_options=None,
                 act_range_method=RangeEstimators.running_minmax, act_range_options=None,
                 scale_domain='linear', **kwargs):
        kwargs.pop('quant_dict', None)
        super().__init__(*args, **kwargs)
        config = dict(
            method=method, act_method=act_method or method,
            n_bits=n_bits, n_bits_act=n_bits_act or n_bits,
            per_channel_weights=per_channel_weights, per_channel_acts=per_channel_acts,
            percentile=percentile,
            weight_range_method=weight_range_method, weight_range_options=weight_range_options or {},
            act_range_method=act_range_method, act_range_options=act_range_options or {},
            scale_domain=scale_domain,
        )
        for name, value in config.items():
            setattr(self, name, value)
        self.cached_params = None
        self._caching = True
        self.quant_params = None
        self._quant_w = self._quant_a = False

    # ---- cache ---------------------------------------------------------------------------------
    @property
    def caching(self):
        return self._caching

    @caching.setter
    def caching(self, value: bool):
        self._caching = value
        if not value:
            self.cached_params = None

    def _set_weight_quant(self, on):
        self.cached_params = None
        self._quant_w = on

    # ---- on / off ------------------------------------------------------------------------------
    def quantized_weights(self):
        self._set_weight_quant(True)

    def full_precision_weights(self):
        self._set_weight_quant(False)

    def quantized_acts(self):
        self._quant_a = True

    def full_precision_acts(self):
        self._quant_a = False

    def quantized(self):
        self.quantized_weights()
        self.quantized_acts()

    def full_precision(self):
        self.full_precision_weights()
        self.full_precision_acts()

    # ---- state of every initialised manager below this module ----------------------------------
    learn_ranges = _apply_to_managers(_set_layer_learn_ranges)
    fix_ranges = _apply_to_managers(_set_layer_fix_ranges)
    estimate_ranges = _apply_to_managers(_set_layer_estimate_ranges)
    estimate_ranges_train = _apply_to_managers(_set_layer_estimate_ranges_train)

    def train(self, mode=True):
        super().train(mode)
        if mode:
            self.cached_params = None
        return self

    def _apply(self, *args, **kwargs):
        self.cached_params = None
        return super(QuantizedModule, self)._apply(*args, **kwargs)

    def extra_repr(self):
        own = f'weight_quant={self._quant_w}, act_quant={self._quant_a}'
        parent = super().extra_repr()
        return f'{parent},\n{own}' if parent else own


class QuantizedActivation(QuantizedModule):
    """A stand-alone activation quantizer site (residual sums, attention scores, ...)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.activation_quantizer = QuantizationManager(
            qmethod=self.act_method, init=self.act_range_method,
            qparams=dict(n_bits=self.n_bits_act, scale_domain=self.scale_domain),
            init_params=self.act_range_options)

    def quantize_activations(self, x):
        return self.activation_quantizer(x) if self._quant_a else x

    def forward(self, x):
        return self.quantize_activations(x)


class FP32Acts(nn.Module):
    """Identity stand-in for a quantizer that is switched off."""

    def forward(self, x):
        return x

    def reset_ranges(self):
        pass
