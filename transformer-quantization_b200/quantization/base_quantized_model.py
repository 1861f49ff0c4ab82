"""QuantizedModel: model-wide convenience switches (mirror of the reference's
quantization/base_quantized_model.py -- same method names, pure control, no arithmetic)."""
from torch import nn

from quantization.base_quantized_classes import (
    QuantizedModule,
    _set_layer_learn_ranges,
    _set_layer_fix_ranges,
    _set_layer_estimate_ranges,
    _set_layer_estimate_ranges_train,
)


def _broadcast(method_name):
    """Bulk switch: call ``method_name`` on every QuantizedModule of the model."""

    def switch(self):
        def visit(layer):
            if isinstance(layer, QuantizedModule):
                getattr(layer, method_name)()

        self.apply(visit)

    switch.__name__ = method_name
    return switch


def _on_sub_manager(attr, visitor):
    """Apply ``visitor`` to ``module.<attr>`` of every QuantizedModule that has one."""

    def switch(self):
        def visit(module):
            if isinstance(module, QuantizedModule) and hasattr(module, attr):
                visitor(getattr(module, attr))

        self.apply(visit)

    return switch


class QuantizedModel(nn.Module):
    """Parent class of the quantized model wrappers (models/quantized_*.py)."""

    quantized_weights = _broadcast('quantized_weights')
    full_precision_weights = _broadcast('full_precision_weights')
    quantized_acts = _broadcast('quantized_acts')
    full_precision_acts = _broadcast('full_precision_acts')
    quantized = _broadcast('quantized')
    full_precision = _broadcast('full_precision')

    # quantizer state switches
    def learn_ranges(self):
        self.apply(_set_layer_learn_ranges)

    def fix_ranges(self):
        self.apply(_set_layer_fix_ranges)

    def estimate_ranges(self):
        self.apply(_set_layer_estimate_ranges)

    def estimate_ranges_train(self):
        self.apply(_set_layer_estimate_ranges_train)

    fix_act_ranges = _on_sub_manager('activation_quantizer', _set_layer_fix_ranges)
    fix_weight_ranges = _on_sub_manager('weight_quantizer', _set_layer_fix_ranges)
    estimate_act_ranges = _on_sub_manager('activation_quantizer', _set_layer_estimate_ranges)
    reset_act_ranges = _on_sub_manager('activation_quantizer', lambda q: q.reset_ranges())

    def set_quant_state(self, weight_quant, act_quant):
        (self.quantized_acts if act_quant else self.full_precision_acts)()
        (self.quantized_weights if weight_quant else self.full_precision_weights)()
