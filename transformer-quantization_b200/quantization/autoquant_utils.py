"""Model auto-conversion: nn.Linear / nn.LayerNorm / nn.Embedding -> hijacked quantized layers.

Mirror of the reference's quantization/autoquant_utils.py: ``quantize_model``,
``quantize_sequential``, ``quantize_module_list``, ``QuantLinear``, ``QuantLayerNorm``,
``QuantEmbedding``, ``QuantizedActivationWrapper`` and ``module_map`` keep their names, signatures
and conversion rules (a Linear is fused with the first activation found later in the same
Sequential, reference :94-105 / quirk A.4-10).
"""
import copy
import warnings

from torch import nn
from torch.nn import functional as F
from torch.nn.modules.pooling import _AdaptiveAvgPoolNd, _AvgPoolNd

from quantization.base_quantized_classes import FP32Acts, QuantizedActivation, QuantizedModule
from quantization.hijacker import QuantizationHijacker, activations_list
from quantization.quantization_manager import QuantizationManager
from quantization import fused_linear


class QuantLinear(QuantizationHijacker, nn.Linear):
    """nn.Linear with fake-quantized weight and output.  On the fixed-range eval path the whole
    layer (GEMM + bias + activation fn + output QDQ) is one tcgen05 kernel, see fused_linear.py."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    def run_forward(self, x, weight, bias, offsets=None):
        return fused_linear.linear(x, weight, bias)

    def fused_forward(self, x, weight, bias):
        return fused_linear.try_fused(self, x, weight, bias)


class QuantizedActivationWrapper(QuantizedActivation):
    """Runs ``layer`` and quantizes its output; optionally re-uses (ties) the quantizer of the
    preceding layer without updating its range (average pooling and friends)."""

    def __init__(self, layer, tie_activation_quantizers=False,
                 input_quantizer: QuantizationManager = None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.tie_activation_quantizers = tie_activation_quantizers
        if input_quantizer:
            assert isinstance(input_quantizer, QuantizationManager)
            self.activation_quantizer = input_quantizer
        self.layer = layer

    def quantize_activations_no_range_update(self, x):
        return self.activation_quantizer.quantizer(x) if self._quant_a else x

    def forward(self, x):
        x = self.layer(x)
        if self.tie_activation_quantizers:
            return self.quantize_activations_no_range_update(x)
        return self.quantize_activations(x)


class QuantLayerNorm(QuantizationHijacker, nn.LayerNorm):
    def __init__(self, *args, activation=None, **kwargs):
        super().__init__(*args, activation=activation, **kwargs)

    def run_forward(self, x, weight, bias, offsets=None):
        return F.layer_norm(input=x.contiguous(), normalized_shape=self.normalized_shape,
                            weight=weight.contiguous(), bias=bias.contiguous(), eps=self.eps)


class QuantEmbedding(QuantizationHijacker, nn.Embedding):
    def __init__(self, *args, activation=None, **kwargs):
        super().__init__(*args, activation=activation, **kwargs)
        # a lookup of an already-quantized table: the output is not quantized again
        self.activation_quantizer = FP32Acts()

    def run_forward(self, x, weight, bias, offsets=None):
        return F.embedding(input=x.contiguous(), weight=weight.contiguous(), padding_idx=self.padding_idx,
                           max_norm=self.max_norm, norm_type=self.norm_type,
                           scale_grad_by_freq=self.scale_grad_by_freq, sparse=self.sparse)


# ---- model conversion ----------------------------------------------------------------------------------
# Rules (the reference's, autoquant_utils.py:88-241): a layer whose exact type is in ``module_map`` becomes its
# hijacked twin with a cloned weight / bias; inside a Sequential / ModuleList it also absorbs the first activation
# module found LATER in the same container (not necessarily adjacent, quirk A.4-10) and the walk continues one (or
# two) positions further; ``specials`` maps user types to factories; parameter-free pooling layers get a
# QuantizedActivationWrapper, optionally sharing the previous layer's activation quantizer; anything else is
# deep-copied and converted child by child.
module_map = {nn.Linear: QuantLinear, nn.LayerNorm: QuantLayerNorm, nn.Embedding: QuantEmbedding}
non_param_modules = (_AdaptiveAvgPoolNd, _AvgPoolNd)

# constructor arguments that re-create a layer of each convertible type
_CTOR_FIELDS = {
    nn.Linear: ('in_features', 'out_features'),
    nn.LayerNorm: ('normalized_shape', 'eps'),
    nn.Embedding: ('num_embeddings', 'embedding_dim', 'padding_idx', 'max_norm', 'norm_type', 'scale_grad_by_freq',
                   'sparse'),
}


def _ctor_args(module, base):
    kwargs = {f: getattr(module, f) for f in _CTOR_FIELDS[base]}
    if base is nn.Linear:
        kwargs['bias'] = module.bias is not None
    return kwargs


def get_linear_args(module):
    return _ctor_args(module, nn.Linear)


def get_layernorm_args(module):
    return _ctor_args(module, nn.LayerNorm)


def get_embedding_args(module):
    return _ctor_args(module, nn.Embedding)


def get_module_args(mod, act):
    for base in _CTOR_FIELDS:
        if isinstance(mod, base):
            return dict(_ctor_args(mod, base), activation=act)
    raise ValueError


def get_act(module, i):
    """(activation module, its index): the first activation after position ``i`` of the container"""
    for j in range(i + 1, len(module)):
        if isinstance(module[j], tuple(activations_list)):
            return module[j], j
    return None, None


def _hijacked_twin(src, act, quant_params, clone):
    twin = module_map[type(src)](**get_module_args(src, act), **quant_params)
    twin.weight.data = src.weight.data.clone() if clone else src.weight.data
    if getattr(src, 'bias', None) is not None:
        twin.bias.data = src.bias.data.clone() if clone else src.bias.data
    return twin


def quant_module(module, i, **quant_params):
    """Convert ``module[i]`` together with the activation it absorbs -> (new layer, index to continue at)"""
    act, _ = get_act(module, i)
    return _hijacked_twin(module[i], act, quant_params, clone=True), i + (2 if act else 1)


def quantize_sequence(model, specials=None, tie_activation_quantizers=False, **quant_params):
    """converted layers of an indexable container, as a python list"""
    specials = specials or {}
    converted, i = [], 0
    while i < len(model):
        layer, kind = model[i], type(model[i])
        if kind in module_map and not isinstance(layer, QuantizedModule):
            twin, i = quant_module(model, i, **quant_params)
            converted.append(twin)
            continue
        if isinstance(layer, QuantizedModule):
            new = layer
        elif kind in specials:
            new = specials[kind](layer, **quant_params)
        elif isinstance(layer, non_param_modules):
            shared = None
            if tie_activation_quantizers and converted and isinstance(converted[-1], QuantizedModule):
                shared = converted[-1].activation_quantizer
                warnings.warn(f'Tying input quantizer {i}^th layer of type {type(converted[-1])} to the '
                              f'quantized {kind} following it')
            new = QuantizedActivationWrapper(layer, tie_activation_quantizers=tie_activation_quantizers,
                                             input_quantizer=shared, **quant_params)
        else:
            new = quantize_model(layer, specials=specials, **quant_params)
        converted.append(new)
        i += 1
    return converted


def quantize_sequential(model, specials=None, tie_activation_quantizers=False, **quant_params):
    return nn.Sequential(*quantize_sequence(model, specials, tie_activation_quantizers, **quant_params))


def quantize_module_list(model, specials=None, tie_activation_quantizers=False, **quant_params):
    return nn.ModuleList(quantize_sequence(model, specials, tie_activation_quantizers, **quant_params))


def quantize_model(model, specials=None, tie_activation_quantizers=False, **quant_params):
    specials = specials or {}
    kind = type(model)
    if isinstance(model, nn.Sequential):
        return quantize_sequential(model, specials, tie_activation_quantizers, **quant_params)
    if kind in specials:
        return specials[kind](model, **quant_params)
    if isinstance(model, non_param_modules):
        return QuantizedActivationWrapper(model, **quant_params)
    if kind in module_map:          # exact type only: subclasses of Linear / LayerNorm / Embedding fall through
        return _hijacked_twin(model, None, quant_params, clone=False)
    # any other container: convert the children of a deep copy
    twin = copy.deepcopy(model)
    for name, child in twin._modules.items():
        new_child = quantize_model(child, specials=specials, **quant_params)
        if new_child is not None:
            setattr(twin, name, new_child)
    return twin
