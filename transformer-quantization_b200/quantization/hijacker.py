"""QuantizationHijacker: mixin that wraps a torch layer's forward as
``weight-quant -> op -> [activation fn] -> activation-quant``.

Mirror of the reference's quantization/hijacker.py (same attribute names: ``activation_quantizer``,
``weight_quantizer``, ``activation_function``, ``cached_params``, ``activation_save_target`` ...).
Usage is unchanged::

    class QuantLinear(QuantizationHijacker, nn.Linear): ...

Differences from the reference, all on the data path:
* the eval-time weight cache keeps the kernel's output tensor on the device instead of the
  reference's GPU -> numpy -> GPU round trip (hijacker.py:81-85);
* subclasses may implement ``fused_forward(x, weight, bias)`` -- op + activation fn + activation
  quantizer in ONE kernel (QuantLinear: TMA-fed tcgen05 GEMM with a QDQ epilogue).  It is used
  when it reports it can handle the current state, otherwise the three-step path below runs.
"""
import copy

from torch import nn

from quantization.base_quantized_classes import QuantizedModule
from quantization.quantization_manager import QuantizationManager
from quantization.range_estimators import RangeEstimators

activations_list = [nn.ReLU, nn.ReLU6, nn.Hardtanh, nn.Sigmoid, nn.Tanh, nn.PReLU, nn.GELU]


class QuantizationHijacker(QuantizedModule):
    """Must be the FIRST base class; the second one must be the nn.Module being hijacked."""

    def __init__(self, *args, activation: nn.Module = None, **kwargs):
        super().__init__(*args, **kwargs)
        if activation:
            assert isinstance(activation, tuple(activations_list))
        self.activation_function = copy.deepcopy(activation) if activation else None

        self.activation_quantizer = QuantizationManager(
            qmethod=self.act_method,
            init=self.act_range_method,
            per_channel=self.per_channel_acts,
            qparams=dict(n_bits=self.n_bits_act, scale_domain=self.scale_domain),
            init_params=self.act_range_options,
        )
        # current_minmax weight ranges take the percentile option, every other estimator its own
        # options dict (reference hijacker.py:52-55)
        if self.weight_range_method == RangeEstimators.current_minmax:
            weight_init_params = dict(percentile=self.percentile)
        else:
            weight_init_params = self.weight_range_options
        self.weight_quantizer = QuantizationManager(
            qmethod=self.method,
            init=self.weight_range_method,
            per_channel=self.per_channel_weights,
            qparams=dict(n_bits=self.n_bits, scale_domain=self.scale_domain),
            init_params=weight_init_params,
        )
        self.activation_save_target = None
        self.activation_save_name = None

    # ---- forward -------------------------------------------------------------------------------
    def forward(self, x, offsets=None):
        weight, bias = self.get_params()
        fused = getattr(self, 'fused_forward', None)
        if fused is not None and self.activation_save_target is None:
            res = fused(x, weight, bias)
            if res is not None:
                return res
        res = self.run_forward(x, weight, bias, offsets=offsets)
        return self.quantize_activations(res)

    def get_params(self):
        """(weight, bias) with the weight fake-quantized if enabled; cached in eval mode."""
        if not self.training and self.cached_params:
            return self.cached_params

        weight, bias = self.get_weight_bias()
        if self._quant_w:
            weight = self.weight_quantizer(weight)

        if self._caching and not self.training and self.cached_params is None:
            # the QDQ kernel already produced a fresh fp32 device tensor: keep it (the reference
            # rebuilds it through numpy).  Copies, like the reference's, so later in-place edits of
            # the parameters do not leak into the cache.
            self.cached_params = (
                weight.detach().clone() if weight is self.weight else weight.detach(),
                bias.detach().clone() if bias is not None else None,
            )
        return weight, bias

    def get_weight_bias(self):
        return self.weight, getattr(self, 'bias', None)

    def run_forward(self, x, weight, bias, offsets=None):
        raise NotImplementedError()

    def quantize_activations(self, activations):
        """[activation fn] -> activation quantizer (one quantizer for the whole output)."""
        if self.activation_function is not None:
            activations = self.activation_function(activations)

        if self.activation_save_target is not None:
            self.activation_save_target[self.activation_save_name] = activations.data.cpu().numpy()

        if self._quant_a:
            activations = self.activation_quantizer(activations)
            if self.activation_save_target is not None:
                self.activation_save_target[self.activation_save_name + '_Q'] = (
                    activations.data.cpu().numpy())
        return activations
