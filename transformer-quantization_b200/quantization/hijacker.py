"""QuantizationHijacker: mixin that turns a torch layer into
``weight-quant -> op -> [activation fn] -> activation-quant``.

Host-side mirror of the reference's quantization/hijacker.py (attribute names kept:
``activation_quantizer``, ``weight_quantizer``, ``activation_function``, ``cached_params``,
``activation_save_target``, ...).  Usage is the reference's::

    class QuantLinear(QuantizationHijacker, nn.Linear): ...

(the hijacker first, the nn.Module being hijacked second).  What differs is the data path:

* the eval-time weight cache keeps the QDQ kernel's output on the device -- the reference rebuilds
  it through numpy (hijacker.py:81-85);
* a subclass may provide ``fused_forward(x, weight, bias)``: op + activation fn + activation
  quantizer in ONE kernel (QuantLinear: TMA-fed tcgen05 GEMM with a QDQ epilogue).  It returns
  None when the current state is outside its support and the three-step path runs instead.
"""
import copy

from torch import nn

from quantization import _dist
from quantization.base_quantized_classes import QuantizedModule
from quantization.quantization_manager import QuantizationManager
from quantization.range_estimators import RangeEstimators

activations_list = [nn.ReLU, nn.ReLU6, nn.Hardtanh, nn.Sigmoid, nn.Tanh, nn.PReLU, nn.GELU]


class QuantizationHijacker(QuantizedModule):

    def __init__(self, *args, activation: nn.Module = None, **kwargs):
        super().__init__(*args, **kwargs)
        if activation:
            assert isinstance(activation, tuple(activations_list))
        self.activation_function = copy.deepcopy(activation) if activation else None
        self.activation_quantizer = self._site_manager('act')
        self.weight_quantizer = self._site_manager('weight')
        self.activation_save_target = None
        self.activation_save_name = None

    def _site_manager(self, kind):
        """the activation / weight QuantizationManager from the module's quantization config"""
        if kind == 'act':
            return QuantizationManager(qmethod=self.act_method, init=self.act_range_method,
                                       per_channel=self.per_channel_acts,
                                       qparams=dict(n_bits=self.n_bits_act, scale_domain=self.scale_domain),
                                       init_params=self.act_range_options)
        # weights: current_minmax takes the percentile option, every other estimator its own dict
        # (reference hijacker.py:52-55)
        if self.weight_range_method == RangeEstimators.current_minmax:
            options = dict(percentile=self.percentile)
        else:
            options = self.weight_range_options
        return QuantizationManager(qmethod=self.method, init=self.weight_range_method,
                                   per_channel=self.per_channel_weights,
                                   qparams=dict(n_bits=self.n_bits, scale_domain=self.scale_domain),
                                   init_params=options)

    # ---- forward -------------------------------------------------------------------------------
    def forward(self, x, offsets=None):
        weight, bias = self.get_params()
        fused = getattr(self, 'fused_forward', None)
        if fused is not None and self.activation_save_target is None:
            out = fused(x, weight, bias)
            if out is not None:
                return out
        return self.quantize_activations(self.run_forward(x, weight, bias, offsets=offsets))

    def get_params(self):
        """(weight, bias): weight fake-quantized when enabled; cached across eval forwards."""
        use_cache = not self.training
        if use_cache and self.cached_params:
            return self.cached_params
        weight, bias = self.get_weight_bias()
        if self._quant_w:
            with _dist.local():                   # weights are replicated: their estimators never all-reduce
                weight = self.weight_quantizer(weight)
        if use_cache and self._caching and self.cached_params is None:
            # the kernel output is already a fresh fp32 device tensor; copy only what aliases the
            # parameters so later in-place edits of them do not leak into the cache
            w_keep = weight.detach().clone() if weight is self.weight else weight.detach()
            b_keep = None if bias is None else bias.detach().clone()
            self.cached_params = (w_keep, b_keep)
        return weight, bias

    def get_weight_bias(self):
        return self.weight, getattr(self, 'bias', None)

    def run_forward(self, x, weight, bias, offsets=None):
        raise NotImplementedError()

    def quantize_activations(self, activations):
        """[activation fn] then the activation quantizer (one quantizer for the whole output)."""
        if self.activation_function is not None:
            activations = self.activation_function(activations)
        dump = self.activation_save_target
        if dump is not None:
            dump[self.activation_save_name] = activations.data.cpu().numpy()
        if self._quant_a:
            activations = self.activation_quantizer(activations)
            if dump is not None:
                dump[self.activation_save_name + '_Q'] = activations.data.cpu().numpy()
        return activations
