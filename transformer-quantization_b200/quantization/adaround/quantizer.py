"""AdaRound quantizers: learned rounding (up / down) per weight.

Host-side mirror of the reference's quantization/adaround/quantizer.py (class names, attributes
``alpha`` / ``round_mode`` / ``soft_targets`` / ``temperature``, ``ADAROUND_QUANTIZER_MAP``).  In the
relaxation modes the per-weight arithmetic -- floor(w / s) + h(alpha) + zero_point, clamp, dequantize,
the alpha initialisation and d / d alpha -- runs in the sm_100a kernels tq_adaround_{init_alpha,fwd,
bwd}_f32 (csrc/tq_qat.cu); ``nearest`` mode is the parent quantizer's kernel.
"""
import logging

import torch
import torch.nn as nn
from torch.autograd import Function

import tq_native
from quantization.adaround.utils import AdaRoundMode
from quantization.quantizers import AsymmetricUniformQuantizer, QuantizerBase, SymmetricUniformQuantizer

logger = logging.getLogger('AdaRound')
logger.setLevel(logging.INFO)


def logit(p, eps=1e-16):
    p = torch.clamp(p, eps, 1 - eps)
    return -torch.log(1 / p - 1)


def hard_sigmoid(x, zeta=1.1, gamma=-0.1):
    return torch.clamp(torch.sigmoid(x) * (zeta - gamma) + gamma, 0.0, 1.0)


def hard_logit(p, zeta=1.1, gamma=-0.1):
    return -torch.log((zeta - p) / (p - gamma))


class _AdaRoundSoftQuant(Function):
    """soft-target AdaRound forward with the gradient of alpha (weights are constants here)"""

    @staticmethod
    def forward(ctx, alpha, w, quantizer, layout, want_int):
        ctx.save_for_backward(alpha, w)
        ctx.quantizer, ctx.layout, ctx.want_int = quantizer, layout, want_int
        ctx.mode, ctx.temperature = quantizer.round_mode.name, quantizer.temperature
        return tq_native.ops().adaround_fwd(w, alpha, quantizer._spec(), *layout, ctx.mode, True, ctx.temperature,
                                            want_int=want_int)

    @staticmethod
    def backward(ctx, grad_out):
        alpha, w = ctx.saved_tensors
        qz = ctx.quantizer
        if ctx.want_int:            # x_int = y / s + zp: rescale so that the kernel's (g * s) factor cancels
            grad_out = grad_out / qz.scale
        ga = tq_native.ops().adaround_bwd(w, alpha, grad_out, qz._spec(), *ctx.layout, ctx.mode, ctx.temperature)
        return ga, None, None, None, None


class AdaRoundQuantizer(QuantizerBase):
    """Mixin in front of a uniform quantizer class (see ADAROUND_QUANTIZER_MAP)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.alpha = None
        self.round_mode = AdaRoundMode.nearest
        self.soft_targets = False
        self.temperature = None          # sigmoid temperature annealing

    def _relaxed(self):
        if self.round_mode == AdaRoundMode.nearest:
            return False
        if self.round_mode not in AdaRoundMode.RELAXATION:
            raise ValueError(f'Unknown rounding mode: {self.round_mode}')
        return True

    def _init_alpha(self, x_float, layout):
        logger.info('Init alpha to be FP32')
        alpha = tq_native.ops().adaround_init_alpha(x_float.detach(), self._spec(), *layout, self.round_mode.name,
                                                    self.temperature)
        self.alpha = nn.Parameter(alpha, requires_grad=True)

    def _adaround(self, x_float, want_int):
        layout = self._layout(x_float)
        if self.alpha is None:
            self._init_alpha(x_float, layout)
        if self.soft_targets and torch.is_grad_enabled() and self.alpha.requires_grad:
            return _AdaRoundSoftQuant.apply(self.alpha, x_float.detach(), self, layout, want_int)
        return tq_native.ops().adaround_fwd(x_float.detach(), self.alpha.detach(), self._spec(), *layout,
                                            self.round_mode.name, self.soft_targets, self.temperature,
                                            want_int=want_int)

    def to_integer_forward(self, x_float):
        if not self._relaxed():
            return super().to_integer_forward(x_float)
        return self._adaround(x_float, want_int=True)

    def forward(self, x_float):
        if not self._relaxed():
            return super().forward(x_float)
        return self._adaround(x_float, want_int=False)

    def get_rest(self):
        """soft target h(alpha) as a differentiable tensor (used by the rounding regulariser)"""
        if self.round_mode == AdaRoundMode.learned_sigmoid:
            return torch.sigmoid(self.alpha)
        if self.round_mode == AdaRoundMode.learned_hard_sigmoid:
            return hard_sigmoid(self.alpha)
        if self.round_mode == AdaRoundMode.sigmoid_temp_decay:
            return torch.sigmoid(self.alpha / self.temperature)
        raise ValueError(f'Unknown rounding mode: {self.round_mode}')

    def extra_repr(self):
        return ', '.join([f'n_bits={self.n_bits}', f'per_channel={self.per_channel}',
                          f'is_initialized={self.is_initialized}', f'round_mode={self.round_mode}',
                          f'soft_targets={self.soft_targets}', f'temperature={self.temperature}'])


class AdaRoundSymmetricUniformQuantizer(AdaRoundQuantizer, SymmetricUniformQuantizer):
    pass


class AdaRoundAsymmetricUniformQuantizer(AdaRoundQuantizer, AsymmetricUniformQuantizer):
    pass


ADAROUND_QUANTIZER_MAP = {
    SymmetricUniformQuantizer: AdaRoundSymmetricUniformQuantizer,
    AsymmetricUniformQuantizer: AdaRoundAsymmetricUniformQuantizer,
}
