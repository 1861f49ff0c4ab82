"""AdaRound option enums, temperature schedule, local loss and layer input / output capture.

Host-side mirror of the reference's quantization/adaround/utils.py (same enum members and class
names / constructor keywords, so configs written for the reference keep working).  Pure control
logic plus scalar maths; the per-weight arithmetic lives in csrc/tq_qat.cu (see quantizer.py).
"""
import logging
import math
from enum import Flag, auto

import torch
import torch.nn.functional as F

from utils.utils import StopForwardException

logger = logging.getLogger('AdaRound')
logger.setLevel(logging.INFO)


def sigmoid(x):
    return 1.0 / (1.0 + math.exp(-x))


class BaseOption(Flag):
    """Flag enum printed by member name, with the list_names() helper the option parsers use"""

    def __str__(self):
        return self.name

    @property
    def cls(self):
        return self.value.cls

    @classmethod
    def list_names(cls):
        return [member.name for member in cls]


# activation handling while rounding is learned: none at all | FP32 during AdaRound, quantized afterwards
AdaRoundActQuantMode = BaseOption('AdaRoundActQuantMode', 'no_act_quant post_adaround')
# where the weight grid comes from: the layer's range estimator | MSE on the weights | MSE on the layer
# output (with FP32 or quantized preceding layers)
AdaRoundInitMode = BaseOption('AdaRoundInitMode', 'range_estimator mse mse_out mse_out_asym')
# regulariser pushing h(alpha) to {0, 1} | annealed sigmoid temperature
AdaRoundLossType = BaseOption('AdaRoundLossType', 'relaxation temp_decay')
AdaRoundTempDecayType = BaseOption('AdaRoundTempDecayType', 'linear cosine sigmoid power exp log')


class AdaRoundMode(BaseOption):
    nearest = auto()                    # plain round-to-nearest (no learned rounding)
    learned_sigmoid = auto()
    learned_hard_sigmoid = auto()
    sigmoid_temp_decay = auto()
    RELAXATION = learned_sigmoid | learned_hard_sigmoid | sigmoid_temp_decay

    @classmethod
    def list_names(cls):
        return [m.name for m in cls if m not in (cls.nearest, cls.RELAXATION)]


MODE_TO_LOSS_TYPE = {AdaRoundMode.sigmoid_temp_decay: AdaRoundLossType.temp_decay}
MODE_TO_LOSS_TYPE.update({m: AdaRoundLossType.relaxation
                          for m in (AdaRoundMode.learned_sigmoid, AdaRoundMode.learned_hard_sigmoid)})


def _sigmoid_progress(u, k):
    off = sigmoid(-k / 2)
    return (sigmoid(k * (u - 0.5)) - off) / (1 - 2 * off)


# progress p(u, shape) in [0, 1] for relative time u in [0, 1]: b(t) = b0 + (b1 - b0) * p
_PROGRESS = {
    'linear': lambda u, k: min(1.0, u),
    'cosine': lambda u, k: 0.5 * (1 - math.cos(u * math.pi)),
    'sigmoid': _sigmoid_progress,
    'power': lambda u, k: u ** k,
    'exp': lambda u, k: (1.0 - math.exp(-k * u)) / (1.0 - math.exp(-k)),
}


class TempDecay:
    """Annealing schedule b(t): ``b_range[0]`` until ``rel_decay_start * t_max``, then down to ``b_range[1]``
    at ``t_max`` along the chosen curve (same curves as the reference, adaround/utils.py:93-133)."""

    def __init__(self, t_max, b_range=(20.0, 2.0), rel_decay_start=0.0,
                 decay_type=AdaRoundTempDecayType.linear, decay_shape=1.0):
        self.t_max = t_max
        self.start_b, self.end_b = b_range
        self.decay_type, self.decay_shape = decay_type, decay_shape
        self.decay_start = rel_decay_start * t_max

    def __call__(self, t):
        b0, b1, k = self.start_b, self.end_b, self.decay_shape
        if t < self.decay_start:
            return b0
        u = (t - self.decay_start) / (self.t_max - self.decay_start)
        name = self.decay_type.name if isinstance(self.decay_type, AdaRoundTempDecayType) else None
        if name == 'log':                   # linear in exp(b / k)
            lo, hi = math.exp(b0 / k), math.exp(b1 / k)
            return k * math.log(lo + (hi - lo) * u)
        if name not in _PROGRESS:
            raise ValueError(f'Unknown temp decay type {self.decay_type}')
        return b0 + (b1 - b0) * _PROGRESS[name](u, k)


class CombinedLoss:
    """Layer reconstruction error + rounding regulariser ``weight * sum(1 - |2 h(alpha) - 1| ** b)`` with
    annealed exponent b; in the temp-decay mode the schedule drives the sigmoid temperature instead
    (reference adaround/utils.py:136-177)."""

    def __init__(self, quantizer, loss_type=AdaRoundLossType.relaxation, weight=0.01, max_count=1000,
                 b_range=(20, 2), warmup=0.0, decay_start=0.0, **temp_decay_kw):
        self.quantizer, self.loss_type, self.weight = quantizer, loss_type, weight
        self.loss_start = max_count * warmup
        self.temp_decay = TempDecay(max_count, b_range=b_range,
                                    rel_decay_start=warmup + (1.0 - warmup) * decay_start, **temp_decay_kw)
        self.iter = 0

    def __call__(self, pred, tgt, *args, **kwargs):
        self.iter += 1
        rec_loss = F.mse_loss(pred, tgt, reduction='none').sum(1).mean()
        b = self.temp_decay(self.iter)
        round_loss = 0
        if self.iter >= self.loss_start:
            if self.loss_type == AdaRoundLossType.temp_decay:
                self.quantizer.temperature = b
            elif self.loss_type == AdaRoundLossType.relaxation:
                h = self.quantizer.get_rest().view(-1)
                round_loss = self.weight * (1 - ((h - 0.5).abs() * 2).pow(b)).sum()
            else:
                raise ValueError(f'Unknown loss type {self.loss_type}')
        total = rec_loss + round_loss
        if self.iter == 1 or self.iter % 100 == 0:
            logger.info(f'Total loss:\t{total:.4f} (rec:{rec_loss:.4f}, round:{round_loss:.3f})\tb={b:.2f}'
                        f'\titer={self.iter}')
        return total


class StopForwardHook:
    def __call__(self, module, *args):
        raise StopForwardException


class DataSaverHook:
    """forward hook that keeps the hooked module's input and / or output and can abort the forward"""

    def __init__(self, store_input=False, store_output=False, stop_forward=False):
        self.store_input, self.store_output, self.stop_forward = store_input, store_output, stop_forward
        self.input_store = self.output_store = None

    def __call__(self, module, input_batch, output_batch):
        if self.store_input:
            self.input_store = input_batch
        if self.store_output:
            self.output_store = output_batch
        if self.stop_forward:
            raise StopForwardException


class GetLayerInpOut:
    """(input, FP32 output) of ``layer`` for a batch of model inputs; with ``asym`` the input is
    re-recorded with the preceding layers quantized (reference adaround/utils.py:201-240)."""

    def __init__(self, model, layer, asym=False, act_quant=False, store_output=True):
        self.model, self.layer, self.asym = model, layer, asym
        self.device = layer.weight.device
        self.act_quant, self.store_output = act_quant, store_output
        self.data_saver = DataSaverHook(store_input=True, store_output=store_output, stop_forward=True)

    def _run(self, model_input):
        try:
            self.model(model_input.to(self.device))
        except StopForwardException:
            pass

    def __call__(self, model_input):
        self.model.full_precision()
        handle = self.layer.register_forward_hook(self.data_saver)
        with torch.no_grad():
            self._run(model_input)
            if self.asym:
                self.data_saver.store_output = False
                self.model.set_quant_state(weight_quant=True, act_quant=self.act_quant)
                self._run(model_input)
                self.data_saver.store_output = True
        handle.remove()
        self.model.full_precision()
        self.layer.quantized_weights()
        return self.data_saver.input_store[0].detach(), self.data_saver.output_store.detach()


class LayerOutputMSE:
    """MSE between the layer's current output and its recorded FP32 output, summed over mini-batches"""

    def __init__(self, layer, get_inp_out, data_tensor, batch_size, name='mse_out'):
        self.input, self.exp_out = get_inp_out(data_tensor)
        self.layer, self.batch_size, self.name = layer, batch_size, name

    def __call__(self):
        loss, bs = 0.0, self.batch_size
        for i in range(math.ceil(self.input.size(0) / bs)):
            loss += F.mse_loss(self.layer(self.input[i * bs:(i + 1) * bs]), self.exp_out[i * bs:(i + 1) * bs]).item()
        return loss


class AdaRoundConfig(dict):
    """option container with attribute access (missing options read as None, like utils.utils.DotDict)"""
    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__

    def __getattr__(self, key):
        return self.get(key)


DEFAULT_ADAROUND_CONFIG = AdaRoundConfig(
    # which layers, how many calibration samples, how the grid is initialised
    layers=('all',), num_samples=1024, init=AdaRoundInitMode.range_estimator,
    # relaxation and its optimiser
    round_mode=AdaRoundMode.learned_hard_sigmoid, asym=True, include_act_func=True, lr=1e-3, iters=1000,
    # regulariser weight and annealing of its exponent
    weight=0.01, annealing=(20, 2), decay_type=AdaRoundTempDecayType.cosine, decay_shape=1.0, decay_start=0.0,
    warmup=0.2,
    act_quant_mode=AdaRoundActQuantMode.post_adaround,
)
