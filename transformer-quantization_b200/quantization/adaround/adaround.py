"""AdaRound for one hijacked layer: swap in the AdaRound weight quantizer, optimise the rounding
variables ``alpha`` against the layer's FP32 output, switch to hard rounding.

Host-side mirror of the reference's quantization/adaround/adaround.py (function names, arguments and
the returned DotDict are the reference's).  The loop is orchestration; per iteration the device runs
tq_adaround_fwd_f32 (soft-quantized weight), the layer's GEMM, and tq_adaround_bwd_f32 (d loss / d alpha)
plus the optimiser's elementwise update.
"""
import logging
from math import ceil

import numpy as np
import torch
import torch.nn.functional as F

from quantization.adaround.quantizer import ADAROUND_QUANTIZER_MAP
from quantization.adaround.utils import (
    MODE_TO_LOSS_TYPE,
    AdaRoundInitMode,
    CombinedLoss,
    GetLayerInpOut,
    LayerOutputMSE,
)
from utils.utils import DotDict

logger = logging.getLogger('AdaRound')
logger.setLevel(logging.INFO)


def _adaround_quantizer_like(org):
    """AdaRound quantizer sharing the grid (``_delta`` / ``_zero_float`` / ``_signed``) of ``org``"""
    cls = ADAROUND_QUANTIZER_MAP.get(org.__class__)
    if cls is None:
        raise NotImplementedError(f'AdaRound is not supported for "{org.__class__}"')
    new = cls(n_bits=org.n_bits, scale_domain=org.scale_domain, per_channel=org.per_channel, eps=org.eps)
    new.register_buffer('_delta', org._delta)
    new.register_buffer('_zero_float', org._zero_float)
    if hasattr(org, '_signed'):
        new.register_buffer('_signed', org._signed)
    return new


def apply_adaround_to_layer(model, layer, data_tensor, batch_size, act_quant, adaround_config, keep_gpu=True):
    """Apply AdaRound to ``layer`` of ``model`` (reference adaround.py:27-137)."""
    cfg = adaround_config
    layer.caching = False

    init = cfg.init
    if init == AdaRoundInitMode.mse:
        apply_mse_init(layer)
    elif init == AdaRoundInitMode.mse_out:
        apply_mse_out_init(model, layer, data_tensor, batch_size)
    elif init == AdaRoundInitMode.mse_out_asym:
        apply_mse_out_init(model, layer, data_tensor, batch_size, asym=True)
    elif init != AdaRoundInitMode.range_estimator:        # range_estimator: grid already initialised
        raise ValueError(f'Unknown initialization for AdaRound: {init}')

    org_act_func = layer.activation_function
    if not cfg.include_act_func:
        layer.activation_function = None

    w_quantizer = _adaround_quantizer_like(layer.weight_quantizer.quantizer)
    layer.weight_quantizer.quantizer = w_quantizer
    w_quantizer.round_mode = cfg.round_mode
    w_quantizer.temperature = cfg.annealing[0]

    get_inp_out = GetLayerInpOut(model, layer, asym=cfg.asym, act_quant=act_quant)
    inp, out = get_inp_out(data_tensor[:batch_size])
    loss_soft_before, loss_hard_before = _compute_and_display_local_losses(w_quantizer, layer, inp, out,
                                                                          infix='before optimization')
    w_quantizer.soft_targets = True

    loss_fn = CombinedLoss(quantizer=w_quantizer, loss_type=MODE_TO_LOSS_TYPE[w_quantizer.round_mode],
                           weight=cfg.weight, max_count=cfg.iters, b_range=cfg.annealing, warmup=cfg.warmup,
                           decay_type=cfg.decay_type, decay_shape=cfg.decay_shape, decay_start=cfg.decay_start)
    optimizer = torch.optim.Adam([w_quantizer.alpha], lr=cfg.lr)
    optimize_local_loss(layer, get_inp_out, data_tensor, optimizer, loss_fn, batch_size, cfg.iters, keep_gpu=keep_gpu)

    logger.info(f'Local loss before optimization (hard quant): {loss_hard_before:.7f}')
    loss_soft_after, loss_hard_after = _compute_and_display_local_losses(w_quantizer, layer, inp, out,
                                                                        infix='after optimization')
    w_quantizer.soft_targets = False          # hard up / down decisions from now on
    layer.activation_function = org_act_func
    layer.caching = True
    return DotDict(loss_soft_before=loss_soft_before, loss_hard_before=loss_hard_before,
                   loss_soft_after=loss_soft_after, loss_hard_after=loss_hard_after)


def _compute_and_display_local_losses(quantizer, layer, inp, out, infix=''):
    keep = quantizer.soft_targets
    losses = []
    for soft, tag in ((True, 'soft'), (False, 'hard')):
        quantizer.soft_targets = soft
        with torch.no_grad():
            loss = float(F.mse_loss(layer(inp), out))
        logger.info(f'Local loss {infix.strip() + " " if infix else ""}({tag} quant): {loss:.7f}')
        losses.append(loss)
    quantizer.soft_targets = keep
    return tuple(losses)


def _search_symmetric_range(layer, score_fn, steps=80):
    """grid initialisation shared by the two MSE inits: shrink the symmetric range from |w|max in 1 % steps
    and keep the best-scoring one (reference adaround.py:160-203)"""
    w, q = layer.weight, layer.weight_quantizer.quantizer
    with torch.no_grad():
        w_absmax = torch.max(w.max(), torch.abs(w.min()))
        best_score, best_max = np.inf, w_absmax
        for i in range(steps):
            s = w_absmax * (1.0 - 0.01 * i)
            q.set_quant_range(-s, s)
            score = score_fn(w, q)
            if score < best_score:
                best_score, best_max = score, s
        logger.info(f'Finished: set max={best_max:.3f} (mse={best_score:.7f})')
        q.set_quant_range(-best_max, best_max)


def apply_mse_init(layer):
    _search_symmetric_range(layer, lambda w, q: F.mse_loss(w, q(w)).item())


def apply_mse_out_init(model, layer, data_tensor, batch_size, asym=False):
    get_inp_out = GetLayerInpOut(model, layer, asym=asym)
    loss_fn = LayerOutputMSE(layer, get_inp_out, data_tensor, batch_size)
    _search_symmetric_range(layer, lambda w, q: loss_fn())


def optimize_local_loss(layer, get_inp_out, data_tensor, optimizer, loss_fn, batch_size, iters,
                        use_cached_data=True, keep_gpu=True):
    """AdaRound optimisation loop (reference adaround.py:206-267): the layer's (input, FP32 output) pairs
    are recorded once and kept on the layer's device; every iteration draws a random mini-batch."""
    cached_inps = cached_outs = None
    device = layer.weight.device
    if use_cached_data:
        logger.info('Caching data for local loss optimization')
        with torch.no_grad():
            pairs = [get_inp_out(data_tensor[i * batch_size:(i + 1) * batch_size])
                     for i in range(ceil(data_tensor.size(0) / batch_size))]
            cached_inps = torch.cat([p[0] for p in pairs])
            cached_outs = torch.cat([p[1] for p in pairs])
            device = pairs[-1][0].device
            del pairs
            if not keep_gpu:
                cached_inps, cached_outs = cached_inps.cpu(), cached_outs.cpu()
    n_cached = cached_inps.size(0) if use_cached_data else data_tensor.size(0)
    for _ in range(iters):
        idx = torch.randperm(n_cached)[:batch_size]
        if use_cached_data:
            idx = idx.to(cached_inps.device)
            cur_inp, cur_out = cached_inps[idx].to(device), cached_outs[idx].to(device)
        else:
            cur_inp, cur_out = get_inp_out(data_tensor[idx])
        optimizer.zero_grad()
        loss = loss_fn(layer(cur_inp), cur_out)
        loss.backward()
        optimizer.step()
