"""QuantLinear data path.

``linear``      -- the plain op of QuantLinear.run_forward (reference autoquant_utils.py:20-21);
                   only reached when fusion does not apply (e.g. weights not quantized).
``try_fused``   -- GEMM + bias + activation fn + output QDQ as ONE tcgen05 kernel
                   (tq_linear_qdq_bf16); returns None when the layer state is not supported.

Operand carriers.  A fake-quantized tensor is ``scale * (x_int - zero_point)``; for n_bits <= 8 the
centred integer ``x_int - zero_point`` (|v| <= 255) and the weight grid are exact in bf16, so the
GEMM runs on the integer grids with fp32 accumulation in TMEM and the scales are applied once in
the epilogue -- more exact than the reference's fp32 GEMM on dequantized values.  The activation's
grid is recovered exactly from the fp32 tensor (``rint(x / scale)``) using the tag the producing
quantizer attached (``_tq_grid``, invalidated by in-place edits through the tensor version
counter); inputs without a per-tensor <= 8-bit tag (FP32 activations, per-embedding-group inputs
whose scale varies along K, 16-bit activations) take the hi|mid|lo bf16 split path (three bf16
planes, fp32-accurate product).
"""
import torch
from torch import nn
from torch.nn import functional as F

import tq_native

ENABLED = True          # tests flip this to compare against the unfused path
CALIBRATION_FUSION = True      # min/max of a calibrating layer from its GEMM epilogue (tests flip it)

_ACT_CODES = ((nn.GELU, 1), (nn.ReLU, 2), (nn.Tanh, 3))


class GridTag:
    """What a per-tensor quantizer knows about the tensor it just produced."""
    __slots__ = ('quantizer', 'version', 'epoch', 'ctr')

    def __init__(self, quantizer, tensor):
        self.quantizer = quantizer
        self.version = tensor._version
        # the quantizer's range buffers are rewritten in place by its next set_quant_range: the tag is only
        # good for the range that produced the tensor
        self.epoch = getattr(quantizer, '_range_epoch', 0)
        self.ctr = None          # bf16 centred integer grid, filled lazily / by a fused producer


def tag_output(quantizer, y):
    """Called by the quantizers after a per-tensor QDQ with n_bits <= 8."""
    y._tq_grid = GridTag(quantizer, y)
    return y


def _valid_tag(x):
    tag = getattr(x, '_tq_grid', None)
    if tag is None or tag.version != x._version or not tag.quantizer.is_initialized:
        return None
    if getattr(tag.quantizer, '_range_epoch', 0) != tag.epoch:
        return None              # the producing quantizer has been given a new range since
    return tag


def linear(x, weight, bias):
    """Plain (unfused) linear: library GEMM."""
    return F.linear(x.contiguous(), weight.contiguous(), bias=bias)


def _act_code(fn):
    if fn is None:
        return 0
    for cls, code in _ACT_CODES:
        if isinstance(fn, cls):
            if cls is nn.GELU and getattr(fn, 'approximate', 'none') != 'none':
                return None
            return code
    return None


def _weight_grid(layer, weight_q):
    """bf16 integer grid of the (cached) fake-quantized weight + its quantizer spec."""
    mgr = layer.weight_quantizer
    qz = getattr(mgr, 'quantizer', None)
    if qz is None or not qz.is_initialized or qz.n_bits > 8 or qz.axis is not None:
        return None
    cache = getattr(layer, '_tq_wgrid', None)
    if cache is not None and cache[0] is weight_q:
        return cache[1], cache[2], cache[3]
    w = layer.weight.detach()
    N = w.shape[0]
    k = qz.delta.numel()
    if k not in (1, N):
        return None
    spec = qz._spec()
    relaxed = getattr(qz, '_relaxed', None)
    if relaxed is not None and relaxed():
        # AdaRound quantizer in a relaxation mode (learned up / down rounding, soft or hard targets): its grid is
        # not round-to-nearest, so tq_quant_int_f32 does not apply -- the layer runs the three-step path (the
        # AdaRound forward kernel quantizes the weight, then the library GEMM)
        return None
    _, w_ctr = tq_native.ops().quant_int(w, spec, 1, k, w.numel() // k if k > 1 else None,
                                         want_f32=False, want_bf16=True)
    layer._tq_wgrid = (weight_q, w_ctr, spec, k)
    return w_ctr, spec, k


def _wants_grad(layer, weight, bias):
    if weight.requires_grad or (bias is not None and bias.requires_grad) or layer.weight.requires_grad:
        return True
    for mgr in (layer.weight_quantizer, layer.activation_quantizer):
        qz = getattr(mgr, 'quantizer', None)
        for name in ('_delta', '_zero_float'):
            t = getattr(qz, name, None) if qz is not None else None
            if torch.is_tensor(t) and t.requires_grad:
                return True
    return False


def _fused_calibration(mgr):
    """range estimation by a per-tensor min/max rule on a per-tensor quantizer whose range is not being trained:
    the min/max can come out of the GEMM epilogue (tq_linear_qdq_bf16 tile_minmax + tq_calib_finalize_f32)"""
    est, qz = mgr.range_estimator, mgr.quantizer
    if not CALIBRATION_FUSION or est is None or est.per_group_range_estimation or not mgr._updates_ranges():
        return False
    if qz.axis is not None or qz.per_channel or est.fused_minmax_mode() is None:
        return False
    return not any(isinstance(getattr(qz, n, None), nn.Parameter) for n in ('_delta', '_zero_float'))


def try_fused(layer, x, weight, bias):
    """Fused QuantLinear forward or None.  ``weight`` is what get_params() returned."""
    if not ENABLED or layer.training or not layer._quant_w or not x.is_cuda or x.dtype != torch.float32:
        return None
    if torch.is_grad_enabled() and (x.requires_grad or _wants_grad(layer, weight, bias)):
        # the input, the layer's parameters or a learnable range is part of an autograd graph (eval-mode
        # fine-tuning, sensitivity / Fisher passes, a first layer behind frozen embeddings): the fused kernel
        # returns a tensor without grad_fn, so the three-step path must build the graph like the reference does
        return None
    act = _act_code(layer.activation_function)
    if act is None:
        return None
    N, K = layer.weight.shape
    if K % 64 != 0 or N % 8 != 0 or x.shape[-1] != K or x.numel() == 0:
        return None
    wg = _weight_grid(layer, weight)
    if wg is None:
        return None
    w_ctr, w_spec, w_params = wg

    # ---- output quantizer state ----
    mgr = layer.activation_quantizer
    out_spec, out_params, calibrate = None, 1, False
    if layer._quant_a and hasattr(mgr, 'quantizer'):
        if mgr.range_estimator is not None and mgr.range_estimator.per_group_range_estimation:
            calibrate = True                       # FP32 ranges pass: manager handles it, no QDQ
        elif mgr._updates_ranges():
            calibrate = True                       # range is needed before quantizing: two steps
        else:
            qz = mgr.quantizer
            if not qz.is_initialized:
                return None
            k = qz.delta.numel()
            if qz.axis is not None and k > 1:
                if qz.axis != x.dim() - 1 or k != N:
                    return None
                qz._adjust_params_per_axis(x)      # keep the reference's [1,..,C] parameter view
                out_params = N
            elif k != 1:
                return None
            out_spec = qz._spec()

    # ---- input operand ----
    ops = tq_native.ops()
    x2 = x.contiguous().view(-1, K)
    M = x2.shape[0]
    tag = _valid_tag(x)
    a_spec = None
    if tag is not None and tag.quantizer.n_bits <= 8 and tag.quantizer.delta.numel() == 1:
        a_spec = tag.quantizer._spec()
        if tag.ctr is None:
            _, tag.ctr = ops.quant_int(x2, a_spec, want_f32=False, want_bf16=True)
        a_ctr, k_split = tag.ctr.view(M, K), 1
    else:
        a_ctr, k_split = ops.split3(x2), 3

    tile_mm = None
    if calibrate and _fused_calibration(mgr):
        # calibration-time fused GEMM: the epilogue reduces min / max of its own output, ONE launch turns them into the
        # estimator update + quantizer range -- no min/max pass over y, no scalar launches; the QDQ then reads y from L2
        tile_mm = getattr(layer, '_tq_tile_mm', None)
        if tile_mm is None or tile_mm.device != x.device:
            tile_mm = layer._tq_tile_mm = torch.zeros(2, dtype=torch.int32, device=x.device)
    y, _ = ops.linear(a_ctr, w_ctr, bias, M, N, K, k_split, a_spec, w_spec, w_params, act, out_spec,
                      out_params, want_f32=True, want_ctr=False, tile_minmax=tile_mm)
    y = y.view(*x.shape[:-1], N)
    if tile_mm is not None:
        mgr.range_estimator.update_from_tile(tile_mm, mgr.quantizer)
        return mgr.quantizer(y)
    if calibrate:
        return mgr(y)                              # estimator update + set_quant_range + QDQ kernels
    if out_spec is not None and out_params == 1 and mgr.quantizer.n_bits <= 8:
        tag_output(mgr.quantizer, y)
    return y
