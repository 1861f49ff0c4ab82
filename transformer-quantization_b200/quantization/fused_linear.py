"""QuantLinear data path.

``linear``      -- the plain op of QuantLinear.run_forward (reference autoquant_utils.py:20-21).
``try_fused``   -- GEMM + bias + activation fn + output QDQ as ONE tcgen05 kernel
                   (tq_linear_qdq_bf16) when the layer state allows it; returns None otherwise.

Operand carriers.  A fake-quantized tensor is ``scale * (x_int - zero_point)``; for n_bits <= 8 the
centred integer ``x_int - zero_point`` (|v| <= 255) and the weight grid are exact in bf16, so the
GEMM runs on the integer grids with fp32 accumulation in TMEM and the scales are applied once in
the epilogue -- more exact than the reference's fp32 GEMM on dequantized values.  The activation's
grid is recovered exactly from the fp32 tensor (``rint(x / scale)``) using the tag the producing
quantizer attached (``_tq_grid``); inputs without a per-tensor 8-bit tag (FP32 activations,
per-embedding-group inputs whose scale varies along K, 16-bit activations) take the hi|mid|lo
bf16 split path (three bf16 planes, fp32-accurate).
"""
import torch
from torch.nn import functional as F

import tq_native

_STATE = {'enabled': True}


def linear(x, weight, bias):
    """Plain (unfused) linear used during calibration / when fusion is not applicable."""
    return F.linear(x.contiguous(), weight.contiguous(), bias=bias)


def try_fused(layer, x, weight, bias):
    return None
