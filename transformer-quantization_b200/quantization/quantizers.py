"""Uniform fake-quantizers backed by the sm_100a kernels of libtq_b200.so.

Host-side mirror of the reference's ``quantization/quantizers.py`` (class / method / buffer names
and error behaviour are kept so ``models/quantized_*.py`` load unchanged); the arithmetic lives in
``csrc/tq_qdq.cu`` / ``csrc/tq_minmax.cu``:

* forward / to_integer_forward  -> tq_qdq_f32 / tq_qdq_axis_f32 / tq_quant_int_f32  (reference
  quantizers.py:172-211: six ATen passes + scalar prologue kernels -> one pass, 8 B / element)
* set_quant_range               -> tq_set_range_{asym,sym}_f32 (reference quantizers.py:234-282,
  334-344: ~10 scalar kernels) -- runs on the device, so calibration never syncs the host
* ``_delta`` / ``_zero_float`` / ``_signed`` stay device buffers with the reference's names,
  dtypes and shapes (state_dicts interchange).  ``signed`` / ``int_min`` / ``int_max`` of the
  symmetric quantizer still return python numbers (one ``.item()``), but nothing on the forward
  path calls them: the kernels read ``_signed`` on the device.
"""
from collections import namedtuple
from enum import Enum

import torch
from torch import nn
from torch.autograd import Function

import tq_native
from quantization import fused_linear


class RoundStraightThrough(Function):
    """round with identity gradient (reference quantizers.py:12-20)."""

    @staticmethod
    def forward(ctx, x):
        return torch.round(x)

    @staticmethod
    def backward(ctx, g):
        return g


class FloorStraightThrough(Function):
    """floor with identity gradient (reference quantizers.py:23-31)."""

    @staticmethod
    def forward(ctx, x):
        return torch.floor(x)

    @staticmethod
    def backward(ctx, g):
        return g


round_ste_func = RoundStraightThrough.apply
floor_ste_func = FloorStraightThrough.apply


class FakeQuantSTE(Function):
    """Quantize-dequantize with the reference's autograd semantics (quantizers.py:172-211 under
    autograd: straight-through round :12-20, clamp masks, gradients of ``_delta`` / ``_zero_float``
    once ``make_range_trainable`` turned them into parameters, :284-288 / :346-349).

    forward  -> tq_qdq_f32 / tq_qdq_axis_f32 (one pass)
    backward -> tq_qdq_bwd_f32: grad_x and the reduced range gradients in one pass over (x, grad_y)
    """

    @staticmethod
    def _frozen(t):
        """Range buffers are rewritten IN PLACE by the next set_quant_range of the same quantizer (stable
        addresses for CUDA graphs), e.g. when one quantizer serves two tensors in a forward (QuantNoNorm:
        weight, then bias) or when ranges keep following the data in train mode.  The reference allocates new
        tensors there, so each autograd node keeps the range it quantized with: do the same with a copy.
        Learnable ranges (requires_grad) are never rewritten by an estimator and must stay the graph's leaves."""
        return t if (t is None or t.requires_grad) else t.clone()

    @staticmethod
    def forward(ctx, x, delta, zero_float, quantizer, layout):
        outer, C, inner = layout
        y = tq_native.ops().qdq(x, quantizer._spec(), outer, C, inner)
        ctx.save_for_backward(x, FakeQuantSTE._frozen(delta), FakeQuantSTE._frozen(zero_float),
                              FakeQuantSTE._frozen(getattr(quantizer, '_signed', None)))
        ctx.quantizer, ctx.layout = quantizer, layout
        return y

    @staticmethod
    def backward(ctx, grad_y):
        x, delta, zero_float, is_signed = ctx.saved_tensors
        qz = ctx.quantizer
        outer, C, inner = ctx.layout
        ops = tq_native.ops()
        spec = ops.spec(delta, zero_float, None if zero_float is not None else is_signed, qz.n_bits,
                        qz.scale_domain == 'log', qz.eps)
        need_x, need_d, need_z = ctx.needs_input_grad[:3]
        gx, gd, gz = ops.qdq_bwd(x, grad_y, spec, delta.numel(), outer, C, inner, want_x=need_x,
                                 want_delta=need_d, want_zero_float=need_z and zero_float is not None)
        return (gx, gd.view_as(delta) if gd is not None else None,
                gz.view_as(zero_float) if gz is not None else None, None, None)


class QuantizerNotInitializedError(Exception):
    """Raised when a quantizer has not initialized (reference quantizers.py:368-372)."""

    def __init__(self):
        super().__init__('Quantizer has not been initialized yet')


def _numel_before(shape, axis):
    n = 1
    for s in shape[:axis]:
        n *= s
    return n


def _numel_after(shape, axis):
    n = 1
    for s in shape[axis + 1:]:
        n *= s
    return n


class QuantizerBase(nn.Module):
    """Abstract quantizer interface (reference quantizers.py:36-78)."""

    def __init__(self, n_bits, per_channel=False, axis=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.n_bits = n_bits
        self.per_channel = per_channel
        self.axis = axis

    def _abstract(self, *_, **__):
        raise NotImplementedError()

    is_initialized = property(_abstract)
    x_max = property(_abstract)
    x_min = property(_abstract)
    symmetric = property(_abstract)
    forward = _abstract
    set_quant_range = _abstract
    _adjust_params_per_axis = _abstract
    _adjust_params_per_channel = _abstract

    def extra_repr(self):
        return (f'n_bits={self.n_bits}, per_channel={self.per_channel}, axis={self.axis}, '
                f'is_initalized={self.is_initialized}')

    def reset(self):
        self._delta = None


class AsymmetricUniformQuantizer(QuantizerBase):
    """Asymmetric uniform fake-quantizer: grid [0, 2^n - 1] with a zero point.

    Parameters follow the reference (quantizers.py:95-107): n_bits, scale_domain ('linear'|'log'),
    per_channel, axis, eps.
    """

    def __init__(self, n_bits, scale_domain='linear', per_channel=False, axis=None, eps=1e-8):
        super().__init__(n_bits, per_channel)
        assert scale_domain in ('linear', 'log')
        self.register_buffer('_delta', None)
        self.register_buffer('_zero_float', None)
        self.n_bits = n_bits
        self.scale_domain = scale_domain
        self.per_channel = per_channel
        self.axis = axis
        self.eps = eps

    # ---- state ------------------------------------------------------------------------------
    @property
    def delta(self):
        if self._delta is None:
            raise QuantizerNotInitializedError()
        return self._delta

    @property
    def zero_float(self):
        if self._zero_float is None:
            raise QuantizerNotInitializedError()
        return self._zero_float

    @property
    def is_initialized(self):
        return self._delta is not None

    @property
    def symmetric(self):
        return False

    @property
    def int_min(self):
        return 0.0

    @property
    def int_max(self):
        return 2.0 ** self.n_bits - 1

    # Introspection only (tiny torch ops on the parameter tensors; the kernels resolve these
    # themselves, see tq::resolve in csrc/tq_common.cuh).
    @property
    def scale(self):
        if self.scale_domain == 'linear':
            return torch.clamp(self.delta, min=self.eps)
        return torch.exp(self.delta)

    @property
    def zero_point(self):
        return torch.clamp(torch.round(self.zero_float), self.int_min, self.int_max)

    @property
    def x_max(self):
        return self.scale * (self.int_max - self.zero_point)

    @property
    def x_min(self):
        return self.scale * (self.int_min - self.zero_point)

    # ---- kernel dispatch --------------------------------------------------------------------
    def _spec(self):
        return tq_native.ops().spec(self.delta, self.zero_float, None, self.n_bits,
                                    self.scale_domain == 'log', self.eps)

    def _layout(self, x):
        """(outer, C, inner) view of x for the parameter broadcast the reference would perform."""
        k = self.delta.numel()
        if self.axis is not None:
            self._adjust_params_per_axis(x)
            if k > 1:
                if x.shape[self.axis] != k:
                    raise RuntimeError(f'The size of tensor a ({x.shape[self.axis]}) must match the size '
                                       f'of tensor b ({k}) at non-singleton dimension {self.axis}')
                return _numel_before(x.shape, self.axis), k, _numel_after(x.shape, self.axis)
        if self.per_channel:
            self._adjust_params_per_channel(x)
            if k > 1:
                if x.shape[0] != k:
                    raise RuntimeError(f'The size of tensor a ({x.shape[0]}) must match the size of '
                                       f'tensor b ({k}) at non-singleton dimension 0')
                return 1, k, x.numel() // k
        if k > 1:                       # vector parameters, no axis: trailing-dim broadcast
            if x.shape[-1] != k:
                raise RuntimeError('quantizer parameters do not broadcast against the input')
            return x.numel() // k, k, 1
        return 1, 1, x.numel()

    def to_integer_forward(self, x_float):
        """x_int = clamp(round(x / scale) + zero_point, int_min, int_max) as an fp32 tensor
        (reference quantizers.py:172-187)."""
        spec = self._spec()
        outer, C, inner = self._layout(x_float)
        yi, _ = tq_native.ops().quant_int(x_float, spec, outer, C, inner, want_f32=True)
        return yi

    def forward(self, x_float):
        """Quantize-dequantize ``x_float`` (reference quantizers.py:189-211)."""
        spec = self._spec()
        outer, C, inner = self._layout(x_float)
        if torch.is_grad_enabled() and (x_float.requires_grad or self._delta.requires_grad or
                                        (self._zero_float is not None and self._zero_float.requires_grad)):
            # training: the output joins the autograd graph (STE + learnable ranges)
            return FakeQuantSTE.apply(x_float, self._delta, self._zero_float, self, (outer, C, inner))
        y = tq_native.ops().qdq(x_float, spec, outer, C, inner)
        if C == 1 and self.n_bits <= 8:
            # remember which grid y lives on: a following QuantLinear feeds the integer grid to
            # the tensor cores (quantization/fused_linear.py)
            fused_linear.tag_output(self, y)
        return y

    def _adjust_params_per_axis(self, x_float):
        """Keep the reference's parameter shape [1,..,C,..,1] (quantizers.py:213-217)."""
        shape = [1] * self.axis + [-1] + [1] * (x_float.dim() - self.axis - 1)
        self._delta = self._delta.view(shape)
        self._zero_float = self._zero_float.view(shape)     # AttributeError for symmetric (A.4-1)

    def _adjust_params_per_channel(self, x):
        """Per-channel parameters are viewed [C, 1, ...] (quantizers.py:219-232)."""
        if x.ndim != self.delta.ndim:
            shape = [-1] + [1] * (x.dim() - 1)
            self._delta = self.delta.view(shape)
            if self._zero_float is not None:
                self._zero_float = self._zero_float.view(shape)

    # ---- range -> parameters ----------------------------------------------------------------
    def _param_device(self):
        for b in (self._delta, self._zero_float):
            if b is not None:
                return b.device
        return tq_native.default_device()

    def _tensorize_min_max(self, x_min, x_max):
        """floats / tensors -> fp32 device tensors (reference quantizers.py:234-261).  The
        ``min(x_min, 0)`` / ``max(x_max, eps)`` clamps of lines 258-259 are applied inside
        tq_set_range_*_f32."""
        if not torch.is_tensor(x_min):
            pair = torch.tensor([float(x_min), float(x_max)], dtype=torch.float32).to(self._param_device())
            return pair[0], pair[1]
        if not x_min.is_cuda:
            dev = tq_native.default_device()
            x_min, x_max = x_min.to(dev), x_max.to(dev)
        if x_min.dim() > 0 and len(x_min) > 1 and not self.per_channel and self.axis is None:
            raise ValueError('x_min and x_max must be a float or 1-D Tensor'
                             ' for per-tensor quantization (per_channel=False)')
        return x_min.detach().float(), x_max.detach().float()

    def _alloc_like(self, name, ref, dtype=torch.float32):
        cur = getattr(self, name)
        if (cur is not None and not isinstance(cur, nn.Parameter) and cur.shape == ref.shape
                and cur.device == ref.device and cur.dtype == dtype):
            return cur                  # updated in place: stable addresses for CUDA graphs
        return torch.empty(ref.shape, dtype=dtype, device=ref.device)

    def set_quant_range(self, x_min, x_max):
        """delta = (x_max - x_min) / (2^n - 1), zero_float = -x_min / delta (quantizers.py:263-282)."""
        x_min, x_max = self._tensorize_min_max(x_min, x_max)
        x_min, x_max = x_min.contiguous(), x_max.contiguous()
        delta = self._alloc_like('_delta', x_min)
        zero_float = self._alloc_like('_zero_float', x_min)
        tq_native.ops().set_range_asym(x_min, x_max, self.n_bits, self.eps, self.scale_domain == 'log',
                                       delta, zero_float)
        self._delta = delta
        self._zero_float = zero_float
        self._range_epoch = getattr(self, '_range_epoch', 0) + 1     # invalidates grid tags of earlier outputs

    def _range_from_tile(self, tile_mm, cur_min, cur_max, mode, momentum, first):
        """set_quant_range fused with the estimator update, fed by a GEMM epilogue's min/max words (tq_calib_finalize_f32)"""
        delta = self._alloc_like('_delta', cur_min)
        zero_float = self._alloc_like('_zero_float', cur_min)
        tq_native.ops().calib_finalize(tile_mm, cur_min, cur_max, mode, momentum, first, False, self.n_bits, self.eps,
                                       self.scale_domain == 'log', delta, zero_float, None)
        self._delta = delta
        self._zero_float = zero_float
        self._range_epoch = getattr(self, '_range_epoch', 0) + 1

    def make_range_trainable(self):
        if self.delta not in self.parameters():
            self._delta = torch.nn.Parameter(self._delta)
            self._zero_float = torch.nn.Parameter(self._zero_float)


class SymmetricUniformQuantizer(AsymmetricUniformQuantizer):
    """Symmetric uniform fake-quantizer: zero point 0; signed grid iff the range reaches below 0
    (reference quantizers.py:291-349)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.register_buffer('_signed', None)

    @property
    def signed(self):
        if self._signed is None:
            raise QuantizerNotInitializedError()
        return self._signed.item()

    @property
    def symmetric(self):
        return True

    @property
    def int_min(self):
        return -(2.0 ** (self.n_bits - 1)) if self.signed else 0

    @property
    def int_max(self):
        return 2.0 ** (self.n_bits - self.signed) - 1

    @property
    def zero_point(self):
        return 0.0

    def _spec(self):
        if self._signed is None:
            raise QuantizerNotInitializedError()
        return tq_native.ops().spec(self.delta, None, self._signed, self.n_bits,
                                    self.scale_domain == 'log', self.eps)

    def set_quant_range(self, x_min, x_max):
        """signed = any(x_min < 0); delta = max(|x_min|, x_max) / int_max (quantizers.py:334-344)."""
        x_min, x_max = self._tensorize_min_max(x_min, x_max)
        x_min, x_max = x_min.contiguous(), x_max.contiguous()
        delta = self._alloc_like('_delta', x_min)
        signed = self._signed
        if signed is None or signed.device != x_min.device:
            signed = torch.empty((), dtype=torch.bool, device=x_min.device)
        tq_native.ops().set_range_sym(x_min, x_max, self.n_bits, self.eps, self.scale_domain == 'log',
                                      delta, signed)
        self._delta = delta
        self._signed = signed
        self._range_epoch = getattr(self, '_range_epoch', 0) + 1

    def _range_from_tile(self, tile_mm, cur_min, cur_max, mode, momentum, first):
        delta = self._alloc_like('_delta', cur_min)
        signed = self._signed
        if signed is None or signed.device != cur_min.device:
            signed = torch.empty((), dtype=torch.bool, device=cur_min.device)
        tq_native.ops().calib_finalize(tile_mm, cur_min, cur_max, mode, momentum, first, True, self.n_bits, self.eps,
                                       self.scale_domain == 'log', delta, None, signed)
        self._delta = delta
        self._signed = signed
        self._range_epoch = getattr(self, '_range_epoch', 0) + 1

    def make_range_trainable(self):
        if self.delta not in self.parameters():
            self._delta = torch.nn.Parameter(self._delta)


QMethodMap = namedtuple('QMethodMap', ['value', 'cls'])


class QMethods(Enum):
    symmetric_uniform = QMethodMap(0, SymmetricUniformQuantizer)
    asymmetric_uniform = QMethodMap(1, AsymmetricUniformQuantizer)

    @property
    def cls(self):
        return self.value.cls

    @classmethod
    def list(cls):
        return [m.name for m in cls]
