"""CPU ORACLE (test infrastructure, NOT product code) -- numpy restatement of the reference's
fake-quantization arithmetic and range estimators.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product path (``transformer-quantization_b200/``) never does; it
fails loudly when the CUDA library is missing.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified reference
(``/root/reference/quantization/*``) in the build container, runs it on seeded inputs and stores
the outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function in
this file against those vectors (bit-exact for integers / ranges, == for fp32 results);
``tests/test_oracle_qat.py`` does the same for the training-time functions (``tests/golden/qat.npz``), and
``tests/test_oracle_vs_reference_live.py`` compares against the reference imported in place on random inputs
(build container only).

Every function cites the reference file:line (paths relative to the reference checkout) whose
arithmetic it restates.  All math is IEEE fp32 (numpy float32), matching torch CPU fp32:
true division, ``np.rint`` == ``torch.round`` (half-to-even), min/max clamp.
"""
import numpy as np

F32 = np.float32


def _f32(a):
    return np.asarray(a, dtype=F32)


# --------------------------------------------------------------------------------------
# quantizer parameter algebra
# --------------------------------------------------------------------------------------
def tensorize_min_max(x_min, x_max, eps=1e-8):
    """quantizers.py:234-261 -- float/array -> fp32, force 0 inside the range, x_max >= eps."""
    x_min = _f32(x_min)
    x_max = _f32(x_max)
    x_min = np.minimum(x_min, np.zeros_like(x_min))                 # :258
    x_max = np.maximum(x_max, np.ones_like(x_max) * F32(eps))       # :259
    return x_min, x_max


def asym_int_max(n_bits):
    """quantizers.py:138-140."""
    return 2.0 ** n_bits - 1


def asym_set_quant_range(x_min, x_max, n_bits, eps=1e-8, scale_domain='linear'):
    """quantizers.py:263-282 -> (_delta, _zero_float) as fp32 arrays."""
    x_min, x_max = tensorize_min_max(x_min, x_max, eps)
    delta = _f32((x_max - x_min) / F32(asym_int_max(n_bits)))       # :276
    zero_float = _f32(-x_min / delta)                               # :277 (uses the linear delta)
    if scale_domain == 'log':
        delta = np.log(delta).astype(F32)                           # :279-280
    return delta, zero_float


def sym_grid(n_bits, signed):
    """quantizers.py:321-328 -> (int_min, int_max)."""
    signed = bool(signed)
    int_min = -(2.0 ** (n_bits - 1)) if signed else 0.0
    int_max = 2.0 ** (n_bits - int(signed)) - 1
    return int_min, int_max


def sym_set_quant_range(x_min, x_max, n_bits, eps=1e-8, scale_domain='linear'):
    """quantizers.py:334-344 -> (_delta fp32, _signed bool)."""
    x_min, x_max = tensorize_min_max(x_min, x_max, eps)
    signed = bool(x_min.min() < 0)                                  # :336
    _, int_max = sym_grid(n_bits, signed)
    x_absmax = np.maximum(np.abs(x_min), x_max)                     # :338
    delta = _f32(x_absmax / F32(int_max))                           # :339
    if scale_domain == 'log':
        delta = np.log(delta).astype(F32)
    return delta, signed


def scale_of(delta, eps=1e-8, scale_domain='linear'):
    """quantizers.py:142-147."""
    delta = _f32(delta)
    if scale_domain == 'linear':
        return np.maximum(delta, F32(eps))
    return np.exp(delta).astype(F32)


def asym_zero_point(zero_float, n_bits):
    """quantizers.py:149-153."""
    zp = np.rint(_f32(zero_float))
    return np.clip(zp, F32(0.0), F32(asym_int_max(n_bits))).astype(F32)


def _view_params(p, x_ndim, axis=None, per_channel=False):
    """quantizers.py:213-232 -- broadcast shape of per-axis / per-channel parameters."""
    p = _f32(p)
    if axis is not None and p.ndim > 0 and p.size > 1:
        shape = [1] * axis + [-1] + [1] * (x_ndim - axis - 1)
        return p.reshape(shape)
    if per_channel and p.ndim > 0 and p.size > 1:
        return p.reshape([-1] + [1] * (x_ndim - 1))
    if p.size == 1:
        return p.reshape(())
    return p


def to_integer(x, scale, zero_point, int_min, int_max):
    """quantizers.py:172-187: clamp(round(x / scale) + zero_point, int_min, int_max)."""
    x = _f32(x)
    x_int = np.rint(x / _f32(scale)) + _f32(zero_point)
    x_int = _f32(x_int)
    # torch.clamp propagates NaN; np.clip does too
    return np.clip(x_int, F32(int_min), F32(int_max)).astype(F32)


def dequantize(x_int, scale, zero_point):
    """quantizers.py:209: scale * (x_int - zero_point)."""
    return (_f32(scale) * (_f32(x_int) - _f32(zero_point))).astype(F32)


def qdq_asym(x, delta, zero_float, n_bits, eps=1e-8, scale_domain='linear', axis=None,
             per_channel=False, return_int=False):
    """AsymmetricUniformQuantizer.forward, quantizers.py:189-211."""
    x = _f32(x)
    scale = _view_params(scale_of(delta, eps, scale_domain), x.ndim, axis, per_channel)
    zp = _view_params(asym_zero_point(zero_float, n_bits), x.ndim, axis, per_channel)
    x_int = to_integer(x, scale, zp, 0.0, asym_int_max(n_bits))
    if return_int:
        return x_int
    return dequantize(x_int, scale, zp)


def qdq_sym(x, delta, signed, n_bits, eps=1e-8, scale_domain='linear', per_channel=False,
            return_int=False):
    """SymmetricUniformQuantizer forward, quantizers.py:189-211 with zero_point == 0.0 (:330-332).
    (axis is unsupported for the symmetric quantizer in the reference: quirk A.4-1.)"""
    x = _f32(x)
    scale = _view_params(scale_of(delta, eps, scale_domain), x.ndim, None, per_channel)
    lo, hi = sym_grid(n_bits, signed)
    x_int = to_integer(x, scale, F32(0.0), lo, hi)
    if return_int:
        return x_int
    return dequantize(x_int, scale, F32(0.0))


# --------------------------------------------------------------------------------------
# min / max range estimators
# --------------------------------------------------------------------------------------
def _rows_along_axis(x, axis):
    """range_estimators.py:82-85: transpose(0, axis).contiguous().view(C, -1)."""
    x = _f32(x)
    if axis != 0:
        x = np.swapaxes(x, 0, axis)
    return np.ascontiguousarray(x).reshape(x.shape[0], -1)


def minmax_tensor(x):
    """range_estimators.py:142-143 (also 159-160, 206-207)."""
    x = _f32(x)
    return F32(x.min()), F32(x.max())


def minmax_axis(x, axis):
    """range_estimators.py:82-85,115-116 / 178-181,196-197 -> ([C], [C])."""
    r = _rows_along_axis(x, axis)
    return r.min(-1), r.max(-1)


def minmax_channel(x):
    """range_estimators.py:118-120,129-130 -- per output channel (dim 0)."""
    r = _f32(x).reshape(np.shape(x)[0], -1)
    return r.min(-1), r.max(-1)


def dim_ranges(x, axis, first=True):
    """FP32 'ranges' pass for the PEG permutation, range_estimators.py:68-80:
    ranges[d] = max_d - min_d.  NB line 78-79: from the second batch on the 'running average' is
    ``0.1 * ranges + (1 - 0.1) * ranges`` of the NEW ranges only (quirk A.4-3) -- the old value is
    dropped, but the fp32 rounding of the two products and the sum is kept (differs from ``ranges``
    by up to 1 ulp)."""
    mn, mx = minmax_axis(x, axis)
    r = (mx - mn).astype(F32)
    if not first:
        momentum = 0.1
        r = (F32(momentum) * r + F32(1 - momentum) * r).astype(F32)
    return r


def stable_order(ranges):
    """range_estimators.py:94 uses torch.argsort (unstable; ties implementation-defined, quirk
    A.4-11).  This build defines ties by ascending index (stable)."""
    return np.argsort(_f32(ranges), kind='stable')


def group_minmax(mn, mx, n_groups, order=None):
    """PEG group statistics, range_estimators.py:87-112 (and 183-193 without permutation).

    The reference permutes rows with a 0/1 permutation matrix (exact), takes min/max over each
    contiguous block of ``C / n_groups`` sorted dims, repeat_interleaves and un-permutes with
    ``P.T.mv`` (exact: one 1.0 per row).  Equivalent gather/scatter on the per-dim vectors:
    """
    mn = _f32(mn)
    mx = _f32(mx)
    C = mn.shape[0]
    assert n_groups > 0 and C % n_groups == 0                        # :89 / :185
    gs = C // n_groups
    if order is None:
        order = np.arange(C)
    order = np.asarray(order)
    gm = mn[order].reshape(n_groups, gs).min(-1)
    gM = mx[order].reshape(n_groups, gs).max(-1)
    out_m = np.empty(C, dtype=F32)
    out_M = np.empty(C, dtype=F32)
    out_m[order] = np.repeat(gm, gs)
    out_M[order] = np.repeat(gM, gs)
    return out_m, out_M


def ema_update(cur, new, momentum):
    """range_estimators.py:209-214: (1 - m) * new + m * cur, in fp32 with python-float factors."""
    new = _f32(new)
    if cur is None:
        return new
    return (F32(1 - momentum) * new + F32(momentum) * _f32(cur)).astype(F32)


class CurrentMinMax:
    """CurrentMinMaxEstimator.forward (range_estimators.py:62-145) without the percentile branch."""

    def __init__(self, per_channel=False, axis=None, n_groups=None):
        self.per_channel, self.axis, self.n_groups = per_channel, axis, n_groups
        self.per_group_range_estimation = False
        self.ranges = None
        self.current_xmin = self.current_xmax = None

    def __call__(self, x):
        if self.per_group_range_estimation:                          # :68-80
            assert self.axis != 0
            self.ranges = dim_ranges(x, self.axis, first=self.ranges is None)
            return None
        if self.axis is not None:
            mn, mx = minmax_axis(x, self.axis)
            if self.n_groups is not None:
                order = stable_order(self.ranges) if self.ranges is not None else None
                mn, mx = group_minmax(mn, mx, self.n_groups, order)
        elif self.per_channel:
            mn, mx = minmax_channel(x)
        else:
            mn, mx = minmax_tensor(x)
        self.current_xmin, self.current_xmax = mn, mx
        return mn, mx


class AllMinMax:
    """AllMinMaxEstimator.forward, range_estimators.py:148-169 (ignores axis, quirk A.4-7)."""

    def __init__(self, per_channel=False, **_):
        self.per_channel = per_channel
        self.current_xmin = self.current_xmax = None

    def __call__(self, x):
        mn, mx = minmax_channel(x) if self.per_channel else minmax_tensor(x)
        if self.current_xmin is None:
            self.current_xmin, self.current_xmax = mn, mx
        else:
            self.current_xmin = np.minimum(self.current_xmin, mn)
            self.current_xmax = np.maximum(self.current_xmax, mx)
        return self.current_xmin, self.current_xmax


class RunningMinMax:
    """RunningMinMaxEstimator.forward, range_estimators.py:172-216 (no permutation: quirk A.4-7)."""

    def __init__(self, momentum=0.9, per_channel=False, axis=None, n_groups=None):
        self.momentum, self.per_channel, self.axis, self.n_groups = momentum, per_channel, axis, n_groups
        self.current_xmin = self.current_xmax = None

    def __call__(self, x):
        if self.axis is not None:
            mn, mx = minmax_axis(x, self.axis)
            if self.n_groups is not None:
                mn, mx = group_minmax(mn, mx, self.n_groups, None)
        elif self.per_channel:
            mn, mx = minmax_channel(x)
        else:
            mn, mx = minmax_tensor(x)
        self.current_xmin = ema_update(self.current_xmin, mn, self.momentum)
        self.current_xmax = ema_update(self.current_xmax, mx, self.momentum)
        return self.current_xmin, self.current_xmax


# --------------------------------------------------------------------------------------
# MSE range estimator
# --------------------------------------------------------------------------------------
def candidate_qparams(neg_thr, pos_thr, n_bits, symmetric, eps=1e-8):
    """MSE_Estimator.quantize, range_estimators.py:287-294: a fresh per-tensor quantizer whose range
    is set from python floats (neg_thr, pos_thr).  Returns (scale, zero_point, int_min, int_max),
    or None when ``if x_min or x_max`` (:292) is falsy (quirk A.4-4: range left untouched)."""
    if not (neg_thr or pos_thr):
        return None
    if symmetric:
        delta, signed = sym_set_quant_range(neg_thr, pos_thr, n_bits, eps)
        lo, hi = sym_grid(n_bits, signed)
        return scale_of(delta, eps), F32(0.0), F32(lo), F32(hi)
    delta, zf = asym_set_quant_range(neg_thr, pos_thr, n_bits, eps)
    return scale_of(delta, eps), asym_zero_point(zf, n_bits), F32(0.0), F32(asym_int_max(n_bits))


def sse(x, scale, zp, lo, hi, per_row=False):
    """MSE_Estimator.loss_fx, range_estimators.py:248-256: sum((x - QDQ(x))**2).  The reference sums
    in fp32 with torch's blocked order; here the squared errors are fp32 and the sum is fp64 (the
    tests compare with a relative tolerance, see tests/test_oracle_golden.py)."""
    x = _f32(x)
    y = dequantize(to_integer(x, scale, zp, lo, hi), scale, zp)
    d = (x - y).astype(F32)
    sq = (d * d).astype(F32)
    if per_row:
        return sq.reshape(len(x), -1).astype(np.float64).sum(1)
    return sq.astype(np.float64).sum()


class MSEGrid:
    """MSE_Estimator with OptMethod.grid, range_estimators.py:228-420, 472-490 (per-tensor)."""

    def __init__(self, n_bits, symmetric, num_candidates=100, range_margin=0.5, eps=1e-8,
                 max_int_skew=None):
        self.n_bits, self.symmetric = n_bits, symmetric
        self.num_candidates, self.range_margin, self.eps = num_candidates, range_margin, eps
        self.max_int_skew = (2 ** n_bits) // 4 if max_int_skew is None else max_int_skew  # :246
        self.loss_array = None
        self.one_sided_dist = None
        self.current_xmin = self.current_xmax = None

    # :329-354
    def _define_search_range(self, x):
        dmin, dmax = float(x.min()), float(x.max())
        if self.one_sided_dist or self.symmetric:
            self.loss_array = np.zeros((1, self.num_candidates + 1))
            self.loss_array[:, 0] = np.inf
            self.max_pos_thr = max(abs(dmin), dmax) + self.range_margin
            self.max_neg_thr = -self.max_pos_thr
            self.max_search_range = self.max_pos_thr
        else:
            self.loss_array = np.zeros([1, self.num_candidates + 1, self.max_int_skew, 2])
            self.loss_array[:, 0, :, :] = np.inf
            self.max_pos_thr = dmax + self.range_margin
            self.max_neg_thr = dmin - self.range_margin
            self.max_search_range = max(abs(self.max_pos_thr), abs(self.max_neg_thr))

    @property
    def step_size(self):                                             # :258-263
        return self.max_search_range / self.num_candidates

    def thresholds_1d(self):
        """(neg_thr, pos_thr) python floats for cand 1..N, range_estimators.py:362-364."""
        out = []
        for c in range(1, self.num_candidates + 1):
            neg = 0 if self.one_sided_dist else -self.step_size * c
            out.append((neg, self.step_size * c))
        return out

    def thresholds_2d(self):
        """[(cand, shift, reverse, neg_thr, pos_thr)], range_estimators.py:390-401."""
        out = []
        for c in range(1, self.num_candidates + 1):
            start = -self.step_size * c
            finish = self.step_size * c
            tdelta = float(finish - start) / (2 ** self.n_bits - 1)
            for shift in range(self.max_int_skew):
                for reverse in range(2):
                    skew = ((-1) ** reverse) * shift * tdelta
                    neg = max(start + skew, self.max_neg_thr)
                    pos = min(finish + skew, self.max_pos_thr)
                    out.append((c, shift, reverse, neg, pos))
        return out

    def _loss(self, x, neg, pos):
        q = candidate_qparams(neg, pos, self.n_bits, self.symmetric, self.eps)
        if q is None:
            raise RuntimeError('quantizer not initialised (x_min == x_max == 0)')
        return sse(x, *q)

    def __call__(self, x):
        x = _f32(x)
        if self.loss_array is None:                                  # :473-481
            if self.one_sided_dist is None:
                self.one_sided_dist = bool(x.min() >= 0)
            self._define_search_range(x)
        if self.one_sided_dist or self.symmetric:                    # :356-376
            for c, (neg, pos) in enumerate(self.thresholds_1d(), start=1):
                self.loss_array[0, c] += self._loss(x, neg, pos)
            min_cand = self.loss_array.argmin(axis=1)
            xmin = (np.zeros(1) if self.one_sided_dist else -self.step_size * min_cand).astype(np.single)
            xmax = (self.step_size * min_cand).astype(np.single)
        else:                                                        # :378-420
            for c, shift, reverse, neg, pos in self.thresholds_2d():
                self.loss_array[0, c, shift, reverse] += self._loss(x, neg, pos)
            mc, ms, mr = np.unravel_index(np.argmin(self.loss_array[0], axis=None),
                                          self.loss_array[0].shape)
            start, finish = -self.step_size * mc, self.step_size * mc
            mdelta = float(finish - start) / (2 ** self.n_bits - 1)
            mskew = ((-1) ** mr) * ms * mdelta
            xmin = np.array([max(start + mskew, self.max_neg_thr)], dtype=np.single)
            xmax = np.array([min(finish + mskew, self.max_pos_thr)], dtype=np.single)
        self.current_xmin, self.current_xmax = xmin, xmax
        return xmin, xmax


class MSEGolden(MSEGrid):
    """MSE_Estimator with OptMethod.golden_section, range_estimators.py:296-327, 422-470.
    scipy.optimize.minimize_scalar(method='Bounded') drives the search exactly as in the
    reference; the objective is the fused QDQ + squared-error sum."""

    def __call__(self, x):
        from scipy.optimize import minimize_scalar
        x = _f32(x)
        if self.loss_array is None:
            if self.one_sided_dist is None:
                self.one_sided_dist = bool(x.min() >= 0)
            self._define_search_range(x)
        lo_b, hi_b = 0.01 * self.max_search_range, self.max_search_range

        def sym_loss(r):                                             # :296-303
            return float(self._loss(x, 0 if self.one_sided_dist else -r, r))

        def shift_loss(shift, r):                                    # :305-312
            return float(self._loss(x, -r + shift, r + shift))

        def range_loss(r):                                           # :314-327
            tdelta = 2 * r / (2 ** self.n_bits - 1)
            ms = tdelta * self.max_int_skew
            return minimize_scalar(shift_loss, args=(r,), bounds=(-ms, ms), method='Bounded').fun

        if self.one_sided_dist or self.symmetric:                    # :422-440
            res = minimize_scalar(sym_loss, bounds=(lo_b, hi_b), method='Bounded')
            xmax = np.array([res.x], dtype=np.single)
            xmin = np.zeros(1, np.single) if self.one_sided_dist else -xmax
        else:                                                        # :442-470
            res = minimize_scalar(range_loss, bounds=(lo_b, hi_b), method='Bounded')
            fr = res.x
            tdelta = 2 * fr / (2 ** self.n_bits - 1)
            ms = tdelta * self.max_int_skew
            sub = minimize_scalar(shift_loss, args=(fr,), bounds=(-ms, ms), method='Bounded')
            xmax = np.array([fr + sub.x], dtype=np.single)
            xmin = np.array([-fr + sub.x], dtype=np.single)
        self.current_xmin, self.current_xmax = xmin, xmax
        return xmin, xmax


# --------------------------------------------------------------------------------------
# training-time path: straight-through backward, learnable ranges, AdaRound soft rounding
# (SURVEY.md section 8(f) ranks 3-4).  Pinned by tests/golden/qat.npz (make_golden_qat.py: the
# reference under torch autograd).
# --------------------------------------------------------------------------------------
def _param_sum(v, C, layout):
    """sum_to_size of an elementwise gradient onto the parameter shape.  ``layout`` = (outer, C,
    inner) view of the tensor; per-tensor parameters: C == 1.  fp64 accumulation (torch sums fp32 in a
    blocked order; tests compare with a tolerance relative to the summed magnitudes)."""
    v = np.asarray(v, np.float64)
    if C == 1:
        return np.array([v.sum()])
    outer, C, inner = layout
    return v.reshape(outer, C, inner).sum(axis=(0, 2))


def qdq_backward(x, g, delta, zero_float, signed, n_bits, eps=1e-8, scale_domain='linear', axis=None,
                 per_channel=False, layout=None):
    """Autograd of AsymmetricUniformQuantizer.forward / SymmetricUniformQuantizer.forward
    (quantizers.py:142-153, 172-211) for the upstream gradient ``g``:

      round_ste            identity gradient                              quantizers.py:12-20
      clamp(u, lo, hi)     gradient where lo <= u <= hi (inclusive)        quantizers.py:185
      x / scale            grad_x = h / s, grad_s += -h * ((x / s) / s)   quantizers.py:184
      scale * (x_int - zp) h = g * s, grad_s += g * (x_int - zp)          quantizers.py:209
      zero_point           used twice (:184 and :209): -sum(h) + sum(h * mask), then the
                           clamp / round_ste of quantizers.py:149-153
      scale                clamp(delta, min=eps): gradient where delta >= eps; exp(delta): * scale

    Returns (grad_x, grad_delta[n_params], grad_zero_float[n_params] | None, mag) where ``mag`` =
    (sum |terms| of grad_scale, sum |terms| of grad_zero_point) per parameter -- the scale of the
    fp32 summation error, used as tolerance base by the tests."""
    x, g = _f32(x), _f32(g)
    delta = _f32(delta).reshape(-1)
    C = delta.size
    s_flat = scale_of(delta, eps, scale_domain)
    asym = zero_float is not None
    if asym:
        zf = _f32(zero_float).reshape(-1)
        zp_flat = asym_zero_point(zf, n_bits)
        lo, hi = 0.0, asym_int_max(n_bits)
    else:
        zp_flat = np.zeros_like(s_flat)
        lo, hi = sym_grid(n_bits, signed)
    shape = x.shape
    if layout is not None:                      # explicit [outer, C, inner] view (the C ABI's convention)
        layout = tuple(int(v) for v in layout)
        x, g = x.reshape(layout), g.reshape(layout)
        s, zp = (s_flat.reshape(1, C, 1), zp_flat.reshape(1, C, 1)) if C > 1 else (s_flat.reshape(()), zp_flat.reshape(()))
    elif C > 1:
        if axis is not None:
            layout = (int(np.prod(x.shape[:axis], dtype=np.int64)), C, int(np.prod(x.shape[axis + 1:], dtype=np.int64)))
        elif per_channel:
            layout = (1, C, x.size // C)
        else:
            layout = (x.size // C, C, 1)
        s = _view_params(s_flat, x.ndim, axis, per_channel) if (axis is not None or per_channel) else s_flat
        zp = _view_params(zp_flat, x.ndim, axis, per_channel) if (axis is not None or per_channel) else zp_flat
    else:
        layout = (1, 1, x.size)
        s, zp = s_flat.reshape(()), zp_flat.reshape(())
    t = (x / s).astype(F32)
    u = (np.rint(t) + zp).astype(F32)
    mask = (u >= F32(lo)) & (u <= F32(hi))
    x_int = np.clip(u, F32(lo), F32(hi)).astype(F32)
    w = (x_int - zp).astype(F32)
    h = (g * s).astype(F32)                       # MulBackward (:209)
    gs1 = (g * w).astype(F32)
    hm = np.where(mask, h, F32(0)).astype(F32)    # ClampBackward (:185)
    grad_x = (hm / s).astype(F32).reshape(shape)  # DivBackward, self (:184)
    gs2 = (-hm * ((x / s).astype(F32) / s).astype(F32)).astype(F32)   # DivBackward, other
    grad_scale = (_param_sum(gs1, C, layout).astype(F32) + _param_sum(gs2, C, layout).astype(F32)).astype(F32)
    mag_s = _param_sum(np.abs(gs1), C, layout) + _param_sum(np.abs(gs2), C, layout)
    if scale_domain == 'linear':
        grad_delta = np.where(delta >= F32(eps), grad_scale, F32(0)).astype(F32)
    else:
        grad_delta = (grad_scale * s_flat).astype(F32)
        mag_s = mag_s * s_flat
    grad_zf, mag_z = None, None
    if asym:
        grad_zp = (_param_sum(-h, C, layout).astype(F32) + _param_sum(hm, C, layout).astype(F32)).astype(F32)
        mag_z = _param_sum(np.abs(h), C, layout) + _param_sum(np.abs(hm), C, layout)
        r = np.rint(zf)
        grad_zf = np.where((r >= F32(lo)) & (r <= F32(hi)), grad_zp, F32(0)).astype(F32)
    return grad_x, grad_delta, grad_zf, (mag_s, mag_z)


ADAROUND_MODES = ('learned_sigmoid', 'learned_hard_sigmoid', 'sigmoid_temp_decay')
ADAROUND_ZETA, ADAROUND_GAMMA = 1.1, -0.1          # adaround/quantizer.py:29,34 defaults


def _sigmoid(a):
    a = _f32(a)
    with np.errstate(over='ignore'):
        return (F32(1) / (F32(1) + np.exp(-a))).astype(F32)


def adaround_alpha_init(x, scale, mode, temperature=None):
    """adaround/quantizer.py:54-71: alpha such that the soft target equals the rounding rest."""
    t = (_f32(x) / _f32(scale)).astype(F32)
    rest = (t - np.floor(t)).astype(F32)
    if mode == 'learned_hard_sigmoid':                                     # hard_logit, :34-36
        return (-np.log((F32(ADAROUND_ZETA) - rest) / (rest - F32(ADAROUND_GAMMA)))).astype(F32)
    p = np.clip(rest, F32(1e-16), F32(1 - 1e-16))                          # logit, :24-26
    a = (-np.log(F32(1) / p - F32(1))).astype(F32)
    if mode == 'sigmoid_temp_decay':
        a = (F32(temperature) * a).astype(F32)
    return a


def adaround_rest(alpha, mode, temperature=None):
    """AdaRoundQuantizer.get_rest, adaround/quantizer.py:84-92."""
    alpha = _f32(alpha)
    if mode == 'learned_sigmoid':
        return _sigmoid(alpha)
    if mode == 'learned_hard_sigmoid':                                     # hard_sigmoid, :29-31
        p = _sigmoid(alpha)
        return np.clip(p * F32(ADAROUND_ZETA - ADAROUND_GAMMA) + F32(ADAROUND_GAMMA), F32(0), F32(1)).astype(F32)
    return _sigmoid((alpha / F32(temperature)).astype(F32))


def adaround_to_integer(x, alpha, scale, zp, lo, hi, mode, soft, temperature=None):
    """AdaRoundQuantizer.to_integer_forward in a relaxation mode, adaround/quantizer.py:46-82:
    floor(x / scale) + (soft target | alpha >= 0) [+ zero_point], clamped to the grid."""
    t = (_f32(x) / _f32(scale)).astype(F32)
    up = adaround_rest(alpha, mode, temperature) if soft else (_f32(alpha) >= 0).astype(F32)
    u = (np.floor(t) + up).astype(F32)
    u = (u + _f32(zp)).astype(F32)
    return np.clip(u, F32(lo), F32(hi)).astype(F32), u


def adaround_qdq(x, alpha, scale, zp, lo, hi, mode, soft, temperature=None):
    x_int, _ = adaround_to_integer(x, alpha, scale, zp, lo, hi, mode, soft, temperature)
    return dequantize(x_int, scale, zp)


def adaround_grad_alpha(x, alpha, g, scale, zp, lo, hi, mode, temperature=None):
    """d / d alpha of scale * (clamp(floor(x / scale) + rest(alpha) + zp, lo, hi) - zp) for the
    upstream gradient g (soft targets)."""
    _, u = adaround_to_integer(x, alpha, scale, zp, lo, hi, mode, True, temperature)
    mask = (u >= F32(lo)) & (u <= F32(hi))
    h = np.where(mask, (_f32(g) * _f32(scale)).astype(F32), F32(0))
    alpha = _f32(alpha)
    if mode == 'sigmoid_temp_decay':
        p = _sigmoid((alpha / F32(temperature)).astype(F32))
        d = (p * (F32(1) - p) / F32(temperature)).astype(F32)
    else:
        p = _sigmoid(alpha)
        d = (p * (F32(1) - p)).astype(F32)
        if mode == 'learned_hard_sigmoid':
            v = p * F32(ADAROUND_ZETA - ADAROUND_GAMMA) + F32(ADAROUND_GAMMA)
            d = np.where((v >= 0) & (v <= 1), d * F32(ADAROUND_ZETA - ADAROUND_GAMMA), F32(0)).astype(F32)
    return (h * d).astype(F32)


# --------------------------------------------------------------------------------------
# grouped integer GEMM for per-embedding-group (PEG) activations -- the formulation planned for the fused
# engine; checker for that kernel, exercised today by tests/test_oracle_qat.py
# --------------------------------------------------------------------------------------
def peg_linear_exact(x_int, zp, scale, w_int, w_scale, bias, n_groups, order=None):
    """y[m, n] = sum_k  scale[k] * (x_int[m, k] - zp[k]) * w_scale[n] * w_int[n, k]  + bias[n]
    for an activation quantized per embedding group (scale / zp constant inside each of the ``n_groups``
    groups of -- optionally range-permuted -- hidden dims) and a symmetric weight grid, computed the way an
    integer tensor-core kernel would: one exact integer accumulator per group,
        acc_g[m, n] = sum_{k in g} x_int[m, k] * w_int[n, k],     rowsum_g[n] = sum_{k in g} w_int[n, k]
        y = w_scale[n] * sum_g  s_g * (acc_g - zp_g * rowsum_g)  + bias
    (int64 here; int32 suffices on the device: |acc_g| <= 255 * 127 * K).  Returns float64."""
    x_int = np.asarray(x_int, np.int64)
    w_int = np.asarray(w_int, np.int64)
    K = x_int.shape[1]
    order = np.arange(K) if order is None else np.asarray(order)
    gs = K // n_groups
    y = np.zeros((x_int.shape[0], w_int.shape[0]), np.float64)
    for g in range(n_groups):
        cols = order[g * gs:(g + 1) * gs]
        s_g, z_g = float(scale[cols[0]]), int(zp[cols[0]])
        assert np.all(scale[cols] == scale[cols[0]]) and np.all(zp[cols] == zp[cols[0]])
        acc = x_int[:, cols] @ w_int[:, cols].T
        rowsum = w_int[:, cols].sum(1)
        y += s_g * (acc - z_g * rowsum[None, :]).astype(np.float64)
    return y * np.asarray(w_scale, np.float64).reshape(1, -1) + np.asarray(bias, np.float64).reshape(1, -1)
