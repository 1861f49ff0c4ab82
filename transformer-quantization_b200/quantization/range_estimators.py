"""Calibration-time range estimators backed by the sm_100a reduction kernels.

Host-side mirror of the reference's ``quantization/range_estimators.py`` (same class names,
constructor signatures, buffers ``current_xmin`` / ``current_xmax`` and attributes ``per_channel,
quantizer, axis, n_groups, per_group_range_estimation, ranges``).  Differences are all "where the
work happens":

* min / max (reference :82-85, :115-116, :142-143, ...) -> one-pass kernels tq_minmax_f32 /
  tq_minmax_axis_f32 (no transpose copy, no second read);
* PEG grouping + range permutation (reference :87-112: argsort, dense CxC permutation matmul,
  repeat_interleave) -> tq_group_minmax_f32 on the [C] vectors;
* EMA / running min-max (reference :162-167, :209-214) -> tq_range_update_f32 on the device;
* MSE grid search (reference :356-420: one deep-copied quantizer + ~14 kernels + a host sync per
  candidate) -> the candidate table is built once on the host exactly like the reference builds
  each temporary quantizer, and tq_mse_sse_f32 evaluates ALL candidates in one read of the tensor;
  losses accumulate in a device fp64 array, argmin + range selection run on the device too.
  After the first batch (which needs min/max on the host to lay out the search grid, as in the
  reference :341-353) the estimator never synchronises.
* golden section keeps scipy's ``minimize_scalar(method='Bounded')`` on the host like the
  reference (:296-327, :422-470); its objective is the same fused kernel with one candidate.

Cross-rank calibration (torch.distributed, NCCL over NVLink): when a process group is
initialised and ``TQ_DIST_CALIBRATION`` is not "0", every estimator all-reduces its statistics
(MIN/MAX for ranges, SUM for MSE losses) before updating its state, so N ranks calibrating on N
shards reproduce single-GPU calibration on the concatenated batch (see ``_dist.py``).
"""
import copy
from collections import namedtuple
from enum import Enum

import math

import numpy as np
import torch
from scipy.optimize import minimize_scalar
from torch import nn
from torch.nn import functional as F

import tq_native
from quantization import _dist
from quantization.utils import to_numpy


def _numel(shape):
    n = 1
    for s in shape:
        n *= s
    return n


class RangeEstimatorBase(nn.Module):
    def __init__(self, per_channel=False, quantizer=None, axis=None, n_groups=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.register_buffer('current_xmin', None)
        self.register_buffer('current_xmax', None)
        self.per_channel = per_channel
        self.quantizer = quantizer
        self.axis = axis
        self.n_groups = n_groups
        self.per_group_range_estimation = False
        self.ranges = None

    def forward(self, x):
        """Update and return (current_xmin, current_xmax) for input ``x``."""
        raise NotImplementedError()

    def reset(self):
        self.current_xmin = None
        self.current_xmax = None

    def __repr__(self):
        # do not list submodules (the quantizer) -- same rendering as the reference (:53-59)
        lines = self.extra_repr().split('\n')
        extra = lines[0] if len(lines) == 1 else '\n  ' + '\n  '.join(lines) + '\n'
        return self._get_name() + '(' + extra + ')'

    # ---- shared kernel-backed statistics -------------------------------------------------------
    def _axis_minmax(self, x):
        """per-dim (mn[C], mx[C]) along ``self.axis`` without materialising a transpose."""
        C = x.shape[self.axis]
        outer, inner = _numel(x.shape[:self.axis]), _numel(x.shape[self.axis + 1:])
        mn, mx = tq_native.ops().minmax_axis(x.detach(), outer, C, inner)
        return _dist.allreduce_minmax(mn, mx)

    def _channel_minmax(self, x):
        C = x.shape[0]
        mn, mx = tq_native.ops().minmax_axis(x.detach(), 1, C, x.numel() // C)
        return _dist.allreduce_minmax(mn, mx)

    def _tensor_minmax(self, x):
        mm = tq_native.ops().minmax(x.detach())
        mn, mx = _dist.allreduce_minmax(mm[0], mm[1])
        return mn, mx

    # ---- calibration-time fused GEMM: the producer's epilogue already reduced min / max (tile_minmax) ----
    def fused_minmax_mode(self):
        """(mode, momentum) if this estimator's update is a per-tensor min/max rule tq_calib_finalize_f32 implements
        (0 current, 1 running EMA, 2 all-time), else None"""
        return None

    def _per_tensor_minmax_rule(self):
        return (self.axis is None and not self.per_channel and not self.per_group_range_estimation
                and not getattr(self, 'percentile', None) and not _dist.enabled())

    def update_from_tile(self, tile_mm, quantizer):
        """estimator update + quantizer.set_quant_range from the GEMM epilogue's ordered-int min/max words, one launch"""
        mode, momentum = self.fused_minmax_mode()
        first = self.current_xmin is None or self.current_xmin.dim() != 0 or self.current_xmin.device != tile_mm.device
        if first:
            self.current_xmin = torch.empty((), dtype=torch.float32, device=tile_mm.device)
            self.current_xmax = torch.empty((), dtype=torch.float32, device=tile_mm.device)
        quantizer._range_from_tile(tile_mm, self.current_xmin, self.current_xmax, mode, momentum, first)

    def _grouped(self, mn, mx, ranges=None):
        ng = self.n_groups
        assert ng > 0 and mn.numel() % ng == 0
        return tq_native.ops().group_minmax(mn, mx, ng, ranges)

    def _store(self, mn, mx, mode, momentum=0.0):
        """current_{xmin,xmax} <- update(mode) on the device; allocates on first use / shape change."""
        first = self.current_xmin is None or self.current_xmin.shape != mn.shape
        if first:
            self.current_xmin = torch.empty_like(mn)
            self.current_xmax = torch.empty_like(mx)
        tq_native.ops().range_update(mn.contiguous(), mx.contiguous(), self.current_xmin, self.current_xmax,
                                     mode, momentum, first)
        return self.current_xmin, self.current_xmax


class CurrentMinMaxEstimator(RangeEstimatorBase):
    """min / max of the current batch (reference :62-145)."""

    def __init__(self, percentile=None, *args, **kwargs):
        self.percentile = percentile
        super().__init__(*args, **kwargs)

    def fused_minmax_mode(self):
        return (0, 0.0) if self._per_tensor_minmax_rule() else None

    def _ranges_pass(self, x):
        # FP32 pass that records per-dim dynamic ranges for the PEG permutation (reference :68-80;
        # the 'running average' there only re-rounds the newest ranges, see tq_dim_ranges_f32)
        assert self.axis != 0
        mn, mx = self._axis_minmax(x)
        self.ranges = tq_native.ops().dim_ranges(mn, mx, first=self.ranges is None)

    @staticmethod
    def _percentiles(rows, qs):
        """np.percentile(rows, qs, axis=-1) (method 'linear') evaluated ON THE DEVICE of ``rows``: sort +
        numpy's own interpolation formula in float64 (numpy promotes float32 data x float64 weights to
        float64, the reference then casts back with torch.Tensor(...)).  Returns len(qs) fp32 vectors."""
        srt, _ = torch.sort(rows, dim=-1)
        n = srt.shape[-1]
        has_nan = torch.isnan(srt[:, -1])                       # numpy: any NaN in a slice -> NaN
        out = []
        for q in qs:
            vi = (n - 1) * (q / 100.0)
            lo = min(max(int(math.floor(vi)), 0), n - 1)
            hi = min(lo + 1, n - 1)
            g = vi - math.floor(vi)
            d = (srt[:, hi] - srt[:, lo]).double()              # numpy: difference in the data's dtype ...
            a, b = srt[:, lo].double(), srt[:, hi].double()      # ... interpolation in float64
            r = b - d * (1.0 - g) if g >= 0.5 else a + d * g     # numpy _lerp
            r = torch.where(has_nan, torch.full_like(r, float('nan')), r)
            out.append(r.float())
        return out

    def _percentile_minmax(self, x):
        # reference :121-127, :133-140 (np.percentile on the host); same numbers, computed where x lives
        if self.per_channel:
            lo, hi = self._percentiles(x.detach().reshape(x.shape[0], -1), (self.percentile, 100 - self.percentile))
            return lo, hi
        lo, hi = self._percentiles(x.detach().reshape(1, -1), (self.percentile, 100))
        return lo, hi

    def forward(self, x):
        if self.per_group_range_estimation:
            self._ranges_pass(x)
            return
        if self.axis is not None:
            mn, mx = self._axis_minmax(x)
            if self.n_groups is not None:
                mn, mx = self._grouped(mn, mx, self.ranges)
        elif self.per_channel:
            if self.percentile:
                self.current_xmin, self.current_xmax = self._percentile_minmax(x)
                return self.current_xmin, self.current_xmax
            mn, mx = self._channel_minmax(x)
        else:
            if self.percentile:
                self.current_xmin, self.current_xmax = self._percentile_minmax(x)
                return self.current_xmin, self.current_xmax
            mn, mx = self._tensor_minmax(x)
        self.current_xmin, self.current_xmax = mn, mx
        return self.current_xmin, self.current_xmax


class AllMinMaxEstimator(RangeEstimatorBase):
    """global min / max over all batches seen so far (reference :148-169; ignores ``axis``)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    def fused_minmax_mode(self):
        return (2, 0.0) if self._per_tensor_minmax_rule() else None

    def forward(self, x):
        mn, mx = self._channel_minmax(x) if self.per_channel else self._tensor_minmax(x)
        return self._store(mn, mx, mode=2)


class RunningMinMaxEstimator(RangeEstimatorBase):
    """exponential moving average of the batch min / max (reference :172-216)."""

    def __init__(self, momentum=0.9, *args, **kwargs):
        self.momentum = momentum
        super().__init__(*args, **kwargs)

    def fused_minmax_mode(self):
        return (1, self.momentum) if self._per_tensor_minmax_rule() else None

    def forward(self, x):
        if self.axis is not None:
            mn, mx = self._axis_minmax(x)
            if self.n_groups is not None:
                mn, mx = self._grouped(mn, mx, None)        # no permutation here (reference :183-193)
        elif self.per_channel:
            mn, mx = self._channel_minmax(x)
        else:
            mn, mx = self._tensor_minmax(x)
        return self._store(mn, mx, mode=1, momentum=self.momentum)


class OptMethod(Enum):
    grid = 1
    golden_section = 2

    @classmethod
    def list(cls):
        return [m.name for m in cls]


class NoDataPassedError(Exception):
    """Raised data has been passed inot the Range Estimator."""

    def __init__(self):
        super().__init__('Data must be pass through the range estimator to be initialized')


class _CandidateTable:
    """Quantizer parameters of a list of (neg_thr, pos_thr) candidates, computed the way
    MSE_Estimator.quantize does (reference :287-294 -> quantizers.py:234-282 / 334-344): a fresh
    per-tensor quantizer whose range is set from the two python floats.  fp32 numpy arithmetic is
    IEEE-identical to the reference's fp32 torch CPU arithmetic."""

    def __init__(self, quantizer, thresholds):
        f32 = np.float32
        n_bits, eps = quantizer.n_bits, f32(quantizer.eps)
        neg = np.array([t[0] for t in thresholds], dtype=np.float64).astype(f32)   # torch.tensor(x).float()
        pos = np.array([t[1] for t in thresholds], dtype=np.float64).astype(f32)
        self.skipped = np.array([not (t[0] or t[1]) for t in thresholds])          # `if x_min or x_max` (:292)
        x_min = np.minimum(neg, f32(0))
        x_max = np.maximum(pos, eps)
        if quantizer.symmetric:
            signed = x_min < 0
            int_max = np.where(signed, 2.0 ** (n_bits - 1) - 1, 2.0 ** n_bits - 1).astype(f32)
            lo = np.where(signed, -(2.0 ** (n_bits - 1)), 0.0).astype(f32)
            delta = (np.maximum(np.abs(x_min), x_max) / int_max).astype(f32)
            zp = np.zeros_like(delta)
            hi = int_max
        else:
            hi = np.full_like(x_min, 2.0 ** n_bits - 1)
            lo = np.zeros_like(x_min)
            delta = ((x_max - x_min) / hi).astype(f32)
            zero_float = (-x_min / delta).astype(f32)
            zp = np.clip(np.rint(zero_float), lo, hi).astype(f32)
        if quantizer.scale_domain == 'log':   # log then exp round trip of the reference
            scale = np.exp(np.log(delta).astype(f32)).astype(f32)
        else:
            scale = np.maximum(delta, eps)
        self.table = np.stack([scale, zp, lo, hi]).astype(f32)                     # [4, n]
        self.n = len(thresholds)


class MSE_Estimator(RangeEstimatorBase):
    """Range search minimising sum((x - QDQ(x))^2) (reference :228-490)."""

    def __init__(self, num_candidates=100, opt_method=OptMethod.grid, range_margin=0.5, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert opt_method in OptMethod
        self.opt_method = opt_method
        self.num_candidates = num_candidates
        self.max_pos_thr = None
        self.max_neg_thr = None
        self.max_search_range = None
        self.one_sided_dist = None
        self.range_margin = range_margin
        if self.quantizer is None:
            raise NotImplementedError('A Quantizer must be given as an argument to the MSE Range' 'Estimator')
        self.max_int_skew = (2 ** self.quantizer.n_bits) // 4  # frozen at construction (quirk A.4-5)
        self._loss_dev = None        # fp64 device accumulator, reference's `loss_array`
        self._tables = None

    # The reference keeps a numpy fp64 `loss_array`; here it lives on the device and is copied out
    # on demand (one sync, only when somebody looks at it).
    @property
    def loss_array(self):
        return None if self._loss_dev is None else self._loss_dev.cpu().numpy()

    @loss_array.setter
    def loss_array(self, value):
        self._loss_dev = None if value is None else torch.as_tensor(value, dtype=torch.float64).to(
            tq_native.default_device())

    @property
    def step_size(self):
        if self.one_sided_dist is None:
            raise NoDataPassedError()
        return self.max_search_range / self.num_candidates

    @property
    def optimization_method(self):
        if self.one_sided_dist is None:
            raise NoDataPassedError()
        one_d = self.one_sided_dist or self.quantizer.symmetric
        if self.opt_method == OptMethod.grid:
            return self._perform_1D_search if one_d else self._perform_2D_search
        if self.opt_method == OptMethod.golden_section:
            return self._golden_section_symmetric if one_d else self._golden_section_asymmetric
        raise NotImplementedError('Optimization Method not Implemented')

    # ---- single-candidate evaluation (golden section, cross entropy) ---------------------------
    def quantize(self, x_float, x_min=None, x_max=None):
        """QDQ with a temporary per-tensor copy of the quantizer (reference :287-294)."""
        temp_q = copy.deepcopy(self.quantizer)
        temp_q.per_channel = False
        if x_min or x_max:
            temp_q.set_quant_range(x_min, x_max)
        return temp_q(x_float)

    def _sse_rows(self, data, neg_thr, pos_thr, per_row):
        """fused QDQ + squared error for ONE candidate; returns numpy (scalar or per-row)."""
        tab = _CandidateTable(self.quantizer, [(neg_thr, pos_thr)])
        if tab.skipped[0]:
            raise RuntimeError('MSE candidate with x_min == x_max == 0 leaves the quantizer uninitialised')
        cand = torch.from_numpy(tab.table).to(data.device)
        rows = data.reshape(len(data), -1) if per_row else data.reshape(1, -1)
        acc = torch.zeros(rows.shape[0], dtype=torch.float64, device=data.device)
        for r in range(rows.shape[0]):
            tq_native.ops().mse_sse(rows[r], cand, 1, acc[r:r + 1])
        acc = _dist.allreduce_sum(acc)
        out = acc.cpu().numpy()
        return out if per_row else out[0]

    def loss_fx(self, data, neg_thr, pos_thr, per_channel_loss=False):
        """sum((data - QDQ(data))^2) for the range (neg_thr, pos_thr) (reference :248-256)."""
        return self._sse_rows(data.detach(), neg_thr, pos_thr, per_channel_loss)

    def golden_sym_loss(self, range, data):
        neg_thr = 0 if self.one_sided_dist else -range
        return self.loss_fx(data, neg_thr, range)

    def golden_asym_shift_loss(self, shift, range, data):
        return self.loss_fx(data, -range + shift, range + shift)

    def golden_asym_range_loss(self, range, data):
        temp_delta = 2 * range / (2 ** self.quantizer.n_bits - 1)
        max_shift = temp_delta * self.max_int_skew
        result = minimize_scalar(self.golden_asym_shift_loss, args=(range, data),
                                 bounds=(-max_shift, max_shift), method='Bounded')
        return result.fun

    # ---- search-space definition ---------------------------------------------------------------
    def _define_search_range(self, data, dmin=None, dmax=None):
        """Lay out the loss array and the search bounds (reference :329-354)."""
        self.channel_groups = len(data) if self.per_channel else 1
        dev = data.device
        self.current_xmax = torch.zeros(self.channel_groups, device=dev)
        self.current_xmin = torch.zeros(self.channel_groups, device=dev)
        if dmin is None:
            mn, mx = self._tensor_minmax(data)
            dmin, dmax = float(mn), float(mx)
        if self.one_sided_dist or self.quantizer.symmetric:
            shape = (self.channel_groups, self.num_candidates + 1)
            self.max_pos_thr = max(abs(dmin), dmax) + self.range_margin
            self.max_neg_thr = -self.max_pos_thr
            self.max_search_range = self.max_pos_thr
        else:
            shape = (self.channel_groups, self.num_candidates + 1, self.max_int_skew, 2)
            self.max_pos_thr = dmax + self.range_margin
            self.max_neg_thr = dmin - self.range_margin
            self.max_search_range = max(abs(self.max_pos_thr), abs(self.max_neg_thr))
        loss = np.zeros(shape)
        loss[:, 0] = np.inf              # candidate 0 (empty interval) is excluded
        self._loss_dev = torch.from_numpy(loss).to(dev)
        self._tables = None

    def _grid_thresholds(self):
        """(neg_thr, pos_thr) python floats of every grid candidate, in loss-array (C) order."""
        out = []
        if self.one_sided_dist or self.quantizer.symmetric:          # reference :361-364
            for c in range(1, self.num_candidates + 1):
                out.append((0 if self.one_sided_dist else -self.step_size * c, self.step_size * c))
            return out
        n_levels = 2 ** self.quantizer.n_bits - 1
        for c in range(1, self.num_candidates + 1):                  # reference :390-401
            start, finish = -self.step_size * c, self.step_size * c
            temp_delta = float(finish - start) / n_levels
            for shift in range(self.max_int_skew):
                for reverse in range(2):
                    skew = ((-1) ** reverse) * shift * temp_delta
                    out.append((max(start + skew, self.max_neg_thr), min(finish + skew, self.max_pos_thr)))
        return out

    def _grid_tables(self, device):
        if self._tables is None:
            thr = self._grid_thresholds()
            tab = _CandidateTable(self.quantizer, thr)
            per_cand = self._loss_dev[0].numel() - len(thr)          # leading dummy entries (cand 0)
            pad = np.zeros(per_cand, dtype=np.float32)
            if self.one_sided_dist or self.quantizer.symmetric:
                # reference :371-374: xmin/xmax = (+-step_size * cand).astype(np.single)
                cxmax = np.array([t[1] for t in thr], dtype=np.float64).astype(np.single)
                cxmin = (np.zeros(len(thr)) if self.one_sided_dist else
                         np.array([t[0] for t in thr], dtype=np.float64)).astype(np.single)
            else:
                cxmin = np.array([t[0] for t in thr], dtype=np.float64).astype(np.single)
                cxmax = np.array([t[1] for t in thr], dtype=np.float64).astype(np.single)
            self._tables = dict(
                cand=torch.from_numpy(tab.table).to(device), n=tab.n, lead=per_cand,
                # ranges of the dummy entries (argmin == 0 only if every loss is inf/NaN)
                cxmin=torch.from_numpy(np.concatenate([pad, cxmin])).to(device),
                cxmax=torch.from_numpy(np.concatenate([pad, cxmax])).to(device))
        return self._tables

    def _grid_search(self, data):
        """Accumulate the losses of all candidates for this batch and re-select the best range
        (reference :356-376 and :378-420).  No host sync."""
        data = data.detach()
        t = self._grid_tables(data.device)
        rows = data.reshape(len(data), -1) if self.per_channel else data.reshape(1, -1)
        batch = torch.zeros(rows.shape[0], t['n'], dtype=torch.float64, device=data.device)
        for ch in range(rows.shape[0]):
            tq_native.ops().mse_sse(rows[ch], t['cand'], t['n'], batch[ch])
        batch = _dist.allreduce_sum(batch)          # ONE collective for all channels
        for ch in range(rows.shape[0]):
            flat = self._loss_dev[ch].view(-1)
            flat[t['lead']:] += batch[ch]
            xmin, xmax, _ = tq_native.ops().mse_argmin(flat, t['cxmin'], t['cxmax'])
            self.current_xmin[ch:ch + 1] = xmin
            self.current_xmax[ch:ch + 1] = xmax

    def _perform_1D_search(self, data):
        self._grid_search(data)

    def _perform_2D_search(self, data):
        self._grid_search(data)

    # ---- golden section (scipy on the host drives the fused objective) -------------------------
    def _segments(self, data):
        for ch in range(self.channel_groups):
            yield ch, (data if (ch == 0 and not self.per_channel) else data[ch])

    def _golden_section_symmetric(self, data):
        for ch, seg in self._segments(data):
            self.result = minimize_scalar(self.golden_sym_loss, args=seg,
                                          bounds=(0.01 * self.max_search_range, self.max_search_range),
                                          method='Bounded')
            self.current_xmax[ch] = torch.tensor(self.result.x).to(device=data.device)
            self.current_xmin[ch] = (torch.tensor(0.0).to(device=data.device) if self.one_sided_dist
                                     else -self.current_xmax[ch])

    def _golden_section_asymmetric(self, data):
        for ch, seg in self._segments(data):
            self.result = minimize_scalar(self.golden_asym_range_loss, args=seg,
                                          bounds=(0.01 * self.max_search_range, self.max_search_range),
                                          method='Bounded')
            self.final_range = self.result.x
            temp_delta = 2 * self.final_range / (2 ** self.quantizer.n_bits - 1)
            max_shift = temp_delta * self.max_int_skew
            self.subresult = minimize_scalar(self.golden_asym_shift_loss, args=(self.final_range, seg),
                                             bounds=(-max_shift, max_shift), method='Bounded')
            self.final_shift = self.subresult.x
            self.current_xmax[ch] = torch.tensor(self.final_range + self.final_shift).to(device=data.device)
            self.current_xmin[ch] = torch.tensor(-self.final_range + self.final_shift).to(device=data.device)

    def forward(self, data):
        if self._loss_dev is None:
            # first batch: one host read of (min, max) to lay out the search grid (reference :473-481)
            mn, mx = self._tensor_minmax(data)
            dmin, dmax = float(mn), float(mx)
            if self.one_sided_dist is None:
                self.one_sided_dist = bool(dmin >= 0)
            self._define_search_range(data, dmin, dmax)
        self.optimization_method(data)
        return self.current_xmin, self.current_xmax

    def reset(self):
        super().reset()
        self._loss_dev = None
        self._tables = None


class CrossEntropyEstimator(MSE_Estimator):
    """Cross-entropy between softmax(x) and softmax(QDQ(x)) as the search objective (reference
    :493-502).  Only ever applied to logits (B, num_labels): the QDQ is our kernel, the tiny
    softmax is a library call."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    def loss_fx(self, data, neg_thr, pos_thr, per_channel_loss=False):
        quantized_data = self.quantize(data, neg_thr, pos_thr)
        log_quantized_probs = F.log_softmax(quantized_data, dim=1)
        unquantized_probs = F.softmax(data, dim=1)
        return to_numpy(torch.sum(-unquantized_probs * log_quantized_probs))

    def _grid_search(self, data):
        # candidate loop on the host: the objective is not a squared error
        thr = self._grid_thresholds()
        lead = self._loss_dev[0].numel() - len(thr)
        losses = np.array([float(self.loss_fx(data, a, b)) for a, b in thr])
        flat = self._loss_dev[0].view(-1)
        flat[lead:] += torch.from_numpy(losses).to(flat.device)
        t = self._grid_tables(data.device)
        xmin, xmax, _ = tq_native.ops().mse_argmin(flat, t['cxmin'], t['cxmax'])
        self.current_xmin[0:1] = xmin
        self.current_xmax[0:1] = xmax


RangeEstimatorMap = namedtuple('RangeEstimatorMap', ['value', 'cls'])


class RangeEstimators(Enum):
    current_minmax = RangeEstimatorMap(0, CurrentMinMaxEstimator)
    allminmax = RangeEstimatorMap(1, AllMinMaxEstimator)
    running_minmax = RangeEstimatorMap(2, RunningMinMaxEstimator)
    MSE = RangeEstimatorMap(3, MSE_Estimator)
    cross_entropy = RangeEstimatorMap(4, CrossEntropyEstimator)

    @property
    def cls(self):
        return self.value.cls

    @classmethod
    def list(cls):
        return [m.name for m in cls]
