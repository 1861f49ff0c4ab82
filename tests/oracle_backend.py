"""TEST-ONLY arithmetic back-end: the CPU oracle behind the ``tq_native.CudaOps`` interface.

Lets the host-side logic of the package (quantizers, estimators, manager state machine, hijacker,
model conversion, PEG wiring) run on CPU tensors in the ``-m "not gpu"`` suite.  It is injected by
the ``oracle_ops`` fixture with ``monkeypatch.setattr(tq_native, '_OPS', OracleOps())`` -- nothing
under ``transformer-quantization_b200/`` imports it, and the product path raises without the
CUDA library.
"""
import numpy as np
import torch

from oracle import fakequant_oracle as O


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


class _Spec:
    def __init__(self, delta, zero_float, is_signed, n_bits, log_domain, eps):
        self.delta, self.zero_float, self.is_signed = delta, zero_float, is_signed
        self.n_bits, self.log_domain, self.eps = int(n_bits), bool(log_domain), float(eps)

    def resolve(self):
        dom = 'log' if self.log_domain else 'linear'
        scale = O.scale_of(_np(self.delta).reshape(-1), self.eps, dom)
        if self.zero_float is not None:
            zp = O.asym_zero_point(_np(self.zero_float).reshape(-1), self.n_bits)
            lo, hi = 0.0, O.asym_int_max(self.n_bits)
        else:
            zp = np.zeros_like(scale)
            lo, hi = O.sym_grid(self.n_bits, bool(_np(self.is_signed)))
        return scale, zp, lo, hi


class OracleOps:
    lib = None

    @staticmethod
    def spec(delta, zero_float, is_signed, n_bits, log_domain=False, eps=1e-8):
        return _Spec(delta, zero_float, is_signed, n_bits, log_domain, eps)

    @staticmethod
    def _bcast(p, outer, C, inner):
        return p.reshape(()) if C == 1 else p.reshape(1, C, 1)

    def _xint(self, x, spec, outer, C, inner):
        scale, zp, lo, hi = spec.resolve()
        n = x.numel()
        if C == 1:
            outer, inner = 1, n
        xv = _np(x).reshape(outer, C, inner) if n else _np(x).reshape(0, C, 1)
        s, z = self._bcast(scale, outer, C, inner), self._bcast(zp, outer, C, inner)
        return O.to_integer(xv, s, z, lo, hi), s, z

    def qdq(self, x, spec, outer=1, C=1, inner=None, out=None):
        xi, s, z = self._xint(x, spec, outer, C, inner)
        y = torch.from_numpy(O.dequantize(xi, s, z).reshape(tuple(x.shape)))
        if out is not None:
            out.copy_(y)
            return out
        return y

    def quant_int(self, x, spec, outer=1, C=1, inner=None, want_f32=True, want_bf16=False):
        xi, s, z = self._xint(x, spec, outer, C, inner)
        yi = torch.from_numpy(xi.reshape(tuple(x.shape)).copy()) if want_f32 else None
        yc = (torch.from_numpy((xi - z).astype(np.float32).reshape(tuple(x.shape))).to(torch.bfloat16)
              if want_bf16 else None)
        return yi, yc

    def minmax(self, x):
        if x.numel() == 0:
            raise RuntimeError('min(): Expected reduction dim to be specified for input.numel() == 0')
        mn, mx = O.minmax_tensor(_np(x))
        return torch.tensor([mn, mx], dtype=torch.float32)

    def minmax_axis(self, x, outer, C, inner):
        v = _np(x).reshape(outer, C, inner)
        r = np.ascontiguousarray(np.swapaxes(v, 0, 1)).reshape(C, -1)
        return torch.from_numpy(r.min(-1).copy()), torch.from_numpy(r.max(-1).copy())

    def group_minmax(self, mn, mx, n_groups, ranges=None):
        order = O.stable_order(_np(ranges)) if ranges is not None else None
        a, b = O.group_minmax(_np(mn), _np(mx), n_groups, order)
        return torch.from_numpy(a), torch.from_numpy(b)

    def dim_ranges(self, mn, mx, first):
        r = (_np(mx) - _np(mn)).astype(np.float32)
        if not first:
            r = (np.float32(0.1) * r + np.float32(1 - 0.1) * r).astype(np.float32)
        return torch.from_numpy(r)

    def range_update(self, new_min, new_max, cur_min, cur_max, mode, momentum=0.0, first=False):
        a, b = _np(new_min), _np(new_max)
        if mode == 0 or first:
            ra, rb = a, b
        elif mode == 1:
            ra, rb = O.ema_update(_np(cur_min), a, momentum), O.ema_update(_np(cur_max), b, momentum)
        else:
            ra, rb = np.minimum(_np(cur_min), a), np.maximum(_np(cur_max), b)
        cur_min.copy_(torch.from_numpy(np.asarray(ra, np.float32)).reshape(cur_min.shape))
        cur_max.copy_(torch.from_numpy(np.asarray(rb, np.float32)).reshape(cur_max.shape))

    def set_range_asym(self, x_min, x_max, n_bits, eps, log_domain, delta, zero_float):
        d, z = O.asym_set_quant_range(_np(x_min), _np(x_max), n_bits, eps, 'log' if log_domain else 'linear')
        delta.copy_(torch.from_numpy(np.asarray(d)).reshape(delta.shape))
        zero_float.copy_(torch.from_numpy(np.asarray(z)).reshape(zero_float.shape))

    def set_range_sym(self, x_min, x_max, n_bits, eps, log_domain, delta, is_signed):
        d, s = O.sym_set_quant_range(_np(x_min), _np(x_max), n_bits, eps, 'log' if log_domain else 'linear')
        delta.copy_(torch.from_numpy(np.asarray(d)).reshape(delta.shape))
        is_signed.fill_(bool(s))

    def mse_sse(self, x, cand, n_cand, loss_accum):
        xv = _np(x).reshape(-1)
        tab = _np(cand).reshape(4, n_cand)
        out = np.array([O.sse(xv, tab[0, c], tab[1, c], tab[2, c], tab[3, c]) for c in range(n_cand)])
        loss_accum += torch.from_numpy(out)

    def mse_argmin(self, loss, cand_xmin, cand_xmax):
        idx = int(np.argmin(_np(loss)))
        return (cand_xmin[idx:idx + 1].clone(), cand_xmax[idx:idx + 1].clone(),
                torch.tensor([idx], dtype=torch.int32))

    # -- training-time path ---------------------------------------------------------------------
    def qdq_bwd(self, x, grad_y, spec, n_params, outer=1, C=1, inner=None, want_x=True, want_delta=True,
                want_zero_float=True):
        n = x.numel()
        if C == 1:
            outer, inner = 1, n
        dom = 'log' if spec.log_domain else 'linear'
        signed = bool(_np(spec.is_signed)) if spec.zero_float is None else None
        gx, gd, gz, _ = O.qdq_backward(_np(x), _np(grad_y), _np(spec.delta), _np(spec.zero_float), signed, spec.n_bits,
                                       spec.eps, dom, layout=(outer, C, inner))
        return (torch.from_numpy(gx.copy()) if want_x else None,
                torch.from_numpy(gd.copy()) if want_delta else None,
                torch.from_numpy(gz.copy()) if (want_zero_float and gz is not None) else None)

    def _ada_grid(self, spec, outer, C, inner):
        scale, zp, lo, hi = spec.resolve()
        return self._bcast(scale, outer, C, inner), self._bcast(zp, outer, C, inner), lo, hi

    def adaround_init_alpha(self, w, spec, outer, C, inner, mode, temperature=None):
        s, _, _, _ = self._ada_grid(spec, outer, C, inner)
        a = O.adaround_alpha_init(_np(w).reshape(outer, C, inner), s, mode, temperature)
        return torch.from_numpy(a.reshape(tuple(w.shape)).copy())

    def adaround_fwd(self, w, alpha, spec, outer, C, inner, mode, soft, temperature=None, want_int=False):
        s, z, lo, hi = self._ada_grid(spec, outer, C, inner)
        xi, _ = O.adaround_to_integer(_np(w).reshape(outer, C, inner), _np(alpha).reshape(outer, C, inner), s, z, lo,
                                      hi, mode, soft, temperature)
        out = xi if want_int else O.dequantize(xi, s, z)
        return torch.from_numpy(out.reshape(tuple(w.shape)).copy())

    def adaround_bwd(self, w, alpha, grad_y, spec, outer, C, inner, mode, temperature=None):
        s, z, lo, hi = self._ada_grid(spec, outer, C, inner)
        ga = O.adaround_grad_alpha(_np(w).reshape(outer, C, inner), _np(alpha).reshape(outer, C, inner),
                                   _np(grad_y).reshape(outer, C, inner), s, z, lo, hi, mode, temperature)
        return torch.from_numpy(ga.reshape(tuple(w.shape)).copy())
