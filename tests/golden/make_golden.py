#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference
(/root/reference/quantization/*) on seeded inputs.  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference has no tests / golden vectors of its own, so these files pin
parity: tests compare the oracle (oracle/fakequant_oracle.py) and the CUDA kernels against them.
Nothing here is read from /root/reference at test time -- only the .npz outputs are.
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get('TQ_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

from quantization.quantizers import QMethods  # noqa: E402
from quantization.range_estimators import RangeEstimators, OptMethod  # noqa: E402
from quantization.quantization_manager import QuantizationManager  # noqa: E402
from quantization.autoquant_utils import QuantLinear  # noqa: E402
from torch import nn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_grad_enabled(False)


def rnd(seed, shape, scale=3.0, outlier_dims=(), shift=0.0):
    rs = np.random.RandomState(seed)
    x = (rs.randn(*shape) * scale + shift).astype(np.float32)
    for d in outlier_dims:
        x[..., d] *= 20.0
    return x


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def n(v):
    if v is None:
        return np.zeros((0,), np.float32)
    if torch.is_tensor(v):
        return v.detach().cpu().numpy().copy()
    return np.array(v)


# --------------------------------------------------------------------------------------
def quantizer_cases():
    cases = []

    def add(name, kind, n_bits, x, xmin, xmax, axis=None, per_channel=False, scale_domain='linear'):
        cases.append(dict(name=name, kind=kind, n_bits=n_bits, x=x, xmin=xmin, xmax=xmax,
                          axis=axis, per_channel=per_channel, scale_domain=scale_domain))

    x = rnd(1, (4, 16, 96), outlier_dims=(7, 50))
    for nb in (2, 4, 8, 16):
        add(f'asym_t_b{nb}', 'asym', nb, x, float(x.min()), float(x.max()))
        add(f'sym_t_b{nb}', 'sym', nb, x, float(x.min()), float(x.max()))
    # clipped range (forces clamping on both sides)
    add('asym_t_clip', 'asym', 8, x, -2.5, 3.75)
    add('sym_t_clip', 'sym', 8, x, -2.5, 3.75)
    # one-sided data -> unsigned symmetric grid
    xp = np.abs(rnd(2, (3, 5, 7, 11)))
    add('sym_t_unsigned', 'sym', 8, xp, float(xp.min()), float(xp.max()))
    add('asym_t_onesided', 'asym', 8, xp, float(xp.min()), float(xp.max()))
    add('sym_t_unsigned_b4', 'sym', 4, xp, 0.0, 2.0)
    # range not containing zero / negative max
    add('asym_t_posmin', 'asym', 8, x, 1.0, 5.0)
    add('asym_t_negmax', 'asym', 8, x, -5.0, -1.0)
    add('sym_t_negmax', 'sym', 8, x, -5.0, -1.0)
    # tiny range -> delta below eps
    add('asym_t_tiny', 'asym', 8, (x * 1e-9).astype(np.float32), -1e-12, 1e-12)
    add('sym_t_tiny', 'sym', 8, (x * 1e-9).astype(np.float32), -1e-12, 1e-12)
    # exact .5 ties (half-to-even) on a power-of-two scale
    k = np.arange(-300, 300, dtype=np.float32)
    ties = ((k + 0.5) * np.float32(0.125)).astype(np.float32)
    add('asym_t_ties', 'asym', 8, ties, -16.0, 15.875)
    add('sym_t_ties', 'sym', 8, ties, -16.0, 15.875)
    # ragged sizes / size 1 / empty
    add('asym_t_ragged', 'asym', 8, rnd(3, (1237,)), -7.0, 9.0)
    add('sym_t_ragged', 'sym', 8, rnd(3, (1237,)), -7.0, 9.0)
    add('asym_t_one', 'asym', 8, rnd(4, (1,)), -1.0, 1.0)
    add('asym_t_empty', 'asym', 8, np.zeros((0, 5), np.float32), -1.0, 1.0)
    # non-finite inputs
    xs = rnd(5, (257,))
    xs[3], xs[77], xs[100], xs[200] = np.inf, -np.inf, np.nan, 1e30
    add('asym_t_nonfinite', 'asym', 8, xs, -4.0, 6.0)
    add('sym_t_nonfinite', 'sym', 8, xs, -4.0, 6.0)
    # log scale domain
    add('asym_t_log', 'asym', 8, x, float(x.min()), float(x.max()), scale_domain='log')
    add('sym_t_log', 'sym', 8, x, float(x.min()), float(x.max()), scale_domain='log')
    # per-axis (per-embedding) asymmetric: vector ranges along the hidden dim
    r = x.reshape(-1, 96)
    add('asym_axis2', 'asym', 8, x, r.min(0), r.max(0), axis=2)
    add('asym_axis2_b4', 'asym', 4, x, r.min(0), r.max(0), axis=2)
    x2 = rnd(6, (32, 96))
    add('asym_axis1_2d', 'asym', 8, x2, x2.min(0), x2.max(0), axis=1)
    x4 = rnd(7, (2, 6, 5, 4))
    r4 = np.swapaxes(x4, 0, 1).reshape(6, -1)
    add('asym_axis1_4d', 'asym', 8, x4, r4.min(1), r4.max(1), axis=1)
    # per-channel weights
    w = rnd(8, (48, 64), scale=0.02)
    add('sym_pc', 'sym', 8, w, w.min(1), w.max(1), per_channel=True)
    add('asym_pc', 'asym', 8, w, w.min(1), w.max(1), per_channel=True)
    add('sym_pc_b4', 'sym', 4, w, w.min(1), w.max(1), per_channel=True)
    w4 = rnd(9, (8, 3, 3, 5), scale=0.1)
    add('asym_pc_4d', 'asym', 8, w4, w4.reshape(8, -1).min(1), w4.reshape(8, -1).max(1), per_channel=True)

    out, manifest = {}, []
    for c in cases:
        cls = QMethods.symmetric_uniform.cls if c['kind'] == 'sym' else QMethods.asymmetric_uniform.cls
        q = cls(n_bits=c['n_bits'], scale_domain=c['scale_domain'], per_channel=c['per_channel'],
                axis=c['axis'])
        xmin, xmax = c['xmin'], c['xmax']
        if isinstance(xmin, np.ndarray):
            q.set_quant_range(t(xmin.astype(np.float32)), t(xmax.astype(np.float32)))
        else:
            q.set_quant_range(xmin, xmax)
        delta0 = n(q._delta).copy()
        zf0 = n(q._zero_float).copy() if c['kind'] == 'asym' else None
        xt = t(c['x'])
        y = q(xt)
        xi = q.to_integer_forward(xt)
        nm = c['name']
        out[f'{nm}.x'] = c['x']
        out[f'{nm}.xmin'] = np.asarray(xmin, np.float64)
        out[f'{nm}.xmax'] = np.asarray(xmax, np.float64)
        out[f'{nm}.delta'] = delta0
        if zf0 is not None:
            out[f'{nm}.zero_float'] = zf0
            out[f'{nm}.zero_point'] = n(q.zero_point).reshape(-1)
        else:
            out[f'{nm}.signed'] = np.asarray(bool(q.signed))
        out[f'{nm}.scale'] = n(q.scale).reshape(-1)
        out[f'{nm}.int_min'] = np.asarray(float(q.int_min))
        out[f'{nm}.int_max'] = np.asarray(float(q.int_max))
        out[f'{nm}.x_int'] = n(xi)
        out[f'{nm}.x_quant'] = n(y)
        out[f'{nm}.q_x_min'] = n(q.x_min).reshape(-1)
        out[f'{nm}.q_x_max'] = n(q.x_max).reshape(-1)
        manifest.append(dict(name=nm, kind=c['kind'], n_bits=c['n_bits'], axis=c['axis'],
                             per_channel=c['per_channel'], scale_domain=c['scale_domain'],
                             vector_range=isinstance(xmin, np.ndarray)))
    np.savez_compressed(os.path.join(HERE, 'quantizers.npz'), **out)
    return manifest


# --------------------------------------------------------------------------------------
def estimator_cases():
    out, manifest = {}, []
    batches = [rnd(20 + i, (4, 16, 96), outlier_dims=(7, 50), shift=0.3 * i) for i in range(3)]
    wb = [rnd(30 + i, (48, 64), scale=0.02) for i in range(3)]

    def run(name, est_enum, kw, data, permute=False, opts=None):
        qz = QMethods.asymmetric_uniform.cls(n_bits=8)
        est = est_enum.cls(quantizer=qz, **kw, **(opts or {}))
        if permute:
            est.per_group_range_estimation = True
            for b in data:
                est(t(b))
            out[f'{name}.ranges'] = n(est.ranges)
            est.per_group_range_estimation = False
        for i, b in enumerate(data):
            mn, mx = est(t(b))
            out[f'{name}.b{i}.xmin'] = n(mn).reshape(-1)
            out[f'{name}.b{i}.xmax'] = n(mx).reshape(-1)
        for i, b in enumerate(data):
            out[f'{name}.x{i}'] = b
        manifest.append(dict(name=name, est=est_enum.name, n_batches=len(data), permute=permute,
                             opts=opts or {}, **{k: v for k, v in kw.items()}))

    E = RangeEstimators
    run('cur_tensor', E.current_minmax, dict(), batches)
    run('cur_axis2', E.current_minmax, dict(axis=2), batches)
    run('cur_ng6', E.current_minmax, dict(axis=2, n_groups=6), batches)
    run('cur_ng96', E.current_minmax, dict(axis=2, n_groups=96), batches)
    run('cur_ng1', E.current_minmax, dict(axis=2, n_groups=1), batches)
    run('cur_ngp6', E.current_minmax, dict(axis=2, n_groups=6), batches, permute=True)
    run('cur_ngp3', E.current_minmax, dict(axis=2, n_groups=3), batches, permute=True)
    run('cur_pc', E.current_minmax, dict(per_channel=True), wb)
    x2 = [rnd(40 + i, (32, 96)) for i in range(2)]
    run('cur_axis1_2d', E.current_minmax, dict(axis=1), x2)
    run('run_tensor', E.running_minmax, dict(), batches)
    run('run_tensor_m5', E.running_minmax, dict(), batches, opts=dict(momentum=0.5))
    run('run_axis2', E.running_minmax, dict(axis=2), batches)
    run('run_ng6', E.running_minmax, dict(axis=2, n_groups=6), batches)
    run('run_pc', E.running_minmax, dict(per_channel=True), wb)
    run('all_tensor', E.allminmax, dict(), batches)
    run('all_pc', E.allminmax, dict(per_channel=True), wb)
    np.savez_compressed(os.path.join(HERE, 'estimators.npz'), **out)
    return manifest


# --------------------------------------------------------------------------------------
def mse_cases():
    out, manifest = {}, []

    def run(name, qm, n_bits, data, opt, num_candidates=100):
        qz = qm.cls(n_bits=n_bits)
        est = RangeEstimators.MSE.cls(quantizer=qz, opt_method=opt, num_candidates=num_candidates)
        for i, b in enumerate(data):
            mn, mx = est(t(b))
            out[f'{name}.b{i}.xmin'] = n(mn).reshape(-1)
            out[f'{name}.b{i}.xmax'] = n(mx).reshape(-1)
            if opt == OptMethod.grid:
                out[f'{name}.b{i}.loss'] = np.array(est.loss_array)
            out[f'{name}.x{i}'] = b
        out[f'{name}.max_pos_thr'] = np.asarray(est.max_pos_thr, np.float64)
        out[f'{name}.max_neg_thr'] = np.asarray(est.max_neg_thr, np.float64)
        out[f'{name}.max_search_range'] = np.asarray(est.max_search_range, np.float64)
        manifest.append(dict(name=name, kind='sym' if qm == QMethods.symmetric_uniform else 'asym',
                             n_bits=n_bits, opt=opt.name, n_batches=len(data),
                             num_candidates=num_candidates, one_sided=bool(est.one_sided_dist),
                             max_int_skew=int(est.max_int_skew)))

    two = [rnd(50 + i, (2, 8, 96), outlier_dims=(7,)) for i in range(2)]
    one = [np.abs(rnd(60 + i, (2, 8, 96))) for i in range(2)]
    S, A = QMethods.symmetric_uniform, QMethods.asymmetric_uniform
    run('grid_sym_b8', S, 8, two, OptMethod.grid)
    run('grid_sym_b4', S, 4, two, OptMethod.grid)
    run('grid_sym_onesided', S, 8, one, OptMethod.grid)
    run('grid_asym_onesided', A, 8, one, OptMethod.grid)
    run('grid_asym_b8', A, 8, two, OptMethod.grid)
    run('grid_asym_b4', A, 4, two, OptMethod.grid)
    run('grid_asym_b8_c20', A, 8, two, OptMethod.grid, num_candidates=20)
    run('gold_sym_b8', S, 8, two[:1], OptMethod.golden_section)
    run('gold_sym_onesided', S, 8, one[:1], OptMethod.golden_section)
    run('gold_asym_b8', A, 8, two[:1], OptMethod.golden_section)
    np.savez_compressed(os.path.join(HERE, 'mse.npz'), **out)
    return manifest


# --------------------------------------------------------------------------------------
def manager_cases():
    out, manifest = {}, []
    batches = [rnd(70 + i, (4, 16, 96), outlier_dims=(11,), shift=0.2 * i) for i in range(4)]

    def run(name, qm, init, n_bits, axis=None, n_groups=None, init_params=None, per_channel=False,
            data=batches):
        m = QuantizationManager(qmethod=qm, init=init, per_channel=per_channel, axis=axis,
                                n_groups=n_groups, qparams=dict(n_bits=n_bits),
                                init_params=init_params or {})
        for i, b in enumerate(data[:-1]):
            out[f'{name}.y{i}'] = n(m(t(b)))
            out[f'{name}.delta{i}'] = n(m.quantizer._delta).reshape(-1)
        m.fix_ranges()
        out[f'{name}.y_fixed'] = n(m(t(data[-1])))
        for i, b in enumerate(data):
            out[f'{name}.x{i}'] = b
        manifest.append(dict(name=name, kind='sym' if qm == QMethods.symmetric_uniform else 'asym',
                             init=init.name, n_bits=n_bits, axis=axis, n_groups=n_groups,
                             per_channel=per_channel, init_params=init_params or {},
                             n_batches=len(data)))

    S, A, E = QMethods.symmetric_uniform, QMethods.asymmetric_uniform, RangeEstimators
    run('mgr_asym_running', A, E.running_minmax, 8)
    run('mgr_sym_running', S, E.running_minmax, 8)
    run('mgr_asym_current', A, E.current_minmax, 8)
    run('mgr_asym_all_b4', A, E.allminmax, 4)
    run('mgr_asym_running_peg6', A, E.running_minmax, 8, axis=2, n_groups=6)
    run('mgr_asym_current_perembd', A, E.current_minmax, 8, axis=2)
    wb = [rnd(80 + i, (48, 64), scale=0.02) for i in range(2)]
    run('mgr_sym_current_w', S, E.current_minmax, 8, data=wb)
    run('mgr_sym_current_w_pc_b4', S, E.current_minmax, 4, per_channel=True, data=wb)
    np.savez_compressed(os.path.join(HERE, 'manager.npz'), **out)
    return manifest


# --------------------------------------------------------------------------------------
def linear_cases():
    out, manifest = {}, []
    rs = np.random.RandomState(90)

    def run(name, act, n_bits, n_bits_act, in_f=64, out_f=96, on_grid=True):
        lin = QuantLinear(in_f, out_f, bias=True, activation=act,
                          method=QMethods.symmetric_uniform, act_method=QMethods.asymmetric_uniform,
                          n_bits=n_bits, n_bits_act=n_bits_act)
        w = (rs.randn(out_f, in_f) * 0.05).astype(np.float32)
        b = (rs.randn(out_f) * 0.1).astype(np.float32)
        lin.weight.data = t(w)
        lin.bias.data = t(b)
        lin.quantized()
        lin.eval()
        xs = []
        for i in range(3):
            x = (rs.randn(8, 16, in_f) * 2).astype(np.float32)
            if on_grid:  # feed an activation that is itself a QDQ output (integer grid * scale)
                qa = QMethods.asymmetric_uniform.cls(n_bits=8)
                qa.set_quant_range(float(x.min()), float(x.max()))
                out[f'{name}.in_delta{i}'] = n(qa._delta).reshape(-1)
                out[f'{name}.in_zero_float{i}'] = n(qa._zero_float).reshape(-1)
                x = n(qa(t(x)))
            xs.append(x)
        for i in range(2):
            out[f'{name}.y{i}'] = n(lin(t(xs[i])))
        lin.fix_ranges()
        out[f'{name}.y_fixed'] = n(lin(t(xs[2])))
        out[f'{name}.w_q'] = n(lin.cached_params[0])
        out[f'{name}.w_delta'] = n(lin.weight_quantizer.quantizer._delta).reshape(-1)
        out[f'{name}.a_delta'] = n(lin.activation_quantizer.quantizer._delta).reshape(-1)
        out[f'{name}.a_zero_float'] = n(lin.activation_quantizer.quantizer._zero_float).reshape(-1)
        out[f'{name}.w'] = w
        out[f'{name}.b'] = b
        for i in range(3):
            out[f'{name}.x{i}'] = xs[i]
        manifest.append(dict(name=name, act=type(act).__name__ if act else None, n_bits=n_bits,
                             n_bits_act=n_bits_act, in_f=in_f, out_f=out_f, on_grid=on_grid))

    run('lin_w8a8', None, 8, 8)
    run('lin_w8a8_gelu', nn.GELU(), 8, 8)
    run('lin_w4a8_relu', nn.ReLU(), 4, 8)
    run('lin_w8a8_tanh', nn.Tanh(), 8, 8, in_f=96, out_f=96)
    run('lin_w8a8_rawin', None, 8, 8, on_grid=False)
    np.savez_compressed(os.path.join(HERE, 'linear.npz'), **out)
    return manifest


if __name__ == '__main__':
    torch.manual_seed(0)
    manifest = dict(
        reference_commit='8dbf3c64',
        torch=torch.__version__,
        numpy=np.__version__,
        quantizers=quantizer_cases(),
        estimators=estimator_cases(),
        mse=mse_cases(),
        manager=manager_cases(),
        linear=linear_cases(),
    )
    with open(os.path.join(HERE, 'manifest.json'), 'w') as f:
        json.dump(manifest, f, indent=1, default=lambda o: o.tolist() if hasattr(o, 'tolist') else str(o))
    for fn in sorted(os.listdir(HERE)):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))
