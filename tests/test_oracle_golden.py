"""Pins the CPU oracle (oracle/fakequant_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only.

Tolerances: integers, ranges, qparams and dequantised floats must be EXACTLY equal (same IEEE fp32
operation chain).  MSE losses: relative 1e-5 (the reference sums fp32 squared errors in torch's
blocked order, the oracle sums in fp64); the selected candidate / resulting range must be equal.
"""
import numpy as np
import pytest

from conftest import golden_cases
from oracle import fakequant_oracle as O


def eq(a, b):
    a = np.asarray(a, np.float32).reshape(-1)
    b = np.asarray(b, np.float32).reshape(-1)
    assert a.shape == b.shape
    ok = (a == b) | (np.isnan(a) & np.isnan(b))
    assert ok.all(), f'{(~ok).sum()} / {a.size} mismatches; first at {np.argmax(~ok)}: {a[~ok][:4]} vs {b[~ok][:4]}'


@pytest.mark.parametrize('case', golden_cases('quantizers'), ids=lambda c: c['name'])
def test_quantizer(case, golden):
    g = golden.file('quantizers')
    nm = case['name']
    x = g[f'{nm}.x']
    xmin, xmax = g[f'{nm}.xmin'], g[f'{nm}.xmax']
    if not case['vector_range']:
        xmin, xmax = float(xmin), float(xmax)
    nb, dom = case['n_bits'], case['scale_domain']
    if case['kind'] == 'asym':
        delta, zf = O.asym_set_quant_range(xmin, xmax, nb, scale_domain=dom)
        eq(delta, g[f'{nm}.delta'])
        eq(zf, g[f'{nm}.zero_float'])
        eq(O.asym_zero_point(zf, nb), g[f'{nm}.zero_point'])
        eq(O.scale_of(delta, scale_domain=dom), g[f'{nm}.scale'])
        xi = O.qdq_asym(x, delta, zf, nb, scale_domain=dom, axis=case['axis'],
                        per_channel=case['per_channel'], return_int=True)
        y = O.qdq_asym(x, delta, zf, nb, scale_domain=dom, axis=case['axis'],
                       per_channel=case['per_channel'])
    else:
        delta, signed = O.sym_set_quant_range(xmin, xmax, nb, scale_domain=dom)
        eq(delta, g[f'{nm}.delta'])
        assert signed == bool(g[f'{nm}.signed'])
        lo, hi = O.sym_grid(nb, signed)
        assert lo == float(g[f'{nm}.int_min']) and hi == float(g[f'{nm}.int_max'])
        xi = O.qdq_sym(x, delta, signed, nb, scale_domain=dom, per_channel=case['per_channel'],
                       return_int=True)
        y = O.qdq_sym(x, delta, signed, nb, scale_domain=dom, per_channel=case['per_channel'])
    eq(xi, g[f'{nm}.x_int'])
    eq(y, g[f'{nm}.x_quant'])


def _make_est(case):
    kw = dict(per_channel=case.get('per_channel', False), axis=case.get('axis'),
              n_groups=case.get('n_groups'))
    if case['est'] == 'current_minmax':
        return O.CurrentMinMax(**kw)
    if case['est'] == 'running_minmax':
        return O.RunningMinMax(momentum=case['opts'].get('momentum', 0.9), **kw)
    return O.AllMinMax(**kw)


@pytest.mark.parametrize('case', golden_cases('estimators'), ids=lambda c: c['name'])
def test_estimator(case, golden):
    g = golden.file('estimators')
    nm = case['name']
    est = _make_est(case)
    data = [g[f'{nm}.x{i}'] for i in range(case['n_batches'])]
    if case['permute']:
        est.per_group_range_estimation = True
        for b in data:
            est(b)
        eq(est.ranges, g[f'{nm}.ranges'])
        est.per_group_range_estimation = False
    for i, b in enumerate(data):
        mn, mx = est(b)
        eq(mn, g[f'{nm}.b{i}.xmin'])
        eq(mx, g[f'{nm}.b{i}.xmax'])


@pytest.mark.parametrize('case', golden_cases('mse'), ids=lambda c: c['name'])
def test_mse(case, golden):
    g = golden.file('mse')
    nm = case['name']
    cls = O.MSEGrid if case['opt'] == 'grid' else O.MSEGolden
    est = cls(case['n_bits'], case['kind'] == 'sym', num_candidates=case['num_candidates'])
    for i in range(case['n_batches']):
        mn, mx = est(g[f'{nm}.x{i}'])
        if case['opt'] == 'grid':
            ref = g[f'{nm}.b{i}.loss']
            fin = np.isfinite(ref)
            assert (np.isfinite(est.loss_array) == fin).all()
            np.testing.assert_allclose(est.loss_array[fin], ref[fin], rtol=1e-5)
            eq(mn, g[f'{nm}.b{i}.xmin'])
            eq(mx, g[f'{nm}.b{i}.xmax'])
        else:
            # golden section: same scipy, objective differs by fp32-vs-fp64 summation only
            np.testing.assert_allclose(mn, g[f'{nm}.b{i}.xmin'], rtol=2e-3, atol=1e-6)
            np.testing.assert_allclose(mx, g[f'{nm}.b{i}.xmax'], rtol=2e-3, atol=1e-6)
    assert est.one_sided_dist == case['one_sided']
    assert est.max_pos_thr == float(g[f'{nm}.max_pos_thr'])
    assert est.max_neg_thr == float(g[f'{nm}.max_neg_thr'])
