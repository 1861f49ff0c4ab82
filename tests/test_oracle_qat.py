"""Pins the training-path part of the CPU oracle (straight-through backward with learnable ranges,
AdaRound soft rounding) against the reference under torch autograd (tests/golden/qat.npz, produced
by tests/golden/make_golden_qat.py).  CPU only.

Tolerances: grad_x is an elementwise fp32 chain -> EXACTLY equal (scale_domain='log': rtol 1e-6, the
scale itself is a libm expf).  grad_delta / grad_zero_float are
sums of N fp32 terms that nearly cancel; torch sums fp32 in a blocked order, the oracle in fp64 ->
|diff| <= 2e-6 * sum|terms| (stated in qat_cases.close_sum).  AdaRound: transcendental functions
(sigmoid / log) differ by a few ulp between numpy and torch -> rtol 2e-6 / atol 1e-6 on alpha and the
soft targets; hard targets (integers) exactly equal.
"""
import numpy as np
import pytest

from oracle import fakequant_oracle as O
from qat_cases import (QAT_MANIFEST, adaround_case_arrays, adaround_grid, backward_case_arrays, close_sum,
                       oracle_backward)


@pytest.mark.parametrize('case', QAT_MANIFEST['backward'], ids=lambda c: c['name'])
def test_backward(case):
    a = backward_case_arrays(case)
    gx, gd, gz, (mag_s, mag_z) = oracle_backward(case, a)
    if case['scale_domain'] == 'log':       # numpy / torch expf differ in the last ulp of the scale
        np.testing.assert_allclose(gx, a['grad_x'], rtol=1e-6, atol=0)
    else:
        assert np.array_equal(gx, a['grad_x']), 'grad_x differs from the reference'
    close_sum(gd, a['grad_delta'], mag_s, 'grad_delta')
    if case['kind'] == 'asym':
        close_sum(gz, a['grad_zero_float'], mag_z, 'grad_zero_float')
    else:
        assert gz is None


def test_backward_masks():
    """the cases built to hit the clamp masks do hit them"""
    by = {c['name']: c for c in QAT_MANIFEST['backward']}
    a = backward_case_arrays(by['bw_asym_t_tiny'])
    assert (a['grad_delta'] == 0).all() and (a['delta'] < 1e-8).all()
    for nm in ('bw_asym_t_zf_out', 'bw_asym_t_zf_neg'):
        a = backward_case_arrays(by[nm])
        assert (a['grad_zero_float'] == 0).all()
    a = backward_case_arrays(by['bw_asym_t_clip'])
    assert (a['grad_x'] == 0).any() and (a['grad_x'] != 0).any() and a['grad_zero_float'][0] != 0


@pytest.mark.parametrize('case', QAT_MANIFEST['adaround'], ids=lambda c: c['name'])
def test_adaround(case):
    a = adaround_case_arrays(case)
    scale, zp, lo, hi = adaround_grid(case, a)
    mode, temp = case['mode'], case['temperature']
    w = a['w']
    alpha0 = O.adaround_alpha_init(w, scale, mode, temp)
    np.testing.assert_allclose(alpha0, a['alpha0'], rtol=2e-5, atol=2e-5)
    y0 = O.adaround_qdq(w, a['alpha0'], scale, zp, lo, hi, mode, True, temp)
    step = float(np.max(scale))
    np.testing.assert_allclose(y0, a['y_soft0'], rtol=0, atol=2e-6 * step * max(abs(lo), hi))
    y1 = O.adaround_qdq(w, a['alpha1'], scale, zp, lo, hi, mode, True, temp)
    np.testing.assert_allclose(y1, a['y_soft1'], rtol=0, atol=2e-6 * step * max(abs(lo), hi))
    xi, _ = O.adaround_to_integer(w, a['alpha1'], scale, zp, lo, hi, mode, False, temp)
    assert np.array_equal(xi, a['x_int_hard1'])
    assert np.array_equal(O.dequantize(xi, scale, zp), a['y_hard1'])
    ga = O.adaround_grad_alpha(w, a['alpha1'], a['g'], scale, zp, lo, hi, mode, temp)
    np.testing.assert_allclose(ga, a['grad_alpha1'], rtol=2e-5, atol=1e-7 * step)
