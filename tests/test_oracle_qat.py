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


@pytest.mark.parametrize('permute', [False, True])
def test_peg_grouped_integer_gemm_is_the_dequantized_matmul(permute):
    """The grouped integer GEMM planned for PEG activations in the fused engine (oracle.peg_linear_exact) equals
    the reference's formulation -- F.linear on the dequantized tensors -- evaluated in float64, to float64
    round-off: per-group integer accumulators + one scale per group lose nothing (the reference's fp32 GEMM on
    dequantized values is the less exact of the two)."""
    rs = np.random.RandomState(7)
    M, K, N, G = 24, 96, 20, 6
    x = (rs.randn(M, K) * (0.5 + 3 * rs.rand(K))).astype(np.float32)
    mn, mx = O.minmax_axis(x.reshape(1, M, K), 2)
    order = O.stable_order((mx - mn).astype(np.float32)) if permute else None
    gmn, gmx = O.group_minmax(mn, mx, G, order)
    delta, zf = O.asym_set_quant_range(gmn, gmx, 8)
    scale, zp = O.scale_of(delta), O.asym_zero_point(zf, 8)
    x_int = O.qdq_asym(x.reshape(1, M, K), delta, zf, 8, axis=2, return_int=True).reshape(M, K)
    w = (rs.randn(N, K) * 0.05).astype(np.float32)
    wd, signed = O.sym_set_quant_range(w.min(), w.max(), 8)
    w_int = O.qdq_sym(w, wd, signed, 8, return_int=True)
    bias = rs.randn(N).astype(np.float32)
    y = O.peg_linear_exact(x_int, zp, scale, w_int, np.full(N, O.scale_of(wd)), bias, G, order)
    xq = O.dequantize(x_int, scale.reshape(1, K), zp.reshape(1, K)).astype(np.float64)
    wq = O.dequantize(w_int, O.scale_of(wd), np.float32(0)).astype(np.float64)
    ref = xq @ wq.T + bias.astype(np.float64)
    np.testing.assert_allclose(y, ref, rtol=1e-6, atol=1e-6 * np.abs(ref).max())
