"""Shared helpers for the training-path tests (golden file tests/golden/qat.npz)."""
import json
import os

import numpy as np

from oracle import fakequant_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
with open(os.path.join(GOLDEN, 'qat_manifest.json')) as f:
    QAT_MANIFEST = json.load(f)
_QAT = None


def qat_file():
    global _QAT
    if _QAT is None:
        _QAT = np.load(os.path.join(GOLDEN, 'qat.npz'))
    return _QAT


def backward_case_arrays(case):
    """-> dict(x, g, delta, zero_float | None, signed | None, golden grads) of one backward case"""
    g, nm = qat_file(), case['name']
    d = dict(x=g[f'{nm}.x'], g=g[f'{nm}.g'], delta=g[f'{nm}.delta'], y=g[f'{nm}.y'], grad_x=g[f'{nm}.grad_x'],
             grad_delta=g[f'{nm}.grad_delta'], zero_float=None, signed=None, grad_zero_float=None)
    if case['kind'] == 'asym':
        d['zero_float'] = g[f'{nm}.zero_float']
        d['grad_zero_float'] = g[f'{nm}.grad_zero_float']
    else:
        d['signed'] = bool(g[f'{nm}.signed'])
    return d


def oracle_backward(case, a):
    return O.qdq_backward(a['x'], a['g'], a['delta'], a['zero_float'], a['signed'], case['n_bits'],
                          scale_domain=case['scale_domain'], axis=case['axis'], per_channel=case['per_channel'])


def close_sum(got, want, mag, what, rtol=2e-6):
    """|got - want| <= rtol * sum|terms|: the fp32 summation-order error bound of a reduced gradient.
    (`within 1e-5 relative` of the north star, with the magnitude of the summed terms as the base --
    the sums themselves cancel to near zero.)"""
    got, want = np.asarray(got, np.float64).reshape(-1), np.asarray(want, np.float64).reshape(-1)
    tol = rtol * np.asarray(mag, np.float64).reshape(-1) + 1e-30
    bad = np.abs(got - want) > tol
    assert not bad.any(), f'{what}: {bad.sum()} / {got.size} outside tolerance; ' \
                          f'got {got[bad][:4]} want {want[bad][:4]} tol {tol[bad][:4]}'


def adaround_case_arrays(case):
    g, nm = qat_file(), case['name']
    keys = ('y_nearest', 'alpha0', 'y_soft0', 'alpha1', 'y_soft1', 'grad_alpha1', 'y_hard1', 'x_int_hard1', 'delta')
    d = {k: g[f'{nm}.{k}'] for k in keys}
    d['w'], d['g'] = g['ada.w'], g['ada.g']
    d['zero_float'] = g[f'{nm}.zero_float'] if case['kind'] == 'asym' else None
    d['signed'] = bool(g[f'{nm}.signed']) if case['kind'] == 'sym' else None
    return d


def adaround_grid(case, a):
    """(scale, zp, lo, hi) broadcastable against the weight"""
    scale = O.scale_of(a['delta'])
    if case['kind'] == 'asym':
        zp = O.asym_zero_point(a['zero_float'], case['n_bits'])
        lo, hi = 0.0, O.asym_int_max(case['n_bits'])
    else:
        zp = np.zeros_like(scale)
        lo, hi = O.sym_grid(case['n_bits'], a['signed'])
    if case['per_channel']:
        scale, zp = scale.reshape(-1, 1), zp.reshape(-1, 1)
    else:
        scale, zp = scale.reshape(()), zp.reshape(())
    return scale, zp, lo, hi
