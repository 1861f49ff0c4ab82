"""Property tests: the CPU oracle against the reference ITSELF, imported in place from the reference
checkout and run on random shapes / bit-widths / ranges (hypothesis).  Complements the committed golden
vectors (which travel to the GPU box): here the inputs are not fixed in advance.  Build container only --
skipped where /root/reference does not exist.
"""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import fakequant_oracle as O

REF = os.environ.get('TQ_REFERENCE', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'quantization')),
                                reason='reference checkout not present')


@pytest.fixture(scope='module')
def ref_quantizers():
    """the reference's quantization.quantizers, imported under a private name so that it cannot shadow (or be
    shadowed by) this repo's package of the same name"""
    saved = {k: v for k, v in sys.modules.items() if k == 'quantization' or k.startswith('quantization.')}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        mod = importlib.import_module('quantization.quantizers')
        assert mod.__file__.startswith(REF)
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == 'quantization' or k.startswith('quantization.')]:
            del sys.modules[k]
        sys.modules.update(saved)
    return mod


shapes = st.lists(st.integers(1, 9), min_size=1, max_size=4).map(tuple)
case = st.fixed_dictionaries(dict(
    shape=shapes, n_bits=st.integers(2, 12), asym=st.booleans(), seed=st.integers(0, 2 ** 31 - 1),
    lo=st.floats(-8, 1), hi=st.floats(-1, 8), scale=st.sampled_from([1e-3, 0.1, 1.0, 30.0]),
    log=st.booleans(), layout=st.sampled_from(['tensor', 'axis', 'channel'])))
SETTINGS = dict(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])


def _build(ref, c):
    rs = np.random.RandomState(c['seed'])
    x = (rs.randn(*c['shape']) * 3 * c['scale']).astype(np.float32)
    g = rs.randn(*c['shape']).astype(np.float32)
    cls = ref.AsymmetricUniformQuantizer if c['asym'] else ref.SymmetricUniformQuantizer
    layout = c['layout'] if len(c['shape']) > 1 else 'tensor'
    if layout == 'axis' and not c['asym']:
        layout = 'tensor'                      # the symmetric quantizer has no per-axis mode (quirk A.4-1)
    axis = len(c['shape']) - 1 if layout == 'axis' else None
    q = cls(n_bits=c['n_bits'], scale_domain='log' if c['log'] else 'linear', per_channel=layout == 'channel', axis=axis)
    lo, hi = min(c['lo'], c['hi']) * c['scale'], max(c['lo'], c['hi']) * c['scale']
    if layout == 'tensor':
        q.set_quant_range(lo, hi)
    else:
        k = c['shape'][-1] if layout == 'axis' else c['shape'][0]
        f = (0.5 + rs.rand(k)).astype(np.float32)
        q.set_quant_range(torch.from_numpy(lo * f), torch.from_numpy(hi * f))
    return q, x, g, axis, layout == 'channel'


@settings(**SETTINGS)
@given(c=case)
def test_forward_and_backward_match_reference(ref_quantizers, c):
    q, x, g, axis, per_channel = _build(ref_quantizers, c)
    delta = q._delta.detach().clone().requires_grad_(True)
    q._delta = delta
    zf = None
    if c['asym']:
        zf = q._zero_float.detach().clone().requires_grad_(True)
        q._zero_float = zf
    xt = torch.from_numpy(x).requires_grad_(True)
    y = q(xt)
    y.backward(torch.from_numpy(g))
    dom = 'log' if c['log'] else 'linear'
    d_np = delta.detach().numpy().reshape(-1)
    z_np = zf.detach().numpy().reshape(-1) if zf is not None else None
    signed = bool(q.signed) if not c['asym'] else None
    if c['asym']:
        yo = O.qdq_asym(x, d_np, z_np, c['n_bits'], scale_domain=dom, axis=axis, per_channel=per_channel)
    else:
        yo = O.qdq_sym(x, d_np, signed, c['n_bits'], scale_domain=dom, per_channel=per_channel)
    gx, gd, gz, (mag_s, mag_z) = O.qdq_backward(x, g, d_np, z_np, signed, c['n_bits'], scale_domain=dom, axis=axis,
                                                per_channel=per_channel)
    if c['log']:        # numpy / torch expf differ in the last ulp of the scale: an element that sits exactly on
        # a rounding tie can move by one grid step
        step = np.exp(d_np).max()
        assert np.abs(yo - y.detach().numpy()).max() <= step * 1.0001
        return
    assert np.array_equal(yo, y.detach().numpy())
    assert np.array_equal(gx, xt.grad.numpy())
    tol_s = 4e-6 * np.asarray(mag_s).reshape(-1) + 1e-30
    assert (np.abs(gd - delta.grad.numpy().reshape(-1)) <= tol_s).all()
    if zf is not None:
        tol_z = 4e-6 * np.asarray(mag_z).reshape(-1) + 1e-30
        assert (np.abs(gz - zf.grad.numpy().reshape(-1)) <= tol_z).all()
