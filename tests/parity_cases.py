"""Shared parity checks: the package's reference-facing Python API (quantizers, estimators,
manager, QuantLinear) against the golden vectors produced by the unmodified reference.

Called twice: from tests/test_host_logic.py on CPU tensors with the oracle injected as back-end
(checks the host-side logic), and from tests/test_gpu_parity.py on CUDA tensors through the C ABI
of libtq_b200.so (the parity tests proper).

Tolerances (north_star): integer round/clamp results bit-exact; dequantised floats, ranges and
quantizer parameters equal (identical IEEE fp32 operation chain); MSE losses rtol 1e-5 (fp32
partial sums in a different order); GEMM-based outputs: see check_linear_case.
"""
import numpy as np
import torch
from torch import nn

from quantization.quantizers import QMethods
from quantization.range_estimators import RangeEstimators, OptMethod
from quantization.quantization_manager import QuantizationManager
from quantization.autoquant_utils import QuantLinear


def T(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def N(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def assert_same(a, b, what=''):
    a = np.asarray(N(a), np.float32).reshape(-1)
    b = np.asarray(N(b), np.float32).reshape(-1)
    assert a.shape == b.shape, f'{what}: shape {a.shape} vs {b.shape}'
    ok = (a == b) | (np.isnan(a) & np.isnan(b))
    assert ok.all(), (f'{what}: {(~ok).sum()} / {a.size} mismatches; first at {np.argmax(~ok)}: '
                      f'{a[~ok][:4]} vs {b[~ok][:4]}')


def qcls(kind):
    return QMethods.symmetric_uniform if kind == 'sym' else QMethods.asymmetric_uniform


def check_quantizer_case(case, g, device):
    nm = case['name']
    q = qcls(case['kind']).cls(n_bits=case['n_bits'], scale_domain=case['scale_domain'],
                               per_channel=case['per_channel'], axis=case['axis'])
    xmin, xmax = g[f'{nm}.xmin'], g[f'{nm}.xmax']
    if case['vector_range']:
        q.set_quant_range(T(xmin.astype(np.float32), device), T(xmax.astype(np.float32), device))
    else:
        q.set_quant_range(float(xmin), float(xmax))
    log = case['scale_domain'] == 'log'
    if log:   # logf/expf on the device vs libm: <= 2 ulp
        np.testing.assert_allclose(N(q._delta).reshape(-1), g[f'{nm}.delta'].reshape(-1), rtol=3e-7)
    else:
        assert_same(q._delta, g[f'{nm}.delta'], 'delta')
    if case['kind'] == 'asym':
        assert_same(q._zero_float, g[f'{nm}.zero_float'], 'zero_float')
    else:
        assert bool(q.signed) == bool(g[f'{nm}.signed'])
        assert float(q.int_min) == float(g[f'{nm}.int_min'])
        assert float(q.int_max) == float(g[f'{nm}.int_max'])
    x = T(g[f'{nm}.x'], device)
    y = q(x)
    xi = q.to_integer_forward(x)
    assert y.shape == x.shape and xi.shape == x.shape
    assert y.data_ptr() != x.data_ptr() or x.numel() == 0      # out of place, fresh allocation
    if log:
        # scale may differ by an ulp -> integers can flip at exact ties only; compare loosely
        d = np.abs(N(xi) - g[f'{nm}.x_int'])
        assert (d <= 1).all() and (d > 0).mean() < 1e-3
        return
    assert_same(xi, g[f'{nm}.x_int'], 'x_int')
    assert_same(y, g[f'{nm}.x_quant'], 'x_quant')
    if case['kind'] == 'asym':
        assert_same(q.zero_point, g[f'{nm}.zero_point'], 'zero_point')
    assert_same(q.scale, g[f'{nm}.scale'], 'scale')
    assert_same(q.x_min, g[f'{nm}.q_x_min'], 'x_min')
    assert_same(q.x_max, g[f'{nm}.q_x_max'], 'x_max')


def check_estimator_case(case, g, device):
    nm = case['name']
    qz = QMethods.asymmetric_uniform.cls(n_bits=8)
    kw = {k: case[k] for k in ('axis', 'n_groups', 'per_channel') if k in case}
    est = RangeEstimators[case['est']].cls(quantizer=qz, **kw, **case['opts'])
    data = [T(g[f'{nm}.x{i}'], device) for i in range(case['n_batches'])]
    if case['permute']:
        est.per_group_range_estimation = True
        for b in data:
            assert est(b) is None
        assert_same(est.ranges, g[f'{nm}.ranges'], 'ranges')
        est.per_group_range_estimation = False
    for i, b in enumerate(data):
        mn, mx = est(b)
        assert_same(mn, g[f'{nm}.b{i}.xmin'], f'xmin batch {i}')
        assert_same(mx, g[f'{nm}.b{i}.xmax'], f'xmax batch {i}')
        assert mn.shape == tuple(np.shape(g[f'{nm}.b{i}.xmin'])) or mn.dim() == 0
    est.reset()
    assert est.current_xmin is None and est.current_xmax is None


def check_mse_case(case, g, device):
    nm = case['name']
    qz = qcls(case['kind']).cls(n_bits=case['n_bits'])
    est = RangeEstimators.MSE.cls(quantizer=qz, opt_method=OptMethod[case['opt']],
                                  num_candidates=case['num_candidates'])
    for i in range(case['n_batches']):
        mn, mx = est(T(g[f'{nm}.x{i}'], device))
        if case['opt'] == 'grid':
            ref = g[f'{nm}.b{i}.loss']
            got = est.loss_array
            fin = np.isfinite(ref)
            assert got.shape == ref.shape and (np.isfinite(got) == fin).all()
            np.testing.assert_allclose(got[fin], ref[fin], rtol=1e-5)
            assert_same(mn, g[f'{nm}.b{i}.xmin'], 'mse xmin')
            assert_same(mx, g[f'{nm}.b{i}.xmax'], 'mse xmax')
        else:
            # scipy's bounded minimiser stops on its own x tolerance inside a flat minimum: the argmin is pinned to 2e-3 ...
            np.testing.assert_allclose(N(mn), g[f'{nm}.b{i}.xmin'], rtol=2e-3, atol=1e-6)
            np.testing.assert_allclose(N(mx), g[f'{nm}.b{i}.xmax'], rtol=2e-3, atol=1e-6)
            # ... and the OBJECTIVE at our optimum equals the objective at the reference's optimum to 1e-4 (the objective
            # itself is pinned against the reference's loss table to 1e-5 by the grid cases above)
            x_i = T(g[f'{nm}.x{i}'], device)
            f0 = lambda v: float(np.asarray(N(v) if torch.is_tensor(v) else v, dtype=np.float64).reshape(-1)[0])
            ours = float(est.loss_fx(x_i, f0(mn), f0(mx)))
            ref = float(est.loss_fx(x_i, f0(g[f'{nm}.b{i}.xmin']), f0(g[f'{nm}.b{i}.xmax'])))
            assert abs(ours - ref) <= 1e-4 * max(abs(ref), 1e-12), (ours, ref)
    assert est.one_sided_dist == case['one_sided']
    assert est.max_pos_thr == float(g[f'{nm}.max_pos_thr'])
    assert est.max_neg_thr == float(g[f'{nm}.max_neg_thr'])
    assert est.max_int_skew == case['max_int_skew']


def check_manager_case(case, g, device):
    nm = case['name']
    m = QuantizationManager(qmethod=qcls(case['kind']), init=RangeEstimators[case['init']],
                            per_channel=case['per_channel'], axis=case['axis'], n_groups=case['n_groups'],
                            qparams=dict(n_bits=case['n_bits']), init_params=case['init_params'])
    nb = case['n_batches']
    for i in range(nb - 1):
        y = m(T(g[f'{nm}.x{i}'], device))
        assert_same(m.quantizer._delta, g[f'{nm}.delta{i}'], f'delta batch {i}')
        assert_same(y, g[f'{nm}.y{i}'], f'y batch {i}')
    m.fix_ranges()
    assert m.state.name == 'fix_ranges'
    assert_same(m(T(g[f'{nm}.x{nb - 1}'], device)), g[f'{nm}.y_fixed'], 'y fixed')


_ACTS = {None: None, 'GELU': nn.GELU, 'ReLU': nn.ReLU, 'Tanh': nn.Tanh}


def check_linear_case(case, g, device, exact_gemm):
    """QuantLinear.  The output quantizer rounds the GEMM result, so a last-bit difference between
    two correct fp32 GEMM implementations (summation order) may move an element to the neighbouring
    grid point.  Bar: |y - y_ref| <= one output quantization step everywhere, and fewer than 0.2 %
    of the elements differ at all.  With ``exact_gemm`` (CPU: same torch GEMM as the reference) the
    result must be identical."""
    nm = case['name']
    act = _ACTS[case['act']]
    lin = QuantLinear(case['in_f'], case['out_f'], bias=True, activation=act() if act else None,
                      method=QMethods.symmetric_uniform, act_method=QMethods.asymmetric_uniform,
                      n_bits=case['n_bits'], n_bits_act=case['n_bits_act'])
    lin.weight.data = T(g[f'{nm}.w'], 'cpu')
    lin.bias.data = T(g[f'{nm}.b'], 'cpu')
    lin.to(device)
    lin.quantized()
    lin.eval()

    def cmp(y, ref, step):
        y, ref = N(y), np.asarray(ref)
        assert y.shape == ref.shape
        if exact_gemm:
            assert_same(y, ref, 'linear out')
            return
        d = np.abs(y - ref)
        assert d.max() <= step * 1.001, f'max diff {d.max()} > step {step}'
        assert (d > step * 1e-3).mean() < 2e-3, f'flip rate {(d > step * 1e-3).mean()}'

    for i in range(2):
        y = lin(T(g[f'{nm}.x{i}'], device))
        step = float(N(lin.activation_quantizer.quantizer._delta).reshape(-1)[0])
        cmp(y, g[f'{nm}.y{i}'], step)
    lin.fix_ranges()
    assert_same(lin.weight_quantizer.quantizer._delta, g[f'{nm}.w_delta'], 'w delta')
    y = lin(T(g[f'{nm}.x2'], device))
    assert_same(lin.cached_params[0], g[f'{nm}.w_q'], 'cached quantized weight')
    cmp(y, g[f'{nm}.y_fixed'], float(g[f'{nm}.a_delta'][0]))
    np.testing.assert_allclose(N(lin.activation_quantizer.quantizer._delta).reshape(-1), g[f'{nm}.a_delta'],
                               rtol=0 if exact_gemm else 1e-5)
