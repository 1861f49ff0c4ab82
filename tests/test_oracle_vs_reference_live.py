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

from reference_path import reference_root
REF = reference_root()
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
# derandomize: the same examples on every run (a test-suite that draws new inputs each time can turn red by chance)
# (exploration: TQ_HYPOTHESIS_RANDOM=1 TQ_HYPOTHESIS_EXAMPLES=5000 pytest tests/test_oracle_vs_reference_live.py)
_RANDOM = os.environ.get('TQ_HYPOTHESIS_RANDOM') == '1'
SETTINGS = dict(max_examples=int(os.environ.get('TQ_HYPOTHESIS_EXAMPLES', 150)), deadline=None,
                derandomize=not _RANDOM, database=None, suppress_health_check=[HealthCheck.function_scoped_fixture])


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


@pytest.fixture(scope='module')
def ref_estimators(ref_quantizers):
    """the reference's quantization.range_estimators (imported like ref_quantizers, private copy)"""
    saved = {k: v for k, v in sys.modules.items() if k == 'quantization' or k.startswith('quantization.')}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        mod = importlib.import_module('quantization.range_estimators')
        qmod = importlib.import_module('quantization.quantizers')
        assert mod.__file__.startswith(REF)
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == 'quantization' or k.startswith('quantization.')]:
            del sys.modules[k]
        sys.modules.update(saved)
    return mod, qmod


est_case = st.fixed_dictionaries(dict(
    B=st.integers(1, 4), T=st.integers(1, 6), groups=st.sampled_from([1, 2, 3, 4, 6]), per_group=st.integers(1, 4),
    seed=st.integers(0, 2 ** 31 - 1), kind=st.sampled_from(['current', 'running']), n_batches=st.integers(1, 3),
    mode=st.sampled_from(['tensor', 'axis', 'groups', 'groups_permuted', 'channel']), momentum=st.sampled_from([0.9, 0.5])))


@settings(**SETTINGS)
@given(c=est_case)
def test_minmax_estimators_match_reference(ref_estimators, c):
    R, Q = ref_estimators
    d = c['groups'] * c['per_group']
    rs = np.random.RandomState(c['seed'])
    scales = (0.5 + 3 * rs.rand(d)).astype(np.float32)          # distinct per-dim ranges: a well-defined permutation
    batches = [(rs.randn(c['B'], c['T'], d) * scales).astype(np.float32) for _ in range(c['n_batches'])]
    mode = c['mode']
    if c['kind'] == 'running' and mode == 'groups_permuted':
        mode = 'groups'                       # the running estimator has no permutation (quirk A.4-7)
    if mode == 'groups_permuted' and c['B'] * c['T'] < 2:
        mode = 'groups'                       # one sample per dim: every range is 0, the permutation is all ties and
        #                                       torch.argsort leaves their order implementation-defined (quirk A.4-11)
    kw = dict(axis=2 if mode in ('axis', 'groups', 'groups_permuted') else None,
              n_groups=c['groups'] if mode.startswith('groups') else None, per_channel=mode == 'channel')
    qz = Q.AsymmetricUniformQuantizer(n_bits=8)
    if c['kind'] == 'current':
        ref = R.CurrentMinMaxEstimator(quantizer=qz, **kw)
        mine = O.CurrentMinMax(**kw)
    else:
        ref = R.RunningMinMaxEstimator(quantizer=qz, momentum=c['momentum'], **kw)
        mine = O.RunningMinMax(momentum=c['momentum'], **kw)
    if mode == 'groups_permuted':
        ref.per_group_range_estimation = mine.per_group_range_estimation = True
        for b in batches:
            ref(torch.from_numpy(b))
            mine(b)
        assert np.array_equal(ref.ranges.numpy(), mine.ranges)
        ref.per_group_range_estimation = mine.per_group_range_estimation = False
    for b in batches:
        rmn, rmx = ref(torch.from_numpy(b))
        omn, omx = mine(b)
        assert np.array_equal(np.asarray(rmn.numpy(), np.float32).reshape(-1), np.asarray(omn, np.float32).reshape(-1))
        assert np.array_equal(np.asarray(rmx.numpy(), np.float32).reshape(-1), np.asarray(omx, np.float32).reshape(-1))


mse_case = st.fixed_dictionaries(dict(
    seed=st.integers(0, 2 ** 31 - 1), n=st.integers(16, 400), n_bits=st.sampled_from([2, 4, 8]), sym=st.booleans(),
    one_sided=st.booleans(), cands=st.sampled_from([5, 20]), n_batches=st.integers(1, 2)))


@settings(**dict(SETTINGS, max_examples=max(40, SETTINGS['max_examples'] // 4)))
@given(c=mse_case)
def test_mse_grid_matches_reference(ref_estimators, c):
    R, Q = ref_estimators
    rs = np.random.RandomState(c['seed'])
    batches = [(rs.randn(c['n']) * 2).astype(np.float32) for _ in range(c['n_batches'])]
    if c['one_sided']:
        batches = [np.abs(b) for b in batches]
    qz = (Q.SymmetricUniformQuantizer if c['sym'] else Q.AsymmetricUniformQuantizer)(n_bits=c['n_bits'])
    ref = R.MSE_Estimator(quantizer=qz, opt_method=R.OptMethod.grid, num_candidates=c['cands'])
    mine = O.MSEGrid(c['n_bits'], c['sym'], num_candidates=c['cands'])
    for b in batches:
        rmn, rmx = ref(torch.from_numpy(b))
        omn, omx = mine(b)
    la, lb = np.asarray(ref.loss_array, np.float64), np.asarray(mine.loss_array, np.float64)
    assert la.shape == lb.shape
    fin = np.isfinite(la)
    assert np.array_equal(fin, np.isfinite(lb))
    np.testing.assert_allclose(lb[fin], la[fin], rtol=2e-5)
    # the selected range is the argmin of those losses: equal unless two candidates tie within the summation tolerance
    srt = np.sort(la[fin])
    if len(srt) < 2 or srt[1] - srt[0] > 1e-4 * max(srt[0], 1e-12):
        assert np.array_equal(np.asarray(rmn).reshape(-1), np.asarray(omn).reshape(-1))
        assert np.array_equal(np.asarray(rmx).reshape(-1), np.asarray(omx).reshape(-1))


pct_case = st.fixed_dictionaries(dict(
    seed=st.integers(0, 2 ** 31 - 1), rows=st.integers(1, 6), cols=st.integers(1, 300), per_channel=st.booleans(),
    pct=st.sampled_from([0.001, 0.01, 0.1, 1.0, 5.0, 25.0, 50.0])))


@settings(**SETTINGS)
@given(c=pct_case)
def test_percentile_estimator_matches_reference(ref_estimators, c, monkeypatch):
    """CurrentMinMaxEstimator(percentile=p): the reference calls np.percentile on the host
    (range_estimators.py:121-127, 133-140; per-tensor: (p, 100), per-channel: (p, 100 - p)); this package's class
    evaluates the same interpolation where the tensor lives.  Whole estimator forward, reference in place vs this
    package's class (oracle back-end for its kernels) -- equal."""
    import tq_native
    from oracle_backend import OracleOps
    monkeypatch.setattr(tq_native, '_OPS', OracleOps())
    monkeypatch.setattr(tq_native, 'default_device', lambda: torch.device('cpu'))
    from quantization.range_estimators import CurrentMinMaxEstimator as Mine
    R, Q = ref_estimators
    rs = np.random.RandomState(c['seed'])
    x = (rs.randn(c['rows'], c['cols']) * (1 + 3 * rs.rand())).astype(np.float32)
    ref = R.CurrentMinMaxEstimator(quantizer=Q.AsymmetricUniformQuantizer(n_bits=8), per_channel=c['per_channel'],
                                   percentile=c['pct'])
    mine = Mine(per_channel=c['per_channel'], percentile=c['pct'])
    rmn, rmx = ref(torch.from_numpy(x))
    omn, omx = mine(torch.from_numpy(x))
    assert np.array_equal(rmn.numpy().reshape(-1), omn.numpy().reshape(-1))
    assert np.array_equal(rmx.numpy().reshape(-1), omx.numpy().reshape(-1))


# ---- model conversion (quantize_model / quantize_sequential / quantize_module_list) ---------------------------
@pytest.fixture(scope='module')
def ref_autoquant(ref_quantizers):
    saved = {k: v for k, v in sys.modules.items() if k == 'quantization' or k.startswith('quantization.')}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        mod = importlib.import_module('quantization.autoquant_utils')
        qmod = importlib.import_module('quantization.quantizers')
        assert mod.__file__.startswith(REF)
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == 'quantization' or k.startswith('quantization.')]:
            del sys.modules[k]
        sys.modules.update(saved)
    return mod, qmod


def _nets():
    from torch import nn

    class Block(nn.Module):                      # an unknown container: converted child by child
        def __init__(self):
            super().__init__()
            self.emb = nn.Embedding(20, 8)
            self.body = nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 16), nn.GELU(), nn.LayerNorm(16))
            self.heads = nn.ModuleList([nn.Linear(16, 4), nn.Linear(16, 4)])

        def forward(self, ids):
            h = self.body(self.emb(ids))
            return self.heads[0](h) + self.heads[1](h)

    torch.manual_seed(11)
    return {
        'mlp': (nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 16), nn.GELU(), nn.LayerNorm(16),
                              nn.Linear(16, 4)), 'float'),
        # the first Linear absorbs the Tanh two positions later and the walk resumes AFTER the next position:
        # the second Linear is skipped (reference autoquant_utils.py:94-105, 143-150 -- quirk A.4-10)
        'skip': (nn.Sequential(nn.Linear(8, 8), nn.Linear(8, 8), nn.Tanh(), nn.Linear(8, 4)), 'float'),
        'pool_tied': (nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.AdaptiveAvgPool1d(4)), 'float'),
        'block': (Block(), 'ids'),
    }


def _tree(m):
    return [(n, type(s).__name__) for n, s in m.named_modules()]


@pytest.mark.parametrize('name', ['mlp', 'skip', 'pool_tied', 'block'])
@pytest.mark.parametrize('tie', [False, True])
def test_model_conversion_matches_reference(ref_autoquant, name, tie, monkeypatch):
    import copy
    import tq_native
    from oracle_backend import OracleOps
    monkeypatch.setattr(tq_native, '_OPS', OracleOps())
    monkeypatch.setattr(tq_native, 'default_device', lambda: torch.device('cpu'))
    import quantization.autoquant_utils as mine_mod
    from quantization.quantizers import QMethods as MineQ
    R, RQ = ref_autoquant
    net, kind = _nets()[name]
    g = torch.Generator().manual_seed(5)
    xs = [torch.randint(0, 20, (3, 6), generator=g) if kind == 'ids' else torch.randn(3, 6, 8, generator=g)
          for _ in range(3)]
    outs = []
    for mod, Q in ((R, RQ.QMethods), (mine_mod, MineQ)):
        qnet = mod.quantize_model(copy.deepcopy(net), tie_activation_quantizers=tie, method=Q.symmetric_uniform,
                                  act_method=Q.asymmetric_uniform, n_bits=4, n_bits_act=8)
        qnet.eval()
        for m in qnet.modules():
            if hasattr(m, 'quantized'):
                m.quantized()
        with torch.no_grad():
            ys = [qnet(x) for x in xs[:2]]
            for m in qnet.modules():
                if hasattr(m, 'fix_ranges') and hasattr(m, 'quantized'):
                    m.fix_ranges()
            ys.append(qnet(xs[2]))
        outs.append((_tree(qnet), [y.numpy().copy() for y in ys]))
    (tree_r, ys_r), (tree_m, ys_m) = outs
    assert tree_r == tree_m                              # same module names, same class names
    for a, b in zip(ys_r, ys_m):
        assert np.array_equal(a, b)


@pytest.mark.parametrize('layer', ['linear_relu', 'layernorm', 'embedding'])
def test_hijacked_layers_match_reference(ref_autoquant, layer, monkeypatch):
    """QuantLinear / QuantLayerNorm / QuantEmbedding through their life cycle (reference hijacker.py:66-116):
    estimate -> fix -> eval with the weight cache, activation dumps (``activation_save_target``), cache reset on
    train() and with ``caching = False``, weight / activation switches.  Outputs and dumps equal."""
    import tq_native
    from torch import nn
    from oracle_backend import OracleOps
    monkeypatch.setattr(tq_native, '_OPS', OracleOps())
    monkeypatch.setattr(tq_native, 'default_device', lambda: torch.device('cpu'))
    import quantization.autoquant_utils as mine_mod
    from quantization.quantizers import QMethods as MineQ
    R, RQ = ref_autoquant
    g = torch.Generator().manual_seed(9)
    if layer == 'embedding':
        xs = [torch.randint(0, 30, (4, 7), generator=g) for _ in range(4)]
    else:
        xs = [torch.randn(4, 7, 12, generator=g) * (1 + i) for i in range(4)]
    results = []
    for mod, Q in ((R, RQ.QMethods), (mine_mod, MineQ)):
        torch.manual_seed(21)
        kw = dict(method=Q.symmetric_uniform, act_method=Q.asymmetric_uniform, n_bits=4, n_bits_act=8)
        if layer == 'linear_relu':
            m = mod.QuantLinear(12, 10, activation=nn.ReLU(), **kw)
        elif layer == 'layernorm':
            m = mod.QuantLayerNorm(12, **kw)
            m.weight.data = 1 + 0.1 * torch.randn(12, generator=torch.Generator().manual_seed(3))
        else:
            m = mod.QuantEmbedding(30, 12, **kw)
        m.eval()
        log = []
        with torch.no_grad():
            log.append(m(xs[0]))                                   # FP32
            m.quantized()
            log.append(m(xs[0]))                                   # estimate ranges
            log.append(m(xs[1]))
            m.fix_ranges()
            dump = {}
            m.activation_save_target, m.activation_save_name = dump, 'site'
            log.append(m(xs[2]))
            m.activation_save_target = None
            cached = m.cached_params is not None
            m.train()
            assert m.cached_params is None
            m.eval()
            m.caching = False
            log.append(m(xs[3]))
            assert m.cached_params is None
            m.full_precision_acts()
            log.append(m(xs[3]))
            m.quantized_acts()
            m.full_precision_weights()
            log.append(m(xs[3]))
        results.append(([y.numpy().copy() for y in log], {k: np.array(v) for k, v in dump.items()}, cached))
    (ys_r, dump_r, c_r), (ys_m, dump_m, c_m) = results
    assert c_r == c_m and set(dump_r) == set(dump_m) and dump_r
    for k in dump_r:
        assert np.array_equal(dump_r[k], dump_m[k]), k
    for i, (a, b) in enumerate(zip(ys_r, ys_m)):
        assert np.array_equal(a, b), f'step {i}'


def test_error_behaviour_matches_reference(ref_estimators, monkeypatch):
    """the exceptions SURVEY.md section 8(b) lists: same exception class (by name) from the reference and from this
    package for the same misuse"""
    import tq_native
    from oracle_backend import OracleOps
    monkeypatch.setattr(tq_native, '_OPS', OracleOps())
    monkeypatch.setattr(tq_native, 'default_device', lambda: torch.device('cpu'))
    import quantization.quantization_manager as mine_mgr
    import quantization.quantizers as mine_q
    import quantization.range_estimators as mine_est
    R, RQ = ref_estimators
    saved = {k: v for k, v in sys.modules.items() if k == 'quantization' or k.startswith('quantization.')}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        ref_mgr = importlib.import_module('quantization.quantization_manager')
        ref_q = importlib.import_module('quantization.quantizers')
        ref_est = importlib.import_module('quantization.range_estimators')
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == 'quantization' or k.startswith('quantization.')]:
            del sys.modules[k]
        sys.modules.update(saved)
    x = torch.randn(2, 3, 12, generator=torch.Generator().manual_seed(0))

    def cases(Q, E, M):
        A, S = Q.QMethods.asymmetric_uniform, Q.QMethods.symmetric_uniform
        yield 'delta before init', lambda: A.cls(n_bits=8).delta
        yield 'zero_float before init', lambda: A.cls(n_bits=8).zero_float
        yield 'signed before init', lambda: S.cls(n_bits=8).signed
        yield 'fix_ranges before init', lambda: M.QuantizationManager(qmethod=A, qparams=dict(n_bits=8)).fix_ranges()
        yield 'vector range on per-tensor quantizer', lambda: A.cls(n_bits=8).set_quant_range(
            torch.tensor([-1.0, -2.0]), torch.tensor([1.0, 2.0]))
        yield 'groups do not divide the dim', lambda: M.QuantizationManager(
            qmethod=A, init=E.RangeEstimators.current_minmax, axis=2, n_groups=5, qparams=dict(n_bits=8))(x)
        yield 'MSE without data', lambda: E.RangeEstimators.MSE.cls(quantizer=A.cls(n_bits=8)).optimization_method
        yield 'MSE without quantizer', lambda: E.RangeEstimators.MSE.cls()

    for (what, ref_call), (_, mine_call) in zip(cases(ref_q, ref_est, ref_mgr), cases(mine_q, mine_est, mine_mgr)):
        names = []
        for call in (ref_call, mine_call):
            try:
                call()
                names.append(None)
            except Exception as e:                         # noqa: BLE001
                names.append(type(e).__name__)
        assert names[0] is not None, f'{what}: the reference raised nothing'
        assert names[0] == names[1], f'{what}: reference {names[0]}, this package {names[1]}'


def test_calibration_loop_and_model_switches_match_reference(ref_autoquant, monkeypatch):
    """utils.pass_data_for_range_estimation (reference utils/utils.py:47-79) + the QuantizedModel bulk switches
    (base_quantized_model.py:15-113) on a small model: batches consumed, quantizer states and outputs equal."""
    import importlib.util
    import types
    import tq_native
    from torch import nn
    from oracle_backend import OracleOps
    monkeypatch.setattr(tq_native, '_OPS', OracleOps())
    monkeypatch.setattr(tq_native, 'default_device', lambda: torch.device('cpu'))
    import quantization.autoquant_utils as mine_auto
    import quantization.base_quantized_model as mine_model
    import quantization.quantization_manager as mine_mgr
    from quantization.quantizers import QMethods as MineQ
    from utils.utils import pass_data_for_range_estimation as mine_pass
    R, RQ = ref_autoquant
    # reference side: base_quantized_model + utils/utils.py loaded from the checkout (utils/__init__ needs the CLI stack)
    saved = {k: v for k, v in sys.modules.items() if k.split('.')[0] in ('quantization', 'utils')}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        ref_model = importlib.import_module('quantization.base_quantized_model')
        ref_mgr = importlib.import_module('quantization.quantization_manager')
        pkg = types.ModuleType('utils')
        pkg.__path__ = []
        sys.modules['utils'] = pkg
        spec = importlib.util.spec_from_file_location('utils.utils', os.path.join(REF, 'utils', 'utils.py'))
        ref_utils = importlib.util.module_from_spec(spec)
        sys.modules['utils.utils'] = ref_utils
        spec.loader.exec_module(ref_utils)
        ref_auto = importlib.import_module('quantization.autoquant_utils')
        ref_Q = importlib.import_module('quantization.quantizers').QMethods
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k.split('.')[0] in ('quantization', 'utils')]:
            del sys.modules[k]
        sys.modules.update(saved)

    def make(auto, model_mod, Q):
        class Net(model_mod.QuantizedModel):
            def __init__(self):
                super().__init__()
                kw = dict(method=Q.symmetric_uniform, act_method=Q.asymmetric_uniform, n_bits=8, n_bits_act=8)
                self.fc1 = auto.QuantLinear(6, 10, activation=nn.GELU(), **kw)
                self.fc2 = auto.QuantLinear(10, 3, **kw)

            def forward(self, x):
                return self.fc2(self.fc1(x))
        torch.manual_seed(31)
        return Net()

    g = torch.Generator().manual_seed(2)
    loader = [(torch.randn(5, 6, generator=g) * (i + 1), torch.zeros(5)) for i in range(4)]
    outs = []
    for auto, model_mod, mgr_mod, Q, pass_fn in ((ref_auto, ref_model, ref_mgr, ref_Q, ref_utils.pass_data_for_range_estimation),
                                                 (mine_auto, mine_model, mine_mgr, MineQ, mine_pass)):
        net = make(auto, model_mod, Q)
        with torch.no_grad():
            pass_fn(loader=loader, model=net, act_quant=True, weight_quant=True, max_num_batches=3)
            states0 = [m.state.name for m in net.modules() if isinstance(m, mgr_mod.QuantizationManager)]
            net.fix_act_ranges()
            states1 = [m.state.name for m in net.modules() if isinstance(m, mgr_mod.QuantizationManager)]
            y_q = net(loader[3][0])
            net.full_precision_acts()
            y_w = net(loader[3][0])
            net.set_quant_state(weight_quant=False, act_quant=True)
            y_a = net(loader[3][0])
            net.reset_act_ranges()
            states2 = [m.state.name for m in net.modules() if isinstance(m, mgr_mod.QuantizationManager)]
        assert not net.training
        outs.append((states0, states1, states2, [t.numpy().copy() for t in (y_q, y_w, y_a)]))
    assert outs[0][:3] == outs[1][:3]
    for a, b in zip(outs[0][3], outs[1][3]):
        assert np.array_equal(a, b)


def test_state_dicts_interchange_with_the_reference(ref_estimators, monkeypatch):
    """calibrated QuantizationManagers expose the same state_dict (keys, dtypes, shapes, values) as the reference's:
    checkpoints written by one side load on the other"""
    import tq_native
    from oracle_backend import OracleOps
    monkeypatch.setattr(tq_native, '_OPS', OracleOps())
    monkeypatch.setattr(tq_native, 'default_device', lambda: torch.device('cpu'))
    import quantization.quantization_manager as mine_mgr
    import quantization.quantizers as mine_q
    import quantization.range_estimators as mine_est
    saved = {k: v for k, v in sys.modules.items() if k == 'quantization' or k.startswith('quantization.')}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        ref_mgr = importlib.import_module('quantization.quantization_manager')
        ref_q = importlib.import_module('quantization.quantizers')
        ref_est = importlib.import_module('quantization.range_estimators')
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == 'quantization' or k.startswith('quantization.')]:
            del sys.modules[k]
        sys.modules.update(saved)
    g = torch.Generator().manual_seed(4)
    x3 = torch.randn(3, 5, 12, generator=g)
    w2 = torch.randn(6, 10, generator=g) * 0.1
    cases = [('asym tensor', 'asymmetric_uniform', 'running_minmax', {}, x3),
             ('sym tensor', 'symmetric_uniform', 'current_minmax', {}, x3),
             ('asym per-embedding groups', 'asymmetric_uniform', 'current_minmax', dict(axis=2, n_groups=3), x3),
             ('sym per-channel', 'symmetric_uniform', 'current_minmax', dict(per_channel=True), w2),
             ('asym per-channel', 'asymmetric_uniform', 'allminmax', dict(per_channel=True), w2)]
    for what, qm, est, kw, data in cases:
        sds = []
        for M, Q, E in ((ref_mgr, ref_q, ref_est), (mine_mgr, mine_q, mine_est)):
            m = M.QuantizationManager(qmethod=Q.QMethods[qm], init=E.RangeEstimators[est], qparams=dict(n_bits=8), **kw)
            with torch.no_grad():
                m(data)
                m(data * 1.5)
            sds.append(m.state_dict())
        ref_sd, mine_sd = sds
        assert list(ref_sd) == list(mine_sd), what
        for k in ref_sd:
            a, b = ref_sd[k], mine_sd[k]
            assert a.dtype == b.dtype and tuple(a.shape) == tuple(b.shape), f'{what}: {k}'
            assert torch.equal(a, b), f'{what}: {k}'


mgr_case = st.fixed_dictionaries(dict(
    seed=st.integers(0, 2 ** 31 - 1), asym=st.booleans(), n_bits=st.sampled_from([2, 4, 8, 16]),
    est=st.sampled_from(['current_minmax', 'running_minmax', 'allminmax']), layout=st.sampled_from(['tensor', 'axis', 'groups', 'channel']),
    groups=st.sampled_from([1, 2, 4]), per_group=st.integers(1, 3), rows=st.integers(2, 5), n_batches=st.integers(1, 3),
    log=st.booleans()))


@settings(**SETTINGS)
@given(c=mgr_case)
def test_quantization_manager_flow_matches_reference(ref_estimators, c, monkeypatch):
    """QuantizationManager end to end on random configurations: estimate over several batches -> fix_ranges ->
    quantize a new batch; the reference in place vs this package's classes (oracle back-end for the kernels).
    Outputs equal (scale_domain='log': within one grid step, libm vs numpy expf on the scale)."""
    import tq_native
    from oracle_backend import OracleOps
    monkeypatch.setattr(tq_native, '_OPS', OracleOps())
    monkeypatch.setattr(tq_native, 'default_device', lambda: torch.device('cpu'))
    import quantization.quantization_manager as mine_mgr
    import quantization.quantizers as mine_q
    import quantization.range_estimators as mine_est
    R, Q = ref_estimators
    saved = {k: v for k, v in sys.modules.items() if k == 'quantization' or k.startswith('quantization.')}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        ref_mgr = importlib.import_module('quantization.quantization_manager')
        ref_q = importlib.import_module('quantization.quantizers')
        ref_est = importlib.import_module('quantization.range_estimators')
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == 'quantization' or k.startswith('quantization.')]:
            del sys.modules[k]
        sys.modules.update(saved)
    layout = c['layout']
    if not c['asym'] and layout in ('axis', 'groups'):
        layout = 'tensor'                      # symmetric quantizer: no per-axis mode (quirk A.4-1)
    if c['est'] == 'allminmax' and layout in ('axis', 'groups'):
        layout = 'tensor'                      # the all-min-max estimator ignores `axis` (quirk A.4-7)
    d = c['groups'] * c['per_group']
    rs = np.random.RandomState(c['seed'])
    sc = (0.5 + 3 * rs.rand(d)).astype(np.float32)
    batches = [(rs.randn(c['rows'], 3, d) * sc * (1 + 0.3 * i)).astype(np.float32) for i in range(c['n_batches'] + 1)]
    kw = dict(per_channel=layout == 'channel', axis=2 if layout in ('axis', 'groups') else None,
              n_groups=c['groups'] if layout == 'groups' else None)
    qparams = dict(n_bits=c['n_bits'], scale_domain='log' if c['log'] else 'linear')
    outs = []
    for M, QQ, E in ((ref_mgr, ref_q, ref_est), (mine_mgr, mine_q, mine_est)):
        qm = QQ.QMethods.asymmetric_uniform if c['asym'] else QQ.QMethods.symmetric_uniform
        m = M.QuantizationManager(qmethod=qm, init=E.RangeEstimators[c['est']], qparams=dict(qparams), **kw)
        ys = []
        with torch.no_grad():
            for b in batches[:-1]:
                ys.append(m(torch.from_numpy(b)).numpy().copy())
            m.fix_ranges()
            ys.append(m(torch.from_numpy(batches[-1])).numpy().copy())
        outs.append((ys, m.quantizer._delta.detach().numpy().reshape(-1).copy()))
    (ys_r, d_r), (ys_m, d_m) = outs
    if c['log']:
        np.testing.assert_allclose(d_m, d_r, rtol=1e-6, atol=1e-7)
        step = float(np.exp(d_r).max())
        for a, b in zip(ys_r, ys_m):
            assert np.abs(a - b).max() <= step * 1.0001
        return
    assert np.array_equal(d_r, d_m)
    for a, b in zip(ys_r, ys_m):
        assert np.array_equal(a, b)


ada_case = st.fixed_dictionaries(dict(
    seed=st.integers(0, 2 ** 31 - 1), rows=st.integers(1, 6), cols=st.integers(1, 24), asym=st.booleans(),
    per_channel=st.booleans(), n_bits=st.sampled_from([2, 3, 4, 8]),
    mode=st.sampled_from(['learned_sigmoid', 'learned_hard_sigmoid', 'sigmoid_temp_decay']),
    temperature=st.sampled_from([0.5, 1.0, 2.5, 20.0]), shrink=st.sampled_from([0.5, 0.8, 1.0])))


@settings(**SETTINGS)
@given(c=ada_case)
def test_adaround_quantizer_matches_reference(ref_quantizers, c):
    """AdaRoundQuantizer in the relaxation modes (reference quantization/adaround/quantizer.py:46-92), oracle vs the
    reference in place on random weights / grids / alpha: alpha initialisation, soft and hard forward, d / d alpha."""
    saved = {k: v for k, v in sys.modules.items() if k.split('.')[0] in ('quantization', 'utils')}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        import importlib.util
        import types
        pkg = types.ModuleType('utils')
        pkg.__path__ = []
        sys.modules['utils'] = pkg
        spec = importlib.util.spec_from_file_location('utils.utils', os.path.join(REF, 'utils', 'utils.py'))
        sub = importlib.util.module_from_spec(spec)
        sys.modules['utils.utils'] = sub
        spec.loader.exec_module(sub)
        aq = importlib.import_module('quantization.adaround.quantizer')
        au = importlib.import_module('quantization.adaround.utils')
        rq = importlib.import_module('quantization.quantizers')
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k.split('.')[0] in ('quantization', 'utils')]:
            del sys.modules[k]
        sys.modules.update(saved)
    rs = np.random.RandomState(c['seed'])
    w = (rs.randn(c['rows'], c['cols']) * 0.05).astype(np.float32)
    base = rq.AsymmetricUniformQuantizer if c['asym'] else rq.SymmetricUniformQuantizer
    q = aq.ADAROUND_QUANTIZER_MAP[base](n_bits=c['n_bits'], per_channel=c['per_channel'])
    if c['per_channel']:
        q.set_quant_range(torch.from_numpy(w.min(1) * c['shrink']), torch.from_numpy(w.max(1) * c['shrink']))
    else:
        q.set_quant_range(float(w.min()) * c['shrink'], float(w.max()) * c['shrink'])
    q.round_mode = au.AdaRoundMode[c['mode']]
    q.temperature = c['temperature']
    q.soft_targets = True
    wt = torch.from_numpy(w)
    y0 = q(wt)
    scale = O.scale_of(q._delta.detach().numpy().reshape(-1))
    if c['asym']:
        zp = O.asym_zero_point(q._zero_float.detach().numpy().reshape(-1), c['n_bits'])
        lo, hi = 0.0, O.asym_int_max(c['n_bits'])
    else:
        zp = np.zeros_like(scale)
        lo, hi = O.sym_grid(c['n_bits'], bool(q.signed))
    shp = (-1, 1) if c['per_channel'] else ()
    scale, zp = scale.reshape(shp), zp.reshape(shp)
    step = float(np.max(scale))
    ytol = 4e-6 * step * max(abs(lo), hi, 1.0)
    temp = c['temperature']
    a0 = O.adaround_alpha_init(w, scale, c['mode'], temp)
    ref_a0 = q.alpha.detach().numpy()
    fin = np.isfinite(ref_a0)
    assert np.array_equal(fin, np.isfinite(a0))
    np.testing.assert_allclose(a0[fin], ref_a0[fin], rtol=5e-5, atol=5e-5 * max(temp, 1.0))
    np.testing.assert_allclose(O.adaround_qdq(w, np.where(fin, ref_a0, 0), scale, zp, lo, hi, c['mode'], True, temp)[fin],
                               y0.detach().numpy()[fin], rtol=0, atol=ytol)
    alpha1 = (np.where(fin, ref_a0, 0) + rs.randn(*w.shape) * 1.5).astype(np.float32)
    with torch.no_grad():
        q.alpha.copy_(torch.from_numpy(alpha1))
    g = rs.randn(*w.shape).astype(np.float32)
    y1 = q(wt)
    y1.backward(torch.from_numpy(g))
    np.testing.assert_allclose(O.adaround_qdq(w, alpha1, scale, zp, lo, hi, c['mode'], True, temp), y1.detach().numpy(),
                               rtol=0, atol=ytol)
    ga = O.adaround_grad_alpha(w, alpha1, g, scale, zp, lo, hi, c['mode'], temp)
    np.testing.assert_allclose(ga, q.alpha.grad.numpy(), rtol=5e-5, atol=2e-7 * step)
    q.soft_targets = False
    with torch.no_grad():
        xi = q.to_integer_forward(wt).numpy()
    mine, _ = O.adaround_to_integer(w, alpha1, scale, zp, lo, hi, c['mode'], False, temp)
    assert np.array_equal(mine, xi)
