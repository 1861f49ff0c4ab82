"""Training-time host logic on CPU (oracle injected as the arithmetic back-end, test-only): the
autograd wiring of the quantizers (FakeQuantSTE), learnable ranges through the manager state machine
and the AdaRound quantizer classes reproduce the reference under torch autograd (tests/golden/qat.npz)."""
import numpy as np
import pytest
import torch

import tq_native
from oracle_backend import OracleOps
from qat_cases import QAT_MANIFEST, check_backward_case

CPU = torch.device('cpu')


@pytest.fixture(autouse=True)
def oracle_ops(monkeypatch):
    monkeypatch.setattr(tq_native, '_OPS', OracleOps())
    monkeypatch.setattr(tq_native, 'default_device', lambda: CPU)


@pytest.mark.parametrize('case', QAT_MANIFEST['backward'], ids=lambda c: c['name'])
def test_backward_api(case):
    check_backward_case(case, CPU)


def test_learn_ranges_state():
    """QuantizationManager.learn_ranges -> nn.Parameters that receive gradients (reference
    quantization_manager.py:82-84, quantizers.py:284-288)"""
    from quantization.quantization_manager import QuantizationManager, Qstates
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators
    rs = np.random.RandomState(0)
    x = torch.from_numpy((rs.randn(4, 16, 32) * 2).astype(np.float32))
    for qm, names in ((QMethods.asymmetric_uniform, {'quantizer._delta', 'quantizer._zero_float'}),
                      (QMethods.symmetric_uniform, {'quantizer._delta'})):
        m = QuantizationManager(qmethod=qm, init=RangeEstimators.running_minmax, qparams=dict(n_bits=4))
        m(x)
        m.learn_ranges()
        assert m.state is Qstates.learn_ranges
        assert {n for n, _ in m.named_parameters()} == names
        d0 = m.quantizer._delta.detach().clone()
        y = m(x * 3)                      # learn_ranges: no estimator update
        assert torch.equal(m.quantizer._delta.detach(), d0)
        y.square().sum().backward()
        for _, p in m.named_parameters():
            assert p.grad is not None and p.grad.shape == p.shape and torch.isfinite(p.grad).all()
        assert m.quantizer._delta.grad.abs().sum() > 0


from qat_cases import check_adaround_case  # noqa: E402


@pytest.mark.parametrize('case', QAT_MANIFEST['adaround'], ids=lambda c: c['name'])
def test_adaround_quantizer_api(case):
    check_adaround_case(case, CPU)


from qat_cases import check_adaround_layer_case  # noqa: E402


@pytest.mark.parametrize('case', QAT_MANIFEST['adaround_layer'], ids=lambda c: c['name'])
def test_adaround_layer_loop(case):
    check_adaround_layer_case(case, CPU)


def test_temp_decay_schedules():
    """every schedule starts at b_range[0], ends at b_range[1] and is monotone"""
    from quantization.adaround.utils import AdaRoundTempDecayType, TempDecay
    for kind in AdaRoundTempDecayType:
        sched = TempDecay(100, b_range=(20.0, 2.0), rel_decay_start=0.2, decay_type=kind, decay_shape=2.0)
        vals = [sched(t) for t in range(0, 101)]
        assert vals[0] == 20.0 and vals[10] == 20.0
        assert abs(vals[-1] - 2.0) < 1e-9, (kind, vals[-1])
        assert all(a >= b - 1e-12 for a, b in zip(vals, vals[1:])), kind


def test_training_step_matches_torch_autograd():
    from qat_cases import check_training_step
    check_training_step(CPU)


def test_estimate_ranges_train_state():
    """Qstates.estimate_ranges_train (the default QAT mode, reference quantization_manager.py:12-16, 94-106 and
    qat_utils.py:36-42): ranges follow the data while the module is in train mode and are frozen in eval mode;
    the forward stays differentiable w.r.t. its input (straight-through) in both."""
    from quantization.quantization_manager import QuantizationManager, Qstates
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators
    rs = np.random.RandomState(1)
    x1 = torch.from_numpy(rs.randn(4, 8, 32).astype(np.float32))
    x2 = (x1 * 4).requires_grad_(True)
    m = QuantizationManager(qmethod=QMethods.asymmetric_uniform, init=RangeEstimators.running_minmax,
                            qparams=dict(n_bits=8))
    m(x1)
    m.estimate_ranges_train()
    assert m.state is Qstates.estimate_ranges_train
    d0 = m.quantizer._delta.clone()
    m.eval()
    y = m(x2)
    assert torch.equal(m.quantizer._delta, d0), 'ranges must be frozen in eval mode'
    y.sum().backward()
    inside = (x2.detach() >= m.quantizer.x_min) & (x2.detach() <= m.quantizer.x_max)
    assert torch.equal(x2.grad != 0, inside) or (x2.grad != 0).sum() >= inside.sum() * 0.98   # edges round in
    m.train()
    x2.grad = None
    y = m(x2)
    assert not torch.equal(m.quantizer._delta, d0), 'ranges must follow the data in train mode'
    y.sum().backward()
    assert x2.grad is not None and (x2.grad != 0).any()
    assert not any(p.requires_grad for p in m.parameters()), 'no learnable ranges in this state'


def test_backward_uses_the_range_of_its_own_forward():
    """One quantizer serving two tensors in a forward (QuantNoNorm quantizes weight, then bias, with the same
    weight quantizer; reference quantized_mobilebert.py:64-68) while its range is being estimated: the range
    buffers are rewritten in place by the second call, the backward of the FIRST call must still use the range
    it quantized with (the reference allocates new range tensors per set_quant_range)."""
    from quantization.quantization_manager import QuantizationManager
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators
    rs = np.random.RandomState(2)
    for qm in (QMethods.asymmetric_uniform, QMethods.symmetric_uniform):
        w = torch.from_numpy((rs.randn(64) * 1.0).astype(np.float32)).requires_grad_(True)
        b = torch.from_numpy((rs.rand(64) * 0.05).astype(np.float32)).requires_grad_(True)   # much smaller, one-sided
        m = QuantizationManager(qmethod=qm, init=RangeEstimators.current_minmax, qparams=dict(n_bits=4))
        m.quantizer.set_quant_range(-0.5, 0.5)          # pre-allocates the buffers that get rewritten in place
        wq = m(w)                                       # range <- w
        d_w = m.quantizer._delta.detach().clone()
        bq = m(b)                                       # range <- b, same buffers
        assert not torch.equal(m.quantizer._delta, d_w)
        (wq.sum() + bq.sum()).backward()
        # STE: every w lies inside its own [min, max] range -> gradient 1 everywhere; with the bias's (tiny) range
        # almost every element of w would be clamped -> gradient 0
        assert torch.equal(w.grad, torch.ones_like(w)), 'backward of the first call used the range of the second call'
        assert torch.equal(b.grad, torch.ones_like(b))


def test_fused_engine_refuses_adaround_weights():
    """the fused engine quantizes weights round-to-nearest: a model whose weights carry learned (AdaRound)
    rounding must be refused (callers then keep the module path), never silently re-rounded"""
    from engine.bert import BertConfig, QuantBertForSequenceClassification
    from engine.fused import FusedBertEngine, UnsupportedByEngine
    from quantization.adaround.adaround import _adaround_quantizer_like
    from quantization.adaround.utils import AdaRoundMode
    from quantization.quantizers import QMethods
    cfg = BertConfig(vocab_size=300, hidden_size=256, num_hidden_layers=1, num_attention_heads=4,
                     intermediate_size=256, max_position_embeddings=128)
    model = QuantBertForSequenceClassification(cfg, method=QMethods.symmetric_uniform,
                                               act_method=QMethods.asymmetric_uniform, n_bits=8, n_bits_act=8)
    model.init_weights(seed=0).eval()
    model.set_quant_state(True, True)
    ids = torch.randint(0, 300, (1, 128), generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        model(ids, torch.ones_like(ids))
    model.fix_ranges()
    FusedBertEngine(model, 1, 128)                          # plain model: accepted
    lin = model.layers[0].ffn_in
    q = _adaround_quantizer_like(lin.weight_quantizer.quantizer)
    q.round_mode = AdaRoundMode.learned_hard_sigmoid
    lin.weight_quantizer.quantizer = q
    with pytest.raises(UnsupportedByEngine):
        FusedBertEngine(model, 1, 128)


@pytest.mark.parametrize('case', QAT_MANIFEST['mse_per_channel'], ids=lambda c: c['name'])
def test_mse_per_channel_weights(case):
    """MSE_Estimator(per_channel=True) on weight matrices (one search per output channel, shared search grid;
    reference range_estimators.py:228-490): per-channel losses rtol 1e-5 (fp64 accumulation here, fp32 blocked
    sums in torch), selected ranges equal (grid) / rtol 2e-3 (golden section, scipy on a slightly different
    objective surface)."""
    from qat_cases import qat_file
    from quantization.quantizers import QMethods
    from quantization.range_estimators import OptMethod, RangeEstimators
    g, nm = qat_file(), case['name']
    qm = QMethods.symmetric_uniform if case['kind'] == 'sym' else QMethods.asymmetric_uniform
    qz = qm.cls(n_bits=case['n_bits'], per_channel=True)
    est = RangeEstimators.MSE.cls(quantizer=qz, per_channel=True, opt_method=OptMethod[case['opt']],
                                  num_candidates=case['num_candidates'])
    mn, mx = est(torch.from_numpy(g[f'{nm}.x']))
    assert bool(est.one_sided_dist) == case['one_sided']
    if case['opt'] == 'grid':
        loss, ref = np.asarray(est.loss_array, np.float64), g[f'{nm}.loss']
        assert loss.shape == ref.shape
        fin = np.isfinite(ref)
        np.testing.assert_allclose(loss[fin], ref[fin], rtol=1e-5)
        got_min, got_max = mn.numpy().reshape(-1), mx.numpy().reshape(-1)
        for ch in range(loss.shape[0]):
            if got_min[ch] == g[f'{nm}.xmin'][ch] and got_max[ch] == g[f'{nm}.xmax'][ch]:
                continue
            # a different candidate may only win on a tie: several skew candidates clamp to the same grid and
            # their fp32 sums coincide exactly in the reference, while the fp64 sums here differ in the last bits
            mine = int(np.argmin(loss[ch]))
            ref_min = float(ref[ch].min())
            assert abs(float(ref[ch].reshape(-1)[mine]) - ref_min) <= 1e-5 * ref_min, f'channel {ch}'
    else:
        np.testing.assert_allclose(mx.numpy().reshape(-1), g[f'{nm}.xmax'], rtol=2e-3)
        np.testing.assert_allclose(mn.numpy().reshape(-1), g[f'{nm}.xmin'], rtol=2e-3)


@pytest.mark.parametrize('case', QAT_MANIFEST['cross_entropy'], ids=lambda c: c['name'])
def test_cross_entropy_estimator(case):
    """CrossEntropyEstimator on logits (reference range_estimators.py:493-502): accumulated loss arrays rtol 1e-5,
    selected range equal unless two candidates tie within that tolerance"""
    from qat_cases import qat_file
    from quantization.quantizers import QMethods
    from quantization.range_estimators import OptMethod, RangeEstimators
    g, nm = qat_file(), case['name']
    qm = QMethods.symmetric_uniform if case['kind'] == 'sym' else QMethods.asymmetric_uniform
    est = RangeEstimators.cross_entropy.cls(quantizer=qm.cls(n_bits=case['n_bits']), opt_method=OptMethod.grid,
                                            num_candidates=case['num_candidates'])
    for i in range(case['n_batches']):
        mn, mx = est(torch.from_numpy(g[f'{nm}.x{i}']))
        loss, ref = np.asarray(est.loss_array, np.float64), g[f'{nm}.b{i}.loss']
        fin = np.isfinite(ref)
        assert loss.shape == ref.shape and np.array_equal(fin, np.isfinite(loss))
        np.testing.assert_allclose(loss[fin], ref[fin], rtol=1e-5)
        if not (np.array_equal(mn.numpy().reshape(-1), g[f'{nm}.b{i}.xmin']) and
                np.array_equal(mx.numpy().reshape(-1), g[f'{nm}.b{i}.xmax'])):
            mine = int(np.argmin(loss[0]))
            ref_min = float(ref[0].min())
            assert abs(float(ref[0].reshape(-1)[mine]) - ref_min) <= 1e-5 * abs(ref_min)
