"""Every BASELINE.json configuration as a named recipe (engine/configs.py) driven end to end on CPU at tiny
dimensions with the oracle arithmetic back-end: build -> [FP32 ranges pass] -> calibrate -> fix -> eval.
Checks the plumbing each configuration adds on top of the per-site parity tests: symmetric activations
(config 1), per-embedding-group ranges with permutation (config 3), MobileBERT W4A8 (config 4), RoBERTa positions
+ MSE grid ranges (config 5)."""
import numpy as np
import pytest
import torch

import tq_native
from oracle_backend import OracleOps


@pytest.fixture(autouse=True)
def oracle_ops(monkeypatch):
    monkeypatch.setattr(tq_native, '_OPS', OracleOps())
    monkeypatch.setattr(tq_native, 'default_device', lambda: torch.device('cpu'))


def _run(name, **kw):
    from engine import configs
    model, recipe = configs.build(name, 'cpu', tiny=True, **kw)
    batches = configs.synthetic_batches(model, recipe, 3, batch=2, seq=16)
    configs.calibrate(model, recipe, batches[:2])
    with torch.no_grad():
        mask = torch.ones_like(batches[2])
        y1 = model(batches[2], mask)
        y2 = model(batches[2], mask)
    assert y1.shape == (2, 2) and torch.isfinite(y1).all() and torch.equal(y1, y2)
    return model, recipe, y1


def _managers(model):
    from quantization.quantization_manager import QuantizationManager
    return [m for m in model.modules() if isinstance(m, QuantizationManager)]


@pytest.mark.parametrize('name', ['bert_w8a8_sym', 'bert_w8a8_asym', 'mobilebert_w4a8'])
def test_per_tensor_configs(name):
    from quantization.quantization_manager import Qstates
    from quantization.quantizers import SymmetricUniformQuantizer
    model, recipe, _ = _run(name)
    mgrs = [m for m in _managers(model) if m.quantizer.is_initialized]
    assert mgrs and all(m.state is Qstates.fix_ranges for m in mgrs)
    assert all(m.quantizer._delta.numel() == 1 for m in mgrs)
    if name == 'bert_w8a8_sym':
        assert all(isinstance(m.quantizer, SymmetricUniformQuantizer) for m in mgrs)
    if name == 'mobilebert_w4a8':
        assert {m.quantizer.n_bits for m in mgrs} == {4, 8}


@pytest.mark.parametrize('name', ['bert_w8a8_peg', 'bert_w8a8_pegp'])
def test_peg_config(name):
    """config 3 (contiguous groups, and the range-permuted variant): the PEG sites carry per-dim parameters with exactly
    K = 6 distinct groups, the others stay per-tensor"""
    model, recipe, _ = _run(name)
    peg = [s.activation_quantizer for s in model.peg_sites()]
    assert len(peg) == 3 + 10 * len(model.layers)
    for mgr in peg:
        d = mgr.quantizer._delta.detach().numpy().reshape(-1)
        assert d.size == model.config.hidden_size and mgr.n_groups == 6
        assert len(np.unique(d)) <= 6
        assert (mgr.range_estimator.ranges is not None) == (name == 'bert_w8a8_pegp')     # permutation ranges collected?
    others = [m for m in _managers(model) if m.quantizer.is_initialized and all(m is not p for p in peg)]
    assert others and all(m.quantizer._delta.numel() == 1 for m in others)


def test_roberta_mse_config():
    """config 5: RoBERTa position ids (offset by the padding id) and MSE-grid activation ranges"""
    from quantization.range_estimators import OptMethod
    model, recipe, _ = _run('roberta_w8a8_mse', act_range_options=dict(opt_method=OptMethod.grid, num_candidates=4))
    assert model.embeddings.roberta_positions and model.config.pad_token_id == 1
    est = model.layers[0].query.activation_quantizer.range_estimator
    assert type(est).__name__ == 'MSE_Estimator' and est.loss_array is not None


@pytest.mark.parametrize('name', ['bert_w8a8_peg', 'mobilebert_w4a8'])
def test_run_config_tool_dry_run(name):
    """tools/run_config.py end to end on CPU (tiny dimensions, oracle back-end, wall-clock timing): the control
    flow a GPU run takes, minus the CUDA graph and the fused engine"""
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location('run_config', os.path.join(ROOT, 'tools', 'run_config.py'))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    r = tool.run(name, steps=1, warmup=1, device='cpu', tiny=True, batch=2, seq=16)
    assert r['config'] == name and r['forward'] == 'module path' and r['logits_finite']
    assert r['max_abs_logit_diff_vs_module_path'] == 0.0 and r['tokens_per_s'] > 0


def test_quant_options_to_model():
    """utils.quant_options: reference option names -> config -> make_qparams -> a quantized model (W4 per-channel
    MSE weights, A8 running min-max with momentum) that calibrates and runs"""
    from engine.bert import BertConfig, QuantBertForSequenceClassification
    from quantization.range_estimators import OptMethod, RangeEstimators
    from utils.quant_options import make_qparams, quant_config
    cfg = quant_config(n_bits=4, n_bits_act=8, qmethod_act='asymmetric_uniform', per_channel=True,
                       weight_quant_method='MSE', num_candidates=10, act_momentum=0.5)
    qp = make_qparams(cfg)
    assert qp['weight_range_method'] is RangeEstimators.MSE and qp['weight_range_options'] == dict(
        opt_method=OptMethod.grid, num_candidates=10)
    assert qp['act_range_options'] == dict(momentum=0.5) and qp['per_channel_weights'] and qp['n_bits'] == 4
    with pytest.raises(ValueError):
        quant_config(act_num_candidates=5)                     # only valid with the MSE activation estimator
    with pytest.raises(TypeError):
        quant_config(nbits=4)
    model = QuantBertForSequenceClassification(
        BertConfig(vocab_size=100, hidden_size=32, num_hidden_layers=1, num_attention_heads=2, intermediate_size=32,
                   max_position_embeddings=16), **qp)
    model.init_weights(seed=0).eval()
    model.set_quant_state(cfg.quant.weight_quant, cfg.quant.act_quant)
    ids = torch.randint(0, 100, (2, 8), generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        model(ids, torch.ones_like(ids))
        model.fix_ranges()
        y = model(ids, torch.ones_like(ids))
    assert torch.isfinite(y).all()
    assert model.layers[0].query.weight_quantizer.quantizer._delta.numel() == 32       # per-channel ranges
