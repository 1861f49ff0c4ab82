"""Drop-in check: the reference's own ``models/quantized_bert.py`` -- UNCHANGED, loaded from the
reference checkout -- runs on top of THIS repo's ``quantization`` / ``utils`` packages and reproduces
the golden outputs it produced on top of the reference's packages.

Needs the reference checkout (build container only; skipped on the GPU box, where /root/reference
does not exist).  Arithmetic back-end: the CPU oracle injected through the test fixture, the
library ops (matmul, layer_norm, ...) are the same torch CPU calls in both runs, so the logits must
be identical.
"""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

import tq_native
from conftest import GOLDEN, PKG

from reference_path import reference_root
REF = reference_root()
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'models')),
                                reason='reference checkout not present')


def _gm():
    spec = importlib.util.spec_from_file_location('make_golden_model', os.path.join(GOLDEN, 'make_golden_model.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture()
def drop_in(monkeypatch):
    from oracle_backend import OracleOps
    monkeypatch.setattr(tq_native, '_OPS', OracleOps())
    monkeypatch.setattr(tq_native, 'default_device', lambda: torch.device('cpu'))
    saved = {k: v for k, v in sys.modules.items() if k.split('.')[0] in ('quantization', 'utils', 'models')}
    gm = _gm()
    qb = gm.import_reference_model(PKG)           # reference model file + THIS package
    yield gm, qb
    for k in [k for k in sys.modules if k.split('.')[0] in ('quantization', 'utils', 'models')]:
        del sys.modules[k]
    sys.modules.update(saved)


@pytest.mark.parametrize('name', ['w8a8_asym', 'w8a8_sym', 'w4a8_asym', 'w8a8_peg4', 'w8a8_pegp4'])
def test_reference_model_file_runs_unchanged_on_this_package(drop_in, name):
    gm, qb = drop_in
    assert qb.__file__.startswith(REF)
    import quantization
    assert quantization.__file__.startswith(PKG)
    G = np.load(os.path.join(GOLDEN, 'bert_tiny.npz'))
    torch.set_grad_enabled(False)
    try:
        res, model = gm.run_config(qb, name, gm.CONFIGS[name], gm.make_hf_model(), gm.make_batches())
    finally:
        torch.set_grad_enabled(True)
    assert np.array_equal(res[f'{name}.logits'], G[f'{name}.logits'])
    assert np.array_equal(res[f'{name}.last_hidden'], G[f'{name}.last_hidden'])
    n = int(G[f'{name}.n_act_quantizers'])
    assert int(res[f'{name}.n_act_quantizers']) == n
    for i in range(n):
        assert np.array_equal(res[f'{name}.q{i}.delta'], G[f'{name}.q{i}.delta']), str(G[f'{name}.q{i}.name'])


@pytest.mark.parametrize('name', ['roberta_w8a8', 'roberta_w8a8_mse'])
def test_reference_roberta_model_file_runs_unchanged_on_this_package(drop_in, name):
    """models/quantized_roberta.py (BASELINE config 5 family; MSE-grid activation ranges in the second
    case) -- unchanged reference file on this repo's quantization / utils packages."""
    gm, qb = drop_in
    assert qb.roberta.__file__.startswith(REF)
    G = np.load(os.path.join(GOLDEN, 'roberta_tiny.npz'))
    torch.set_grad_enabled(False)
    try:
        res, model = gm.run_roberta_config(qb, name, gm.ROBERTA_CONFIGS[name], gm.make_hf_roberta(), gm.make_batches())
    finally:
        torch.set_grad_enabled(True)
    assert int(res[f'{name}.n_act_quantizers']) == int(G[f'{name}.n_act_quantizers'])
    if name.endswith('_mse'):
        # MSE losses are accumulated in a different (deterministic fp64) order than torch's blocked fp32
        # sums: the selected grid candidate can differ by one on near-ties -> logits within 2 output steps
        step = float(model.classifier.out_proj.activation_quantizer.quantizer.scale)
        assert np.abs(res[f'{name}.logits'] - G[f'{name}.logits']).max() <= 2 * step + 1e-7
    else:
        assert np.array_equal(res[f'{name}.logits'], G[f'{name}.logits'])
        assert np.array_equal(res[f'{name}.last_hidden'], G[f'{name}.last_hidden'])


@pytest.mark.parametrize('name', ['mobilebert_w4a8', 'mobilebert_w8a8'])
def test_reference_mobilebert_model_file_runs_unchanged_on_this_package(drop_in, name):
    """models/quantized_mobilebert.py (BASELINE config 4 family: W4A8, QuantNoNorm, bottlenecks, stacked
    FFNs) -- unchanged reference file, HF MobileBERT building blocks, this repo's quantization package."""
    gm, qb = drop_in
    qm = gm.import_reference_mobilebert(qb)
    assert qm.__file__.startswith(REF)
    G = np.load(os.path.join(GOLDEN, 'mobilebert_tiny.npz'))
    torch.set_grad_enabled(False)
    try:
        res, model = gm.run_mobilebert_config(qm, name, gm.MOBILEBERT_CONFIGS[name], gm.make_hf_mobilebert(),
                                              gm.make_batches())
    finally:
        torch.set_grad_enabled(True)
    assert int(res[f'{name}.n_quantizers']) == int(G[f'{name}.n_quantizers'])
    assert np.array_equal(res[f'{name}.logits'], G[f'{name}.logits'])
    assert np.array_equal(res[f'{name}.last_hidden'], G[f'{name}.last_hidden'])


@pytest.mark.parametrize('name', ['w8a8_asym', 'w4a8_asym', 'w8a8_sym'])
def test_reference_model_qat_step_on_this_package(drop_in, name):
    """One quantization-aware training step with learnable ranges: the unchanged
    reference model file on this package's quantizers (FakeQuantSTE -> tq_qdq_bwd semantics, oracle back-end
    here) vs the reference's autograd.  Forward, loss and weight gradients are the same fp32 chain ->
    equal; range gradients are sums in a different order -> rtol 1e-3 relative to the largest one."""
    gm, qb = drop_in
    G = np.load(os.path.join(GOLDEN, 'bert_tiny_qat.npz'))
    batches = gm.make_batches()
    torch.set_grad_enabled(False)
    try:
        _, model = gm.run_config(qb, name, gm.CONFIGS[name], gm.make_hf_model(), batches)
        res = gm.run_qat_step(model, name, batches[-1], gm.qat_labels())
    finally:
        torch.set_grad_enabled(True)
    assert int(res[f'{name}.qat.n_range_params']) == int(G[f'{name}.qat.n_range_params'])
    assert np.array_equal(res[f'{name}.qat.logits'], G[f'{name}.qat.logits'])
    assert float(res[f'{name}.qat.loss']) == float(G[f'{name}.qat.loss'])
    keys = [k for k in G.files if k.startswith(f'{name}.qat.grad.')]
    assert keys and set(keys) == {k for k in res if k.startswith(f'{name}.qat.grad.')}
    rng = [k for k in keys if k.split('.')[-1] in ('_delta', '_zero_float')]
    for k in keys:
        if k in rng:
            continue
        np.testing.assert_allclose(res[k], G[k], rtol=1e-4, atol=1e-6 * np.abs(G[k]).max(), err_msg=k)
    for leaf in ('_delta', '_zero_float'):
        sel = [k for k in rng if k.endswith(leaf)]
        if not sel:
            continue
        top = max(np.abs(G[k]).max() for k in sel)
        for k in sel:
            assert np.abs(res[k] - G[k]).max() <= 1e-3 * top + 1e-3 * np.abs(G[k]).max(), k
