"""Range-permuted per-embedding-group quantization (PEG with permutation, BASELINE config 3) at model level with
the reference's real calibration protocol (main.py:519-537: the FP32 pass runs with the activation quantizers ON
and the weights FP32, so every PEG site records its per-dim ranges and the groups are formed on range-sorted
dims).  Golden: tests/golden/bert_tiny_pegp.npz (make_golden_model.py, TQ_GOLDEN_ONLY=pegp; 23 estimators with
permutation ranges).  CPU, oracle back-end: calibrated parameters, logits and hidden states exactly equal, for

  * this package's own caller (engine/bert.py through engine/configs.calibrate's protocol), and
  * the UNCHANGED reference model file on this package's quantization / utils modules.

(tests/golden/bert_tiny.npz::w8a8_pegp4 keeps the quantizers off during that pass -- no ranges, no permutation --
and stays as the GPU model-level case until the permuted one has been run on hardware.)"""
import os
import sys

import numpy as np
import pytest
import torch

import tq_native
from conftest import GOLDEN, PKG

GP = np.load(os.path.join(GOLDEN, 'bert_tiny_pegp.npz'))
GW = np.load(os.path.join(GOLDEN, 'bert_tiny.npz'))
NAME = 'w8a8_pegp4'
from reference_path import reference_root
REF = reference_root()


@pytest.fixture()
def oracle_ops(monkeypatch):
    from oracle_backend import OracleOps
    monkeypatch.setattr(tq_native, '_OPS', OracleOps())
    monkeypatch.setattr(tq_native, 'default_device', lambda: torch.device('cpu'))


def test_own_caller_with_permutation(oracle_ops):
    from engine import configs
    from engine.bert import BertConfig, QuantBertForSequenceClassification
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators
    cfg = BertConfig(vocab_size=1000, hidden_size=128, num_hidden_layers=2, num_attention_heads=4,
                     intermediate_size=512, max_position_embeddings=64)
    model = QuantBertForSequenceClassification(cfg, method=QMethods.symmetric_uniform,
                                               act_method=QMethods.asymmetric_uniform, n_bits=8, n_bits_act=8,
                                               act_range_method=RangeEstimators.current_minmax)
    model.load_hf_state_dict({k[2:]: torch.from_numpy(GW[k]) for k in GW.files if k.startswith('w.')})
    model.eval()
    model.set_quant_state(weight_quant=True, act_quant=True)
    ids = [torch.from_numpy(GW['ids'][i]) for i in range(3)]
    recipe = configs.RECIPES['bert_w8a8_peg']._replace(peg=('ngp', 4))
    configs.calibrate(model, recipe, ids[:-1])
    n_perm = sum(1 for m in model.modules() if getattr(m, 'ranges', None) is not None)
    assert n_perm == int(GP[f'{NAME}.n_range_vectors'])
    with torch.no_grad():
        mask = torch.ones_like(ids[-1])
        logits, hidden = model(ids[-1], mask), model.encode(ids[-1], mask)
    qs = model.act_quantizers()
    assert len(qs) == int(GP[f'{NAME}.n_act_quantizers'])
    for i, mgr in enumerate(qs):
        assert np.array_equal(mgr.quantizer._delta.detach().numpy().reshape(-1), GP[f'{NAME}.q{i}.delta']), \
            f'site {i} ({GP[f"{NAME}.q{i}.name"]})'
    assert np.array_equal(logits.numpy(), GP[f'{NAME}.logits'])
    assert np.array_equal(hidden.numpy(), GP[f'{NAME}.last_hidden'])
    assert not np.array_equal(logits.numpy(), GW[f'{NAME}.logits'])       # the permutation changes the result


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'models')), reason='reference checkout not present')
def test_reference_model_file_with_permutation(oracle_ops):
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_golden_model', os.path.join(GOLDEN, 'make_golden_model.py'))
    gm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gm)
    saved = {k: v for k, v in sys.modules.items() if k.split('.')[0] in ('quantization', 'utils', 'models')}
    try:
        qb = gm.import_reference_model(PKG)
        torch.set_grad_enabled(False)
        res, model = gm.run_config(qb, NAME, gm.CONFIGS[NAME], gm.make_hf_model(), gm.make_batches(),
                                   main_py_ranges_pass=True)
    finally:
        torch.set_grad_enabled(True)
        for k in [k for k in sys.modules if k.split('.')[0] in ('quantization', 'utils', 'models')]:
            del sys.modules[k]
        sys.modules.update(saved)
    assert np.array_equal(res[f'{NAME}.logits'], GP[f'{NAME}.logits'])
    assert np.array_equal(res[f'{NAME}.last_hidden'], GP[f'{NAME}.last_hidden'])
    for i in range(int(GP[f'{NAME}.n_act_quantizers'])):
        assert np.array_equal(res[f'{NAME}.q{i}.delta'], GP[f'{NAME}.q{i}.delta'])
