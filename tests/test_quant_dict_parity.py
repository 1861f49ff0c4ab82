"""Mixed-precision / per-site control through the reference's ``--quant-dict`` grammar (README.md:160-173,
main.py:442-498; helpers utils/per_embd_quant_utils.py:7-52): the paper's MP-PTQ recipe (16-bit FFN input, output
and residual sum), PEG on the FFN sites only, and a dictionary that uses every value kind (bits, 'fp32',
'per_embd', 'ng<K>', per-layer keys, the all-modules key, weight keys).  Golden: tests/golden/
bert_tiny_quant_dict.npz (reference model + the reference's hijack helpers wired like main.py).  CPU, oracle
back-end: per-site bit-widths and calibrated parameters, logits and hidden states exactly equal, for this
package's own caller (engine.bert.apply_quant_dict) and for the UNCHANGED reference model file on this package."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

import tq_native
from conftest import GOLDEN, PKG

G = np.load(os.path.join(GOLDEN, 'bert_tiny_quant_dict.npz'))
GW = np.load(os.path.join(GOLDEN, 'bert_tiny.npz'))
from reference_path import reference_root
REF = reference_root()
NAMES = ['mp16_ffn', 'peg_ffn', 'mixed']


def _gm():
    spec = importlib.util.spec_from_file_location('make_golden_model', os.path.join(GOLDEN, 'make_golden_model.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture()
def oracle_ops(monkeypatch):
    from oracle_backend import OracleOps
    monkeypatch.setattr(tq_native, '_OPS', OracleOps())
    monkeypatch.setattr(tq_native, 'default_device', lambda: torch.device('cpu'))


@pytest.mark.parametrize('name', NAMES)
def test_own_caller_quant_dict(oracle_ops, name):
    from engine.bert import BertConfig, QuantBertForSequenceClassification
    from quantization.base_quantized_classes import FP32Acts
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators
    quant_dict = _gm().QUANT_DICTS[name]
    cfg = BertConfig(vocab_size=1000, hidden_size=128, num_hidden_layers=2, num_attention_heads=4,
                     intermediate_size=512, max_position_embeddings=64)
    model = QuantBertForSequenceClassification(cfg, method=QMethods.symmetric_uniform,
                                               act_method=QMethods.asymmetric_uniform, n_bits=8, n_bits_act=8,
                                               act_range_method=RangeEstimators.running_minmax)
    model.load_hf_state_dict({k[2:]: torch.from_numpy(GW[k]) for k in GW.files if k.startswith('w.')})
    model.eval()
    model.apply_quant_dict(quant_dict)
    model.set_quant_state(weight_quant=True, act_quant=True)
    ids = [torch.from_numpy(GW['ids'][i]) for i in range(3)]
    with torch.no_grad():
        for b in ids[:-1]:
            model(b, torch.ones_like(b))
        model.fix_ranges()
        mask = torch.ones_like(ids[-1])
        logits, hidden = model(ids[-1], mask), model.encode(ids[-1], mask)
    sites = []
    for mname, m in model.named_modules():              # same filter as the golden generator
        q = getattr(m, 'quantizer', None)
        if q is not None and mname.endswith('activation_quantizer') and q.is_initialized:
            sites.append(q)
    assert len(sites) == int(G[f'{name}.n_act_quantizers'])
    # module order differs between the two callers (attribute names): compare as multisets of (bits, parameters)
    mine = sorted((q.n_bits, tuple(q._delta.detach().numpy().reshape(-1).tolist())) for q in sites)
    ref = sorted((int(G[f'{name}.q{i}.n_bits']), tuple(G[f'{name}.q{i}.delta'].tolist())) for i in range(len(sites)))
    assert mine == ref
    assert np.array_equal(logits.numpy(), G[f'{name}.logits'])
    assert np.array_equal(hidden.numpy(), G[f'{name}.last_hidden'])
    if name == 'mixed':
        assert isinstance(model.classifier.activation_quantizer, FP32Acts)
        assert isinstance(model.classifier.weight_quantizer, FP32Acts)
        assert model.embeddings.word.weight_quantizer.quantizer.n_bits == 4


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'models')), reason='reference checkout not present')
@pytest.mark.parametrize('name', NAMES)
def test_reference_model_file_quant_dict(oracle_ops, name):
    gm = _gm()
    saved = {k: v for k, v in sys.modules.items() if k.split('.')[0] in ('quantization', 'utils', 'models')}
    try:
        qb = gm.import_reference_model(PKG)
        import utils
        assert utils.__file__.startswith(PKG)
        torch.set_grad_enabled(False)
        res, _ = gm.run_quant_dict_config(qb, name, gm.QUANT_DICTS[name], gm.make_hf_model(), gm.make_batches())
    finally:
        torch.set_grad_enabled(True)
        for k in [k for k in sys.modules if k.split('.')[0] in ('quantization', 'utils', 'models')]:
            del sys.modules[k]
        sys.modules.update(saved)
    n = int(G[f'{name}.n_act_quantizers'])
    assert int(res[f'{name}.n_act_quantizers']) == n
    for i in range(n):
        assert int(res[f'{name}.q{i}.n_bits']) == int(G[f'{name}.q{i}.n_bits'])
        assert np.array_equal(res[f'{name}.q{i}.delta'], G[f'{name}.q{i}.delta']), str(G[f'{name}.q{i}.name'])
    assert np.array_equal(res[f'{name}.logits'], G[f'{name}.logits'])
    assert np.array_equal(res[f'{name}.last_hidden'], G[f'{name}.last_hidden'])
