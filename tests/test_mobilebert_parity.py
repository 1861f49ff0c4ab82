"""This package's own MobileBERT caller (engine/mobilebert.py, BASELINE config 4 family: W4A8 / W8A8 with
QuantNoNorm, bottlenecks, shared key/query bottleneck, trigram embeddings, stacked FFNs) against the golden
outputs of the UNMODIFIED reference model file (tests/golden/mobilebert_tiny.npz, produced by
tests/golden/make_golden_model.py from models/quantized_mobilebert.py).

CPU: oracle arithmetic back-end + the same torch library ops as the reference -> logits and final hidden
states must be EXACTLY equal (same quantizers at the same sites in the same order).  The weights come from
tests/hf41_shim.make_tiny_mobilebert (seeded HuggingFace MobileBERT; needs `transformers`, not the reference).
"""
import os

import numpy as np
import pytest
import torch

import tq_native
from conftest import GOLDEN
from quantization.quantizers import QMethods
from quantization.range_estimators import RangeEstimators

G = np.load(os.path.join(GOLDEN, 'mobilebert_tiny.npz'))
CONFIGS = {'mobilebert_w4a8': 4, 'mobilebert_w8a8': 8}


def tiny_model(n_bits, device):
    import hf41_shim
    from engine.mobilebert import MobileBertConfig, QuantMobileBertForSequenceClassification
    hf = hf41_shim.make_tiny_mobilebert()
    cfg = MobileBertConfig(vocab_size=1000, hidden_size=128, num_hidden_layers=2, num_attention_heads=4,
                           intermediate_size=128, embedding_size=32, intra_bottleneck_size=32,
                           num_feedforward_networks=2, max_position_embeddings=64)
    model = QuantMobileBertForSequenceClassification(
        cfg, method=QMethods.symmetric_uniform, act_method=QMethods.asymmetric_uniform, n_bits=n_bits, n_bits_act=8,
        act_range_method=RangeEstimators.running_minmax)
    model.load_hf_state_dict(hf.state_dict())
    return model.to(device).eval()


def run(name, device):
    model = tiny_model(CONFIGS[name], device)
    ids = [torch.from_numpy(G['ids'][i]).to(device) for i in range(3)]
    mask = torch.ones_like(ids[0])
    model.set_quant_state(weight_quant=True, act_quant=True)
    with torch.no_grad():
        for b in ids[:-1]:
            model(b, mask)
        model.fix_ranges()
        return model, model(ids[-1], mask), model.encode(ids[-1], mask)


@pytest.mark.parametrize('name', list(CONFIGS))
def test_mobilebert_caller_matches_reference_cpu(name, monkeypatch):
    from oracle_backend import OracleOps
    monkeypatch.setattr(tq_native, '_OPS', OracleOps())
    monkeypatch.setattr(tq_native, 'default_device', lambda: torch.device('cpu'))
    model, logits, hidden = run(name, 'cpu')
    n = sum(1 for m in model.modules() if getattr(m, 'quantizer', None) is not None and m.quantizer.is_initialized)
    assert n == int(G[f'{name}.n_quantizers'])
    assert np.array_equal(logits.numpy(), G[f'{name}.logits'])
    assert np.array_equal(hidden.numpy(), G[f'{name}.last_hidden'])


@pytest.mark.gpu
@pytest.mark.parametrize('name', list(CONFIGS))
def test_mobilebert_caller_matches_reference_gpu(name):
    """the same caller on the CUDA back-end (QuantNoNorm's uncached weight / bias QDQs, W4 grids, ReLU and
    bottleneck GEMMs on tcgen05): GEMM tolerance of DESIGN.md section 3"""
    model, logits, hidden = run(name, 'cuda')
    n = sum(1 for m in model.modules() if getattr(m, 'quantizer', None) is not None and m.quantizer.is_initialized)
    assert n == int(G[f'{name}.n_quantizers'])
    ref_logits, ref_hidden = G[f'{name}.logits'], G[f'{name}.last_hidden']
    step = float(model.classifier.activation_quantizer.quantizer.delta.max())
    assert np.abs(logits.cpu().numpy() - ref_logits).max() <= 3 * step + 1e-6
    hstep = float((ref_hidden.max() - ref_hidden.min()) / 255.0)
    dh = np.abs(hidden.cpu().numpy() - ref_hidden)
    assert dh.max() <= 6 * hstep and (dh > 0.5 * hstep).mean() < 0.05
