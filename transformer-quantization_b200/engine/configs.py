"""The five BASELINE.json configurations as named recipes on this package's own callers:

    bert_w8a8_sym      config 1  BERT-base W8A8 per-tensor symmetric, seq 128 batch 4 (plumbing check)
    bert_w8a8_asym     config 2  BERT-base W8A8 per-tensor asymmetric, seq 128 batch 32      <- bench.py's workload
    bert_w8a8_peg      config 3  BERT-base W8A8 per-embedding-group activations (K = 6 contiguous groups), batch 32
    bert_w8a8_pegp               the same with range-based permutation of the groups (module path)
    mobilebert_w4a8    config 4  MobileBERT W4A8, seq 128 batch 64
    roberta_w8a8_mse   config 5  RoBERTa-base W8A8, MSE (grid) activation ranges; calibration batches shard over
                                 the ranks, statistics all-reduced in quantization/_dist.py

``build(name, device)`` returns a random-init model of the configuration's architecture (there is no network
for checkpoints) with both quantizer kinds switched on; ``calibrate(model, recipe, batches)`` runs the
reference's calibration protocol (main.py:512-563: optional FP32 pass that collects the per-dim ranges for the
PEG permutation, then range estimation, then ``fix_ranges``).  ``tiny=True`` shrinks every dimension so the CPU
suite can drive the same code with the oracle back-end (tests/test_baseline_configs.py).
"""
from collections import namedtuple

import torch

from engine.bert import BertConfig, QuantBertForSequenceClassification
from engine.mobilebert import MobileBertConfig, QuantMobileBertForSequenceClassification
from quantization.quantizers import QMethods
from quantization.range_estimators import OptMethod, RangeEstimatorBase, RangeEstimators

Recipe = namedtuple('Recipe', 'family act_method n_bits n_bits_act act_range_method act_range_options peg batch seq')

S, A = QMethods.symmetric_uniform, QMethods.asymmetric_uniform
RECIPES = {
    'bert_w8a8_sym': Recipe('bert', S, 8, 8, RangeEstimators.running_minmax, {}, None, 4, 128),
    'bert_w8a8_asym': Recipe('bert', A, 8, 8, RangeEstimators.running_minmax, {}, None, 32, 128),
    'bert_w8a8_peg': Recipe('bert', A, 8, 8, RangeEstimators.current_minmax, {}, ('ng', 6), 32, 128),
    'bert_w8a8_pegp': Recipe('bert', A, 8, 8, RangeEstimators.current_minmax, {}, ('ngp', 6), 32, 128),
    'mobilebert_w4a8': Recipe('mobilebert', A, 4, 8, RangeEstimators.running_minmax, {}, None, 64, 128),
    'roberta_w8a8_mse': Recipe('roberta', A, 8, 8, RangeEstimators.MSE,
                               dict(opt_method=OptMethod.grid, num_candidates=100), None, 32, 128),
}


def _arch(family, tiny):
    if family == 'mobilebert':
        if tiny:
            return MobileBertConfig(vocab_size=500, hidden_size=48, num_hidden_layers=2, num_attention_heads=2,
                                    intermediate_size=48, embedding_size=16, intra_bottleneck_size=24,
                                    num_feedforward_networks=2, max_position_embeddings=40)
        return MobileBertConfig()
    kw = {}
    if family == 'roberta':
        kw = dict(vocab_size=50265, max_position_embeddings=514, type_vocab_size=1, pad_token_id=1,
                  roberta_positions=True)
    if tiny:
        kw.update(vocab_size=500, hidden_size=48, num_hidden_layers=2, num_attention_heads=2, intermediate_size=96,
                  max_position_embeddings=40)
    return BertConfig(**kw)


def build(name, device, tiny=False, seed=0, act_range_options=None):
    """-> (model, recipe): random-init weights, eval mode, weight and activation quantizers on"""
    r = RECIPES[name]
    cls = QuantMobileBertForSequenceClassification if r.family == 'mobilebert' else QuantBertForSequenceClassification
    model = cls(_arch(r.family, tiny), method=S, act_method=r.act_method, n_bits=r.n_bits, n_bits_act=r.n_bits_act,
                weight_range_method=RangeEstimators.current_minmax, act_range_method=r.act_range_method,
                act_range_options=dict(r.act_range_options if act_range_options is None else act_range_options))
    model.init_weights(seed=seed)
    model.to(device).eval()
    model.set_quant_state(weight_quant=True, act_quant=True)
    return model, r


def synthetic_batches(model, recipe, n_batches, seed=1234, batch=None, seq=None):
    """token ids of the configuration's shape; RoBERTa's padding id (1) is avoided (it would shift positions)"""
    c = model.config
    g = torch.Generator().manual_seed(seed)
    lo = 2 if getattr(c, 'roberta_positions', False) else 0
    B, T = batch or recipe.batch, seq or recipe.seq
    return [torch.randint(lo, c.vocab_size, (B, T), generator=g) for _ in range(n_batches)]


def calibrate(model, recipe, batches):
    """the reference's protocol: [FP32 ranges pass for the PEG permutation] -> estimate ranges -> fix ranges"""
    device = next(model.parameters()).device
    with torch.no_grad():
        if recipe.peg is not None:
            kind, k = recipe.peg
            model.set_per_embedding_groups(k, permute=(kind == 'ngp'))
            if kind == 'ngp':
                # main.py:519-530: weights FP32, activation quantizers ON -- the managers are invoked, record the
                # per-dim ranges of their input and pass it through unquantized
                model.full_precision()
                model.set_quant_state(weight_quant=False, act_quant=True)
                for b in batches:
                    b = b.to(device)
                    model(b, torch.ones_like(b))
                model.set_quant_state(weight_quant=True, act_quant=True)
                for m in model.modules():
                    if isinstance(m, RangeEstimatorBase):
                        m.per_group_range_estimation = False
        for b in batches:
            b = b.to(device)
            model(b, torch.ones_like(b))
        model.fix_ranges()
    return model
