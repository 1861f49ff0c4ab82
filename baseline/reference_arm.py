"""Reference arm of bench.py: the UNMODIFIED reference (``baseline/_ref``, installed by
tools/install_reference.sh -- git-ignored, travels to the GPU box) on the host CPU.

Its own ``quantization`` package (quantizers, estimators, manager, hijacker) and its own
``models/quantized_bert.py`` are loaded from ``baseline/_ref`` and driven through the reference's public API:
``QuantizedBertForSequenceClassification(hf_model, **qparams)`` -> ``set_quant_state`` -> calibration
forward -> ``fix_ranges`` -> eval forwards.  Nothing of this repo's package is on that path.  The only
foreign code is tests/hf41_shim.py: HuggingFace-4.1-style container modules, because the reference was
written against transformers 4.1 and the image has 5.5.

Weights: ``oracle.bert_oracle.random_bert_state_dict(seed=0)`` -- the same tensors, drawn in the same
order, as ``engine.bert.QuantBertForSequenceClassification.init_weights(seed=0)`` gives the GPU arm.
"""
import importlib.util
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('TQ_REFERENCE') or os.path.join(ROOT, 'baseline', '_ref')
if not os.path.isdir(os.path.join(REF, 'models')) and os.path.isdir('/root/reference/models'):
    REF = '/root/reference'


def available():
    return os.path.isdir(os.path.join(REF, 'models')) and os.path.isdir(os.path.join(REF, 'quantization'))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def import_reference():
    """-> the reference's models.quantized_bert module, with ``quantization`` / ``utils`` / ``models`` all
    resolved inside the reference tree (this repo's same-named packages are evicted from sys.modules)."""
    tests = os.path.join(ROOT, 'tests')
    if tests not in sys.path:
        sys.path.insert(0, tests)
    import hf41_shim
    hf41_shim.install()
    for k in [k for k in sys.modules if k.split('.')[0] in ('quantization', 'utils', 'models')]:
        del sys.modules[k]
    pkg = os.path.join(ROOT, 'transformer-quantization_b200')
    saved_path = list(sys.path)
    sys.path[:] = [REF] + [p for p in sys.path if os.path.abspath(p or '.') != pkg]
    try:
        import quantization  # noqa: F401
        assert os.path.abspath(quantization.__file__).startswith(os.path.abspath(REF)), quantization.__file__
        # the reference's utils/__init__ pulls in datasets / click glue that needs the network-era
        # dependencies: expose only the two modules its model files import
        u = types.ModuleType('utils')
        u.__path__ = [os.path.join(REF, 'utils')]
        sys.modules['utils'] = u
        tb = _load('utils.tb_utils', os.path.join(REF, 'utils', 'tb_utils.py'))
        pe = _load('utils.per_embd_quant_utils', os.path.join(REF, 'utils', 'per_embd_quant_utils.py'))
        for n in ('_tb_advance_global_step', '_tb_advance_token_counters', '_tb_hist'):
            setattr(u, n, getattr(tb, n))
        for n in ('set_act_quant_axis_and_groups', 'hijack_act_quant', 'hijack_weight_quant', 'hijack_act_quant_modules'):
            setattr(u, n, getattr(pe, n))
        m = types.ModuleType('models')
        m.__path__ = [os.path.join(REF, 'models')]
        sys.modules['models'] = m
        qb = _load('models.quantized_bert', os.path.join(REF, 'models', 'quantized_bert.py'))
    finally:
        sys.path[:] = saved_path
    return qb, hf41_shim


class ReferenceBert:
    """BERT-base W8A8 (sym weights current_minmax / asym activations running_minmax) on the reference's code."""

    def __init__(self, state_dict, n_layers=12):
        qb, shim = import_reference()
        from quantization.quantizers import QMethods
        from quantization.range_estimators import RangeEstimators
        cfg = shim.TinyBertConfig(vocab_size=30522, hidden_size=768, num_hidden_layers=n_layers, num_attention_heads=12,
                                  intermediate_size=3072, max_position_embeddings=512, hidden_dropout_prob=0.0,
                                  attention_probs_dropout_prob=0.0)
        hf = shim.BertForSequenceClassification(cfg)
        missing, unexpected = hf.load_state_dict(state_dict, strict=False)
        assert not unexpected and all('position_ids' in k for k in missing), (missing, unexpected)
        qparams = dict(method=QMethods.symmetric_uniform, act_method=QMethods.asymmetric_uniform, n_bits=8, n_bits_act=8,
                       per_channel_weights=False, percentile=None, quant_setup='all',
                       weight_range_method=RangeEstimators.current_minmax, weight_range_options={},
                       act_range_method=RangeEstimators.running_minmax, act_range_options={}, quant_dict={})
        self.model = qb.QuantizedBertForSequenceClassification(hf, **qparams)
        self.model.eval()
        self.model.set_quant_state(weight_quant=True, act_quant=True)
        self.source = os.path.abspath(qb.__file__)

    def __call__(self, ids, mask):
        return self.model(input_ids=ids, attention_mask=mask, return_dict=True).logits

    def fix_ranges(self):
        self.model.fix_ranges()
