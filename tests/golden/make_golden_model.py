#!/usr/bin/env python
"""Model-level golden outputs: the UNMODIFIED reference ``models/quantized_bert.py`` +
``quantization/*`` (from /root/reference) run on CPU through the HF-4.1 compatibility shim
(tests/hf41_shim.py) on a tiny BERT.  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_model.py

Writes tests/golden/bert_tiny.npz: weights, token ids and, per configuration, the logits, the
final hidden states and every activation quantizer's (delta, zero_float) after calibration; and
tests/golden/roberta_tiny.npz: the same for ``models/quantized_roberta.py`` (TQ_GOLDEN_ONLY=roberta
regenerates only that file); tests/golden/bert_tiny_qat.npz (TQ_GOLDEN_ONLY=qat): loss and
gradients of one training step with learnable ranges; tests/golden/bert_tiny_pegp.npz (TQ_GOLDEN_ONLY=pegp): the
range-permuted PEG configuration with main.py's FP32 ranges pass (activation quantizers on);
tests/golden/bert_tiny_quant_dict.npz (TQ_GOLDEN_ONLY=quant_dict): mixed-precision / per-site `quant_dict` recipes.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.dont_write_bytecode = True
sys.path.insert(0, TESTS)
from reference_path import reference_root  # noqa: E402
REF = reference_root()


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def import_reference_model(package_root):
    """import ``models.quantized_bert`` from the reference with ``quantization`` / ``utils``
    resolved under ``package_root`` (the reference itself, or this repo's package)."""
    import hf41_shim
    hf41_shim.install()
    for k in [k for k in sys.modules if k.split('.')[0] in ('quantization', 'utils', 'models')]:
        del sys.modules[k]
    sys.path.insert(0, package_root)
    try:
        import quantization  # noqa: F401
        if package_root == REF:
            # the reference's utils/__init__ drags in datasets / click glue: expose only what the
            # model files use
            u = types.ModuleType('utils')
            u.__path__ = [os.path.join(REF, 'utils')]
            sys.modules['utils'] = u
            tb = load_by_path('utils.tb_utils', os.path.join(REF, 'utils', 'tb_utils.py'))
            pe = load_by_path('utils.per_embd_quant_utils', os.path.join(REF, 'utils', 'per_embd_quant_utils.py'))
            for n in ('_tb_advance_global_step', '_tb_advance_token_counters', '_tb_hist'):
                setattr(u, n, getattr(tb, n))
            for n in ('set_act_quant_axis_and_groups', 'hijack_act_quant', 'hijack_weight_quant',
                      'hijack_act_quant_modules'):
                setattr(u, n, getattr(pe, n))
        else:
            import utils  # noqa: F401
        m = types.ModuleType('models')
        m.__path__ = [os.path.join(REF, 'models')]
        sys.modules['models'] = m
        qb = load_by_path('models.quantized_bert', os.path.join(REF, 'models', 'quantized_bert.py'))
        qb.roberta = load_by_path('models.quantized_roberta', os.path.join(REF, 'models', 'quantized_roberta.py'))
    finally:
        sys.path.remove(package_root)
    return qb


CONFIGS = {
    # name: (weight method, act method, n_bits, n_bits_act, act estimator, quant_dict-style PEG spec)
    'w8a8_asym': dict(method='symmetric_uniform', act_method='asymmetric_uniform', n_bits=8, n_bits_act=8,
                      act_range_method='running_minmax', peg=None),
    'w8a8_sym': dict(method='symmetric_uniform', act_method='symmetric_uniform', n_bits=8, n_bits_act=8,
                     act_range_method='running_minmax', peg=None),
    'w4a8_asym': dict(method='symmetric_uniform', act_method='asymmetric_uniform', n_bits=4, n_bits_act=8,
                      act_range_method='running_minmax', peg=None),
    'w8a8_peg4': dict(method='symmetric_uniform', act_method='asymmetric_uniform', n_bits=8, n_bits_act=8,
                      act_range_method='running_minmax', peg=('ng', 4)),
    'w8a8_pegp4': dict(method='symmetric_uniform', act_method='asymmetric_uniform', n_bits=8, n_bits_act=8,
                       act_range_method='current_minmax', peg=('ngp', 4)),
    'w8a8_mse': dict(method='symmetric_uniform', act_method='asymmetric_uniform', n_bits=8, n_bits_act=8,
                     act_range_method='MSE', peg=None),
}

# the reference's models/quantized_roberta.py (BASELINE config 5 family): running min-max and MSE ranges
ROBERTA_CONFIGS = {
    'roberta_w8a8': dict(method='symmetric_uniform', act_method='asymmetric_uniform', n_bits=8, n_bits_act=8,
                         act_range_method='running_minmax', peg=None),
    'roberta_w8a8_mse': dict(method='symmetric_uniform', act_method='asymmetric_uniform', n_bits=8, n_bits_act=8,
                             act_range_method='MSE', peg=None),
}


def peg_sites(model):
    """the sites main.py:378-434 switches to per-embedding(-group) quantization"""
    E = model.bert.embeddings
    sites = [E.sum_input_token_type_embd_act_quantizer, E.sum_pos_embd_act_quantizer, E.LayerNorm]
    for L in model.bert.encoder.layer:
        A, S, O = L.attention.self, L.attention.output, L.output
        sites += [A.query, A.key, A.value, A.context_act_quantizer, S.dense, S.res_act_quantizer,
                  S.LayerNorm, O.dense, O.res_act_quantizer, O.LayerNorm]
    return sites


def run_config(qb, name, cfg, hf_model, batches, main_py_ranges_pass=False):
    """quantize -> (FP32 ranges pass) -> calibrate on batches[:-1] -> fix -> eval batches[-1].
    ``main_py_ranges_pass``: run the FP32 pass exactly like main.py:519-530 does -- weights FP32, activation
    quantizers ON (pass_data_for_range_estimation(act_quant=True, weight_quant=False)), so the managers are
    invoked and record the per-dim ranges; without it (the protocol of bert_tiny.npz) the quantizers are off
    during that pass, no ranges are recorded and the groups are formed without permutation."""
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators, RangeEstimatorBase
    from utils import set_act_quant_axis_and_groups

    qparams = dict(method=QMethods[cfg['method']], act_method=QMethods[cfg['act_method']],
                   n_bits=cfg['n_bits'], n_bits_act=cfg['n_bits_act'], per_channel_weights=False,
                   percentile=None, quant_setup='all',
                   weight_range_method=RangeEstimators.current_minmax, weight_range_options={},
                   act_range_method=RangeEstimators[cfg['act_range_method']], act_range_options={},
                   quant_dict={})
    model = qb.QuantizedBertForSequenceClassification(hf_model, **qparams)
    model.eval()
    model.set_quant_state(weight_quant=True, act_quant=True)
    if cfg['peg']:
        kind, k = cfg['peg']
        for s in peg_sites(model):
            set_act_quant_axis_and_groups(s, axis=2, n_groups=k, permute=(kind == 'ngp'))
        if kind == 'ngp':           # main.py:519-537: FP32 pass to collect per-dim ranges
            model.full_precision()
            if main_py_ranges_pass:
                model.set_quant_state(weight_quant=False, act_quant=True)
            for b in batches[:-1]:
                model(input_ids=b, attention_mask=torch.ones_like(b))
            model.set_quant_state(weight_quant=True, act_quant=True)
            for m in model.modules():
                if isinstance(m, RangeEstimatorBase):
                    m.per_group_range_estimation = False
    with torch.no_grad():
        for b in batches[:-1]:
            model(input_ids=b, attention_mask=torch.ones_like(b))
        model.fix_ranges()
        out = model(input_ids=batches[-1], attention_mask=torch.ones_like(batches[-1]), return_dict=True)
        hidden = model.bert(batches[-1], attention_mask=torch.ones_like(batches[-1]), return_dict=True)
    res = {f'{name}.logits': out.logits.numpy().copy(),
           f'{name}.last_hidden': hidden.last_hidden_state.numpy().copy(),
           f'{name}.pooled': hidden.pooler_output.numpy().copy()}
    i = 0
    for mname, m in model.named_modules():
        q = getattr(m, 'quantizer', None)
        if q is not None and mname.endswith('activation_quantizer') and q.is_initialized:
            res[f'{name}.q{i}.delta'] = q._delta.detach().numpy().reshape(-1).copy()
            if getattr(q, '_zero_float', None) is not None:
                res[f'{name}.q{i}.zero_float'] = q._zero_float.detach().numpy().reshape(-1).copy()
            res[f'{name}.q{i}.name'] = np.array(mname)
            i += 1
    res[f'{name}.n_act_quantizers'] = np.array(i)
    return res, model


def run_roberta_config(qb, name, cfg, hf_model, batches):
    """models/quantized_roberta.py: quantize -> calibrate on batches[:-1] -> fix -> eval batches[-1]"""
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators

    qparams = dict(method=QMethods[cfg['method']], act_method=QMethods[cfg['act_method']],
                   n_bits=cfg['n_bits'], n_bits_act=cfg['n_bits_act'], per_channel_weights=False,
                   percentile=None, quant_setup='all',
                   weight_range_method=RangeEstimators.current_minmax, weight_range_options={},
                   act_range_method=RangeEstimators[cfg['act_range_method']], act_range_options={},
                   quant_dict={})
    model = qb.roberta.QuantizedRobertaForSequenceClassification(hf_model, **qparams)
    model.eval()
    model.set_quant_state(weight_quant=True, act_quant=True)
    with torch.no_grad():
        for b in batches[:-1]:
            model(input_ids=b, attention_mask=torch.ones_like(b))
        model.fix_ranges()
        out = model(input_ids=batches[-1], attention_mask=torch.ones_like(batches[-1]), return_dict=True)
        hidden = model.roberta(batches[-1], attention_mask=torch.ones_like(batches[-1]), return_dict=True)
    res = {f'{name}.logits': out.logits.numpy().copy(),
           f'{name}.last_hidden': hidden.last_hidden_state.numpy().copy()}
    n = 0
    for mname, m in model.named_modules():
        q = getattr(m, 'quantizer', None)
        if q is not None and mname.endswith('activation_quantizer') and q.is_initialized:
            n += 1
    res[f'{name}.n_act_quantizers'] = np.array(n)
    return res, model


MOBILEBERT_CONFIGS = {
    # BASELINE config 4 family: 4-bit symmetric weights, 8-bit asymmetric activations
    'mobilebert_w4a8': dict(method='symmetric_uniform', act_method='asymmetric_uniform', n_bits=4, n_bits_act=8,
                            act_range_method='running_minmax'),
    'mobilebert_w8a8': dict(method='symmetric_uniform', act_method='asymmetric_uniform', n_bits=8, n_bits_act=8,
                            act_range_method='running_minmax'),
}


def import_reference_mobilebert(qb):
    """models/quantized_mobilebert.py from the reference, on whatever quantization / utils packages
    ``import_reference_model`` has put in place"""
    import hf41_shim
    hf41_shim.install_mobilebert()
    u = sys.modules['utils']
    if not hasattr(u, 'DotDict'):
        ut = load_by_path('utils.utils', os.path.join(REF, 'utils', 'utils.py'))
        u.DotDict = ut.DotDict
    return load_by_path('models.quantized_mobilebert', os.path.join(REF, 'models', 'quantized_mobilebert.py'))


def run_mobilebert_config(qm, name, cfg, hf_model, batches):
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators

    qparams = dict(method=QMethods[cfg['method']], act_method=QMethods[cfg['act_method']],
                   n_bits=cfg['n_bits'], n_bits_act=cfg['n_bits_act'], per_channel_weights=False,
                   percentile=None, quant_setup='all',
                   weight_range_method=RangeEstimators.current_minmax, weight_range_options={},
                   act_range_method=RangeEstimators[cfg['act_range_method']], act_range_options={},
                   quant_dict={})
    model = qm.QuantizedMobileBertForSequenceClassification(hf_model, **qparams)
    model.eval()
    model.set_quant_state(weight_quant=True, act_quant=True)
    with torch.no_grad():
        for b in batches[:-1]:
            model(input_ids=b, attention_mask=torch.ones_like(b))
        model.fix_ranges()
        out = model(input_ids=batches[-1], attention_mask=torch.ones_like(batches[-1]), return_dict=True)
        hidden = model.mobilebert(batches[-1], attention_mask=torch.ones_like(batches[-1]), return_dict=True)
    n = sum(1 for mname, m in model.named_modules()
            if getattr(m, 'quantizer', None) is not None and m.quantizer.is_initialized)
    return {f'{name}.logits': out.logits.numpy().copy(),
            f'{name}.last_hidden': hidden.last_hidden_state.numpy().copy(),
            f'{name}.n_quantizers': np.array(n)}, model


# mixed-precision / per-site control through `quant_dict` (reference README.md:160-173, main.py:442-498)
QUANT_DICTS = {
    'mp16_ffn': {'y': 16, 'h': 16, 'x': 16},                                  # README: MP-PTQ, 16-bit FFN sites
    'peg_ffn': {'y': 'ng4', 'h': 'ng4', 'x': 'ng4'},                          # README: PEG on the FFN sites only
    # every value kind; the all-modules key (L0) on a layer without FP32 sites (the reference cannot re-bit them)
    'mixed': {'s1': 'fp32', 'p1': 6, 'c': 'per_embd', 'g1': 12, 'u': 'ng2', 'z1': 'fp32', 'e': 16, 'Et': 4,
              'L0': 12, 'P': 16, 'C': 'fp32', 'wC': 'fp32'},
}


def wire_quant_dict(model, quant_dict, utils_mod):
    """what main.py:442-498 does with --quant-dict (layer count taken from the model instead of the literal 12)"""
    hijack_act_quant, hijack_weight_quant = utils_mod.hijack_act_quant, utils_mod.hijack_weight_quant
    hijack_act_quant_modules = utils_mod.hijack_act_quant_modules
    E = model.bert.embeddings
    hijack_act_quant(quant_dict, 'e', E.sum_input_token_type_embd_act_quantizer)
    hijack_act_quant(quant_dict, 'e', E.sum_pos_embd_act_quantizer)
    hijack_weight_quant(quant_dict, 'Et', E.word_embeddings)
    for i, L in enumerate(model.bert.encoder.layer):
        A, S, O = L.attention.self, L.attention.output, L.output
        for key, site in (('s', A.attn_scores_act_quantizer), ('p', A.attn_probs_act_quantizer),
                          ('c', A.context_act_quantizer), ('g', S.dense), ('u', S.res_act_quantizer),
                          ('x', S.LayerNorm), ('h', O.dense), ('y', O.res_act_quantizer), ('z', O.LayerNorm)):
            hijack_act_quant(quant_dict, f'{key}{i}', site)
            hijack_act_quant(quant_dict, key, site)
        hijack_act_quant_modules(quant_dict, f'L{i}', L)
        hijack_act_quant_modules(quant_dict, 'L', L)
    hijack_act_quant(quant_dict, 'P', model.bert.pooler.dense_act[0])
    hijack_act_quant(quant_dict, 'C', model.classifier)
    hijack_act_quant(quant_dict, 'wP', model.bert.pooler.dense_act[0])
    hijack_weight_quant(quant_dict, 'wC', model.classifier)


def run_quant_dict_config(qb, name, quant_dict, hf_model, batches):
    """W8A8 asymmetric running-minmax model with per-site overrides: quantize -> wire -> calibrate -> fix -> eval"""
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators
    import utils as utils_mod
    qparams = dict(method=QMethods.symmetric_uniform, act_method=QMethods.asymmetric_uniform, n_bits=8, n_bits_act=8,
                   per_channel_weights=False, percentile=None, quant_setup='all',
                   weight_range_method=RangeEstimators.current_minmax, weight_range_options={},
                   act_range_method=RangeEstimators.running_minmax, act_range_options={}, quant_dict={})
    model = qb.QuantizedBertForSequenceClassification(hf_model, **qparams)
    model.eval()
    wire_quant_dict(model, quant_dict, utils_mod)
    model.set_quant_state(weight_quant=True, act_quant=True)
    with torch.no_grad():
        for b in batches[:-1]:
            model(input_ids=b, attention_mask=torch.ones_like(b))
        model.fix_ranges()
        out = model(input_ids=batches[-1], attention_mask=torch.ones_like(batches[-1]), return_dict=True)
        hidden = model.bert(batches[-1], attention_mask=torch.ones_like(batches[-1]), return_dict=True)
    res = {f'{name}.logits': out.logits.numpy().copy(), f'{name}.last_hidden': hidden.last_hidden_state.numpy().copy()}
    n = 0
    for mname, m in model.named_modules():
        q = getattr(m, 'quantizer', None)
        if q is not None and mname.endswith('activation_quantizer') and q.is_initialized:
            res[f'{name}.q{n}.delta'] = q._delta.detach().numpy().reshape(-1).copy()
            res[f'{name}.q{n}.n_bits'] = np.array(q.n_bits)
            res[f'{name}.q{n}.name'] = np.array(mname)
            n += 1
    res[f'{name}.n_act_quantizers'] = np.array(n)
    return res, model


QAT_CONFIGS = ('w8a8_asym', 'w4a8_asym', 'w8a8_sym')


def run_qat_step(model, name, batch, labels, seed=77):
    """one quantization-aware training step on a calibrated model with learnable ranges (README
    `train-quantized --learn-ranges`; quantizers.py:284-288): loss and gradients"""
    model.learn_ranges()
    model.train()
    model.zero_grad()
    torch.manual_seed(seed)                       # dropout masks
    with torch.enable_grad():
        out = model(input_ids=batch, attention_mask=torch.ones_like(batch), labels=labels, return_dict=True)
        out.loss.backward()
    res = {f'{name}.qat.loss': out.loss.detach().numpy().copy(),
           f'{name}.qat.logits': out.logits.detach().numpy().copy()}
    n_q = 0
    for pname, p in model.named_parameters():
        if p.grad is None:
            continue
        leaf = pname.split('.')[-1]
        if leaf in ('_delta', '_zero_float'):
            res[f'{name}.qat.grad.{pname}'] = p.grad.numpy().reshape(-1).copy()
            n_q += 1
        elif pname in ('classifier.weight', 'bert.pooler.dense_act.weight', 'bert.encoder.layer.0.attention.self.query.weight',
                       'bert.encoder.layer.1.output.dense.weight', 'bert.embeddings.LayerNorm.weight',
                       'bert.embeddings.position_embeddings.weight'):
            res[f'{name}.qat.grad.{pname}'] = p.grad.numpy().copy()
    res[f'{name}.qat.n_range_params'] = np.array(n_q)
    return res


def make_hf_mobilebert():
    import hf41_shim
    return hf41_shim.make_tiny_mobilebert()


def make_hf_roberta():
    import hf41_shim
    cfg = hf41_shim.TinyBertConfig(pad_token_id=1)
    m = hf41_shim.RobertaForSequenceClassification(cfg)
    hf41_shim.init_weights(m, seed=2)
    hf41_shim.perturb(m, seed=3)
    m.eval()
    return m


def make_hf_model():
    import hf41_shim
    cfg = hf41_shim.TinyBertConfig()
    m = hf41_shim.BertForSequenceClassification(cfg)
    hf41_shim.init_weights(m, seed=0)
    hf41_shim.perturb(m, seed=1)
    m.eval()
    return m


def make_batches(n=3, B=4, T=32, vocab=1000):
    g = torch.Generator().manual_seed(1234)
    return [torch.randint(0, vocab, (B, T), generator=g) for _ in range(n)]


def qat_labels(B=4):
    return torch.tensor([0, 1, 1, 0][:B])


if __name__ == '__main__':
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    qb = import_reference_model(REF)
    hf = make_hf_model()
    batches = make_batches()
    if os.environ.get('TQ_GOLDEN_ONLY', '') == 'quant_dict':
        outd = {}
        for name, qd in QUANT_DICTS.items():
            res, _ = run_quant_dict_config(qb, name, qd, hf, batches)
            outd.update(res)
            print(name, 'logits', res[f'{name}.logits'][0], 'act quantizers', int(res[f'{name}.n_act_quantizers']))
        np.savez_compressed(os.path.join(HERE, 'bert_tiny_quant_dict.npz'), **outd)
        print('bert_tiny_quant_dict.npz', os.path.getsize(os.path.join(HERE, 'bert_tiny_quant_dict.npz')))
        sys.exit(0)
    if os.environ.get('TQ_GOLDEN_ONLY', '') == 'pegp':
        outp = {}
        for name in ('w8a8_pegp4',):
            res, model = run_config(qb, name, CONFIGS[name], hf, batches, main_py_ranges_pass=True)
            outp.update(res)
            n_perm = sum(1 for m in model.modules() if getattr(m, 'ranges', None) is not None)
            outp[f'{name}.n_range_vectors'] = np.array(n_perm)
            print(name, 'logits', res[f'{name}.logits'][0], 'estimators with permutation ranges', n_perm)
        np.savez_compressed(os.path.join(HERE, 'bert_tiny_pegp.npz'), **outp)
        print('bert_tiny_pegp.npz', os.path.getsize(os.path.join(HERE, 'bert_tiny_pegp.npz')))
        sys.exit(0)
    if os.environ.get('TQ_GOLDEN_ONLY', '') == 'qat':
        outq = {}
        for name in QAT_CONFIGS:
            _, model = run_config(qb, name, CONFIGS[name], hf, batches)
            outq.update(run_qat_step(model, name, batches[-1], qat_labels()))
            print(name, 'qat loss', float(outq[f'{name}.qat.loss']), 'range params', int(outq[f'{name}.qat.n_range_params']))
        np.savez_compressed(os.path.join(HERE, 'bert_tiny_qat.npz'), **outq)
        print('bert_tiny_qat.npz', os.path.getsize(os.path.join(HERE, 'bert_tiny_qat.npz')))
        sys.exit(0)
    out = {'ids': np.stack([b.numpy() for b in batches])}
    for k, v in hf.state_dict().items():
        out['w.' + k] = v.numpy().copy()
    for name, cfg in ({} if os.environ.get('TQ_GOLDEN_ONLY', '') in ('roberta', 'mobilebert') else CONFIGS).items():
        res, _ = run_config(qb, name, cfg, hf, batches)
        out.update(res)
        print(name, 'logits', res[f'{name}.logits'][0], 'act quantizers', int(res[f'{name}.n_act_quantizers']))
    if os.environ.get('TQ_GOLDEN_ONLY', '') not in ('roberta', 'mobilebert'):
        np.savez_compressed(os.path.join(HERE, 'bert_tiny.npz'), **out)
        print('bert_tiny.npz', os.path.getsize(os.path.join(HERE, 'bert_tiny.npz')))
    hr = make_hf_roberta()
    outr = {'ids': np.stack([b.numpy() for b in batches])}
    for k, v in hr.state_dict().items():
        outr['w.' + k] = v.numpy().copy()
    for name, cfg in ({} if os.environ.get('TQ_GOLDEN_ONLY', '') == 'mobilebert' else ROBERTA_CONFIGS).items():
        res, _ = run_roberta_config(qb, name, cfg, hr, batches)
        outr.update(res)
        print(name, 'logits', res[f'{name}.logits'][0], 'act quantizers', int(res[f'{name}.n_act_quantizers']))
    if os.environ.get('TQ_GOLDEN_ONLY', '') != 'mobilebert':
        np.savez_compressed(os.path.join(HERE, 'roberta_tiny.npz'), **outr)
        print('roberta_tiny.npz', os.path.getsize(os.path.join(HERE, 'roberta_tiny.npz')))
    qm = import_reference_mobilebert(qb)
    hm = make_hf_mobilebert()
    outm = {'ids': np.stack([b.numpy() for b in batches])}
    for name, cfg in MOBILEBERT_CONFIGS.items():
        res, _ = run_mobilebert_config(qm, name, cfg, hm, batches)
        outm.update(res)
        print(name, 'logits', res[f'{name}.logits'][0], 'quantizers', int(res[f'{name}.n_quantizers']))
    np.savez_compressed(os.path.join(HERE, 'mobilebert_tiny.npz'), **outm)
    print('mobilebert_tiny.npz', os.path.getsize(os.path.join(HERE, 'mobilebert_tiny.npz')))
