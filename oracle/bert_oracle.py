"""CPU ORACLE (test infrastructure, NOT product code) -- the reference's quantized BERT forward
restated for the host CPU, used as the model-level checker and as the CPU baseline / reference arm
of bench.py.

What is restated, and from where:
* the quantize -> round -> clamp -> dequantize chain exactly as the reference issues it -- one
  torch CPU op per step, ``x / scale``, ``round``, ``+ zero_point``, ``clamp``, ``- zero_point``,
  ``* scale`` (quantization/quantizers.py:142-153, 184-185, 209) -- so its cost is the cost of the
  reference's own CPU path (6 passes + temporaries);
* set_quant_range (quantizers.py:263-282 asymmetric, 334-344 symmetric) and the running / current
  min-max estimators (range_estimators.py:142-143, 205-214);
* the placement of the 13 activation quantizers per encoder layer + 3 embedding + pooler +
  classifier sites (models/quantized_bert.py:79-86, 135-213, 238-248, 264-280, 283-291, 378-386,
  597) and the cached fake-quantized weights (quantization/hijacker.py:72-86).
GEMMs, LayerNorm, softmax, GELU, tanh and the embedding lookup are the same torch CPU library
calls the reference makes (autoquant_utils.py:20-21, 58-66, 77-85).

Parity status: PINNED by tests/test_model_parity.py::test_oracle_model_matches_golden against
tests/golden/bert_tiny.npz (outputs of the unmodified reference model), bit-exact.
"""
import math

import torch
from torch.nn import functional as F


class Site:
    """One quantizer site: asymmetric or symmetric per-tensor fake quantizer + min-max estimator."""

    def __init__(self, n_bits=8, symmetric=False, estimator='running_minmax', momentum=0.9, eps=1e-8):
        self.n_bits, self.symmetric, self.estimator, self.momentum, self.eps = n_bits, symmetric, estimator, momentum, eps
        self.xmin = self.xmax = None
        self.delta = self.zero_float = None
        self.signed = None
        self.fixed = False
        self.record = None          # (dict, key): tools/parity_fullsize.py captures the site's output

    def _set_range(self, x_min, x_max):
        x_min = torch.min(x_min, torch.zeros_like(x_min))                      # quantizers.py:258
        x_max = torch.max(x_max, torch.ones_like(x_max) * self.eps)           # :259
        if self.symmetric:
            self.signed = bool((x_min.min() < 0).item())                       # :336
            int_max = 2.0 ** (self.n_bits - int(self.signed)) - 1
            self.delta = torch.max(x_min.abs(), x_max) / int_max               # :338-339
        else:
            self.delta = (x_max - x_min) / (2.0 ** self.n_bits - 1)            # :276
            self.zero_float = -x_min / self.delta                              # :277

    def _estimate(self, x):
        mn, mx = torch.min(x), torch.max(x)                                    # range_estimators.py:206-207
        if self.estimator == 'running_minmax' and self.xmin is not None:
            mn = (1 - self.momentum) * mn + self.momentum * self.xmin          # :213-214
            mx = (1 - self.momentum) * mx + self.momentum * self.xmax
        self.xmin, self.xmax = mn, mx
        self._set_range(mn, mx)

    def __call__(self, x):
        if not self.fixed:
            self._estimate(x)
        scale = torch.clamp(self.delta, min=self.eps)                          # quantizers.py:144
        if self.symmetric:
            int_min = -(2.0 ** (self.n_bits - 1)) if self.signed else 0
            int_max = 2.0 ** (self.n_bits - int(self.signed)) - 1
            zero_point = 0.0
        else:
            int_min, int_max = 0.0, 2.0 ** self.n_bits - 1
            zero_point = torch.clamp(torch.round(self.zero_float), int_min, int_max)   # :151-152
        x_int = torch.round(x / scale) + zero_point                            # :184
        x_int = torch.clamp(x_int, int_min, int_max)                           # :185
        y = scale * (x_int - zero_point)                                       # :209
        if self.record is not None:
            self.record[0][self.record[1]] = y
        return y


class OracleBert:
    """Functional BERT-for-sequence-classification over a HuggingFace-named state dict."""

    def __init__(self, sd, n_layers, n_heads, n_bits=8, n_bits_act=8, sym_acts=False, eps_ln=1e-12,
                 act_estimator='running_minmax', device=None):
        self.sd = {k: v.float() for k, v in sd.items()}
        self.device = device        # None: host CPU (the baseline); a CUDA device runs the same op chain on cuBLAS
        self.L, self.H = n_layers, n_heads
        self.n_bits, self.eps_ln = n_bits, eps_ln
        mk = lambda: Site(n_bits_act, sym_acts, act_estimator)
        self.act = {}
        for name in ['e_tok', 'e_pos', 'e_ln', 'pool', 'cls']:
            self.act[name] = mk()
        for i in range(n_layers):
            for s in ['q', 'k', 'v', 's', 'p', 'c', 'g', 'u', 'x', 'f', 'h', 'y', 'z']:
                self.act[f'{i}.{s}'] = mk()
        self.wq = {}          # cached fake-quantized weights (hijacker.py:72-86)
        self.fixed = False

    def _w(self, key):
        if key not in self.wq:
            s = Site(self.n_bits, True, 'current_minmax')       # weights: symmetric, current min-max
            self.wq[key] = s(self.sd[key])
        return self.wq[key]

    def fix_ranges(self):
        for s in self.act.values():
            s.fixed = True

    def _lin(self, x, p):
        return F.linear(x.contiguous(), self._w(p + '.weight').contiguous(), self.sd.get(p + '.bias'))

    def _ln(self, x, p):
        d = x.shape[-1]
        return F.layer_norm(x.contiguous(), (d,), self._w(p + '.weight').contiguous(),
                            self.sd[p + '.bias'].contiguous(), self.eps_ln)

    def encode(self, ids, mask=None):
        A = self.act
        B, T = ids.shape
        tt = torch.zeros_like(ids)
        e = F.embedding(ids, self._w('bert.embeddings.word_embeddings.weight')) + \
            F.embedding(tt, self._w('bert.embeddings.token_type_embeddings.weight'))
        e = A['e_tok'](e)
        e = e + F.embedding(torch.arange(T, device=self.device).unsqueeze(0), self._w('bert.embeddings.position_embeddings.weight'))
        e = A['e_pos'](e)
        h = A['e_ln'](self._ln(e, 'bert.embeddings.LayerNorm'))
        ext = None if mask is None else (1.0 - mask[:, None, None, :].float()) * -10000.0
        d = h.shape[-1]
        hd = d // self.H
        split = lambda t: t.view(B, T, self.H, hd).permute(0, 2, 1, 3)
        for i in range(self.L):
            p = f'bert.encoder.layer.{i}.'
            q = split(A[f'{i}.q'](self._lin(h, p + 'attention.self.query')))
            k = split(A[f'{i}.k'](self._lin(h, p + 'attention.self.key')))
            v = split(A[f'{i}.v'](self._lin(h, p + 'attention.self.value')))
            s = A[f'{i}.s'](torch.matmul(q, k.transpose(-1, -2)))
            s = s / math.sqrt(hd)
            if ext is not None:
                s = s + ext
            pr = A[f'{i}.p'](torch.softmax(s, dim=-1))
            c = torch.matmul(pr, v).permute(0, 2, 1, 3).contiguous().view(B, T, d)
            c = A[f'{i}.c'](c)
            g = A[f'{i}.g'](self._lin(c, p + 'attention.output.dense'))
            a = A[f'{i}.x'](self._ln(A[f'{i}.u'](g + h), p + 'attention.output.LayerNorm'))
            f = A[f'{i}.f'](F.gelu(self._lin(a, p + 'intermediate.dense')))
            o = A[f'{i}.h'](self._lin(f, p + 'output.dense'))
            h = A[f'{i}.z'](self._ln(A[f'{i}.y'](o + a), p + 'output.LayerNorm'))
        return h

    def __call__(self, ids, mask=None):
        h = self.encode(ids, mask)
        pooled = self.act['pool'](torch.tanh(self._lin(h[:, 0], 'bert.pooler.dense')))
        return self.act['cls'](self._lin(pooled, 'classifier'))


def random_bert_state_dict(vocab=30522, hidden=768, layers=12, inter=3072, max_pos=512, type_vocab=2,
                           num_labels=2, seed=0, std=0.02):
    """HF-style random init in the exact parameter order of engine.bert.QuantBertForSequenceClassification
    .init_weights, so both sides draw identical weights from the same seed."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, o, i):
        sd[name + '.weight'] = torch.randn(o, i, generator=g) * std
        sd[name + '.bias'] = torch.zeros(o)

    def emb(name, n, d):
        sd[name + '.weight'] = torch.randn(n, d, generator=g) * std

    def ln(name, d):
        sd[name + '.weight'] = torch.ones(d)
        sd[name + '.bias'] = torch.zeros(d)

    emb('bert.embeddings.word_embeddings', vocab, hidden)
    emb('bert.embeddings.position_embeddings', max_pos, hidden)
    emb('bert.embeddings.token_type_embeddings', type_vocab, hidden)
    ln('bert.embeddings.LayerNorm', hidden)
    for i in range(layers):
        p = f'bert.encoder.layer.{i}.'
        lin(p + 'attention.self.query', hidden, hidden)
        lin(p + 'attention.self.key', hidden, hidden)
        lin(p + 'attention.self.value', hidden, hidden)
        lin(p + 'attention.output.dense', hidden, hidden)
        ln(p + 'attention.output.LayerNorm', hidden)
        lin(p + 'intermediate.dense', inter, hidden)
        lin(p + 'output.dense', hidden, inter)
        ln(p + 'output.LayerNorm', hidden)
    lin('bert.pooler.dense', hidden, hidden)
    lin('classifier', num_labels, hidden)
    return sd
