"""Quantized MobileBERT assembled from this package's own hijacked layers (BASELINE config 4: W4A8).

Like ``engine/bert.py`` this is a CALLER for machines without the reference checkout, written from the
site census of the reference's models/quantized_mobilebert.py (:58-72 QuantNoNorm, :75-163 embeddings,
:166-270 self-attention, :273-311 self-output, :327-362 output bottleneck, :365-404 output, :407-448
bottleneck layer / FFN output, :451-462 stacked FFN, :465-545 layer, :548-566 pooler, :669-760 head) on top of
HuggingFace's MobileBERT structure (bottlenecks, shared key/query bottleneck, trigram embeddings, stacked
FFNs, NoNorm).  One quantizer per site, same order, same classes, so calibrated ranges and logits compare
one-to-one with the reference's goldens (tests/test_mobilebert_parity.py).

Per layer (``use_bottleneck`` and ``key_query_shared_bottleneck``, the published configuration):

    b_in   = NoNorm(dense(h))                      layer input      (hidden -> true_hidden)
    b_att  = NoNorm(dense(h))                      shared q/k input (hidden -> true_hidden)
    q, k = dense(b_att); v = dense(h)              scores QDQ -> / sqrt(d) + mask -> softmax -> probs QDQ
    ctx    = QDQ(probs @ v)
    a      = NoNorm(QDQ(dense(ctx) + b_in))
    a      = NoNorm(QDQ(dense(act(dense(a))) + a))           x (num_feedforward_networks - 1)
    o      = NoNorm(QDQ(dense(act(dense(a))) + a))
    h_next = NoNorm(QDQ(dense(o) + h))             output bottleneck (true_hidden -> hidden)
"""
import math

import torch
from torch import nn
from torch.nn import functional as F

from quantization.autoquant_utils import QuantEmbedding, QuantLinear
from quantization.base_quantized_classes import QuantizedActivation
from quantization.base_quantized_model import QuantizedModel
from quantization.hijacker import QuantizationHijacker


class MobileBertConfig:
    """google/mobilebert-uncased defaults"""

    def __init__(self, vocab_size=30522, hidden_size=512, num_hidden_layers=24, num_attention_heads=4,
                 intermediate_size=512, embedding_size=128, intra_bottleneck_size=128, num_feedforward_networks=4,
                 max_position_embeddings=512, type_vocab_size=2, num_labels=2, pad_token_id=0, trigram_input=True,
                 hidden_act='relu', classifier_activation=False):
        self.vocab_size, self.hidden_size, self.num_hidden_layers = vocab_size, hidden_size, num_hidden_layers
        self.num_attention_heads, self.intermediate_size = num_attention_heads, intermediate_size
        self.embedding_size, self.true_hidden_size = embedding_size, intra_bottleneck_size
        self.num_feedforward_networks = num_feedforward_networks
        self.max_position_embeddings, self.type_vocab_size = max_position_embeddings, type_vocab_size
        self.num_labels, self.pad_token_id, self.trigram_input = num_labels, pad_token_id, trigram_input
        self.hidden_act, self.classifier_activation = hidden_act, classifier_activation


class QuantNoNorm(QuantizationHijacker):
    """MobileBERT's NoNorm (elementwise ``x * weight + bias``) with fake-quantized parameters and output.
    Both parameters go through the SAME weight quantizer, weight first, bias second (reference
    quantized_mobilebert.py:58-72): while ranges are being estimated the range seen last -- the bias's -- is the
    one that gets fixed."""

    def __init__(self, features, **quant_params):
        super().__init__(**quant_params)
        self.weight = nn.Parameter(torch.ones(features))
        self.bias = nn.Parameter(torch.zeros(features))

    def forward(self, x, offsets=None):
        weight, bias = self.weight, self.bias
        if self._quant_w:
            weight = self.weight_quantizer(weight)
            bias = self.weight_quantizer(bias)
        return self.quantize_activations(x * weight + bias)


def _act(name):
    return {'relu': nn.ReLU, 'gelu': nn.GELU}[name]()


class _DenseNoNorm(QuantizedModel):
    """dense -> NoNorm (bottleneck layers)"""

    def __init__(self, d_in, d_out, **qp):
        super().__init__()
        self.dense = QuantLinear(d_in, d_out, **qp)
        self.norm = QuantNoNorm(d_out, **qp)

    def forward(self, x):
        return self.norm(self.dense(x))


class _ResidualNoNorm(QuantizedModel):
    """dense -> + residual -> QDQ -> NoNorm (self-output, FFN output, layer output, output bottleneck)"""

    def __init__(self, d_in, d_out, **qp):
        super().__init__()
        self.dense = QuantLinear(d_in, d_out, **qp)
        self.res = QuantizedActivation(**qp)
        self.norm = QuantNoNorm(d_out, **qp)

    def forward(self, x, residual):
        return self.norm(self.res(self.dense(x) + residual))


class _FFN(QuantizedModel):
    def __init__(self, c, **qp):
        super().__init__()
        self.intermediate = QuantLinear(c.true_hidden_size, c.intermediate_size, activation=_act(c.hidden_act), **qp)
        self.output = _ResidualNoNorm(c.intermediate_size, c.true_hidden_size, **qp)

    def forward(self, a):
        return self.output(self.intermediate(a), a)


class Embeddings(QuantizedModel):
    def __init__(self, c, **qp):
        super().__init__()
        self.trigram = c.trigram_input
        self.word = QuantEmbedding(c.vocab_size, c.embedding_size, padding_idx=c.pad_token_id, **qp)
        self.position = QuantEmbedding(c.max_position_embeddings, c.hidden_size, **qp)
        self.token_type = QuantEmbedding(c.type_vocab_size, c.hidden_size, **qp)
        self.transformation = QuantLinear(c.embedding_size * (3 if c.trigram_input else 1), c.hidden_size, **qp)
        self.e_pos = QuantizedActivation(**qp)        # input + position embeddings   (:159)
        self.e_tok = QuantizedActivation(**qp)        # + token-type embeddings       (:160)
        self.norm = QuantNoNorm(c.hidden_size, **qp)
        self.register_buffer('position_ids', torch.arange(c.max_position_embeddings).unsqueeze(0), persistent=False)

    def forward(self, input_ids, token_type_ids=None):
        T = input_ids.shape[1]
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        e = self.word(input_ids)
        if self.trigram:                  # token t sees the embeddings of t+1 | t | t-1 (zero padded at the ends)
            e = torch.cat([F.pad(e[:, 1:], [0, 0, 0, 1, 0, 0], value=0), e,
                           F.pad(e[:, :-1], [0, 0, 1, 0, 0, 0], value=0)], dim=2)
        e = self.transformation(e)
        pos = self.position(self.position_ids[:, :T])
        tok = self.token_type(token_type_ids)
        return self.norm(self.e_tok(self.e_pos(e + pos) + tok))


class Layer(QuantizedModel):
    def __init__(self, c, **qp):
        super().__init__()
        d, t = c.hidden_size, c.true_hidden_size
        self.heads, self.head_dim = c.num_attention_heads, c.true_hidden_size // c.num_attention_heads
        self.b_in = _DenseNoNorm(d, t, **qp)
        self.b_att = _DenseNoNorm(d, t, **qp)
        self.query = QuantLinear(t, t, **qp)
        self.key = QuantLinear(t, t, **qp)
        self.value = QuantLinear(d, t, **qp)
        self.s = QuantizedActivation(**qp)
        self.p = QuantizedActivation(**qp)
        self.c = QuantizedActivation(**qp)
        self.attn_out = _ResidualNoNorm(t, t, **qp)
        self.ffn = nn.ModuleList([_FFN(c, **qp) for _ in range(c.num_feedforward_networks - 1)])
        self.intermediate = QuantLinear(t, c.intermediate_size, activation=_act(c.hidden_act), **qp)
        self.output = _ResidualNoNorm(c.intermediate_size, t, **qp)
        self.out_bottleneck = _ResidualNoNorm(t, d, **qp)

    def _split(self, x):
        B, T, _ = x.shape
        return x.view(B, T, self.heads, self.head_dim).permute(0, 2, 1, 3)

    def forward(self, h, ext_mask):
        layer_input = self.b_in(h)
        shared = self.b_att(h)
        q, k, v = self._split(self.query(shared)), self._split(self.key(shared)), self._split(self.value(h))
        scores = self.s(torch.matmul(q, k.transpose(-1, -2))) / math.sqrt(self.head_dim)
        if ext_mask is not None:
            scores = scores + ext_mask
        probs = self.p(torch.softmax(scores, dim=-1))
        ctx = self.c(torch.matmul(probs, v))                    # quantized in (B, H, T, d) layout (:257-258)
        ctx = ctx.permute(0, 2, 1, 3).contiguous().view(layer_input.shape)
        a = self.attn_out(ctx, layer_input)
        for ffn in self.ffn:
            a = ffn(a)
        o = self.output(self.intermediate(a), a)
        return self.out_bottleneck(o, h)


class QuantMobileBertForSequenceClassification(QuantizedModel):
    """Embeddings -> N layers -> first-token pooling (optionally dense + tanh) -> classifier."""

    def __init__(self, config, **quant_params):
        super().__init__()
        qp = dict(quant_params)
        qp.pop('quant_setup', None)
        qp.pop('quant_dict', None)
        self.config = config
        self.embeddings = Embeddings(config, **qp)
        self.layers = nn.ModuleList([Layer(config, **qp) for _ in range(config.num_hidden_layers)])
        self.pooler = (QuantLinear(config.hidden_size, config.hidden_size, activation=nn.Tanh(), **qp)
                       if config.classifier_activation else None)
        self.classifier = QuantLinear(config.hidden_size, config.num_labels, **qp)

    def encode(self, input_ids, attention_mask=None, token_type_ids=None):
        ext = None
        if attention_mask is not None:
            ext = (1.0 - attention_mask[:, None, None, :].to(torch.float32)) * -10000.0
        h = self.embeddings(input_ids, token_type_ids)
        for layer in self.layers:
            h = layer(h, ext)
        return h

    def forward(self, input_ids, attention_mask=None, token_type_ids=None):
        first = self.encode(input_ids, attention_mask, token_type_ids)[:, 0]
        pooled = self.pooler(first) if self.pooler is not None else first
        return self.classifier(pooled)

    # ---- helpers -------------------------------------------------------------------------------------------
    def load_hf_state_dict(self, sd):
        """weights stored under HuggingFace MobileBertForSequenceClassification names"""
        def put(mod, prefix):
            mod.weight.data = torch.as_tensor(sd[prefix + '.weight']).clone().to(mod.weight.device)
            if getattr(mod, 'bias', None) is not None and prefix + '.bias' in sd:
                mod.bias.data = torch.as_tensor(sd[prefix + '.bias']).clone().to(mod.bias.device)

        def put_pair(block, prefix):          # dense + NoNorm
            put(block.dense, prefix + '.dense')
            put(block.norm, prefix + '.LayerNorm')

        E, p = self.embeddings, 'mobilebert.embeddings.'
        put(E.word, p + 'word_embeddings')
        put(E.position, p + 'position_embeddings')
        put(E.token_type, p + 'token_type_embeddings')
        put(E.transformation, p + 'embedding_transformation')
        put(E.norm, p + 'LayerNorm')
        for i, L in enumerate(self.layers):
            p = f'mobilebert.encoder.layer.{i}.'
            put_pair(L.b_in, p + 'bottleneck.input')
            put_pair(L.b_att, p + 'bottleneck.attention')
            put(L.query, p + 'attention.self.query')
            put(L.key, p + 'attention.self.key')
            put(L.value, p + 'attention.self.value')
            put_pair(L.attn_out, p + 'attention.output')
            for j, ffn in enumerate(L.ffn):
                put(ffn.intermediate, p + f'ffn.{j}.intermediate.dense')
                put_pair(ffn.output, p + f'ffn.{j}.output')
            put(L.intermediate, p + 'intermediate.dense')
            put_pair(L.output, p + 'output')
            put_pair(L.out_bottleneck, p + 'output.bottleneck')
        if self.pooler is not None:
            put(self.pooler, 'mobilebert.pooler.dense')
        put(self.classifier, 'classifier')
        return self

    def init_weights(self, seed=0, std=0.02):
        """seeded random init on the CPU generator (no network for checkpoints): normal(0, std) for Linear /
        Embedding weights, zero biases, NoNorm (1, 0)"""
        g = torch.Generator().manual_seed(seed)
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Embedding)):
                m.weight.data = (torch.randn(m.weight.shape, generator=g) * std).to(m.weight.device)
                if isinstance(m, nn.Linear) and m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, QuantNoNorm):
                # NOT the (1, 0) of an untrained NoNorm: both parameters go through ONE quantizer whose fixed range is
                # the one seen last -- the bias's (reference quantized_mobilebert.py:58-72).  An all-zero bias would give
                # a degenerate range that clamps the weight to ~0 and with it every activation of the network; trained
                # checkpoints have biases of the weights' magnitude, so the synthetic model gets them too.
                m.weight.data = (1.0 + 0.1 * torch.randn(m.weight.shape, generator=g)).to(m.weight.device)
                m.bias.data = (0.5 * torch.randn(m.bias.shape, generator=g)).to(m.bias.device)
        return self
