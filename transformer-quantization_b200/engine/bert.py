"""Quantized BERT-family encoder assembled from this package's own hijacked layers.

This is the CALLER the benchmark and the parity tests drive on machines where the reference
checkout (and its ``models/quantized_bert.py``) is not present.  It is not a copy of that file: it
is written directly from the act-quant site census (reference
models/quantized_bert.py:79-86, 135-213, 238-248, 264-280, 283-291, 378-386, 597) and places one
quantizer at each of those sites, in the same order, using the same classes
(QuantLinear / QuantLayerNorm / QuantEmbedding / QuantizedActivation), so per-site ranges and
logits can be compared one-to-one with the reference (tests/test_model_parity.py).

Site letters follow the reference's ``quant_dict`` grammar (main.py:452-491):
  s scores, p probs, c context, g attn-out dense, u residual 1, x LayerNorm 1 (= FFN input),
  h FFN-out dense, y residual 2, z LayerNorm 2; P pooler, C classifier, e embedding sums.
"""
import math

import torch
from torch import nn

from quantization.autoquant_utils import QuantEmbedding, QuantLayerNorm, QuantLinear
from quantization.base_quantized_classes import QuantizedActivation
from quantization.base_quantized_model import QuantizedModel
from utils.per_embd_quant_utils import (hijack_act_quant, hijack_act_quant_modules, hijack_weight_quant,
                                        set_act_quant_axis_and_groups)


class BertConfig:
    """BERT-base defaults (RoBERTa-base: vocab 50265, max_pos 514, type_vocab 1, pad 1)."""

    def __init__(self, vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2, num_labels=2,
                 layer_norm_eps=1e-12, pad_token_id=0, roberta_positions=False):
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.max_position_embeddings = max_position_embeddings
        self.type_vocab_size = type_vocab_size
        self.num_labels = num_labels
        self.layer_norm_eps = layer_norm_eps
        self.pad_token_id = pad_token_id
        self.roberta_positions = roberta_positions


class Embeddings(QuantizedModel):
    def __init__(self, c, **qp):
        super().__init__()
        self.word = QuantEmbedding(c.vocab_size, c.hidden_size, padding_idx=c.pad_token_id, **qp)
        self.position = QuantEmbedding(c.max_position_embeddings, c.hidden_size, **qp)
        self.token_type = QuantEmbedding(c.type_vocab_size, c.hidden_size, **qp)
        self.e_tok = QuantizedActivation(**qp)       # word + token-type sum   (quantized_bert.py:79)
        self.e_pos = QuantizedActivation(**qp)       # + position embeddings   (quantized_bert.py:84)
        self.norm = QuantLayerNorm(c.hidden_size, eps=c.layer_norm_eps, **qp)
        self.pad = c.pad_token_id
        self.roberta_positions = c.roberta_positions
        self.register_buffer('position_ids', torch.arange(c.max_position_embeddings).unsqueeze(0),
                             persistent=False)

    def forward(self, input_ids, token_type_ids=None):
        B, T = input_ids.shape
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        if self.roberta_positions:        # quantized_roberta.py:26-41: positions count non-pad tokens
            mask = input_ids.ne(self.pad).int()
            pos = (torch.cumsum(mask, dim=1).type_as(mask) * mask).long() + self.pad
        else:
            pos = self.position_ids[:, :T]
        e = self.e_tok(self.word(input_ids) + self.token_type(token_type_ids))
        e = self.e_pos(e + self.position(pos))
        return self.norm(e)


class EncoderBlock(QuantizedModel):
    def __init__(self, c, **qp):
        super().__init__()
        d = c.hidden_size
        self.heads = c.num_attention_heads
        self.head_dim = d // c.num_attention_heads
        self.query = QuantLinear(d, d, **qp)
        self.key = QuantLinear(d, d, **qp)
        self.value = QuantLinear(d, d, **qp)
        self.s = QuantizedActivation(**qp)
        self.p = QuantizedActivation(**qp)
        self.c = QuantizedActivation(**qp)
        self.g = QuantLinear(d, d, **qp)
        self.u = QuantizedActivation(**qp)
        self.x = QuantLayerNorm(d, eps=c.layer_norm_eps, **qp)
        self.ffn_in = QuantLinear(d, c.intermediate_size, activation=nn.GELU(), **qp)
        self.h = QuantLinear(c.intermediate_size, d, **qp)
        self.y = QuantizedActivation(**qp)
        self.z = QuantLayerNorm(d, eps=c.layer_norm_eps, **qp)

    def _split(self, t):
        B, T, _ = t.shape
        return t.view(B, T, self.heads, self.head_dim).permute(0, 2, 1, 3)

    def forward(self, hidden, ext_mask):
        q, k, v = self._split(self.query(hidden)), self._split(self.key(hidden)), self._split(self.value(hidden))
        scores = self.s(torch.matmul(q, k.transpose(-1, -2)))       # quantized BEFORE scaling / masking
        scores = scores / math.sqrt(self.head_dim)
        if ext_mask is not None:
            scores = scores + ext_mask
        probs = self.p(torch.softmax(scores, dim=-1))
        ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous()
        ctx = self.c(ctx.view(hidden.shape))
        a = self.x(self.u(self.g(ctx) + hidden))
        return self.z(self.y(self.h(self.ffn_in(a)) + a))


class QuantBertForSequenceClassification(QuantizedModel):
    """Embeddings -> N encoder blocks -> pooler (first token, dense + tanh) -> classifier."""

    def __init__(self, config, **quant_params):
        super().__init__()
        qp = dict(quant_params)
        qp.pop('quant_setup', None)
        qp.pop('quant_dict', None)
        self.config = config
        self.embeddings = Embeddings(config, **qp)
        self.layers = nn.ModuleList([EncoderBlock(config, **qp) for _ in range(config.num_hidden_layers)])
        self.pooler = QuantLinear(config.hidden_size, config.hidden_size, activation=nn.Tanh(), **qp)
        self.classifier = QuantLinear(config.hidden_size, config.num_labels, **qp)

    def encode(self, input_ids, attention_mask=None, token_type_ids=None):
        ext = None
        if attention_mask is not None:
            ext = (1.0 - attention_mask[:, None, None, :].to(torch.float32)) * -10000.0
        h = self.embeddings(input_ids, token_type_ids)
        for blk in self.layers:
            h = blk(h, ext)
        return h

    def forward(self, input_ids, attention_mask=None, token_type_ids=None):
        h = self.encode(input_ids, attention_mask, token_type_ids)
        pooled = self.pooler(h[:, 0])
        logits = self.classifier(pooled)
        if self.config.num_labels == 1:
            logits = torch.clamp(logits, 0.0, 5.0)
        return logits

    # ---- helpers ---------------------------------------------------------------------------------
    def init_weights(self, seed=0, std=0.02):
        """HF-style random init (there is no network for checkpoints): normal(0, std) for Linear /
        Embedding weights, zero biases, LayerNorm (1, 0).  Generated on the CPU generator so the
        same seed gives the same model everywhere."""
        g = torch.Generator().manual_seed(seed)
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Embedding)):
                m.weight.data = (torch.randn(m.weight.shape, generator=g) * std).to(m.weight.device)
                if isinstance(m, nn.Linear) and m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.LayerNorm):
                m.weight.data.fill_(1.0)
                m.bias.data.zero_()
        return self

    def peg_sites(self):
        """sites main.py:378-434 switches to per-embedding(-group) quantization"""
        E = self.embeddings
        sites = [E.e_tok, E.e_pos, E.norm]
        for L in self.layers:
            sites += [L.query, L.key, L.value, L.c, L.g, L.u, L.x, L.h, L.y, L.z]
        return sites

    def set_per_embedding_groups(self, n_groups, permute=False):
        for s in self.peg_sites():
            set_act_quant_axis_and_groups(s, axis=2, n_groups=n_groups, permute=permute)

    def apply_quant_dict(self, quant_dict):
        """Mixed-precision / per-site control with the reference's ``--quant-dict`` grammar (main.py:442-498,
        README.md:160-173): keys are the site letters of this module's docstring, optionally followed by a layer
        index (``'x'`` = every layer, ``'x3'`` = layer 3), ``L`` / ``L<i>`` = every activation quantizer of the
        layer(s), ``Et`` / ``wC`` = weight quantizers of the token embedding / classifier; values: an int (bits),
        ``'fp32'``, ``'per_embd'``, ``'ng<K>'``, ``'ngp<K>'``.  Call before calibration.  E.g. the paper's
        MP-PTQ recipe is ``{'y': 16, 'h': 16, 'x': 16}``, PEG on the FFN sites ``{'y': 'ng6', 'h': 'ng6', 'x': 'ng6'}``."""
        E = self.embeddings
        for site in (E.e_tok, E.e_pos):
            hijack_act_quant(quant_dict, 'e', site)
        hijack_weight_quant(quant_dict, 'Et', E.word)
        for i, L in enumerate(self.layers):
            for letter in 'spcguxhyz':                       # same order as the reference applies them
                site = getattr(L, letter)
                hijack_act_quant(quant_dict, f'{letter}{i}', site)
                hijack_act_quant(quant_dict, letter, site)
            hijack_act_quant_modules(quant_dict, f'L{i}', L)
            hijack_act_quant_modules(quant_dict, 'L', L)
        hijack_act_quant(quant_dict, 'P', self.pooler)
        hijack_act_quant(quant_dict, 'C', self.classifier)
        hijack_act_quant(quant_dict, 'wP', self.pooler)      # sic: the reference routes 'wP' to the ACTIVATION quantizer
        hijack_weight_quant(quant_dict, 'wC', self.classifier)
        return self

    def load_hf_state_dict(self, sd):
        """weights stored under HuggingFace BertForSequenceClassification names"""
        def put(mod, prefix):
            mod.weight.data = torch.as_tensor(sd[prefix + '.weight']).clone().to(mod.weight.device)
            if getattr(mod, 'bias', None) is not None and prefix + '.bias' in sd:
                mod.bias.data = torch.as_tensor(sd[prefix + '.bias']).clone().to(mod.bias.device)

        E = self.embeddings
        put(E.word, 'bert.embeddings.word_embeddings')
        put(E.position, 'bert.embeddings.position_embeddings')
        put(E.token_type, 'bert.embeddings.token_type_embeddings')
        put(E.norm, 'bert.embeddings.LayerNorm')
        for i, L in enumerate(self.layers):
            p = f'bert.encoder.layer.{i}.'
            put(L.query, p + 'attention.self.query')
            put(L.key, p + 'attention.self.key')
            put(L.value, p + 'attention.self.value')
            put(L.g, p + 'attention.output.dense')
            put(L.x, p + 'attention.output.LayerNorm')
            put(L.ffn_in, p + 'intermediate.dense')
            put(L.h, p + 'output.dense')
            put(L.z, p + 'output.LayerNorm')
        put(self.pooler, 'bert.pooler.dense')
        put(self.classifier, 'classifier')
        return self

    def act_quantizers(self):
        """activation QuantizationManagers in the reference's module order (for per-site checks)"""
        out = []
        E = self.embeddings
        out += [E.e_tok, E.e_pos, E.norm]
        for L in self.layers:
            out += [L.query, L.key, L.value, L.s, L.p, L.c, L.g, L.u, L.x, L.ffn_in, L.h, L.y, L.z]
        out += [self.pooler, self.classifier]
        return [m.activation_quantizer for m in out]
