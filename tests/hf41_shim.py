"""TEST-ONLY compatibility shim: HuggingFace-4.1-style BERT container modules.

The reference's ``models/quantized_bert.py`` was written against transformers~=4.1; the installed
transformers 5.5 dropped ``apply_chunking_to_forward`` from ``modeling_utils``, changed
``get_extended_attention_mask`` / ``get_head_mask`` and rewrote the container forwards (SURVEY.md
Appendix C).  This module provides attribute-compatible containers and patches the few names the
reference imports, so the UNCHANGED reference model file can be executed (a) against the
reference's own ``quantization`` package to produce golden outputs and (b) against this repo's
``quantization`` package to prove it is a drop-in.  Used by tests/golden/make_golden_model.py and
tests/test_reference_models.py only.
"""
import math
import types

import torch
from torch import nn
from torch.nn import functional as F


class TinyBertConfig:
    def __init__(self, vocab_size=1000, hidden_size=128, num_hidden_layers=2, num_attention_heads=4,
                 intermediate_size=512, max_position_embeddings=64, type_vocab_size=2, num_labels=2,
                 layer_norm_eps=1e-12, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                 pad_token_id=0):
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.max_position_embeddings = max_position_embeddings
        self.type_vocab_size = type_vocab_size
        self.num_labels = num_labels
        self.layer_norm_eps = layer_norm_eps
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.pad_token_id = pad_token_id
        self.output_attentions = False
        self.output_hidden_states = False
        self.use_return_dict = True
        self.is_decoder = False
        self.add_cross_attention = False
        self.chunk_size_feed_forward = 0
        self.position_embedding_type = 'absolute'
        self.initializer_range = 0.02


class BertEmbeddings(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.word_embeddings = nn.Embedding(c.vocab_size, c.hidden_size, padding_idx=c.pad_token_id)
        self.position_embeddings = nn.Embedding(c.max_position_embeddings, c.hidden_size)
        self.token_type_embeddings = nn.Embedding(c.type_vocab_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)
        self.register_buffer('position_ids', torch.arange(c.max_position_embeddings).expand((1, -1)))
        self.position_embedding_type = 'absolute'


class BertSelfAttention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.num_attention_heads = c.num_attention_heads
        self.attention_head_size = c.hidden_size // c.num_attention_heads
        self.all_head_size = c.hidden_size
        self.query = nn.Linear(c.hidden_size, c.hidden_size)
        self.key = nn.Linear(c.hidden_size, c.hidden_size)
        self.value = nn.Linear(c.hidden_size, c.hidden_size)
        self.dropout = nn.Dropout(c.attention_probs_dropout_prob)
        self.position_embedding_type = 'absolute'


class BertSelfOutput(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)


class BertAttention(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.self = BertSelfAttention(c)
        self.output = BertSelfOutput(c)

    def forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, past_key_value=None, output_attentions=False):
        self_outputs = self.self(hidden_states, attention_mask, head_mask, encoder_hidden_states,
                                 encoder_attention_mask, past_key_value, output_attentions)
        attention_output = self.output(self_outputs[0], hidden_states)
        return (attention_output,) + self_outputs[1:]


class BertIntermediate(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.intermediate_size)
        self.intermediate_act_fn = F.gelu


class BertOutput(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.intermediate_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)


class BertLayer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.chunk_size_feed_forward = 0
        self.seq_len_dim = 1
        self.is_decoder = False
        self.add_cross_attention = False
        self.attention = BertAttention(c)
        self.intermediate = BertIntermediate(c)
        self.output = BertOutput(c)


class _EncoderOut(tuple):
    hidden_states = None
    attentions = None
    cross_attentions = None


class BertEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer(c) for _ in range(c.num_hidden_layers)])

    def forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, output_attentions=False, output_hidden_states=False,
                return_dict=True):
        for i, layer in enumerate(self.layer):
            hidden_states = layer(hidden_states, attention_mask, head_mask[i] if head_mask else None,
                                  encoder_hidden_states, encoder_attention_mask, None, output_attentions)[0]
        return _EncoderOut((hidden_states,))


class BertPooler(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.hidden_size)
        self.activation = nn.Tanh()


class BertModel(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.config = c
        self.embeddings = BertEmbeddings(c)
        self.encoder = BertEncoder(c)
        self.pooler = BertPooler(c)


class BertForSequenceClassification(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.config = c
        self.num_labels = c.num_labels
        self.bert = BertModel(c)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)
        self.classifier = nn.Linear(c.hidden_size, c.num_labels)


def init_weights(model, seed=0, std=0.02):
    """HF-style random init: normal(0, std) Linear / Embedding weights, zero biases, LN (1, 0)."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, (nn.Linear, nn.Embedding)):
            m.weight.data = torch.randn(m.weight.shape, generator=g) * std
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.LayerNorm):
            m.weight.data.fill_(1.0)
            m.bias.data.zero_()
    return model


def perturb(model, seed=1):
    """make biases / LN parameters non-trivial so parity tests exercise them"""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.Linear) and m.bias is not None:
            m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.05
        elif isinstance(m, nn.LayerNorm):
            m.weight.data = 1.0 + torch.randn(m.weight.shape, generator=g) * 0.1
            m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.05
    return model


def install():
    """Patch the names the reference's models/quantized_bert.py imports from transformers."""
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    import transformers.models.bert.modeling_bert as mb

    if not hasattr(mu, 'apply_chunking_to_forward'):
        mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    mb.BertLayer = BertLayer
    mb.BertSelfAttention = BertSelfAttention
    mb.BertSelfOutput = BertSelfOutput

    def get_extended_attention_mask(self, attention_mask, input_shape, device=None, dtype=None):
        m = attention_mask[:, None, None, :].to(torch.float32)
        return (1.0 - m) * -10000.0

    def get_head_mask(self, head_mask, num_hidden_layers, is_attention_chunked=False):
        return [None] * num_hidden_layers

    mu.ModuleUtilsMixin.get_extended_attention_mask = get_extended_attention_mask
    mu.ModuleUtilsMixin.get_head_mask = get_head_mask
