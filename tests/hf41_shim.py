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


# ---- RoBERTa (reference models/quantized_roberta.py): same containers under the Roberta* type names
# (the reference's ``specials`` tables match on the exact type), pad-aware position ids, and the
# two-layer classification head.
class RobertaSelfAttention(BertSelfAttention):
    pass


class RobertaSelfOutput(BertSelfOutput):
    pass


class RobertaAttention(BertAttention):
    def __init__(self, c):
        nn.Module.__init__(self)
        self.self = RobertaSelfAttention(c)
        self.output = RobertaSelfOutput(c)


class RobertaLayer(BertLayer):
    def __init__(self, c):
        nn.Module.__init__(self)
        self.chunk_size_feed_forward = 0
        self.seq_len_dim = 1
        self.is_decoder = False
        self.add_cross_attention = False
        self.attention = RobertaAttention(c)
        self.intermediate = BertIntermediate(c)
        self.output = BertOutput(c)


class RobertaEncoder(BertEncoder):
    def __init__(self, c):
        nn.Module.__init__(self)
        self.layer = nn.ModuleList([RobertaLayer(c) for _ in range(c.num_hidden_layers)])


class RobertaEmbeddings(BertEmbeddings):
    def __init__(self, c):
        super().__init__(c)
        self.padding_idx = c.pad_token_id


class RobertaClassificationHead(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.hidden_size)
        self.dropout = nn.Dropout(c.hidden_dropout_prob)
        self.out_proj = nn.Linear(c.hidden_size, c.num_labels)

    def forward(self, features, **kwargs):
        x = features[:, 0, :]
        x = self.dropout(x)
        x = self.dense(x)
        x = torch.tanh(x)
        x = self.dropout(x)
        return self.out_proj(x)


class RobertaModel(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.config = c
        self.embeddings = RobertaEmbeddings(c)
        self.encoder = RobertaEncoder(c)
        self.pooler = None                     # RobertaForSequenceClassification: add_pooling_layer=False


class RobertaForSequenceClassification(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.config = c
        self.num_labels = c.num_labels
        self.roberta = RobertaModel(c)
        self.classifier = RobertaClassificationHead(c)


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

    import transformers.models.roberta.modeling_roberta as mr
    mr.RobertaLayer = RobertaLayer
    mr.RobertaSelfAttention = RobertaSelfAttention
    mr.RobertaSelfOutput = RobertaSelfOutput


def install_mobilebert():
    """MobileBERT (reference models/quantized_mobilebert.py): the installed transformers still has the
    building blocks the reference wraps (NoNorm, BottleneckLayer, FFNLayer, ...), but the container
    forwards lost ``head_mask`` / ``output_attentions``.  Put the 4.1-style container forwards back."""
    install()
    import transformers.models.mobilebert.modeling_mobilebert as mm
    from transformers.modeling_outputs import BaseModelOutput

    def attention_forward(self, query_tensor, key_tensor, value_tensor, layer_input, attention_mask=None,
                          head_mask=None, output_attentions=None):
        self_outputs = self.self(query_tensor, key_tensor, value_tensor, attention_mask, head_mask, output_attentions)
        attention_output = self.output(self_outputs[0], layer_input)
        return (attention_output,) + tuple(self_outputs[1:])

    def encoder_forward(self, hidden_states, attention_mask=None, head_mask=None, output_attentions=False,
                        output_hidden_states=False, return_dict=True):
        for i, layer_module in enumerate(self.layer):
            layer_outputs = layer_module(hidden_states, attention_mask, head_mask[i] if head_mask else None,
                                         output_attentions)
            hidden_states = layer_outputs[0]
        return BaseModelOutput(last_hidden_state=hidden_states, hidden_states=None, attentions=None)

    mm.MobileBertAttention.forward = attention_forward
    mm.MobileBertEncoder.forward = encoder_forward
    return mm


def make_tiny_mobilebert(seed=4):
    """a 2-layer MobileBERT with every structural feature of the real one (bottlenecks, shared key/query
    bottleneck, trigram embeddings, stacked FFNs, NoNorm), HF classes, seeded random weights"""
    mm = install_mobilebert()
    from transformers import MobileBertConfig
    cfg = MobileBertConfig(vocab_size=1000, hidden_size=128, num_hidden_layers=2, num_attention_heads=4,
                           intermediate_size=128, embedding_size=32, intra_bottleneck_size=32,
                           num_feedforward_networks=2, max_position_embeddings=64, trigram_input=True,
                           use_bottleneck=True, key_query_shared_bottleneck=True, normalization_type='no_norm',
                           classifier_activation=False, hidden_dropout_prob=0.0, num_labels=2)
    torch.manual_seed(seed)
    m = mm.MobileBertForSequenceClassification(cfg)
    g = torch.Generator().manual_seed(seed)
    for p_ in m.parameters():                   # explicit values: independent of HF's init code
        p_.data = torch.randn(p_.shape, generator=g) * (0.05 if p_.dim() > 1 else 0.1) + (1.0 if p_.dim() == 1 and p_.shape[0] in (32, 128) and False else 0.0)
    for mod in m.modules():
        if isinstance(mod, mm.NoNorm):
            mod.weight.data = 1.0 + 0.1 * torch.randn(mod.weight.shape, generator=g)
            mod.bias.data = 0.05 * torch.randn(mod.bias.shape, generator=g)
    return m.eval()
