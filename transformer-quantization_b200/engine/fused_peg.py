"""Fused fixed-range engine for the per-embedding-group (PEG) configuration (BASELINE config 3).

Reference: ``main.py:378-434`` switches the embedding sums / LayerNorm outputs / Q, K, V / context / dense outputs /
residual sums of every layer to per-embedding-group activation quantizers (``utils/per_embd_quant_utils.py:54-68``:
``axis=2, n_groups=K``), scores, probabilities, the FFN intermediate, pooler and classifier stay per-tensor.

Built from a calibrated ``engine.bert.QuantBertForSequenceClassification`` whose PEG sites use K CONTIGUOUS groups
(no range-based permutation) of d / K hidden dimensions, d / K a multiple of 128 (BERT-base: K in {1, 2, 3, 6}).
Five kernels per encoder layer, every tensor between them as x_int bytes:

    tq_linear_peg_qdq_i8          x (PEG) -> Q | K | V, 3 K output segments               -> bf16 centred grid
    tq_attention_peg_qdq_i8       per-head scales (a group holds whole heads), context PEG  -> bytes
    tq_linear_peg_res_ln_qdq_i8   c (PEG) -> dense (PEG) + x (PEG) -> u (PEG) -> LayerNorm -> x' (PEG)
    tq_linear_peg_qdq_i8          x' (PEG) -> FFN-in + GELU (per-tensor)
    tq_linear_peg_res_ln_qdq_i8   f (per-tensor) -> dense (PEG) + x' (PEG) -> y (PEG) -> LayerNorm -> z (PEG)

The A operand's per-group scale / zero point is handled inside the GEMM (group-by-group accumulation, see
csrc/tq_linear.cu namespace peg); per-group OUTPUT quantizers are per-tile constants.  Range-permuted groups
(``ngp``) and anything else outside this shape raise ``UnsupportedByEngine`` -- callers keep the module path.
"""
import torch
from torch import nn

import tq_native
from engine.fused import UnsupportedByEngine, _Site, _Weight, _ln_params, _mgr
from quantization.base_quantized_classes import FP32Acts


class _GroupSite:
    """per-embedding-group activation quantizer with contiguous groups -> one parameter slot per group"""

    def __init__(self, module, D):
        m = module.activation_quantizer
        if isinstance(m, FP32Acts) or not module._quant_a:
            raise UnsupportedByEngine('every activation site must be quantized')
        q = m.quantizer
        if not q.is_initialized or q.n_bits > 8 or q.symmetric:
            raise UnsupportedByEngine('PEG engine needs initialised asymmetric <= 8-bit activation quantizers')
        est = m.range_estimator
        if est is not None and getattr(est, 'ranges', None) is not None:
            raise UnsupportedByEngine('range-permuted groups (ngp): use the module path')
        d = q._delta.detach().reshape(-1).float()
        z = q._zero_float.detach().reshape(-1).float()
        if d.numel() != D:
            raise UnsupportedByEngine('expected a per-embedding (-group) quantizer at this site')
        G = int(m.n_groups) if getattr(m, 'n_groups', None) else D
        if G < 1 or D % G != 0 or (D // G) % 128 != 0 or G > 8:
            raise UnsupportedByEngine('PEG engine needs <= 8 groups of a multiple of 128 hidden dimensions')
        gw = D // G
        dg, zg = d.view(G, gw), z.view(G, gw)
        if not (bool((dg == dg[:, :1]).all()) and bool((zg == zg[:, :1]).all())):
            raise UnsupportedByEngine('quantizer parameters vary inside a group (permuted groups?)')
        ops = tq_native.ops()
        self.q, self.G, self.gw = q, G, gw
        self.delta_g = dg[:, 0].contiguous()
        self.zero_g = zg[:, 0].contiguous()
        self.gspec = ops.spec(self.delta_g, self.zero_g, None, q.n_bits, q.scale_domain == 'log', q.eps)
        self.delta_c, self.zero_c = d.contiguous(), z.contiguous()         # per-column form (embedding kernel)
        self.cspec = ops.spec(self.delta_c, self.zero_c, None, q.n_bits, q.scale_domain == 'log', q.eps)


def _cat_specs(sites, ops):
    q0 = sites[0].q
    delta = torch.cat([s.delta_g for s in sites]).contiguous()
    zero = torch.cat([s.zero_g for s in sites]).contiguous()
    return (delta, zero), ops.spec(delta, zero, None, q0.n_bits, q0.scale_domain == 'log', q0.eps)


class FusedBertPegEngine:
    def __init__(self, model, batch, seq):
        if model.training:
            raise UnsupportedByEngine('engine runs the eval forward')
        cfg = model.config
        self.B, self.T, self.D, self.H = batch, seq, cfg.hidden_size, cfg.num_attention_heads
        self.hd = self.D // self.H
        if seq != 128 or self.hd != 64:
            raise UnsupportedByEngine('engine supports seq 128, head_dim 64')
        self.model = model
        self.num_labels = cfg.num_labels
        self.ops = ops = tq_native.ops()
        self.dev = dev = next(model.parameters()).device
        E = model.embeddings
        if E.roberta_positions:
            raise UnsupportedByEngine('RoBERTa position ids: use the module path')
        D = self.D
        with torch.no_grad():
            self.word_q = E.word.get_params()[0].detach().float().contiguous()
            self.pos_q = E.position.get_params()[0].detach().float().contiguous()
            self.type_q = E.token_type.get_params()[0].detach().float().contiguous()
            self.e_tok, self.e_pos = _GroupSite(E.e_tok, D), _GroupSite(E.e_pos, D)
            self.e_gamma, self.e_beta, self.e_eps = _ln_params(E.norm)
            self.e_out = _GroupSite(E.norm, D)
            G, gw = self.e_out.G, self.e_out.gw
            if self.H % G != 0:
                raise UnsupportedByEngine('a group must hold whole attention heads')
            self.G, self.gw = G, gw
            self.layers = []
            for L in model.layers:
                d = {}
                for name, mod in (('q', L.query), ('k', L.key), ('v', L.value), ('c', L.c), ('g', L.g), ('u', L.u), ('x', L.x),
                                  ('h', L.h), ('y', L.y), ('z', L.z)):
                    d[name] = _GroupSite(mod, D)
                    if d[name].G != G:
                        raise UnsupportedByEngine('all PEG sites must use the same number of groups')
                d['s'], d['p'], d['f'] = _Site(_mgr(L.s)), _Site(_mgr(L.p)), _Site(_mgr(L.ffn_in))
                if not isinstance(L.ffn_in.activation_function, nn.GELU):
                    raise UnsupportedByEngine('FFN activation must be nn.GELU')
                d['wqkv'], d['wg'], d['wf'], d['wh'] = _Weight([L.query, L.key, L.value]), _Weight([L.g]), _Weight([L.ffn_in]), _Weight([L.h])
                for key, groups in (('wqkv', G), ('wg', G), ('wf', G), ('wh', 1)):
                    w = d[key]
                    gi = w.grid.to(torch.int32)
                    w.grp_rowsum = gi.view(w.N, groups, w.K // groups).sum(dim=2, dtype=torch.int32).t().contiguous()
                # Q | K | V: 3 G output segments, the weight scale of each projection repeated over its G segments
                d['qkv_keep'], d['qkv_out'] = _cat_specs([d['q'], d['k'], d['v']], ops)
                w = d['wqkv']
                d['wqkv_delta'] = w.seg_delta.repeat_interleave(G).contiguous()
                q0 = L.query.weight_quantizer.quantizer
                d['wqkv_spec'] = ops.spec(d['wqkv_delta'], None, w._signed, q0.n_bits, q0.scale_domain == 'log', q0.eps)
                d['ln1'], d['ln2'] = _ln_params(L.x), _ln_params(L.z)
                self.layers.append(d)
        M = batch * seq
        self.M = M
        u8 = dict(dtype=torch.uint8, device=dev)
        self.x8 = torch.empty(M, D, **u8)
        self.c8 = torch.empty(M, D, **u8)
        self.a8 = torch.empty(M, D, **u8)
        self.f8 = torch.empty(M, cfg.intermediate_size, **u8)
        self.qkv = torch.empty(M, 3 * D, dtype=torch.bfloat16, device=dev)
        if cfg.intermediate_size % 128 != 0 or D % 128 != 0 or D // 128 > 8:
            raise UnsupportedByEngine('hidden / intermediate sizes must be multiples of 128 (hidden <= 1024)')
        self.i8 = True

    @torch.no_grad()
    def forward(self, input_ids, attention_mask=None, token_type_ids=None):
        ops = self.ops
        B, T, D, H, M, G, gw = self.B, self.T, self.D, self.H, self.M, self.G, self.gw
        assert tuple(input_ids.shape) == (B, T)
        mask = None
        if attention_mask is not None:
            mask = ((1.0 - attention_mask.to(torch.float32)) * -10000.0).contiguous()
        ids = input_ids.reshape(-1).contiguous()
        tt = token_type_ids.reshape(-1).contiguous() if token_type_ids is not None else None
        x, c, a, f = self.x8, self.c8, self.a8, self.f8
        ops.embed_ln_qdq_i8(ids, tt, None, T, self.word_q, self.type_q, self.pos_q, self.e_tok.cspec, D, self.e_pos.cspec, D,
                            self.e_gamma, self.e_beta, self.e_eps, self.e_out.cspec, D, x)
        x_site = self.e_out
        for d in self.layers:
            w = d['wqkv']
            ops.linear_peg_i8(x, w.grid8, w.grp_rowsum, w.bias, M, w.N, w.K, x_site.gspec, G, d['wqkv_spec'], 3 * G, d['qkv_out'],
                              3 * G, gw, 0, out_ctr=self.qkv)
            ops.attention_peg_i8(self.qkv, B, T, H, self.hd, d['q'].gspec, d['k'].gspec, d['v'].gspec, G, d['s'].spec,
                                 d['p'].spec, d['c'].gspec, G, mask, c)
            w = d['wg']
            g1, b1, e1 = d['ln1']
            ops.linear_peg_res_ln_i8(c, w.grid8, w.grp_rowsum, w.bias, M, w.N, w.K, d['c'].gspec, G, w.seg_spec, 1, d['g'].gspec, G,
                                     x, x_site.gspec, G, d['u'].gspec, G, g1, b1, e1, d['x'].gspec, G, gw, a)
            w = d['wf']
            ops.linear_peg_i8(a, w.grid8, w.grp_rowsum, w.bias, M, w.N, w.K, d['x'].gspec, G, w.seg_spec, 1, d['f'].spec, 1, w.N, 1,
                              out_i8=f)
            w = d['wh']
            g2, b2, e2 = d['ln2']
            ops.linear_peg_res_ln_i8(f, w.grid8, w.grp_rowsum, w.bias, M, w.N, w.K, d['f'].spec, 1, w.seg_spec, 1, d['h'].gspec, G,
                                     a, d['x'].gspec, G, d['y'].gspec, G, g2, b2, e2, d['z'].gspec, G, gw, x)
            x_site = d['z']
        # head: the first token of every sequence through the module path's pooler / classifier (2 tiny GEMMs)
        first = self.hidden_states()[:, 0].contiguous()
        logits = self.model.classifier(self.model.pooler(first))
        if self.num_labels == 1:
            logits = torch.clamp(logits, 0.0, 5.0)
        return logits

    __call__ = forward

    def hidden_states(self):
        """dequantized output of the last encoder block of the most recent forward"""
        z = self.layers[-1]['z']
        zp = torch.clamp(torch.round(z.zero_c), 0, 2 ** z.q.n_bits - 1)
        scale = torch.clamp(z.delta_c, min=z.q.eps)
        return ((self.x8.float() - zp) * scale).view(self.B, self.T, self.D)
