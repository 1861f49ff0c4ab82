"""Fused fixed-range inference engine for the quantized BERT encoder.

Built FROM a calibrated ``engine.bert.QuantBertForSequenceClassification`` (ranges fixed, eval
mode): it reads the fake-quantized weights and the per-site quantizer buffers of that model and
runs the same forward -- the same 161 quantizer sites in the same order -- as 5 kernels per encoder
layer, every tensor between them carried as the centred integer grid in bf16 (2 B / element):

    embeddings : tq_embed_ln_qdq_bf16                       (3 sites)
    per layer  : tq_linear_qdq_bf16   fused Q|K|V GEMM, per-column output quantizers      (3 sites)
                 tq_attention_qdq_bf16 scores / probs / context                            (3 sites)
                 tq_linear_res_ln_qdq_bf16 attention-output dense + residual + LayerNorm    (3 sites)
                 tq_linear_qdq_bf16   FFN-in + GELU                                          (1 site)
                 tq_linear_res_ln_qdq_bf16 FFN-out dense + residual + LayerNorm             (3 sites)
    head       : pooler (tanh) and classifier through tq_linear_qdq_bf16                   (2 sites)

(``TQ_ENGINE_FUSE_LN=0`` or a ``trace`` request splits the two residual blocks back into
tq_linear_res_qdq_bf16 + tq_ln_qdq_bf16: 7 kernels per layer.)

8-bit operand mode (the default whenever every activation grid is unsigned, i.e. the BASELINE configuration): the
tensors between the kernels are x_int BYTES, every GEMM runs on the int8 tensor cores with exact int32 accumulation,
and the launches are merged further -- per layer the attention kernel plus ONE encoder-chain launch
(``tq_chain_plan_*``: attention-output + LayerNorm -> FFN-in + GELU -> FFN-out + LayerNorm -> the next layer's Q | K | V,
a 4-CTA cluster per sequence; ``TQ_ENGINE_CHAIN`` = 0 / 1 / 2 selects separate kernels / per-layer chains / the whole
encoder in one launch), and one launch for the classification head (``tq_head_qdq_i8``): 27 launches per forward of
BERT-base.  Every merged form is bit-identical to the separate kernels (tests/test_gpu_engine.py).

The module-level path (quantization.* classes, one kernel per site + library ops) stays the
reference-facing API and the calibration path; this engine is what the throughput benchmark runs.
Supported: per-tensor quantizers with n_bits <= 8 at every site, seq 128, head_dim 64, hidden % 256
== 0.  Anything else raises ``UnsupportedByEngine`` and callers keep using the module path.
"""
import os

import torch
from torch import nn

import tq_native
from quantization.base_quantized_classes import FP32Acts


class UnsupportedByEngine(RuntimeError):
    pass


def _mgr(module):
    m = module.activation_quantizer
    if isinstance(m, FP32Acts) or not module._quant_a:
        raise UnsupportedByEngine('every activation site must be quantized')
    q = m.quantizer
    if not q.is_initialized or q.n_bits > 8 or q.delta.numel() != 1 or q.axis is not None:
        raise UnsupportedByEngine('engine needs initialised per-tensor <= 8-bit activation quantizers')
    return q


def _unsigned_grid(q):
    """integer grid [0, 2^n - 1]: asymmetric quantizers and symmetric ones over non-negative data"""
    return (not q.symmetric) or (not bool(q.signed))


class _Site:
    """per-tensor quantizer -> tq_qspec.  The spec holds RAW device pointers into the quantizer's buffers: the tensors
    are kept here (so the memory cannot be freed under the engine) and ``check()`` detects a quantizer whose buffers
    were re-allocated since (reset_ranges + recalibration, model.to(), make_range_trainable, a shape-changing
    set_quant_range) -- the engine must be rebuilt then."""

    def __init__(self, quantizer):
        self.q = quantizer
        self.spec = quantizer._spec()
        self.keep = (quantizer._delta, getattr(quantizer, '_zero_float', None), getattr(quantizer, '_signed', None))

    def check(self):
        q = self.q
        now = (q._delta, getattr(q, '_zero_float', None), getattr(q, '_signed', None))
        for a, b in zip(self.keep, now):
            if (a is None) != (b is None) or (a is not None and a.data_ptr() != b.data_ptr()):
                raise RuntimeError('fused engine: a quantizer\'s range buffers were re-allocated after the engine was built '
                                   '(recalibration / .to() / trainable ranges) -- build a new engine from the model')


class _ColSite:
    """several per-tensor quantizers side by side along the output columns -> one [N] spec"""

    def __init__(self, quantizers, widths):
        ops = tq_native.ops()
        sym = [q.symmetric for q in quantizers]
        if any(sym) and not all(sym):
            raise UnsupportedByEngine('mixed symmetric / asymmetric quantizers in one fused GEMM')
        self.delta = torch.cat([q.delta.reshape(1).expand(w) for q, w in zip(quantizers, widths)]).contiguous()
        q0 = quantizers[0]
        if all(sym):
            flags = [bool(q.signed) for q in quantizers]
            if len(set(flags)) != 1:
                raise UnsupportedByEngine('symmetric quantizers with different signedness in one fused GEMM')
            self.zero_float = None
            self.spec = ops.spec(self.delta, None, q0._signed, q0.n_bits, q0.scale_domain == 'log', q0.eps)
        else:
            self.zero_float = torch.cat([q.zero_float.reshape(1).expand(w)
                                         for q, w in zip(quantizers, widths)]).contiguous()
            self.spec = ops.spec(self.delta, self.zero_float, None, q0.n_bits, q0.scale_domain == 'log', q0.eps)
        if len({(q.n_bits, q.scale_domain, q.eps) for q in quantizers}) != 1:
            raise UnsupportedByEngine('fused sites must share n_bits / scale_domain / eps')
        self.n = int(sum(widths))
        # one parameter slot per SEGMENT (= per fused quantizer) for tq_linear_seg_qdq_i8
        self.nseg = len(quantizers)
        self.equal_widths = len(set(widths)) == 1
        self.seg_delta = torch.cat([q.delta.reshape(1) for q in quantizers]).contiguous()
        if all(sym):
            self.seg_zero_float = None
            self.seg_spec = ops.spec(self.seg_delta, None, q0._signed, q0.n_bits, q0.scale_domain == 'log', q0.eps)
        else:
            self.seg_zero_float = torch.cat([q.zero_float.reshape(1) for q in quantizers]).contiguous()
            self.seg_spec = ops.spec(self.seg_delta, self.seg_zero_float, None, q0.n_bits, q0.scale_domain == 'log', q0.eps)


class _Weight:
    """bf16 integer grid of one or more fake-quantized Linear weights stacked along N"""

    def __init__(self, layers, pad_to=None):
        ops = tq_native.ops()
        grids, deltas, biases = [], [], []
        q0 = None
        for lin in layers:
            if not lin._quant_w:
                raise UnsupportedByEngine('weights must be quantized')
            qz = lin.weight_quantizer.quantizer
            if not qz.is_initialized or qz.n_bits > 8 or not qz.symmetric or qz.delta.numel() != 1:
                raise UnsupportedByEngine('engine needs per-tensor symmetric <= 8-bit weight quantizers')
            relaxed = getattr(qz, '_relaxed', None)
            if relaxed is not None and relaxed():
                raise UnsupportedByEngine('AdaRound-rounded weights (learned up / down rounding): use the module path')
            q0 = q0 or qz
            if bool(qz.signed) != bool(q0.signed) or qz.n_bits != q0.n_bits:
                raise UnsupportedByEngine('stacked weights must share the integer grid')
            w = lin.weight.detach()
            _, g = ops.quant_int(w, qz._spec(), want_f32=False, want_bf16=True)
            grids.append(g)
            deltas.append(qz.delta.reshape(1).expand(w.shape[0]))
            biases.append(lin.bias.detach() if lin.bias is not None else torch.zeros(w.shape[0], device=w.device))
        self.grid = torch.cat(grids).contiguous()
        self.delta = torch.cat(deltas).contiguous()
        self.bias = torch.cat(biases).contiguous().float()
        if pad_to is not None and self.grid.shape[0] < pad_to:
            n, k = self.grid.shape
            self.grid = torch.cat([self.grid, torch.zeros(pad_to - n, k, dtype=self.grid.dtype, device=self.grid.device)])
            self.delta = torch.cat([self.delta, self.delta[-1:].expand(pad_to - n)]).contiguous()
            self.bias = torch.cat([self.bias, torch.zeros(pad_to - n, device=self.bias.device)])
        self.N, self.K = self.grid.shape
        self._signed = q0._signed
        self.spec = ops.spec(self.delta, None, self._signed, q0.n_bits, q0.scale_domain == 'log', q0.eps)
        # one scale per stacked layer (segment of output columns) for tq_linear_seg_qdq_i8 / the per-tensor form
        self.nseg = len(layers)
        self.seg_delta = torch.cat([lin.weight_quantizer.quantizer.delta.reshape(1) for lin in layers]).contiguous()
        self.seg_spec = ops.spec(self.seg_delta, None, self._signed, q0.n_bits, q0.scale_domain == 'log', q0.eps)
        self.equal_widths = len({lin.weight.shape[0] for lin in layers}) == 1 and pad_to is None
        # 8-bit operand mode: the same integers, one byte each (two's complement when the grid is signed),
        # and their row sums for the zero-point correction of the activations
        gi = self.grid.to(torch.int32)
        self.grid8 = (gi.to(torch.int8) if bool(q0.signed) else gi.to(torch.uint8)).contiguous()
        self.rowsum = gi.sum(dim=1, dtype=torch.int32).contiguous()


def _ln_params(ln):
    """(fake-quantized gamma, beta, eps) of a QuantLayerNorm in eval mode"""
    w, b = ln.get_params()
    return w.detach().float().contiguous(), b.detach().float().contiguous(), float(ln.eps)


class FusedBertEngine:
    def __init__(self, model, batch, seq):
        if model.training:
            raise UnsupportedByEngine('engine runs the eval forward')
        cfg = model.config
        self.B, self.T, self.D, self.H = batch, seq, cfg.hidden_size, cfg.num_attention_heads
        self.hd = self.D // self.H
        if seq != 128 or self.hd != 64 or self.D % 256 != 0:
            raise UnsupportedByEngine('engine supports seq 128, head_dim 64, hidden % 256 == 0')
        self.num_labels = cfg.num_labels
        self.ops = tq_native.ops()
        dev = next(model.parameters()).device
        self.dev = dev
        E = model.embeddings
        if E.roberta_positions:
            raise UnsupportedByEngine('RoBERTa position ids: use the module path')
        with torch.no_grad():
            self.word_q = E.word.get_params()[0].detach().float().contiguous()
            self.pos_q = E.position.get_params()[0].detach().float().contiguous()
            self.type_q = E.token_type.get_params()[0].detach().float().contiguous()
            self.e_tok, self.e_pos = _Site(_mgr(E.e_tok)), _Site(_mgr(E.e_pos))
            self.e_gamma, self.e_beta, self.e_eps = _ln_params(E.norm)
            self.e_out = _Site(_mgr(E.norm))
            self.layers = []
            for L in model.layers:
                d = {}
                d['wqkv'] = _Weight([L.query, L.key, L.value])
                d['qkv_out'] = _ColSite([_mgr(L.query), _mgr(L.key), _mgr(L.value)], [self.D] * 3)
                d['q'], d['k'], d['v'] = _Site(_mgr(L.query)), _Site(_mgr(L.key)), _Site(_mgr(L.value))
                d['s'], d['p'], d['c'] = _Site(_mgr(L.s)), _Site(_mgr(L.p)), _Site(_mgr(L.c))
                d['wg'], d['g'], d['u'] = _Weight([L.g]), _Site(_mgr(L.g)), _Site(_mgr(L.u))
                d['ln1'] = _ln_params(L.x)
                d['x'] = _Site(_mgr(L.x))
                if not isinstance(L.ffn_in.activation_function, nn.GELU):
                    raise UnsupportedByEngine('FFN activation must be nn.GELU')
                d['wf'], d['f'] = _Weight([L.ffn_in]), _Site(_mgr(L.ffn_in))
                d['wh'], d['h'], d['y'] = _Weight([L.h]), _Site(_mgr(L.h)), _Site(_mgr(L.y))
                d['ln2'] = _ln_params(L.z)
                d['z'] = _Site(_mgr(L.z))
                self.layers.append(d)
            self.w_pool, self.pool_out = _Weight([model.pooler]), _Site(_mgr(model.pooler))
            self.w_cls, self.cls_out = _Weight([model.classifier], pad_to=16), _Site(_mgr(model.classifier))
        M, D = batch * seq, self.D
        bf = dict(dtype=torch.bfloat16, device=dev)
        self.x = torch.empty(M, D, **bf)
        self.qkv = torch.empty(M, 3 * D, **bf)
        self.c = torch.empty(M, D, **bf)
        self.u = torch.empty(M, D, **bf)
        self.a = torch.empty(M, D, **bf)
        self.f = torch.empty(M, cfg.intermediate_size, **bf)
        self.yb = torch.empty(M, D, **bf)
        self.M = M
        # residual blocks: LayerNorm fused into the GEMM epilogue (cluster kernel) unless switched off
        self.fuse_ln = os.environ.get('TQ_ENGINE_FUSE_LN', '1') != '0' and D <= 2048
        # 8-bit operand mode: activations travel as x_int bytes and every GEMM runs on the int8 tensor
        # cores (exact int32 accumulation).  Needs unsigned activation grids (asymmetric quantizers, the
        # BASELINE configuration) -- anything else keeps the bf16 carriers.
        sites = [self.e_out, self.pool_out] + [d[k] for d in self.layers for k in ('c', 'x', 'f', 'z')]
        self.i8 = (os.environ.get('TQ_ENGINE_I8', '1') != '0' and self.fuse_ln and D % 128 == 0
                   and cfg.intermediate_size % 128 == 0 and all(_unsigned_grid(st.q) for st in sites))
        if self.i8:
            u8 = dict(dtype=torch.uint8, device=dev)
            self.x8 = torch.empty(M, D, **u8)
            self.c8 = torch.empty(M, D, **u8)
            self.a8 = torch.empty(M, D, **u8)
            self.f8 = torch.empty(M, cfg.intermediate_size, **u8)
            self.first8 = torch.empty(batch, D, **u8)
            self.pooled8 = torch.empty(batch, D, **u8)
        self.ffn_in_bf16 = os.environ.get('TQ_ENGINE_FFN_IN_BF16', '1') != '0'
        # lean int8 kernels (tq_linear_seg_qdq_i8 + the lean form of tq_linear_res_ln_qdq_i8): shapes they cover
        self.lean = (self.i8 and os.environ.get('TQ_ENGINE_LEAN', '1') != '0' and D % 128 == 0
                     and cfg.intermediate_size % 128 == 0
                     and all(st.q.n_bits <= 8 for d in self.layers for st in (d['q'], d['k'], d['v'], d['f'])))
        # classification head (first token -> pooler -> classifier) as one dp4a kernel; needs per-tensor weight quantizers
        self.head = (self.i8 and os.environ.get('TQ_ENGINE_HEAD', '1') != '0' and D % 128 == 0 and D <= 1024
                     and self.w_pool.nseg == 1 and self.w_cls.nseg == 1)
        # chain kernel (tq_chain_plan_*): a cluster of D / 192 CTAs carries one sequence through a list of stages.
        #   TQ_ENGINE_CHAIN=1 (default)  one launch per layer for the four GEMM stages (attention-output + LN, FFN-in, FFN-out
        #                                + LN, next Q|K|V), attention as its own kernel (three CTAs per SM hide its latencies)
        #   TQ_ENGINE_CHAIN=2            the whole encoder in ONE launch, attention as a chain stage (bit-identical; measured
        #                                1.13 ms against 1.07 ms per step: with the attention code in the same kernel every
        #                                GEMM epilogue runs 15-20 % slower and three heads per CTA run one after the other)
        #   TQ_ENGINE_CHAIN=0            every stage its own kernel
        I = cfg.intermediate_size
        mode = int(os.environ.get('TQ_ENGINE_CHAIN', '1'))
        ok = (self.lean and D % 192 == 0 and D // 192 <= 8 and I % 256 == 0 and (I // 256) % (D // 192) == 0
              and (3 * D // 192) % (D // 192) == 0)
        self.chain = mode if ok else 0
        if self.chain == 2 and self.H % (D // 192) != 0:
            self.chain = 1
        if self.chain:
            self.mask_buf = torch.zeros(batch, seq, dtype=torch.float32, device=dev)
            self._build_chains()
        self._last_i8 = False

    def _build_chains(self):
        ops = self.ops
        cs = ops.chain_stage
        D, M, x, c, a, f = self.D, self.M, self.x8, self.c8, self.a8, self.f8
        per_layer, whole = [], []
        x_site = self.e_out
        d0 = self.layers[0]
        w = d0['wqkv']
        whole.append(cs(0, x, w.grid8, w.rowsum, w.bias, self.qkv, w.N, w.K, x_site.spec, w.seg_spec, d0['qkv_out'].seg_spec, 3))
        for li, d in enumerate(self.layers):
            wg, wf, wh = d['wg'], d['wf'], d['wh']
            g1, b1, e1 = d['ln1']
            g2, b2, e2 = d['ln2']
            att = ops.chain_attention_stage(self.qkv, c, D, self.H, d['q'].spec, d['k'].spec, d['v'].spec, d['s'].spec, d['p'].spec,
                                            d['c'].spec, self.mask_buf)
            st = [cs(2, c, wg.grid8, wg.rowsum, wg.bias, a, wg.N, wg.K, d['c'].spec, wg.seg_spec, d['g'].spec, 1, x, x_site.spec,
                     d['u'].spec, d['x'].spec, g1, b1, e1),
                  cs(1, a, wf.grid8, wf.rowsum, wf.bias, f, wf.N, wf.K, d['x'].spec, wf.seg_spec, d['f'].spec),
                  cs(2, f, wh.grid8, wh.rowsum, wh.bias, x, wh.N, wh.K, d['f'].spec, wh.seg_spec, d['h'].spec, 1, a, d['x'].spec,
                     d['y'].spec, d['z'].spec, g2, b2, e2)]
            if li + 1 < len(self.layers):
                n = self.layers[li + 1]
                w = n['wqkv']
                st.append(cs(0, x, w.grid8, w.rowsum, w.bias, self.qkv, w.N, w.K, d['z'].spec, w.seg_spec, n['qkv_out'].seg_spec, 3))
            per_layer.append(st)
            whole.append(att)
            whole.extend(st)
            x_site = d['z']
        if self.chain == 2:
            self.plan = ops.chain_plan(whole, M)
        else:
            self.plans = [ops.chain_plan(st, M) for st in per_layer]

    def _linear(self, a_ctr, a_site, w, act, out_spec, out_params, out_ctr=None, want_f32=False, M=None):
        M = a_ctr.shape[0] if M is None else M
        ops = self.ops
        dev = a_ctr.device
        y = torch.empty(M, w.N, dtype=torch.float32, device=dev) if want_f32 else None
        yc = out_ctr if out_ctr is not None else (None if want_f32 else torch.empty(M, w.N, dtype=torch.bfloat16, device=dev))
        null = tq_native.QSpec(None, None, None, 8, 0, 1e-8)
        ops._run('linear_qdq', 2 * M * w.N * w.K, 1, ops.lib.tq_linear_qdq_bf16, a_ctr.data_ptr(), w.grid.data_ptr(),
                 w.bias.data_ptr(), tq_native._ptr(y), tq_native._ptr(yc), M, w.N, w.K, 1, a_site.spec, w.spec, w.N,
                 int(act), out_spec if out_spec is not None else null, int(out_params), None, None, 0,
                 tq_native._stream())
        return y, yc

    @torch.no_grad()
    def forward(self, input_ids, attention_mask=None, token_type_ids=None, trace=None):
        """``trace`` (tests): dict that receives the dequantized output of every block site, under the
        names of the module path, plus ``<name>.step`` = its quantization step."""
        ops = self.ops

        def rec(name, ctr, site, cols=None):
            if trace is not None:
                t = ctr if cols is None else ctr[:, cols[0]:cols[1]]
                sc = site.q.scale.reshape(())
                trace[name] = (t.float() * sc).view(self.B, self.T, -1).clone()
                trace[name + '.step'] = float(sc)

        B, T, D, H = self.B, self.T, self.D, self.H
        assert tuple(input_ids.shape) == (B, T)
        self.validate()
        # the pre-LayerNorm sums (sites u / y) only exist in the unfused chain: tracing keeps it
        fuse_ln = self.fuse_ln and trace is None
        self._last_i8 = self.i8 and fuse_ln
        if self._last_i8:
            return self._forward_i8(input_ids, attention_mask, token_type_ids)
        mask = None
        if attention_mask is not None:
            mask = ((1.0 - attention_mask.to(torch.float32)) * -10000.0).contiguous()
        ids = input_ids.reshape(-1).contiguous()
        tt = token_type_ids.reshape(-1).contiguous() if token_type_ids is not None else None
        x = self.x
        ops.embed_ln_qdq(ids, tt, None, T, self.word_q, self.type_q, self.pos_q, self.e_tok.spec, 1,
                         self.e_pos.spec, 1, self.e_gamma, self.e_beta, self.e_eps, self.e_out.spec, 1, out_ctr=x)
        x_site = self.e_out
        rec('emb', x, x_site)
        for li, d in enumerate(self.layers):
            self._linear(x, x_site, d['wqkv'], 0, d['qkv_out'].spec, d['qkv_out'].n, out_ctr=self.qkv)
            rec(f'{li}.query', self.qkv, d['q'], (0, D))
            rec(f'{li}.key', self.qkv, d['k'], (D, 2 * D))
            rec(f'{li}.value', self.qkv, d['v'], (2 * D, 3 * D))
            ops.attention(self.qkv, B, T, H, self.hd, d['q'].spec, d['k'].spec, d['v'].spec, d['s'].spec,
                          d['p'].spec, d['c'].spec, mask, out_ctr=self.c)
            rec(f'{li}.c', self.c, d['c'])
            w = d['wg']
            g1, b1, e1 = d['ln1']
            if fuse_ln:      # dense + residual + LayerNorm in one cluster kernel: only the LN output leaves the chip
                ops.linear_res_ln(self.c, w.grid, w.bias, self.M, w.N, w.K, d['c'].spec, w.spec, w.N, d['g'].spec,
                                  x, x_site.spec, d['u'].spec, g1, b1, e1, d['x'].spec, out_ctr=self.a)
            else:
                ops.linear_res(self.c, w.grid, w.bias, self.M, w.N, w.K, d['c'].spec, w.spec, w.N, d['g'].spec, 1,
                               x, x_site.spec, d['u'].spec, 1, out_ctr=self.u)
                rec(f'{li}.u', self.u, d['u'])
                ops.ln_qdq(self.u, d['u'].spec, 1, g1, b1, e1, d['x'].spec, 1, out_ctr=self.a)
            rec(f'{li}.x', self.a, d['x'])
            self._linear(self.a, d['x'], d['wf'], 1, d['f'].spec, 1, out_ctr=self.f)
            rec(f'{li}.ffn_in', self.f, d['f'])
            w = d['wh']
            g2, b2, e2 = d['ln2']
            if fuse_ln:
                ops.linear_res_ln(self.f, w.grid, w.bias, self.M, w.N, w.K, d['f'].spec, w.spec, w.N, d['h'].spec,
                                  self.a, d['x'].spec, d['y'].spec, g2, b2, e2, d['z'].spec, out_ctr=x)
            else:
                ops.linear_res(self.f, w.grid, w.bias, self.M, w.N, w.K, d['f'].spec, w.spec, w.N, d['h'].spec, 1,
                               self.a, d['x'].spec, d['y'].spec, 1, out_ctr=self.yb)
                rec(f'{li}.y', self.yb, d['y'])
                ops.ln_qdq(self.yb, d['y'].spec, 1, g2, b2, e2, d['z'].spec, 1, out_ctr=x)
            x_site = d['z']
            rec(f'{li}.z', x, x_site)
        first = x.view(B, T, D)[:, 0].contiguous()                       # pooler input: first token
        _, pooled = self._linear(first, x_site, self.w_pool, 3, self.pool_out.spec, 1)
        logits, _ = self._linear(pooled, self.pool_out, self.w_cls, 0, self.cls_out.spec, 1, want_f32=True)
        logits = logits[:, :self.num_labels]
        if self.num_labels == 1:
            logits = torch.clamp(logits, 0.0, 5.0)
        return logits

    def _forward_i8(self, input_ids, attention_mask, token_type_ids):
        """the same chain with x_int byte carriers and int8 tensor-core GEMMs (per layer: attention + one encoder-chain launch)"""
        ops = self.ops
        B, T, D, H, M = self.B, self.T, self.D, self.H, self.M
        mask = None
        if attention_mask is not None:
            mask = ((1.0 - attention_mask.to(torch.float32)) * -10000.0).contiguous()
        ids = input_ids.reshape(-1).contiguous()
        tt = token_type_ids.reshape(-1).contiguous() if token_type_ids is not None else None
        x, c, a, f = self.x8, self.c8, self.a8, self.f8
        ops.embed_ln_qdq_i8(ids, tt, None, T, self.word_q, self.type_q, self.pos_q, self.e_tok.spec, 1, self.e_pos.spec, 1,
                            self.e_gamma, self.e_beta, self.e_eps, self.e_out.spec, 1, x)
        x_site = self.e_out
        lean = self.lean
        if self.chain:
            if attention_mask is not None:
                self.mask_buf.copy_(mask.view(B, T))
            else:
                self.mask_buf.zero_()
        if self.chain == 2:          # the whole encoder in one launch
            ops.chain_run(self.plan)
            x_site = self.layers[-1]['z']
        elif self.chain == 1:        # QKV(0), then per layer: attention + one chain launch (attn-out + LN, FFN-in, FFN-out + LN, next QKV)
            d = self.layers[0]
            w = d['wqkv']
            ops.linear_seg_i8(x, w.grid8, w.rowsum, w.bias, M, w.N, w.K, x_site.spec, w.seg_spec, d['qkv_out'].seg_spec, 3, 0,
                              out_ctr=self.qkv)
            for d, plan in zip(self.layers, self.plans):
                ops.attention_i8(self.qkv, B, T, H, self.hd, d['q'].spec, d['k'].spec, d['v'].spec, d['s'].spec, d['p'].spec,
                                 d['c'].spec, self.mask_buf, c)
                ops.chain_run(plan)
            x_site = self.layers[-1]['z']
        for d in (() if self.chain else self.layers):
            w = d['wqkv']
            if lean:     # per-segment quantizers (Q | K | V), lean int8 kernel
                ops.linear_seg_i8(x, w.grid8, w.rowsum, w.bias, M, w.N, w.K, x_site.spec, w.seg_spec, d['qkv_out'].seg_spec,
                                  3, 0, out_ctr=self.qkv)
            else:
                ops.linear_i8(x, w.grid8, w.rowsum, w.bias, M, w.N, w.K, x_site.spec, w.spec, w.N, 0, d['qkv_out'].spec,
                              d['qkv_out'].n, out_ctr=self.qkv)
            ops.attention_i8(self.qkv, B, T, H, self.hd, d['q'].spec, d['k'].spec, d['v'].spec, d['s'].spec, d['p'].spec,
                             d['c'].spec, mask, c)
            w = d['wg']
            g1, b1, e1 = d['ln1']
            # FFN-in (K = hidden, GELU epilogue) measured faster with bf16 operands: the block before it writes
            # its output in both carrier formats
            mixed = self.ffn_in_bf16 and not lean
            ops.linear_res_ln_i8(c, w.grid8, w.rowsum, w.bias, M, w.N, w.K, d['c'].spec, w.seg_spec if lean else w.spec,
                                 1 if lean else w.N, d['g'].spec, x,
                                 x_site.spec, d['u'].spec, g1, b1, e1, d['x'].spec, a, out_ctr=self.a if mixed else None)
            w = d['wf']
            if lean:
                ops.linear_seg_i8(a, w.grid8, w.rowsum, w.bias, M, w.N, w.K, d['x'].spec, w.seg_spec, d['f'].spec, 1, 1, out_i8=f)
            elif mixed:
                ops.linear_bf16_o8(self.a, w.grid, w.bias, M, w.N, w.K, d['x'].spec, w.spec, w.N, 1, d['f'].spec, 1, f)
            else:
                ops.linear_i8(a, w.grid8, w.rowsum, w.bias, M, w.N, w.K, d['x'].spec, w.spec, w.N, 1, d['f'].spec, 1, out_i8=f)
            w = d['wh']
            g2, b2, e2 = d['ln2']
            ops.linear_res_ln_i8(f, w.grid8, w.rowsum, w.bias, M, w.N, w.K, d['f'].spec, w.seg_spec if lean else w.spec,
                                 1 if lean else w.N, d['h'].spec, a,
                                 d['x'].spec, d['y'].spec, g2, b2, e2, d['z'].spec, x)
            x_site = d['z']
        if self.head:            # first token -> pooler -> classifier in one launch
            wp, wc = self.w_pool, self.w_cls
            # (a fresh output tensor per call, like every other forward path: callers may keep the logits of several batches)
            logits = ops.head_i8(x, T * D, B, D, wc.N, wp.grid8, wp.rowsum, wp.bias, x_site.spec, wp.seg_spec, self.pool_out.spec,
                                 wc.grid8, wc.rowsum, wc.bias, wc.seg_spec, self.cls_out.spec,
                                 torch.empty(B, wc.N, dtype=torch.float32, device=x.device))
        else:
            self.first8.copy_(x.view(B, T, D)[:, 0])                     # pooler input: first token
            w = self.w_pool
            ops.linear_i8(self.first8, w.grid8, w.rowsum, w.bias, B, w.N, w.K, x_site.spec, w.spec, w.N, 3, self.pool_out.spec,
                          1, out_i8=self.pooled8)
            w = self.w_cls
            logits = ops.linear_i8(self.pooled8, w.grid8, w.rowsum, w.bias, B, w.N, w.K, self.pool_out.spec, w.spec, w.N, 0,
                                   self.cls_out.spec, 1, want_f32=True)
        logits = logits[:, :self.num_labels]
        if self.num_labels == 1:
            logits = torch.clamp(logits, 0.0, 5.0)
        return logits

    __call__ = forward

    def validate(self):
        """the raw pointers baked into the specs still point at the model's live quantizer buffers?"""
        for st in self._all_sites():
            st.check()

    def _all_sites(self):
        yield from (self.e_tok, self.e_pos, self.e_out, self.pool_out, self.cls_out)
        for d in self.layers:
            for k in ('q', 'k', 'v', 's', 'p', 'c', 'g', 'u', 'x', 'f', 'h', 'y', 'z'):
                yield d[k]

    def i8_flop_share(self, kernel_class='linear_qdq'):
        """share of the tensor-core flops of one forward (of the given bench kernel class) that runs on kind::i8
        (bench.py: flop-weighted tensor peak); the chain class of the one-launch encoder includes the bf16 attention products"""
        if not self.i8:
            return 0.0
        i8 = bf = 0
        if kernel_class == 'chain' and self.chain == 2:
            bf += len(self.layers) * 2 * self.T * self.D          # per row: QK^T and PV over T keys (x 2 flops, as N * K below)
        for d in self.layers:
            for key in ('wqkv', 'wg', 'wh'):
                i8 += d[key].N * d[key].K
            if self.ffn_in_bf16 and not self.lean:
                bf += d['wf'].N * d['wf'].K
            else:
                i8 += d['wf'].N * d['wf'].K
        return i8 / float(i8 + bf)

    def hidden_states(self):
        """dequantized output of the last encoder block of the most recent forward (for tests)"""
        z = self.layers[-1]['z'].q
        if self._last_i8:
            zp = z.zero_point
            zp = zp.reshape(()) if torch.is_tensor(zp) else zp
            return ((self.x8.float() - zp) * z.scale.reshape(())).view(self.B, self.T, self.D)
        return (self.x.float() * z.scale.reshape(())).view(self.B, self.T, self.D)
