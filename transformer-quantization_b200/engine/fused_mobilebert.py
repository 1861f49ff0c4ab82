"""Fused fixed-range engine for the quantized MobileBERT encoder (BASELINE config 4: W4A8, batch 64, seq 128).

Built from a calibrated ``engine.mobilebert.QuantMobileBertForSequenceClassification`` (ranges fixed, eval mode;
the same site census as the reference's models/quantized_mobilebert.py:58-72, 166-270, 273-311, 327-462, 465-545).
Every activation between kernels is carried as x_int bytes, every GEMM runs on the int8 tensor cores (4-bit weight
grids are int8 grids with a short range), 15 kernels per layer instead of ~60 launches on the module path:

    tq_linear_nonorm_qdq_i8   bottleneck input      dense 512 -> 128 -> QDQ -> NoNorm -> QDQ
    tq_linear_nonorm_qdq_i8   shared q/k bottleneck  dense 512 -> 128 -> QDQ -> NoNorm -> QDQ
    tq_linear_seg_qdq_i8      Q | K (two segments) and V, written into one [M, 3 * H * 64] buffer: every 32-wide head sits
                              in a 64-column slot whose upper half is produced by ZERO weight rows (exact zeros)
    tq_attention_pad_qdq_i8   scores / probs / context (divided by sqrt(32) exactly), context as bytes
    tq_linear_nonorm_qdq_i8   attention output      dense -> QDQ -> + bottleneck input -> QDQ -> NoNorm -> QDQ
    4 x { tq_linear_seg_qdq_i8 (128 -> 512, ReLU), tq_linear_nonorm_qdq_i8 (512 -> 128 + residual + NoNorm) }
    tq_linear_nonorm_qdq_i8   output bottleneck     dense 128 -> 512 -> QDQ -> + layer input -> QDQ -> NoNorm -> QDQ

QuantNoNorm sends its weight AND its bias through one weight quantizer on every forward (reference :58-72); with fixed
ranges both are constants, so they are evaluated once here through the module's own quantizer.  The embedding block
(once per forward) and the classifier run on the module path.
"""
import torch
from torch import nn

import tq_native
from engine.fused import UnsupportedByEngine, _Site, _mgr

SLOT = 64            # head slot width of the attention kernel


def _weight_q(lin):
    if not lin._quant_w:
        raise UnsupportedByEngine('weights must be quantized')
    qz = lin.weight_quantizer.quantizer
    if not qz.is_initialized or qz.n_bits > 8 or not qz.symmetric or qz.delta.numel() != 1:
        raise UnsupportedByEngine('engine needs per-tensor symmetric <= 8-bit weight quantizers')
    return qz


class _W:
    """int8 grid of one fake-quantized Linear weight, optionally with its output rows / input columns spread into
    64-wide head slots (zero padding), its row sums and its bias"""

    def __init__(self, lin, heads=0, pad_rows=False, pad_cols=False):
        ops = tq_native.ops()
        qz = _weight_q(lin)
        w = lin.weight.detach()
        gi, _ = ops.quant_int(w, qz._spec(), want_f32=True, want_bf16=False)
        gi = gi.to(torch.int32)
        bias = lin.bias.detach().float() if lin.bias is not None else torch.zeros(w.shape[0], device=w.device)
        N, K = gi.shape
        if pad_rows:
            hd = N // heads
            g2 = torch.zeros(heads * SLOT, K, dtype=torch.int32, device=w.device)
            b2 = torch.zeros(heads * SLOT, device=w.device)
            for h in range(heads):
                g2[h * SLOT:h * SLOT + hd] = gi[h * hd:(h + 1) * hd]
                b2[h * SLOT:h * SLOT + hd] = bias[h * hd:(h + 1) * hd]
            gi, bias = g2, b2
        if pad_cols:
            hd = K // heads
            g2 = torch.zeros(gi.shape[0], heads * SLOT, dtype=torch.int32, device=w.device)
            for h in range(heads):
                g2[:, h * SLOT:h * SLOT + hd] = gi[:, h * hd:(h + 1) * hd]
            gi = g2
        self.N, self.K = gi.shape
        self.grid8 = (gi.to(torch.int8) if bool(qz.signed) else gi.to(torch.uint8)).contiguous()
        self.rowsum = gi.sum(dim=1, dtype=torch.int32).contiguous()
        self.bias = bias.contiguous()
        self.delta = qz.delta.reshape(1).contiguous()
        self._signed = qz._signed
        self.spec = ops.spec(self.delta, None, self._signed, qz.n_bits, qz.scale_domain == 'log', qz.eps)


def _stack(ws):
    """stack weights along N (fused Q | K projection): grids, row sums, biases; one weight scale per segment"""
    ops = tq_native.ops()
    out = _W.__new__(_W)
    out.grid8 = torch.cat([w.grid8 for w in ws]).contiguous()
    out.rowsum = torch.cat([w.rowsum for w in ws]).contiguous()
    out.bias = torch.cat([w.bias for w in ws]).contiguous()
    out.N, out.K = out.grid8.shape
    out.delta = torch.cat([w.delta for w in ws]).contiguous()
    out._signed = ws[0]._signed
    q = ws[0].spec
    out.spec = ops.spec(out.delta, None, out._signed, q.n_bits, bool(q.log_domain), q.eps)
    return out


def _seg_out(quantizers, ops):
    q0 = quantizers[0]
    if any(q.symmetric for q in quantizers):
        raise UnsupportedByEngine('engine needs asymmetric activation quantizers')
    delta = torch.cat([q.delta.reshape(1) for q in quantizers]).contiguous()
    zero = torch.cat([q.zero_float.reshape(1) for q in quantizers]).contiguous()
    return (delta, zero), ops.spec(delta, zero, None, q0.n_bits, q0.scale_domain == 'log', q0.eps)


class _NoNorm:
    """fake-quantized parameters (fixed ranges -> constants) and output quantizer of a QuantNoNorm"""

    def __init__(self, mod):
        with torch.no_grad():
            w, b = mod.weight, mod.bias
            if mod._quant_w:
                w = mod.weight_quantizer(w)          # weight first, bias second -- like every forward of the module
                b = mod.weight_quantizer(b)
        self.gamma = w.detach().float().contiguous().clone()
        self.beta = b.detach().float().contiguous().clone()
        self.site = _Site(_mgr(mod))


class _Block:
    """dense [+ residual quantizer] + NoNorm"""

    def __init__(self, blk, heads=0, pad_cols=False):
        self.w = _W(blk.dense, heads=heads, pad_cols=pad_cols)
        self.dense = _Site(_mgr(blk.dense))
        self.res = _Site(_mgr(blk.res)) if hasattr(blk, 'res') else None
        self.nn = _NoNorm(blk.norm)


class FusedMobileBertEngine:
    def __init__(self, model, batch, seq):
        if model.training:
            raise UnsupportedByEngine('engine runs the eval forward')
        c = model.config
        self.model = model
        self.B, self.T = batch, seq
        self.D, self.t, self.H = c.hidden_size, c.true_hidden_size, c.num_attention_heads
        self.hd = self.t // self.H
        if seq != 128 or self.hd > SLOT or self.t % 128 != 0 or self.D % 128 != 0 or c.intermediate_size % 128 != 0:
            raise UnsupportedByEngine('engine supports seq 128, head_dim <= 64, widths that are multiples of 128')
        if c.hidden_act != 'relu':
            raise UnsupportedByEngine('engine supports the ReLU feed-forward activation')
        self.ops = ops = tq_native.ops()
        self.dev = dev = next(model.parameters()).device
        H = self.H
        self.e_out = _Site(_mgr(model.embeddings.norm))
        if self.e_out.q.symmetric:
            raise UnsupportedByEngine('engine needs asymmetric activation quantizers')
        self.layers = []
        with torch.no_grad():
            for L in model.layers:
                d = {'b_in': _Block(L.b_in), 'b_att': _Block(L.b_att)}
                wq, wk = _W(L.query, heads=H, pad_rows=True), _W(L.key, heads=H, pad_rows=True)
                d['wqk'] = _stack([wq, wk])
                d['qk_keep'], d['qk_out'] = _seg_out([_mgr(L.query), _mgr(L.key)], ops)
                d['wv'] = _W(L.value, heads=H, pad_rows=True)
                d['q'], d['k'], d['v'] = _Site(_mgr(L.query)), _Site(_mgr(L.key)), _Site(_mgr(L.value))
                d['s'], d['p'], d['c'] = _Site(_mgr(L.s)), _Site(_mgr(L.p)), _Site(_mgr(L.c))
                d['attn_out'] = _Block(L.attn_out, heads=H, pad_cols=True)
                d['ffn'] = []
                for ffn in list(L.ffn) + [None]:
                    inter = ffn.intermediate if ffn is not None else L.intermediate
                    outb = ffn.output if ffn is not None else L.output
                    if not isinstance(inter.activation_function, nn.ReLU):
                        raise UnsupportedByEngine('engine supports the ReLU feed-forward activation')
                    d['ffn'].append((_W(inter), _Site(_mgr(inter)), _Block(outb)))
                d['out_b'] = _Block(L.out_bottleneck)
                self.layers.append(d)
        M = batch * seq
        self.M = M
        u8 = dict(dtype=torch.uint8, device=dev)
        self.h8 = [torch.empty(M, self.D, **u8) for _ in range(2)]
        self.li8 = torch.empty(M, self.t, **u8)
        self.sh8 = torch.empty(M, self.t, **u8)
        self.qkv = torch.empty(M, 3 * H * SLOT, dtype=torch.bfloat16, device=dev)
        self.c8 = torch.empty(M, H * SLOT, **u8)
        self.a8 = [torch.empty(M, self.t, **u8) for _ in range(2)]
        self.i8 = torch.empty(M, c.intermediate_size, **u8)
        self._last = self.h8[0]
        self._last_site = self.e_out

    def _nonorm(self, a8, a_site, blk, res8, res_site, out8):
        ops = self.ops
        w = blk.w
        ops.linear_nonorm_i8(a8, w.grid8, w.rowsum, w.bias, self.M, w.N, w.K, a_site.spec, w.spec, blk.dense.spec, res8,
                             res_site.spec if res8 is not None else None, blk.res.spec if res8 is not None else None,
                             blk.nn.gamma, blk.nn.beta, blk.nn.site.spec, out8)
        return blk.nn.site

    @torch.no_grad()
    def forward(self, input_ids, attention_mask=None, token_type_ids=None):
        ops, model = self.ops, self.model
        B, T, H, M = self.B, self.T, self.H, self.M
        assert tuple(input_ids.shape) == (B, T)
        mask = None
        if attention_mask is not None:
            mask = ((1.0 - attention_mask.to(torch.float32)) * -10000.0).contiguous()
        # embedding block on the module path (trigram lookup, one GEMM, three quantizer sites, NoNorm): once per forward
        e = model.embeddings(input_ids, token_type_ids)
        xi, _ = ops.quant_int(e.reshape(M, self.D), self.e_out.spec, want_f32=True, want_bf16=False)
        h, h_site, flip = self.h8[0], self.e_out, 0
        h.copy_(xi)
        for d in self.layers:
            li_site = self._nonorm(h, h_site, d['b_in'], None, None, self.li8)
            sh_site = self._nonorm(h, h_site, d['b_att'], None, None, self.sh8)
            w = d['wqk']
            ops.linear_seg_i8(self.sh8, w.grid8, w.rowsum, w.bias, M, w.N, w.K, sh_site.spec, w.spec, d['qk_out'], 2, 0,
                              out_ctr=self.qkv, ldc=3 * H * SLOT)
            w = d['wv']
            ops.linear_seg_i8(h, w.grid8, w.rowsum, w.bias, M, w.N, w.K, h_site.spec, w.spec, d['v'].spec, 1, 0,
                              out_ctr=self.qkv[:, 2 * H * SLOT:], ldc=3 * H * SLOT)
            ops.attention_pad_i8(self.qkv, B, T, H, SLOT, self.hd, d['q'].spec, d['k'].spec, d['v'].spec, d['s'].spec,
                                 d['p'].spec, d['c'].spec, mask, self.c8)
            a, a_site, af = self.a8[0], self._nonorm(self.c8, d['c'], d['attn_out'], self.li8, li_site, self.a8[0]), 0
            for wi, i_site, outb in d['ffn']:
                ops.linear_seg_i8(a, wi.grid8, wi.rowsum, wi.bias, M, wi.N, wi.K, a_site.spec, wi.spec, i_site.spec, 1, 2,
                                  out_i8=self.i8)
                nxt = self.a8[af ^ 1]
                a_site = self._nonorm(self.i8, i_site, outb, a, a_site, nxt)
                a, af = nxt, af ^ 1
            nxt = self.h8[flip ^ 1]
            h_site = self._nonorm(a, a_site, d['out_b'], h, h_site, nxt)
            h, flip = nxt, flip ^ 1
        self._last, self._last_site = h, h_site
        first = self.hidden_states()[:, 0].contiguous()
        pooled = model.pooler(first) if model.pooler is not None else first
        return model.classifier(pooled)

    __call__ = forward

    def hidden_states(self):
        """dequantized output of the last layer of the most recent forward"""
        q = self._last_site.q
        zp = q.zero_point
        zp = zp.reshape(()) if torch.is_tensor(zp) else zp
        return ((self._last.float() - zp) * q.scale.reshape(())).view(self.B, self.T, self.D)
