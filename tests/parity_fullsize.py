#!/usr/bin/env python
"""Model-level parity at BASELINE size (BERT-base, 12 layers, B=32, T=128, seed-0 weights, ids seed 1234 --
exactly bench.py's model): where do the fused engine, the module path and the reference arithmetic part
ways, site by site and source by source?   GPU tool; writes gpurun_out/parity_fullsize.json.

    python tests/parity_fullsize.py [--layers 12] [--out gpurun_out/parity_fullsize.json]

Four comparisons, all in units of the quantization step of the site that is compared:

  floor    the REFERENCE arithmetic (oracle/bert_oracle.py: the reference's op chain in torch fp32) run twice,
           on the host CPU (MKL fp32 GEMM) and on the GPU (cuBLAS fp32 GEMM): same formulas, two GEMM
           summation orders.  This is the drift the reference has against ITSELF on two machines.
  module   this repo's module path (one kernel per site, exact integer tensor-core GEMMs) vs the CPU oracle.
  engine   the fused engine vs the module path: traced chain (unfused LayerNorm, bf16 carriers) per site, and
           the default int8 / fused-LayerNorm chain on the logits and the last hidden state.
  local    every fused stage fed with the MODULE PATH's own tensors of layer L (no accumulated drift): flips
           of each kernel in isolation -- the per-source budget.
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'transformer-quantization_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

SITE_NAMES = ['e_tok', 'e_pos', 'e_ln']
LAYER_SITES = ['q', 'k', 'v', 's', 'p', 'c', 'g', 'u', 'x', 'f', 'h', 'y', 'z']


def site_names(n_layers):
    out = list(SITE_NAMES)
    for i in range(n_layers):
        out += [f'{i}.{s}' for s in LAYER_SITES]
    return out + ['pool', 'cls']


def site_modules(model):
    """the modules whose outputs are the 161 quantized activations, in engine.bert's act_quantizers() order"""
    E = model.embeddings
    out = [E.e_tok, E.e_pos, E.norm]
    for L in model.layers:
        out += [L.query, L.key, L.value, L.s, L.p, L.c, L.g, L.u, L.x, L.ffn_in, L.h, L.y, L.z]
    return out + [model.pooler, model.classifier]


def diff_stats(a, b, step):
    """a, b fp32 tensors on one device; step = quantization step: -> (max |d| in steps, share of elements off by >= half a step)"""
    d = (a.double() - b.double()).abs() / step
    return float(d.max()), float((d >= 0.5).double().mean())


def build_calibrated_model(n_layers, dev):
    """bench.py's model: BERT-base (n_layers), weights seed 0, W8 sym / A8 asym, calibrated on the ids-seed-1234
    batch, ranges fixed -> (model on dev, ids, mask)"""
    import bench
    from engine.bert import BertConfig, QuantBertForSequenceClassification
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators
    model = QuantBertForSequenceClassification(
        BertConfig(num_hidden_layers=n_layers), method=QMethods.symmetric_uniform,
        act_method=QMethods.asymmetric_uniform, n_bits=8, n_bits_act=8,
        weight_range_method=RangeEstimators.current_minmax, act_range_method=RangeEstimators.running_minmax)
    model.init_weights(seed=0)
    model.to(dev).eval()
    model.set_quant_state(weight_quant=True, act_quant=True)
    ids = bench.synthetic_ids(1234)[0]
    mask = torch.ones_like(ids)
    with torch.no_grad():
        model(ids.to(dev), mask.to(dev))
        model.fix_ranges()
    return model, ids, mask


def capture_module_path(model, ids, mask):
    """one module-path forward; returns {site: fp32 output of that activation quantizer}, logits"""
    # the site's OUTPUT is the output of the module that owns the quantizer (a fused QuantLinear quantizes in
    # its GEMM epilogue and never calls the manager)
    mods = site_modules(model)
    names = site_names(len(model.layers))
    assert len(mods) == len(names)
    got = {}
    hooks = [m.register_forward_hook(lambda mod, inp, out, n=n: got.__setitem__(n, out)) for n, m in zip(names, mods)]
    try:
        with torch.no_grad():
            logits = model(ids, mask)
    finally:
        for h in hooks:
            h.remove()
    return got, logits


def compare_with_golden(mod, logits, names, G=None):
    """module-path site outputs (dict name -> fp32 tensor) vs tests/golden/bert_base_fullsize.npz (strided samples
    of the UNMODIFIED reference's outputs on the host CPU) -> per-site (max |d| in steps, share off by >= 0.5 step),
    max relative difference of the ranges, logit error in classifier steps"""
    import numpy as np
    if G is None:
        G = np.load(os.path.join(ROOT, 'tests', 'golden', 'bert_base_fullsize.npz'))
    stride = int(G['stride'])
    assert int(G['n_sites']) == len(names)
    per_site = {}
    for i, n in enumerate(names):
        ref = torch.from_numpy(G[f'q{i}.sample'])
        got = mod[n].detach().reshape(-1)[::stride].float().cpu()
        per_site[n] = diff_stats(got, ref, float(G[f'q{i}.delta'].reshape(-1)[0]))
    cls = float(G[f'q{len(names) - 1}.delta'].reshape(-1)[0])
    err = float(np.abs(logits.detach().float().cpu().numpy() - G['logits']).max() / cls)
    return per_site, err


def local_stage_flips(eng, model, mod, layers):
    """Every fused stage of the engine fed with the MODULE PATH's own tensors of layer L (teacher forcing: no
    accumulated drift) -> {layer: {stage: (max |d| in steps, share off by >= half a step)}} against the module
    path's output of the same stage."""
    import tq_native
    ops = tq_native.ops()
    dev = eng.dev
    B, T, L = eng.B, eng.T, len(eng.layers)
    names = site_names(L)
    mgrs = dict(zip(names, model.act_quantizers()))
    step = {n: float(m.quantizer.scale.reshape(-1)[0]) for n, m in mgrs.items()}

    def grid(name):
        """bf16 centred integer grid of the module path's tensor at `name`"""
        q = mgrs[name].quantizer
        _, g = ops.quant_int(mod[name].reshape(-1, mod[name].shape[-1]), q._spec(), want_f32=False, want_bf16=True)
        return g

    def deq(ctr, name):
        return ctr.float() * step[name]

    D, H, M = eng.D, eng.H, eng.M
    local = {}
    if getattr(eng, '_ids', None) is not None:
        # embedding block (three sites in one kernel) on the same token ids
        x = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
        ops.embed_ln_qdq(eng._ids.reshape(-1).contiguous(), None, None, T, eng.word_q, eng.type_q, eng.pos_q, eng.e_tok.spec, 1,
                         eng.e_pos.spec, 1, eng.e_gamma, eng.e_beta, eng.e_eps, eng.e_out.spec, 1, out_ctr=x)
        local['emb'] = {'embed_ln->e_ln': diff_stats(deq(x, 'e_ln').reshape(mod['e_ln'].shape), mod['e_ln'], step['e_ln'])}
    for li in layers:
        d = eng.layers[li]
        prev = 'e_ln' if li == 0 else f'{li - 1}.z'
        prev_site = eng.e_out if li == 0 else eng.layers[li - 1]['z']
        x_in = grid(prev)
        r = {}
        # QKV GEMM (same GEMM kernel as the module path: expect no flips)
        qkv = torch.empty(M, 3 * D, dtype=torch.bfloat16, device=dev)
        eng._linear(x_in, prev_site, d['wqkv'], 0, d['qkv_out'].spec, d['qkv_out'].n, out_ctr=qkv)
        for j, s in enumerate('qkv'):
            r[f'qkv_gemm->{s}'] = diff_stats(deq(qkv[:, j * D:(j + 1) * D], f'{li}.{s}').reshape(mod[f'{li}.{s}'].shape),
                                             mod[f'{li}.{s}'], step[f'{li}.{s}'])
        # attention on the module path's q | k | v
        qkv_m = torch.cat([grid(f'{li}.q'), grid(f'{li}.k'), grid(f'{li}.v')], dim=1).contiguous()
        c = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
        ops.attention(qkv_m, B, T, H, eng.hd, d['q'].spec, d['k'].spec, d['v'].spec, d['s'].spec, d['p'].spec, d['c'].spec,
                      None, out_ctr=c)
        r['attention->c'] = diff_stats(deq(c, f'{li}.c').reshape(mod[f'{li}.c'].shape), mod[f'{li}.c'], step[f'{li}.c'])
        # ... and the same formulation in torch on the same inputs, exp / sum order of the library softmax
        qm, km, vm = (mod[f'{li}.{s}'].view(B, T, H, eng.hd).permute(0, 2, 1, 3) for s in 'qkv')
        s_ref = mgrs[f'{li}.s'](torch.matmul(qm, km.transpose(-1, -2)))
        r['torch scores (cuBLAS fp32) vs module s'] = diff_stats(s_ref, mod[f'{li}.s'], step[f'{li}.s'])
        # attention-out block
        c_in = grid(f'{li}.c')
        w = d['wg']
        g1, b1, e1 = d['ln1']
        u = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
        ops.linear_res(c_in, w.grid, w.bias, M, w.N, w.K, d['c'].spec, w.spec, w.N, d['g'].spec, 1, x_in, prev_site.spec,
                       d['u'].spec, 1, out_ctr=u)
        r['attn_out_res->u'] = diff_stats(deq(u, f'{li}.u').reshape(mod[f'{li}.u'].shape), mod[f'{li}.u'], step[f'{li}.u'])
        a, _ = ops.ln_qdq(grid(f'{li}.u'), d['u'].spec, 1, g1, b1, e1, d['x'].spec, 1)
        r['ln_kernel->x'] = diff_stats(deq(a, f'{li}.x').reshape(mod[f'{li}.x'].shape), mod[f'{li}.x'], step[f'{li}.x'])
        if eng.fuse_ln:
            _, a2 = ops.linear_res_ln(c_in, w.grid, w.bias, M, w.N, w.K, d['c'].spec, w.spec, w.N, d['g'].spec, x_in,
                                      prev_site.spec, d['u'].spec, g1, b1, e1, d['x'].spec)
            r['attn_out_res_ln_fused->x'] = diff_stats(deq(a2, f'{li}.x').reshape(mod[f'{li}.x'].shape), mod[f'{li}.x'],
                                                       step[f'{li}.x'])
        # FFN
        a_in = grid(f'{li}.x')
        f = torch.empty(M, d['wf'].N, dtype=torch.bfloat16, device=dev)
        eng._linear(a_in, d['x'], d['wf'], 1, d['f'].spec, 1, out_ctr=f)
        r['ffn_in_gelu->f'] = diff_stats(deq(f, f'{li}.f').reshape(mod[f'{li}.f'].shape), mod[f'{li}.f'], step[f'{li}.f'])
        w = d['wh']
        g2, b2, e2 = d['ln2']
        y = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
        ops.linear_res(grid(f'{li}.f'), w.grid, w.bias, M, w.N, w.K, d['f'].spec, w.spec, w.N, d['h'].spec, 1, a_in,
                       d['x'].spec, d['y'].spec, 1, out_ctr=y)
        r['ffn_out_res->y'] = diff_stats(deq(y, f'{li}.y').reshape(mod[f'{li}.y'].shape), mod[f'{li}.y'], step[f'{li}.y'])
        z, _ = ops.ln_qdq(grid(f'{li}.y'), d['y'].spec, 1, g2, b2, e2, d['z'].spec, 1)
        r['ln_kernel->z'] = diff_stats(deq(z, f'{li}.z').reshape(mod[f'{li}.z'].shape), mod[f'{li}.z'], step[f'{li}.z'])
        if eng.fuse_ln:
            _, z2 = ops.linear_res_ln(grid(f'{li}.f'), w.grid, w.bias, M, w.N, w.K, d['f'].spec, w.spec, w.N, d['h'].spec,
                                      a_in, d['x'].spec, d['y'].spec, g2, b2, e2, d['z'].spec)
            r['ffn_out_res_ln_fused->z'] = diff_stats(deq(z2, f'{li}.z').reshape(mod[f'{li}.z'].shape), mod[f'{li}.z'],
                                                      step[f'{li}.z'])
        local[li] = r
    return local


def run_oracle(sd, ids, mask, device, n_layers):
    """reference arithmetic (torch fp32 op chain) on `device`: calibrate on the batch, fix, forward; per-site outputs"""
    from oracle.bert_oracle import OracleBert
    sd = {k: v.to(device) for k, v in sd.items()}
    m = OracleBert(sd, n_layers=n_layers, n_heads=12, n_bits=8, n_bits_act=8, sym_acts=False, device=device)
    ids, mask = ids.to(device), mask.to(device)
    with torch.no_grad():
        m(ids, mask)
        m.fix_ranges()
        got = {}
        for name, site in m.act.items():
            site.record = (got, name)
        logits = m(ids, mask)
        for site in m.act.values():
            site.record = None
    steps = {n: float(torch.clamp(s.delta, min=s.eps)) for n, s in m.act.items()}
    return got, logits, steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--layers', type=int, default=12)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'parity_fullsize.json'))
    ap.add_argument('--local-layers', default=','.join(str(i) for i in range(12)))
    args = ap.parse_args()
    import bench
    import tq_native
    from engine.fused import FusedBertEngine
    from oracle.bert_oracle import random_bert_state_dict
    dev = torch.device('cuda')
    ops = tq_native.ops()
    torch.backends.cuda.matmul.allow_tf32 = False
    B, T = bench.BATCH, bench.SEQ
    L = args.layers
    report = {'config': f'BERT-base {L} layers, B={B}, T={T}, weights seed 0, ids seed 1234, W8 sym / A8 asym, '
                        'one calibration batch'}

    # ---- module path + engine on the GPU --------------------------------------------------------------
    model, ids, mask = build_calibrated_model(L, dev)
    ids_d, mask_d = ids.to(dev), mask.to(dev)
    mod, mod_logits = capture_module_path(model, ids_d, mask_d)
    names = site_names(L)
    mgrs = dict(zip(names, model.act_quantizers()))
    step = {n: float(m.quantizer.scale.reshape(-1)[0]) for n, m in mgrs.items()}
    cls_step = step['cls']
    report['classifier_step'] = cls_step
    report['logit_spread'] = [float(mod_logits.min()), float(mod_logits.max())]

    # ---- floor: the reference arithmetic on two GEMM libraries -----------------------------------------
    sd = random_bert_state_dict(layers=L, seed=0)
    torch.set_num_threads(bench.usable_cpus())
    cpu, cpu_logits, cpu_steps = run_oracle(sd, ids, mask, torch.device('cpu'), L)
    gpu, gpu_logits, _ = run_oracle(sd, ids, mask, dev, L)
    floor = {}
    for n in names:
        floor[n] = diff_stats(gpu[n].cpu(), cpu[n], cpu_steps[n])
    report['floor_reference_cpu_vs_reference_cublas'] = {
        'logits_max_abs': float((gpu_logits.cpu() - cpu_logits).abs().max()),
        'logits_in_classifier_steps': float((gpu_logits.cpu() - cpu_logits).abs().max() / cpu_steps['cls']),
        'per_site': {n: {'max_steps': v[0], 'flip_rate': v[1]} for n, v in floor.items()}}
    del gpu

    # ---- module path vs the CPU oracle ------------------------------------------------------------------
    mvo = {n: diff_stats(mod[n].cpu().reshape(cpu[n].shape), cpu[n], cpu_steps[n]) for n in names}
    report['module_path_vs_cpu_oracle'] = {
        'logits_max_abs': float((mod_logits.cpu() - cpu_logits).abs().max()),
        'logits_in_classifier_steps': float((mod_logits.cpu() - cpu_logits).abs().max() / cls_step),
        'range_rel_diff_max': max(abs(step[n] - cpu_steps[n]) / cpu_steps[n] for n in names),
        'per_site': {n: {'max_steps': v[0], 'flip_rate': v[1]} for n, v in mvo.items()}}
    del cpu
    if L == 12:
        per_site, err = compare_with_golden(mod, mod_logits, names)
        report['module_path_vs_reference_golden (strided samples)'] = {
            'logits_in_classifier_steps': err,
            'per_site': {n: {'max_steps': v[0], 'flip_rate': v[1]} for n, v in per_site.items()}}

    # ---- engine vs module path ---------------------------------------------------------------------------
    eng = FusedBertEngine(model, B, T)
    trace = {}
    tr_logits = eng(ids_d, mask_d, trace=trace)
    tmap = {'emb': 'e_ln'}
    for i in range(L):
        tmap.update({f'{i}.query': f'{i}.q', f'{i}.key': f'{i}.k', f'{i}.value': f'{i}.v', f'{i}.c': f'{i}.c',
                     f'{i}.u': f'{i}.u', f'{i}.x': f'{i}.x', f'{i}.ffn_in': f'{i}.f', f'{i}.y': f'{i}.y', f'{i}.z': f'{i}.z'})
    evm = {}
    for tn, n in tmap.items():
        evm[n] = diff_stats(trace[tn].reshape(mod[n].shape), mod[n], step[n])
    fast_logits = eng(ids_d, mask_d)
    hid = eng.hidden_states()
    report['engine_vs_module_path'] = {
        'traced_chain (bf16 carriers, unfused LayerNorm)': {
            'logits_in_classifier_steps': float((tr_logits - mod_logits).abs().max() / cls_step),
            'per_site': {n: {'max_steps': v[0], 'flip_rate': v[1]} for n, v in evm.items()}},
        'default_chain (int8 operands, fused LayerNorm)': {
            'i8': bool(eng._last_i8),
            'logits_in_classifier_steps': float((fast_logits - mod_logits).abs().max() / cls_step),
            'logits_vs_traced_chain_steps': float((fast_logits - tr_logits).abs().max() / cls_step),
            'last_hidden': dict(zip(('max_steps', 'flip_rate'),
                                    diff_stats(hid.reshape(mod[f'{L - 1}.z'].shape), mod[f'{L - 1}.z'], step[f'{L - 1}.z'])))}}
    del trace

    # ---- local: each fused stage on the module path's own tensors -----------------------------------------
    eng._ids = ids_d
    loc = local_stage_flips(eng, model, mod, [int(v) for v in args.local_layers.split(',') if int(v) < L])
    local = {f'layer {li}': {k: {'max_steps': v[0], 'flip_rate': v[1]} for k, v in r.items()} for li, r in loc.items()}
    report['local_stage_on_module_tensors'] = local
    torch.cuda.synchronize()

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, 'w') as fjs:
        json.dump(report, fjs, indent=1)

    # ---- compact summary --------------------------------------------------------------------------------
    def worst(per_site):
        k = max(per_site, key=lambda n: per_site[n]['flip_rate'])
        return f"worst site {k}: {per_site[k]['flip_rate']:.4%} off by >= 0.5 step, max {per_site[k]['max_steps']:.1f} steps"

    fl = report['floor_reference_cpu_vs_reference_cublas']
    mv = report['module_path_vs_cpu_oracle']
    ev = report['engine_vs_module_path']
    print(f"classifier step {cls_step:.6f}, logit spread {report['logit_spread']}")
    print(f"floor  (reference arithmetic, CPU MKL vs cuBLAS fp32): logits {fl['logits_in_classifier_steps']:.1f} steps; {worst(fl['per_site'])}")
    print(f"module path vs CPU oracle: logits {mv['logits_in_classifier_steps']:.1f} steps; {worst(mv['per_site'])}")
    t = ev['traced_chain (bf16 carriers, unfused LayerNorm)']
    print(f"engine (traced) vs module path: logits {t['logits_in_classifier_steps']:.1f} steps; {worst(t['per_site'])}")
    dflt = ev['default_chain (int8 operands, fused LayerNorm)']
    print(f"engine (default) vs module path: logits {dflt['logits_in_classifier_steps']:.1f} steps; vs traced {dflt['logits_vs_traced_chain_steps']:.1f}; last hidden {dflt['last_hidden']}")
    for lname, r in local.items():
        for k, v in r.items():
            print(f"local {lname:9s} {k:42s} flips {v['flip_rate']:.5%}  max {v['max_steps']:.2f}")


if __name__ == '__main__':
    main()
