#!/usr/bin/env python
"""Run one BASELINE.json configuration (engine/configs.py) on cuda:0: build the random-init model, calibrate on
two synthetic batches (incl. the FP32 ranges pass for the PEG permutation), fix the ranges, capture the eval
forward in a CUDA graph and time it with CUDA events.  One JSON line per configuration.

    python tools/run_config.py [--config NAME | --all] [--steps 20] [--warmup 3] [--no-graph]

BERT-base per-tensor asymmetric (config 2) goes through the fused engine exactly like bench.py; every other
configuration runs the module path (one kernel per quantizer site, fused QuantLinear) today -- see DESIGN.md
section 10 for what is planned for them."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))

from engine import configs  # noqa: E402
import tq_native  # noqa: E402


def _profile_live(forward, ids, mask, ops, reps=5):
    """per-kernel-class device time of one eager forward: every call of the library is re-issued `reps` times
    between two CUDA events right after it ran (the kernels are idempotent) -- kernel time, not launch overhead"""
    orig = ops._run
    agg = {}

    def timed(name, work, kernels, fn, *args):
        orig(name, work, kernels, fn, *args)
        fn(*args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn(*args)
        e1.record()
        a = agg.setdefault(name, [[], 0.0, 0])
        a[0].append((e0, e1))
        a[1] += work
        a[2] += 1

    ops._run = timed
    try:
        with torch.no_grad():
            forward(ids, mask)
        torch.cuda.synchronize()
    finally:
        ops._run = orig
    return {k: {'seconds': sum(a.elapsed_time(b) for a, b in evs) * 1e-3 / reps, 'work': work, 'launches': n}
            for k, (evs, work, n) in agg.items()}


def run(name, steps, warmup, use_graph=True, device='cuda:0', tiny=False, batch=None, seq=None, profile=False):
    """``device`` / ``tiny`` / ``batch`` / ``seq`` exist for the CPU dry run of this tool in the test-suite (oracle
    back-end injected by the test, wall-clock timing); on a GPU the defaults are the BASELINE shapes."""
    dev = torch.device(device)
    on_gpu = dev.type == 'cuda'
    sync = torch.cuda.synchronize if on_gpu else (lambda: None)
    use_graph = use_graph and on_gpu
    ops = tq_native.ops()
    model, recipe = configs.build(name, dev, tiny=tiny)
    if batch or seq:
        recipe = recipe._replace(batch=batch or recipe.batch, seq=seq or recipe.seq)
    batches = configs.synthetic_batches(model, recipe, 3)
    t0 = time.perf_counter()
    configs.calibrate(model, recipe, batches[:2])
    sync()
    t_cal = time.perf_counter() - t0
    ids = batches[2].to(dev)
    mask = torch.ones_like(ids)
    forward, kind = model, 'module path'
    if on_gpu and recipe.family == 'bert':
        from engine.fused import FusedBertEngine, UnsupportedByEngine
        from engine.fused_peg import FusedBertPegEngine
        try:
            if recipe.peg is None:
                forward, kind = FusedBertEngine(model, recipe.batch, recipe.seq), 'fused engine'
            else:
                forward, kind = FusedBertPegEngine(model, recipe.batch, recipe.seq), 'fused PEG engine (engine/fused_peg.py)'
        except UnsupportedByEngine as e:
            kind = f'module path ({e})'
    elif on_gpu and recipe.family == 'mobilebert':
        from engine.fused import UnsupportedByEngine
        from engine.fused_mobilebert import FusedMobileBertEngine
        try:
            forward, kind = FusedMobileBertEngine(model, recipe.batch, recipe.seq), 'fused MobileBERT engine (engine/fused_mobilebert.py)'
        except UnsupportedByEngine as e:
            kind = f'module path ({e})'
    with torch.no_grad():
        for _ in range(2):
            ref = forward(ids, mask)
        sync()
        l0 = getattr(ops, 'launches', 0)
        if use_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = forward(ids, mask)
            step = graph.replay
        else:
            def step():
                return forward(ids, mask)
            out = step()
        launches = getattr(ops, 'launches', 0) - l0
        for _ in range(max(warmup, 3)):
            step()
        sync()
        if on_gpu:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        if on_gpu:
            e1.record()
            sync()
            ms = e0.elapsed_time(e1) / steps
        else:
            ms = (time.perf_counter() - t0) * 1e3 / steps
        # cross-check of whatever ran against the module path (one kernel per quantizer site)
        module_logits = model(ids, mask)
        out_step = float(model.classifier.activation_quantizer.quantizer.scale.reshape(-1)[0])
        vs_module = float((out.float() - module_logits.float()).abs().max())
        prof = None
        if profile and on_gpu:
            prof = _profile_live(forward, ids, mask, ops)
    extra = {'kernel_profile': prof} if prof is not None else {}
    hid = None
    if hasattr(forward, 'hidden_states') and on_gpu:
        with torch.no_grad():
            ref_h = model.encode(ids, mask)
            hs = float(ref_h.abs().max()) / 128.0
            dh = (forward.hidden_states().float() - ref_h).abs()
            hid = {'max_abs_diff_in_approx_steps': float(dh.max()) / hs, 'share_off_by_half_a_step': float((dh > 0.5 * hs).float().mean())}
    return dict(extra, last_hidden_vs_module_path=hid, config=name, forward=kind, batch=recipe.batch, seq=recipe.seq, ms_per_step=ms,
                tokens_per_s=recipe.batch * recipe.seq / ms * 1e3, library_launches_per_step=launches,
                calibration_s=t_cal, cuda_graph=use_graph, logits_finite=bool(torch.isfinite(out).all()),
                graph_equals_eager=bool(torch.equal(out, ref)) if use_graph else None,
                max_abs_logit_diff_vs_module_path=vs_module, logit_quantization_step=out_step)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='bert_w8a8_asym', choices=sorted(configs.RECIPES))
    ap.add_argument('--all', action='store_true')
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--no-graph', action='store_true')
    a = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit('tools/run_config.py: no CUDA device -- the product path has no CPU fallback')
    for name in (sorted(configs.RECIPES) if a.all else [a.config]):
        print(json.dumps(run(name, a.steps, a.warmup, not a.no_graph)), flush=True)
