#!/usr/bin/env python
"""Timing of the training-time kernels on one B200 (CUDA events, 3 warm-ups, tensors larger than L2 or an
L2 flush between iterations): tq_qdq_bwd_f32 in its three layouts against its 12 B / element HBM roofline,
next to the same backward written with torch ops (the reference's autograd formulation) on the same GPU.
Writes gpurun_out/qat_bench.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))
import tq_native  # noqa: E402

ops = tq_native.ops()
dev = 'cuda'
pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
PEAK = json.load(open(pk))['hbm_gbs'] if os.path.exists(pk) else 6650.0
flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3, flush=False, batch=1):
    """median / best time of one call; `batch` back-to-back calls between the two events amortise the launch
    and event overhead for tensors larger than L2 (the small BASELINE shapes are flushed and timed singly)"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(batch):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / batch)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


res = []


def report(name, layout, bytes_, t, **extra):
    r = dict(kernel=name, layout=list(layout), alg_bytes=bytes_, ms_median=t[0], ms_best=t[1],
             gbs_median=bytes_ / t[0] / 1e6, frac_of_measured_peak=bytes_ / t[0] / 1e6 / PEAK, **extra)
    res.append(r)
    print(json.dumps(r), flush=True)


def torch_backward(x, g, delta, zf, C, layout):
    """the reference's formulation: autograd through clamp(round_ste(x / s) + zp) etc."""
    outer, C, inner = layout
    xv = x.view(outer, C, inner).requires_grad_(True)
    d = delta.view(1, C, 1).clone().requires_grad_(True)
    z = zf.view(1, C, 1).clone().requires_grad_(True)
    s = torch.clamp(d, min=1e-8)
    zp_f = z + (torch.round(z) - z).detach()
    zp = torch.clamp(zp_f, 0, 255)
    t = xv / s
    xi = torch.clamp(t + (torch.round(t) - t).detach() + zp, 0, 255)
    y = s * (xi - zp)
    y.backward(g.view(outer, C, inner))
    return xv.grad, d.grad, z.grad


for layout, flush in [((1, 1, 256 * 1024 * 1024), False), ((1, 1, 64 * 1024 * 1024), False), ((1, 1, 32 * 128 * 768), True), ((1, 1, 32 * 128 * 3072), True),
                      ((16 * 4096, 768, 1), False), ((32 * 128, 768, 1), True), ((32 * 128, 3072, 1), True),
                      ((1, 3072, 768), True), ((1, 30522, 768), True), ((1, 8 * 30522, 768), False)]:
    outer, C, inner = layout
    n = outer * C * inner
    x = torch.randn(n, device=dev) * 3
    g = torch.randn(n, device=dev)
    delta = torch.full((C,), 0.03, device=dev)
    zf = torch.full((C,), 120.3, device=dev)
    spec = ops.spec(delta, zf, None, 8)
    batch = 1 if flush else 5
    report('qdq_bwd', layout, 12 * n, timeit(lambda: ops.qdq_bwd(x, g, spec, C, outer, C, inner), flush=flush, batch=batch),
           launches_per_timing=batch)
    report('qdq_bwd(range gradients only, no grad_x)', layout, 8 * n,
           timeit(lambda: ops.qdq_bwd(x, g, spec, C, outer, C, inner, want_x=False), flush=flush, batch=batch),
           launches_per_timing=batch)
    if n <= 32 * 128 * 3072:
        report('torch_autograd(reference formulation)', layout, 12 * n,
               timeit(lambda: torch_backward(x, g, delta, zf, C, layout), iters=5, warm=2, flush=flush))
    del x, g

# AdaRound kernels on a BERT-base FFN weight
w = torch.randn(3072, 768, device=dev) * 0.04
delta = torch.full((1,), 0.04 * 3 / 7, device=dev)
sg = torch.tensor(True, device=dev)
spec = ops.spec(delta, None, sg, 4)
n = w.numel()
alpha = ops.adaround_init_alpha(w, spec, 1, 1, n, 'learned_hard_sigmoid')
gy = torch.randn_like(w)
report('adaround_fwd(soft)', (1, 1, n), 12 * n, timeit(lambda: ops.adaround_fwd(w, alpha, spec, 1, 1, n, 'learned_hard_sigmoid', True), flush=True))
report('adaround_bwd', (1, 1, n), 16 * n, timeit(lambda: ops.adaround_bwd(w, alpha, gy, spec, 1, 1, n, 'learned_hard_sigmoid'), flush=True))

os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'qat_bench.json'), 'w'), indent=1)
