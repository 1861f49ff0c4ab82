#!/usr/bin/env python
"""Per-kernel timing on one B200: CUDA events, >=3 warm-ups, inputs larger than L2 (or an L2 flush
between iterations for the small BASELINE shapes).  Writes gpurun_out/kernel_bench.json."""
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
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0}
PEAK = peaks['hbm_gbs']
flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2


def timeit(fn, iters=10, warm=3, flush=False):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


res = []


def report(name, shape, bytes_, t):
    med, best = t
    r = dict(kernel=name, shape=list(shape), alg_bytes=bytes_, ms_median=med, ms_best=best,
             gbs_median=bytes_ / med / 1e6, gbs_best=bytes_ / best / 1e6, frac_of_measured_peak=bytes_ / med / 1e6 / PEAK)
    res.append(r)
    print(json.dumps(r), flush=True)


def qspec(x, n_bits=8):
    mm = ops.minmax(x)
    d, z = torch.empty(1, device=dev), torch.empty(1, device=dev)
    ops.set_range_asym(mm[0:1], mm[1:2], n_bits, 1e-8, False, d, z)
    return ops.spec(d, z, None, n_bits), (d, z)


for shape, flush in [((256 * 1024 * 1024,), False), ((64 * 1024 * 1024,), False), ((32, 128, 768), True),
                     ((32, 12, 128, 128), True), ((32, 128, 3072), True)]:
    x = torch.randn(shape, device=dev)
    n = x.numel()
    spec, keep = qspec(x)
    y = torch.empty_like(x)
    report('qdq_tensor', shape, 8 * n, timeit(lambda: ops.qdq(x, spec, out=y), flush=flush))
    report('minmax_tensor', shape, 4 * n, timeit(lambda: ops.minmax(x), flush=flush))
    report('torch_copy(reference point)', shape, 8 * n, timeit(lambda: y.copy_(x), flush=flush))
    if len(shape) == 3 or n % 768 == 0:
        C = shape[-1] if len(shape) == 3 else 768
        rows = n // C
        report('minmax_axis', (rows, C), 4 * n, timeit(lambda: ops.minmax_axis(x, rows, C, 1), flush=flush))
        mn, mx = ops.minmax_axis(x, rows, C, 1)
        gm, gM = ops.group_minmax(mn, mx, 6)
        dv, zv = torch.empty(C, device=dev), torch.empty(C, device=dev)
        ops.set_range_asym(gm.contiguous(), gM.contiguous(), 8, 1e-8, False, dv, zv)
        sp = ops.spec(dv, zv, None, 8)
        report('qdq_peg6', (rows, C), 8 * n, timeit(lambda: ops.qdq(x, sp, rows, C, 1, out=y), flush=flush))
    del x, y

# MSE grid: one read, all candidates
import numpy as np
from quantization.quantizers import QMethods
from quantization.range_estimators import RangeEstimators, OptMethod
x = torch.randn(32, 128, 768, device=dev) * 2
for kind, opt in [('symmetric_uniform', 'grid'), ('asymmetric_uniform', 'grid')]:
    qz = QMethods[kind].cls(n_bits=8)
    est = RangeEstimators.MSE.cls(quantizer=qz, opt_method=OptMethod[opt])
    est(x)
    torch.cuda.synchronize()
    t = timeit(lambda: est(x), iters=3, warm=1)
    n_cand = est._tables['n']
    r = dict(kernel=f'mse_{opt}_{kind}', shape=[32, 128, 768], n_cand=n_cand, ms_median=t[0], ms_best=t[1],
             gcand_elem_per_s=x.numel() * n_cand / t[0] / 1e6)
    res.append(r)
    print(json.dumps(r), flush=True)

os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'kernel_bench.json'), 'w'), indent=1)

# ---- fused linear (tcgen05) ---------------------------------------------------------------------
try:
    tf_peak = peaks.get('bf16_tflops', 1590.0)
    for (M, N, K, act) in [(4096, 768, 768, 0), (4096, 2304, 768, 0), (4096, 3072, 768, 1), (4096, 768, 3072, 0),
                           (8192, 128, 512, 0), (8192, 512, 128, 2), (16384, 4096, 4096, 0)]:
        a = torch.randint(-255, 256, (M, K), device=dev).to(torch.bfloat16)
        w = torch.randint(-128, 128, (N, K), device=dev).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        d = torch.tensor([0.02], device=dev); z = torch.tensor([128.0], device=dev)
        od = torch.tensor([50.0], device=dev); oz = torch.tensor([120.0], device=dev)
        a_spec = ops.spec(d, z, None, 8); o_spec = ops.spec(od, oz, None, 8)
        ws = torch.tensor([0.001], device=dev); sg = torch.tensor(True, device=dev)
        w_spec = ops.spec(ws, None, sg, 8)
        fn = lambda: ops.linear(a, w, bias, M, N, K, 1, a_spec, w_spec, 1, act, o_spec, 1)
        t = timeit(fn, iters=10, warm=3, flush=True)
        fl = 2.0 * M * N * K
        r = dict(kernel='linear_qdq_bf16', shape=[M, N, K], act=act, ms_median=t[0], ms_best=t[1],
                 tflops_median=fl / t[0] / 1e9, tflops_best=fl / t[1] / 1e9, frac_of_measured_bf16_peak=fl / t[0] / 1e9 / tf_peak)
        res.append(r)
        print(json.dumps(r), flush=True)
        t2 = timeit(lambda: torch.matmul(a, w.T), iters=10, warm=3, flush=True)
        r = dict(kernel='cublas_bf16_matmul(reference point)', shape=[M, N, K], ms_median=t2[0], tflops_median=fl / t2[0] / 1e9)
        res.append(r)
        print(json.dumps(r), flush=True)
except Exception as e:  # noqa
    print('linear bench failed:', repr(e), flush=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'kernel_bench.json'), 'w'), indent=1)
