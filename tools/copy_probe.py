#!/usr/bin/env python
"""Attainable copy bandwidth of this library's streaming access pattern vs torch's copy kernel (one B200,
CUDA events, 5 launches per timing, tensors larger than L2).  Writes gpurun_out/copy_probe.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))
import tq_native  # noqa: E402

ops = tq_native.ops()
lib = ops.lib
st = torch.cuda.current_stream().cuda_stream
res = []


def timeit(fn, batch=5, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(batch):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / batch)
    ts.sort()
    return ts[len(ts) // 2]


for n in (256 * 1024 * 1024, 64 * 1024 * 1024):
    x = torch.randn(n, device='cuda')
    y = torch.empty_like(x)
    t = timeit(lambda: y.copy_(x))
    res.append(dict(kernel='torch_copy', n=n, us=t * 1e3, gbs=8.0 * n / t / 1e6))
    for flags in range(8):
        def fn():
            rc = lib.tq_probe_copy_f32(x.data_ptr(), y.data_ptr(), n, flags, st)
            assert rc == 0
        t = timeit(fn)
        assert torch.equal(x, y)
        res.append(dict(kernel=f'probe ld_na={flags & 1} st_na={(flags >> 1) & 1} persistent={(flags >> 2) & 1}', n=n,
                        us=t * 1e3, gbs=8.0 * n / t / 1e6))
    del x, y
for r in res:
    print(f"{r['kernel']:44s} n={r['n']:>10d} {r['us']:9.1f} us {r['gbs']:8.0f} GB/s", flush=True)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'copy_probe.json'), 'w'), indent=1)
