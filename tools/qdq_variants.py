#!/usr/bin/env python
"""Per-tensor quant-dequant kernel variants (TQ_QDQ_VARIANT = ldg | deep | bulk | chunk, one process each -- the
switch is read once) on one B200: CUDA events, 3 warm-ups, 5 back-to-back launches per timing for the tensors
larger than L2, L2 flush + single launch for the BASELINE activation shapes.  Every variant must produce the
same bits (checksum of the output words).  Writes gpurun_out/qdq_variants.json."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import torch
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))
    import tq_native
    ops = tq_native.ops()
    dev = 'cuda'
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = []
    for n, flush in [(256 * 1024 * 1024, False), (64 * 1024 * 1024, False), (32 * 128 * 3072, True), (32 * 128 * 768, True)]:
        g = torch.Generator(device=dev).manual_seed(1234)
        x = torch.randn(n, device=dev, generator=g) * 3
        y = torch.empty_like(x)
        mm = ops.minmax(x)
        d, z = torch.empty(1, device=dev), torch.empty(1, device=dev)
        ops.set_range_asym(mm[0:1] * 0.5, mm[1:2] * 0.5, 8, 1e-8, False, d, z)
        spec = ops.spec(d, z, None, 8)
        for _ in range(3):
            ops.qdq(x, spec, out=y)
        torch.cuda.synchronize()
        batch = 1 if flush else 5
        ts = []
        for _ in range(10):
            if flush:
                flush_buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(batch):
                ops.qdq(x, spec, out=y)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / batch)
        ts.sort()
        out.append(dict(variant=os.environ.get('TQ_QDQ_VARIANT', 'default'), n=n, us_median=ts[len(ts) // 2] * 1e3,
                        gbs=8.0 * n / ts[len(ts) // 2] / 1e6, checksum=int(y.view(torch.int32).sum(dtype=torch.int64).item())))
        del x, y
    print(json.dumps(out))
    sys.exit(0)

res = []
for variant in ('default', 'ldg', 'deep', 'bulk', 'chunk'):
    env = dict(os.environ)
    env.pop('TQ_QDQ_VARIANT', None)
    if variant != 'default':
        env['TQ_QDQ_VARIANT'] = variant
    p = subprocess.run([sys.executable, os.path.abspath(__file__), 'child'], env=env, capture_output=True, text=True)
    if p.returncode != 0:
        print(variant, 'failed:', p.stderr[-400:])
        continue
    rows = json.loads(p.stdout.strip().splitlines()[-1])
    res.extend(rows)
    for r in rows:
        print(f"{r['variant']:8s} n={r['n']:>10d} {r['us_median']:9.1f} us {r['gbs']:8.0f} GB/s  checksum {r['checksum']}", flush=True)
sums = {}
for r in res:
    sums.setdefault(r['n'], set()).add(r['checksum'])
print('bit-identical across variants:', all(len(v) == 1 for v in sums.values()))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'qdq_variants.json'), 'w'), indent=1)
