"""Tiny driver for ncu: a few launches of the per-tensor QDQ kernel on a 1 GiB tensor."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))
import tq_native
ops = tq_native.ops()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256 * 1024 * 1024
x = torch.randn(n, device='cuda'); y = torch.empty_like(x)
mm = ops.minmax(x)
d, z = torch.empty(1, device='cuda'), torch.empty(1, device='cuda')
ops.set_range_asym(mm[0:1], mm[1:2], 8, 1e-8, False, d, z)
spec = ops.spec(d, z, None, 8)
for _ in range(4):
    ops.qdq(x, spec, out=y)
torch.cuda.synchronize()
