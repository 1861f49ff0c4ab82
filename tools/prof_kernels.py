"""Driver for ncu: one launch of every streaming / reduction kernel of the library that has no capture yet
(VERDICT r1 weak 8): per-tensor min/max, per-axis min/max, per-embedding-group QDQ, MSE grid, LayerNorm + QDQ,
embedding + LayerNorm + QDQ -- at > L2 sizes where the kernel is a streaming one."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))
import tq_native
ops = tq_native.ops()
dev = 'cuda'
rows, C = 87381, 768                       # 67.1 M elements = 268 MB fp32 (> 126 MB L2)
x = torch.randn(rows, C, device=dev)
y = torch.empty_like(x)
mn, mx = ops.minmax_axis(x, rows, C, 1)
gm, gM = ops.group_minmax(mn, mx, 6)
dv, zv = torch.empty(C, device=dev), torch.empty(C, device=dev)
ops.set_range_asym(gm.contiguous(), gM.contiguous(), 8, 1e-8, False, dv, zv)
sp = ops.spec(dv, zv, None, 8)
d1, z1 = torch.empty(1, device=dev), torch.empty(1, device=dev)
mm = ops.minmax(x)
ops.set_range_asym(mm[0:1], mm[1:2], 8, 1e-8, False, d1, z1)
sp1 = ops.spec(d1, z1, None, 8)
xc = torch.randint(-128, 128, (rows, C), device=dev).to(torch.bfloat16)
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
from quantization.quantizers import QMethods
from quantization.range_estimators import RangeEstimators, OptMethod
xm = torch.randn(32, 128, 768, device=dev) * 2
est = RangeEstimators.MSE.cls(quantizer=QMethods.asymmetric_uniform.cls(n_bits=8), opt_method=OptMethod.grid)
est(xm)
for _ in range(2):
    ops.minmax(x); ops.minmax_axis(x, rows, C, 1); ops.qdq(x, sp, rows, C, 1, out=y); ops.qdq(x, sp1, out=y)
    ops.ln_qdq(xc, sp1, 1, gamma, beta, 1e-12, sp1, 1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.minmax(x)
ops.minmax_axis(x, rows, C, 1)
ops.qdq(x, sp, rows, C, 1, out=y)
ops.qdq(x, sp1, out=y)
ops.ln_qdq(xc, sp1, 1, gamma, beta, 1e-12, sp1, 1)
est(xm)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
