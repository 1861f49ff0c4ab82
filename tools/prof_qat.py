#!/usr/bin/env python
"""One launch of each tq_qdq_bwd_f32 kernel variant at an HBM-sized shape, for
  ncu --set full --clock-control none --import-source on -k regex:qdq_bwd -c 4 -o gpurun_out/<name> python tools/prof_qat.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))
import tq_native  # noqa: E402

ops = tq_native.ops()
dev = 'cuda'
for outer, C, inner in [(1, 1, 64 * 1024 * 1024), (65536, 768, 1), (1, 30522, 768), (64, 96, 4096)]:
    n = outer * C * inner
    x = torch.randn(n, device=dev) * 3
    g = torch.randn(n, device=dev)
    delta = torch.full((C,), 0.03, device=dev)
    zf = torch.full((C,), 120.3, device=dev)
    ops.qdq_bwd(x, g, ops.spec(delta, zf, None, 8), C, outer, C, inner)
    torch.cuda.synchronize()
    del x, g
print('done')
