"""Driver for ncu: the fused linear at the four BERT-base shapes (M = 32*128 tokens)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))
import tq_native
ops = tq_native.ops()
dev = 'cuda'
shapes = [(4096, 768, 768, 0), (4096, 2304, 768, 0), (4096, 3072, 768, 1), (4096, 768, 3072, 0)]
if len(sys.argv) > 1:
    shapes = [shapes[int(sys.argv[1])]]
for (M, N, K, act) in shapes:
    a = torch.randint(-255, 256, (M, K), device=dev).to(torch.bfloat16)
    w = torch.randint(-128, 128, (N, K), device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    d = torch.tensor([0.02], device=dev); z = torch.tensor([128.0], device=dev)
    od = torch.tensor([50.0], device=dev); oz = torch.tensor([120.0], device=dev)
    a_spec = ops.spec(d, z, None, 8); o_spec = ops.spec(od, oz, None, 8)
    ws = torch.tensor([0.001], device=dev); sg = torch.tensor(True, device=dev)
    w_spec = ops.spec(ws, None, sg, 8)
    for _ in range(3):
        ops.linear(a, w, bias, M, N, K, 1, a_spec, w_spec, 1, act, o_spec, 1)
    torch.cuda.synchronize()
