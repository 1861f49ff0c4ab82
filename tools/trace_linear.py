"""clock64 timeline of CTA 0 of the fused linear (debug aid): slots written by csrc/tq_linear.cu."""
import os, sys, torch, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))
import tq_native
from tq_native import QSpec
ops = tq_native.ops()
dev = 'cuda'
names = ['start', 'setup_done', 'prod_first', 'prod_done', 'mma_first_full', 'mma_second_full', 'mma_done', 'epi_params_done', 'epi_acc_ready', 'epi_done', 'teardown']
for (M, N, K, act) in [(4096, 768, 768, 0), (4096, 3072, 768, 1), (4096, 768, 3072, 0), (4096, 2304, 768, 0)]:
    a = torch.randint(-255, 256, (M, K), device=dev).to(torch.bfloat16)
    w = torch.randint(-128, 128, (N, K), device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    d = torch.tensor([0.02], device=dev); z = torch.tensor([128.0], device=dev)
    od = torch.tensor([50.0], device=dev); oz = torch.tensor([120.0], device=dev)
    ws_ = torch.tensor([0.001], device=dev); sg = torch.tensor(True, device=dev)
    a_spec = ops.spec(d, z, None, 8); o_spec = ops.spec(od, oz, None, 8); w_spec = ops.spec(ws_, None, sg, 8)
    y = torch.empty(M, N, device=dev)
    yc = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for label, yp, ycp in (('fp32-out', y.data_ptr(), None), ('bf16-out', None, yc.data_ptr()), ('both', y.data_ptr(), yc.data_ptr())):
        trace = torch.zeros(16, dtype=torch.int64, device=dev)
        for it in range(3):
            rc = ops.lib.tq_linear_qdq_bf16(a.data_ptr(), w.data_ptr(), bias.data_ptr(), yp, ycp, M, N, K, 1,
                                            a_spec, w_spec, 1, act, o_spec, 1, None, trace.data_ptr(), 128,
                                            torch.cuda.current_stream().cuda_stream)
            assert rc == 0
        torch.cuda.synchronize()
        t = trace.tolist()
        print((M, N, K, act), label, ' '.join(f'{n}={t[i] - t[0]}' for i, n in enumerate(names) if n in ('mma_done', 'epi_acc_ready', 'epi_done', 'teardown')))
