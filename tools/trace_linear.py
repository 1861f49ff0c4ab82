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


# residual + fused LayerNorm (cluster kernel): the timeline buffer is passed through TQ_LINEAR_TRACE_PTR
names_ln = {4: 'first_full', 6: 'mma_issued', 8: 'acc_ready', 11: 'pass1_done', 12: 'stats_exchanged', 9: 'epi_done', 10: 'teardown'}
for (M, N, K) in [(4096, 768, 768), (4096, 768, 3072)]:
    a = torch.randint(-255, 256, (M, K), device=dev).to(torch.bfloat16)
    w = torch.randint(-128, 128, (N, K), device=dev).to(torch.bfloat16)
    r = torch.randint(-128, 128, (M, N), device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    gamma = torch.ones(N, device=dev); beta = torch.zeros(N, device=dev)
    d = torch.tensor([0.02], device=dev); z = torch.tensor([128.0], device=dev)
    od = torch.tensor([0.05], device=dev); oz = torch.tensor([120.0], device=dev)
    ws_ = torch.tensor([0.001], device=dev); sg = torch.tensor(True, device=dev)
    a_spec = ops.spec(d, z, None, 8); o_spec = ops.spec(od, oz, None, 8); w_spec = ops.spec(ws_, None, sg, 8)
    yc = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    trace = torch.zeros(16, dtype=torch.int64, device=dev)
    os.environ['TQ_LINEAR_TRACE_PTR'] = hex(trace.data_ptr())
    for bn in (None, '256'):
        if bn:
            os.environ['TQ_LINEAR_BN'] = bn
        for it in range(3):
            ops.linear_res_ln(a, w, bias, M, N, K, a_spec, w_spec, 1, o_spec, r, a_spec, o_spec, gamma, beta, 1e-12, o_spec,
                              out_ctr=yc)
        torch.cuda.synchronize()
        t = trace.tolist()
        print((M, N, K), 'res+LN bn=%s' % (bn or 'auto'), ' '.join(f'{n}={t[i] - t[0]}' for i, n in names_ln.items()))
    os.environ.pop('TQ_LINEAR_BN', None)
    os.environ.pop('TQ_LINEAR_TRACE_PTR', None)


# 8-bit operand mode (kind::i8): same timeline through TQ_LINEAR_TRACE_PTR
for (M, N, K, act) in [(4096, 2304, 768, 0), (4096, 3072, 768, 1)]:
    a8 = torch.randint(0, 256, (M, K), device=dev).to(torch.uint8)
    w8 = torch.randint(-128, 128, (N, K), device=dev).to(torch.int8)
    rsum = w8.to(torch.int32).sum(dim=1, dtype=torch.int32).contiguous()
    bias = torch.randn(N, device=dev)
    d = torch.tensor([0.02], device=dev); z = torch.tensor([128.0], device=dev)
    od = torch.tensor([50.0], device=dev); oz = torch.tensor([120.0], device=dev)
    ws_ = torch.tensor([0.001], device=dev); sg = torch.tensor(True, device=dev)
    a_spec = ops.spec(d, z, None, 8); o_spec = ops.spec(od, oz, None, 8); w_spec = ops.spec(ws_, None, sg, 8)
    y8 = torch.empty(M, N, device=dev, dtype=torch.uint8)
    trace = torch.zeros(16, dtype=torch.int64, device=dev)
    os.environ['TQ_LINEAR_TRACE_PTR'] = hex(trace.data_ptr())
    os.environ['TQ_PDL'] = '0'
    for kind in ('i8', 'bf16'):
        a_bf = (a8.float() - 128).to(torch.bfloat16); w_bf = w8.to(torch.bfloat16)
        yc = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for it in range(3):
            if kind == 'i8':
                ops.linear_i8(a8, w8, rsum, bias, M, N, K, a_spec, w_spec, 1, act, o_spec, 1, out_i8=y8)
            else:
                ops.linear(a_bf, w_bf, bias, M, N, K, 1, a_spec, w_spec, 1, act, o_spec, 1, want_f32=False, want_ctr=True)
        torch.cuda.synchronize()
        t = trace.tolist()
        print((M, N, K, act), kind, ' '.join(f'{n}={t[i] - t[0]}' for i, n in enumerate(names) if i in (1, 2, 4, 5, 6, 7, 8, 9, 10)))
    os.environ.pop('TQ_LINEAR_TRACE_PTR', None)
