"""Main-loop rate of the fused linear with 8-bit operands (tcgen05 kind::i8): cycles per 128-byte k-block
(= 128 K columns, four K=32 MMAs) from the kernel's clock64 timeline, a lone CTA vs every SM streaming.
Compare with tools/mainloop_probe.py (bf16: 64 K columns per k-block).  K = 8192, one tile per CTA."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))
import tq_native
ops = tq_native.ops()
dev = 'cuda'
K = 8192
d = torch.tensor([0.02], device=dev); z = torch.tensor([128.0], device=dev)
od = torch.tensor([5000.0], device=dev); oz = torch.tensor([120.0], device=dev)
wsd = torch.tensor([0.001], device=dev); sg = torch.tensor(True, device=dev)
a_spec = ops.spec(d, z, None, 8); o_spec = ops.spec(od, oz, None, 8); w_spec = ops.spec(wsd, None, sg, 8)
trace = torch.zeros(16, dtype=torch.int64, device=dev)
os.environ['TQ_LINEAR_TRACE_PTR'] = str(trace.data_ptr())


def probe(M, N, bn, stages):
    os.environ['TQ_LINEAR_BN'] = str(bn)
    if stages:
        os.environ['TQ_LINEAR_STAGES'] = str(stages)
    else:
        os.environ.pop('TQ_LINEAR_STAGES', None)
    a = torch.randint(0, 256, (M, K), device=dev, dtype=torch.uint8)
    w = torch.randint(-128, 128, (N, K), device=dev, dtype=torch.int8)
    rs = w.to(torch.int32).sum(dim=1, dtype=torch.int32).contiguous()
    y8 = torch.empty(M, N, device=dev, dtype=torch.uint8)
    for _ in range(3):
        rc = ops.lib.tq_linear_qdq_i8(a.data_ptr(), w.data_ptr(), rs.data_ptr(), None, None, None, y8.data_ptr(), M, N, K,
                                      a_spec, w_spec, 1, 0, o_spec, 1, torch.cuda.current_stream().cuda_stream)
        assert rc == 0, rc
    torch.cuda.synchronize()
    t = trace.tolist()
    return (t[8] - t[4]) / (K // 128), t[4] - t[0], t[10] - t[0]


print('%-34s %10s %12s %10s' % ('config (int8)', 'cyc/kblock', 'first_full', 'total'))
for grid_name in ('alone', 'all SMs'):
    for bn in (256, 192, 128, 64):
        for stages in (0, 3):
            n_cta = 1 if grid_name == 'alone' else 148
            r = probe(128 * n_cta, bn, bn, stages)
            print('%-34s %10.0f %12d %10d' % ('%s bn=%d stages=%s' % (grid_name, bn, stages or 'max'), r[0], r[1], r[2]),
                  flush=True)
