"""Main-loop rate of the fused linear: cycles per 64-wide k-block from the kernel's clock64 timeline
(first smem stage full -> accumulator complete), for each tile width, CTA pairing, ring depth and
grid size (one CTA / pair alone vs every SM streaming).  K = 4096, one tile per CTA."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))
import tq_native
ops = tq_native.ops()
dev = 'cuda'
K = 4096
keep = []
d = torch.tensor([0.02], device=dev); z = torch.tensor([128.0], device=dev)
od = torch.tensor([50.0], device=dev); oz = torch.tensor([120.0], device=dev)
wsd = torch.tensor([0.001], device=dev); sg = torch.tensor(True, device=dev)
a_spec = ops.spec(d, z, None, 8); o_spec = ops.spec(od, oz, None, 8); w_spec = ops.spec(wsd, None, sg, 8)


def probe(M, N, ctas, bn, stages):
    os.environ['TQ_LINEAR_BN'] = str(bn)
    os.environ['TQ_LINEAR_CTAS'] = str(ctas)
    if stages:
        os.environ['TQ_LINEAR_STAGES'] = str(stages)
    else:
        os.environ.pop('TQ_LINEAR_STAGES', None)
    a = torch.randint(-255, 256, (M, K), device=dev).to(torch.bfloat16)
    w = torch.randint(-128, 128, (N, K), device=dev).to(torch.bfloat16)
    yc = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    trace = torch.zeros(16, dtype=torch.int64, device=dev)
    for _ in range(3):
        rc = ops.lib.tq_linear_qdq_bf16(a.data_ptr(), w.data_ptr(), None, None, yc.data_ptr(), M, N, K, 1, a_spec,
                                        w_spec, 1, 0, o_spec, 1, None, trace.data_ptr(), 128,
                                        torch.cuda.current_stream().cuda_stream)
        assert rc == 0, rc
    torch.cuda.synchronize()
    t = trace.tolist()
    return (t[8] - t[4]) / (K // 64), t[4] - t[0], t[10] - t[0]


print('%-34s %10s %12s %10s' % ('config', 'cyc/kblock', 'first_full', 'total'))
for grid_name, ctas_list in (('alone', (1, 2)), ('all SMs', (1, 2))):
    for ctas in ctas_list:
        for bn in (256, 192, 128, 64):
            if ctas == 2 and bn < 128:
                continue
            for stages in (0, 2, 3):
                n_cta = ctas if grid_name == 'alone' else 148
                M, N = 128 * n_cta, bn
                r = probe(M, N, ctas, bn, stages)
                print('%-34s %10.0f %12d %10d' % ('%s ctas=%d bn=%d stages=%s' % (grid_name, ctas, bn, stages or 'max'),
                                                  r[0], r[1], r[2]), flush=True)
