"""Per-CTA start / end times (globaltimer) of the fused linear on the BERT-base GEMM shapes: launch
skew, tail, and the kernel span compared with the in-graph time per launch (tools/sweep_linear.py)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))
import tq_native
ops = tq_native.ops()
dev = 'cuda'
M = 4096
d = torch.tensor([0.02], device=dev); z = torch.tensor([128.0], device=dev)
od = torch.tensor([0.05], device=dev); oz = torch.tensor([120.0], device=dev)
wsd = torch.tensor([0.001], device=dev); sg = torch.tensor(True, device=dev)
a_spec = ops.spec(d, z, None, 8); o_spec = ops.spec(od, oz, None, 8); w_spec = ops.spec(wsd, None, sg, 8)
for (N, K, act) in [(768, 768, 0), (2304, 768, 0), (3072, 768, 1), (768, 3072, 0)]:
    a = torch.randint(-255, 256, (M, K), device=dev).to(torch.bfloat16)
    w = torch.randint(-128, 128, (N, K), device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev) * 0.1
    yc = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ws = torch.zeros(16 + 4 * 160, dtype=torch.int64, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for it in range(4):
        if it == 3:
            ev[0].record()
        rc = ops.lib.tq_linear_qdq_bf16(a.data_ptr(), w.data_ptr(), bias.data_ptr(), None, yc.data_ptr(), M, N, K, 1,
                                        a_spec, w_spec, 1, act, o_spec, 1, None, ws.data_ptr(), ws.numel() * 8,
                                        torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    ev[1].record()
    torch.cuda.synchronize()
    t = ws[16:].view(-1, 4).cpu()
    t = t[t[:, 1] > 0]
    s0 = t[:, 0].min().item()
    starts = (t[:, 0] - s0).float() / 1e3
    ends = (t[:, 1] - s0).float() / 1e3
    life = (t[:, 1] - t[:, 0]).float() / 1e3
    print('N=%d K=%d act=%d: %d CTAs | start skew max %.2f us | end min %.2f max %.2f us | CTA life min %.2f med %.2f '
          'max %.2f us | clk span med %d | event time %.2f us' %
          (N, K, act, t.shape[0], starts.max().item(), ends.min().item(), ends.max().item(), life.min().item(),
           life.median().item(), life.max().item(), int(t[:, 3].median().item()), ev[0].elapsed_time(ev[1]) * 1e3))
    order = torch.argsort(t[:, 1])
    print('   last CTAs to finish (smid, start, end):', [(int(t[i, 2]), round(starts[i].item(), 2), round(ends[i].item(), 2)) for i in order[-5:].tolist()])
    print('   first CTAs to finish:', [(int(t[i, 2]), round(starts[i].item(), 2), round(ends[i].item(), 2)) for i in order[:5].tolist()])
