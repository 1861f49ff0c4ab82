"""Per-stage clock64 timeline of tq_linear_chain_i8 on the four BERT-base GEMM stages of one encoder layer
(csrc/tq_linear.cu namespace chain, Params.trace): for every CTA and stage -- stage top, first accumulator ready,
epilogue done, past the stage barrier.  Prints the mean / max over CTAs relative to each CTA's own first stamp."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))
import tq_native
ops = tq_native.ops()
dev = 'cuda'
M, D, I = 4096, 768, 3072
keep = []


def spec(scale, zp=None, signed=None, n=1):
    d = torch.full((n,), scale, device=dev)
    z = None if zp is None else torch.full((n,), float(zp), device=dev)
    s = None if signed is None else torch.tensor(signed, device=dev)
    keep.extend([d, z, s])
    return ops.spec(d, z, s, 8)


def weight(N, K):
    w8 = torch.randint(-128, 128, (N, K), device=dev).to(torch.int8)
    return w8, w8.to(torch.int32).sum(dim=1, dtype=torch.int32).contiguous(), torch.randn(N, device=dev) * 0.1


u8 = lambda *s: torch.randint(0, 256, s, device=dev).to(torch.uint8)
c, x, a, f = u8(M, D), u8(M, D), u8(M, D), u8(M, I)
qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
wg, wf, wh, wq = weight(D, D), weight(I, D), weight(D, I), weight(3 * D, D)
gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
a_sp, w_sp, w3_sp, o_sp, o3_sp = spec(0.02, 128), spec(0.001, None, True), spec(0.001, None, True, 3), spec(0.05, 120), spec(0.05, 120, None, 3)
cs = ops.chain_stage
stages = [cs(2, c, wg[0], wg[1], wg[2], a, D, D, a_sp, w_sp, o_sp, 1, x, a_sp, o_sp, a_sp, gamma, beta, 1e-12),
          cs(1, a, wf[0], wf[1], wf[2], f, I, D, a_sp, w_sp, o_sp),
          cs(2, f, wh[0], wh[1], wh[2], x, D, I, a_sp, w_sp, o_sp, 1, a, a_sp, o_sp, a_sp, gamma, beta, 1e-12),
          cs(0, x, wq[0], wq[1], wq[2], qkv, 3 * D, D, a_sp, w3_sp, o3_sp, 3)]
n_cta = (M // 128) * (D // 192)
trace = torch.zeros(n_cta * 16, dtype=torch.int64, device=dev)
os.environ['TQ_LINEAR_TRACE_CHAIN'] = hex(trace.data_ptr())
os.environ['TQ_PDL'] = '0'
for _ in range(3):
    ops.linear_chain_i8(stages, M)
torch.cuda.synchronize()
t = trace.view(n_cta, 4, 4).double()
t0 = t[:, 0, 0].clone()
rel = t - t0.view(-1, 1, 1)
names = ['attn-out + LN', 'FFN-in GELU', 'FFN-out + LN', 'next QKV']
print('cycles relative to each CTA\'s first stage top: mean (max) over %d CTAs' % n_cta)
for s in range(4):
    r = rel[:, s]
    print('  %-14s top=%7.0f (%7.0f)  first_acc=%7.0f (%7.0f)  epi_done=%7.0f (%7.0f)  past_barrier=%7.0f (%7.0f)   stage=%6.0f  main=%6.0f  epi=%6.0f  wait=%6.0f' % (
        names[s], r[:, 0].mean(), r[:, 0].max(), r[:, 1].mean(), r[:, 1].max(), r[:, 2].mean(), r[:, 2].max(), r[:, 3].mean(), r[:, 3].max(),
        (r[:, 3] - r[:, 0]).mean(), (r[:, 1] - r[:, 0]).mean(), (r[:, 2] - r[:, 1]).mean(), (r[:, 3] - r[:, 2]).mean()))
del os.environ['TQ_LINEAR_TRACE_CHAIN']
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.linear_chain_i8(stages, M)
e1.record()
torch.cuda.synchronize()
print('chain launch: %.1f us' % (e0.elapsed_time(e1) * 50))
