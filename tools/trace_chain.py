"""Per-stage clock64 timeline of the encoder chain kernel (tq_chain_plan_run) on BERT-base layer stages
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
LAYERS = int(os.environ.get('TRACE_LAYERS', '2'))
mask = torch.zeros(M // 128, 128, device=dev)
s_sp, p_sp = spec(0.08, 130), spec(1.0 / 255, 0)
att = ops.chain_attention_stage(qkv, c, D, 12, o_sp, o_sp, o_sp, s_sp, p_sp, a_sp, mask)
layer = [att,
         cs(2, c, wg[0], wg[1], wg[2], a, D, D, a_sp, w_sp, o_sp, 1, x, a_sp, o_sp, a_sp, gamma, beta, 1e-12),
         cs(1, a, wf[0], wf[1], wf[2], f, I, D, a_sp, w_sp, o_sp),
         cs(2, f, wh[0], wh[1], wh[2], x, D, I, a_sp, w_sp, o_sp, 1, a, a_sp, o_sp, a_sp, gamma, beta, 1e-12),
         cs(0, x, wq[0], wq[1], wq[2], qkv, 3 * D, D, a_sp, w3_sp, o3_sp, 3)]
names = ['attention', 'attn-out + LN', 'FFN-in GELU', 'FFN-out + LN', 'next QKV']
if os.environ.get('TRACE_ATT', '1') == '0':          # GEMM stages only (the kernel instantiation without attention code)
    layer, names = layer[1:], names[1:]
stages = [cs(0, x, wq[0], wq[1], wq[2], qkv, 3 * D, D, a_sp, w3_sp, o3_sp, 3)] + layer * LAYERS
names = ['QKV(0)'] + names * LAYERS
n = len(stages)
plan = ops.chain_plan(stages, M)
n_cta = (M // 128) * (D // 192)
trace = torch.zeros(n_cta * n * 4, dtype=torch.int64, device=dev)
os.environ['TQ_LINEAR_TRACE_CHAIN'] = hex(trace.data_ptr())
os.environ['TQ_PDL'] = '0'
for _ in range(3):
    ops.chain_run(plan)
torch.cuda.synchronize()
t = trace.view(n_cta, n, 4).double()
t0 = t[:, 0, 0].clone()
rel = t - t0.view(-1, 1, 1)
print('cycles relative to each CTA\'s first stage top: mean (max) over %d CTAs; %d stages in one launch' % (n_cta, n))
for s in range(n):
    r = rel[:, s]
    print('  %-14s top=%7.0f (%7.0f)  first_acc=%7.0f  epi_done=%7.0f (%7.0f)  past_barrier=%7.0f   stage=%6.0f  main=%6.0f  epi=%6.0f  wait=%6.0f' % (
        names[s], r[:, 0].mean(), r[:, 0].max(), r[:, 1].mean(), r[:, 2].mean(), r[:, 2].max(), r[:, 3].mean(),
        (r[:, 3] - r[:, 0]).mean(), (r[:, 1] - r[:, 0]).mean(), (r[:, 2] - r[:, 1]).mean(), (r[:, 3] - r[:, 2]).mean()))
del os.environ['TQ_LINEAR_TRACE_CHAIN']
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.chain_run(plan)
e1.record()
torch.cuda.synchronize()
print('chain launch (%d layers): %.1f us' % (LAYERS, e0.elapsed_time(e1) * 50))
