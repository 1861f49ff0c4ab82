"""Tile-width sweep of the fused linear on the four BERT-base GEMMs with the engine's epilogues
(QKV: per-column quantizers, attention-out / FFN-out: residual + second quantizer, FFN-in: GELU).
TQ_LINEAR_BN forces the tile width; the last column is the library's own pick."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))
import tq_native
ops = tq_native.ops()
dev = 'cuda'
M = 4096
keep = []


def spec(scale, zp=None, signed=None, n=1):
    d = torch.full((n,), scale, device=dev) * (1 + 0.01 * torch.arange(n, device=dev) / max(n, 1))
    z = None if zp is None else torch.full((n,), float(zp), device=dev)
    s = None if signed is None else torch.tensor(signed, device=dev)
    keep.extend([d, z, s])
    return ops.spec(d, z, s, 8)


def bench(fn, per_graph=20, replays=5):
    """average us per launch; launches are replayed from a CUDA graph (no host launch cost)"""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(per_graph):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (per_graph * replays) * 1e3


cases = [('qkv   2304x768 percol', 2304, 768, 0, True, False), ('attout 768x768 res', 768, 768, 0, False, True),
         ('ffn_in 3072x768 gelu', 3072, 768, 1, False, False), ('ffnout 768x3072 res', 768, 3072, 0, False, True),
         ('attout 768x768 res+LN', 768, 768, 0, False, 'ln'), ('ffnout 768x3072 res+LN', 768, 3072, 0, False, 'ln'),
         ('i8 qkv 2304x768 percol', 2304, 768, 0, True, 'i8'), ('i8 ffn_in 3072x768 gelu', 3072, 768, 1, False, 'i8'),
         ('i8 attout 768x768 res+LN', 768, 768, 0, False, 'i8ln'), ('i8 ffnout 768x3072 res+LN', 768, 3072, 0, False, 'i8ln'),
         ('i8 ffn_in 3072x768 plain u8', 3072, 768, 0, False, 'i8'), ('i8 ffn_in gelu bf16-out', 3072, 768, 1, False, 'i8bf'),
         ('i8 ffn_in plain bf16-out', 3072, 768, 0, False, 'i8bf')]
combos = [(1, '256'), (1, '192'), (1, '128'), (1, '96'), (2, '256'), (2, '192'), (2, '128'), (None, None)]
if os.environ.get('SWEEP_QUICK'):
    combos = [(1, '256'), (1, '192'), (None, None)]
print('%-24s' % 'case (ctas x bn)', ' '.join('%8s' % ('auto' if c is None else '%dx%s' % (c, b)) for c, b in combos))
for name, N, K, act, percol, res in cases:
    a = torch.randint(-128, 128, (M, K), device=dev).to(torch.bfloat16)
    w = torch.randint(-128, 128, (N, K), device=dev).to(torch.bfloat16)
    bias = torch.randn(N, device=dev) * 0.1
    r = torch.randint(-128, 128, (M, N), device=dev).to(torch.bfloat16)
    a_sp = spec(0.02, 128); w_sp = spec(0.001, None, True, N)
    o_sp = spec(0.05, 120, None, N if percol else 1); r_sp = spec(0.03, 128); o2_sp = spec(0.06, 125)
    yc = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    if res in ('i8', 'i8ln', 'i8bf'):
        a8 = torch.randint(0, 256, (M, K), device=dev).to(torch.uint8)
        w8 = torch.randint(-128, 128, (N, K), device=dev).to(torch.int8)
        rsum = w8.to(torch.int32).sum(dim=1, dtype=torch.int32).contiguous()
        r8 = torch.randint(0, 256, (M, N), device=dev).to(torch.uint8)
        y8 = torch.empty(M, N, device=dev, dtype=torch.uint8)
        gamma = torch.ones(N, device=dev); beta = torch.zeros(N, device=dev); ln_sp = spec(0.03, 120)
        if res == 'i8ln':
            fn = lambda: ops.linear_res_ln_i8(a8, w8, rsum, bias, M, N, K, a_sp, w_sp, N, o_sp, r8, r_sp, o2_sp, gamma, beta,
                                              1e-12, ln_sp, y8)
        elif percol or res == 'i8bf':
            fn = lambda: ops.linear_i8(a8, w8, rsum, bias, M, N, K, a_sp, w_sp, N, act, o_sp, N if percol else 1, out_ctr=yc)
        else:
            fn = lambda: ops.linear_i8(a8, w8, rsum, bias, M, N, K, a_sp, w_sp, N, act, o_sp, 1, out_i8=y8)
    elif res == 'ln':
        gamma = torch.ones(N, device=dev); beta = torch.zeros(N, device=dev); ln_sp = spec(0.03, 120)
        fn = lambda: ops.linear_res_ln(a, w, bias, M, N, K, a_sp, w_sp, N, o_sp, r, r_sp, o2_sp, gamma, beta, 1e-12,
                                       ln_sp, out_ctr=yc)
    elif res:
        fn = lambda: ops.linear_res(a, w, bias, M, N, K, a_sp, w_sp, N, o_sp, 1, r, r_sp, o2_sp, 1, out_ctr=yc)
    else:
        fn = lambda: ops.linear(a, w, bias, M, N, K, 1, a_sp, w_sp, N, act, o_sp, N if percol else 1,
                                want_f32=False, want_ctr=True)
    row = []
    for ctas, bn in combos:
        if bn is None:
            os.environ.pop('TQ_LINEAR_BN', None)
            os.environ.pop('TQ_LINEAR_CTAS', None)
        else:
            os.environ['TQ_LINEAR_BN'] = bn
            os.environ['TQ_LINEAR_CTAS'] = str(ctas)
        row.append(bench(fn))
        print('   %s %s x %s: %.1f us' % (name, ctas, bn, row[-1]), file=sys.stderr, flush=True)
    flops = 2.0 * M * N * K
    print('%-24s' % name, ' '.join('%8.1f' % t for t in row), '  us   best %.0f TFLOP/s' % (flops / min(row) / 1e6))
