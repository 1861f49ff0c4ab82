"""Per-tile clock64 timeline of CTA 0 of the fused linear on the four BERT-base GEMMs in the engine's modes
(csrc/tq_linear.cu TQ_TTRACE): when do the parameters, the accumulator and the epilogue of every tile start and
end, relative to the start of the CTA?  Shows which of {main loop, parameter set-up, epilogue} bounds a tile."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))
import tq_native
ops = tq_native.ops()
dev = 'cuda'
M = 4096
keep = []
os.environ['TQ_PDL'] = '0'


def spec(scale, zp=None, signed=None, n=1):
    d = torch.full((n,), scale, device=dev)
    z = None if zp is None else torch.full((n,), float(zp), device=dev)
    s = None if signed is None else torch.tensor(signed, device=dev)
    keep.extend([d, z, s])
    return ops.spec(d, z, s, 8)


trace = torch.zeros(16, dtype=torch.int64, device=dev)
tiles = torch.zeros(64, dtype=torch.int64, device=dev)
os.environ['TQ_LINEAR_TRACE_PTR'] = hex(trace.data_ptr())
os.environ['TQ_LINEAR_TRACE_TILES'] = hex(tiles.data_ptr())
names = ['par0', 'par1', 'acc', 'epiE', 'mmaS', 'ld1|drained', 'mmaE', 'prodE']


def run(label, fn):
    for _ in range(3):
        tiles.zero_()
        fn()
    torch.cuda.synchronize()
    t0 = trace.tolist()
    tt = tiles.view(8, 8).tolist()
    extra = f' ln_pass1_done={t0[11] - t0[0]} ln_stats_exchanged={t0[12] - t0[0]}' if 'LayerNorm' in label else ''
    print(f'{label}: setup_done={t0[1] - t0[0]} teardown={t0[10] - t0[0]}{extra}')
    for i, row in enumerate(tt):
        if not any(row):
            continue
        print('   tile %d  ' % i + ' '.join('%s=%6d' % (n, v - t0[0]) if v else '%s=     -' % n for n, v in zip(names, row)))


for N, K, label in [(2304, 768, 'qkv'), (768, 768, 'attn_out'), (3072, 768, 'ffn_in'), (768, 3072, 'ffn_out')]:
    a8 = torch.randint(0, 256, (M, K), device=dev).to(torch.uint8)
    a_bf = (a8.float() - 128).to(torch.bfloat16)
    w8 = torch.randint(-128, 128, (N, K), device=dev).to(torch.int8)
    w_bf = w8.to(torch.bfloat16)
    rsum = w8.to(torch.int32).sum(dim=1, dtype=torch.int32).contiguous()
    bias = torch.randn(N, device=dev) * 0.1
    r8 = torch.randint(0, 256, (M, N), device=dev).to(torch.uint8)
    gamma, beta = torch.ones(N, device=dev), torch.zeros(N, device=dev)
    a_sp, w_sp = spec(0.02, 128), spec(0.001, None, True)
    y8 = torch.empty(M, N, device=dev, dtype=torch.uint8)
    yc = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    if label == 'qkv':
        o_sp = spec(0.05, 120, None, N)
        run('i8 QKV 2304x768 per-column quantizers, bf16 out',
            lambda: ops.linear_i8(a8, w8, rsum, bias, M, N, K, a_sp, w_sp, 1, 0, o_sp, N, out_ctr=yc))
        o3, w3 = spec(0.05, 120, None, 3), spec(0.001, None, True, 3)
        run('LEAN i8 QKV 2304x768 three segments, bf16 out',
            lambda: ops.linear_seg_i8(a8, w8, rsum, bias, M, N, K, a_sp, w3, o3, 3, 0, out_ctr=yc))
    elif label == 'ffn_in':
        o_sp = spec(0.05, 120)
        run('bf16 FFN-in 3072x768 GELU, u8 out',
            lambda: ops.linear_bf16_o8(a_bf, w_bf, bias, M, N, K, a_sp, w_sp, 1, 1, o_sp, 1, y8))
        run('i8 FFN-in 3072x768 GELU, u8 out',
            lambda: ops.linear_i8(a8, w8, rsum, bias, M, N, K, a_sp, w_sp, 1, 1, o_sp, 1, out_i8=y8))
        run('LEAN i8 FFN-in 3072x768 GELU, u8 out',
            lambda: ops.linear_seg_i8(a8, w8, rsum, bias, M, N, K, a_sp, w_sp, o_sp, 1, 1, out_i8=y8))
        run('LEAN i8 FFN-in 3072x768 no activation, u8 out',
            lambda: ops.linear_seg_i8(a8, w8, rsum, bias, M, N, K, a_sp, w_sp, o_sp, 1, 0, out_i8=y8))
    else:
        o_sp, r_sp, u_sp, z_sp = spec(0.05, 120), spec(0.03, 128), spec(0.06, 125), spec(0.03, 128)
        for lean_flag in ('0', '1'):
            os.environ['TQ_LINEAR_LEAN'] = lean_flag
            run(f'{"LEAN" if lean_flag == "1" else "general"} i8 {label} {N}x{K} residual + LayerNorm, u8 out',
                lambda: ops.linear_res_ln_i8(a8, w8, rsum, bias, M, N, K, a_sp, w_sp, 1, o_sp, r8, r_sp, u_sp, gamma, beta, 1e-12,
                                             z_sp, y8))


# ---- PEG kernels (group-by-group accumulation): slot 5 = all groups drained
os.environ['TQ_LINEAR_LEAN'] = '1'
G = 6
for N, K, label, ln in [(2304, 768, 'PEG qkv', False), (3072, 768, 'PEG ffn_in', False), (768, 768, 'PEG attn_out + LN', True), (768, 3072, 'PEG ffn_out + LN (A per-tensor)', True)]:
    g = 1 if K == 3072 else G
    a8 = torch.randint(0, 256, (M, K), device=dev).to(torch.uint8)
    w8 = torch.randint(-128, 128, (N, K), device=dev).to(torch.int8)
    grs = w8.to(torch.int32).view(N, g, K // g).sum(dim=2, dtype=torch.int32).t().contiguous()
    bias = torch.randn(N, device=dev) * 0.1
    r8 = torch.randint(0, 256, (M, N), device=dev).to(torch.uint8)
    gamma, beta = torch.ones(N, device=dev), torch.zeros(N, device=dev)
    nseg = N // 128
    a_sp, w_sp = spec(0.02, 128, None, g), spec(0.001, None, True)
    o_sp, r_sp, u_sp, z_sp = spec(0.05, 120, None, nseg), spec(0.03, 128, None, nseg), spec(0.06, 125, None, nseg), spec(0.03, 128, None, nseg)
    y8 = torch.empty(M, N, device=dev, dtype=torch.uint8)
    if ln:
        run(f'{label} {N}x{K}', lambda: ops.linear_peg_res_ln_i8(a8, w8, grs, bias, M, N, K, a_sp, g, w_sp, 1, o_sp, nseg, r8, r_sp, nseg,
                                                                 u_sp, nseg, gamma, beta, 1e-12, z_sp, nseg, 128, y8))
    else:
        run(f'{label} {N}x{K}', lambda: ops.linear_peg_i8(a8, w8, grs, bias, M, N, K, a_sp, g, w_sp, 1, o_sp, nseg, 128, 1 if N == 3072 else 0,
                                                          out_i8=y8))
