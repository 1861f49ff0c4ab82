"""GPU parity tests proper: every call goes through the C ABI of libtq_b200.so on cuda:0.

1. golden vectors of the unmodified reference (tests/golden/*.npz), through the reference-facing
   Python API -- bit-exact integers, equal floats (tolerances in tests/parity_cases.py);
2. CUDA kernels vs the CPU oracle on seeded random inputs at sizes the oracle finishes in seconds;
3. BASELINE.json full-size tensors through size-independent properties (idempotence of QDQ,
   x_int on the integer grid and inside [int_min, int_max], min/max equal to a second
   implementation, linearity of the MSE loss accumulator);
4. edge cases: empty / ragged / misaligned inputs, NaN / inf, argument errors of the C ABI.
"""
import numpy as np
import pytest
import torch

import tq_native
from conftest import golden_cases
import parity_cases as P
from oracle import fakequant_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def ops():
    return tq_native.ops()


# ---- 1. golden vectors ------------------------------------------------------------------------
@pytest.mark.parametrize('case', golden_cases('quantizers'), ids=lambda c: c['name'])
def test_quantizer_golden(case, golden):
    P.check_quantizer_case(case, golden.file('quantizers'), DEV)


@pytest.mark.parametrize('case', golden_cases('estimators'), ids=lambda c: c['name'])
def test_estimator_golden(case, golden):
    P.check_estimator_case(case, golden.file('estimators'), DEV)


@pytest.mark.parametrize('case', golden_cases('mse'), ids=lambda c: c['name'])
def test_mse_golden(case, golden):
    P.check_mse_case(case, golden.file('mse'), DEV)


@pytest.mark.parametrize('case', golden_cases('manager'), ids=lambda c: c['name'])
def test_manager_golden(case, golden):
    P.check_manager_case(case, golden.file('manager'), DEV)


@pytest.mark.parametrize('case', golden_cases('linear'), ids=lambda c: c['name'])
def test_quant_linear_golden(case, golden):
    P.check_linear_case(case, golden.file('linear'), DEV, exact_gemm=False)


def test_division_free_quotient_is_exact(ops):
    """tq::div_rn (two FMA corrections of x * RN(1/s)) == IEEE division, bit for bit, on 1.2e10
    adversarial pairs (ties of the integer grid +- 4 ulp, random floats, random scales, 2-16 bits)."""
    out = torch.zeros(5, dtype=torch.int64, device=DEV)
    for seed in range(3):
        code = ops.lib.tq_selftest_div(1234 + seed, 148 * 16, 4096 * 4, out.data_ptr(), None)
        assert code == 0
    torch.cuda.synchronize()
    assert out.tolist()[:2] == [0, 0], (f'quotient mismatches {out[0].item()}, grid mismatches {out[1].item()}, '
                                        f'sub-2^-60 quotient mismatches {out[2].item()}')
    # the packed (FFMA2) forms the fused epilogues use: quot2 bit for bit, quant_int2_finite / quant_ctr2_finite integers
    assert out.tolist()[3:] == [0, 0], f'packed quotient mismatches {out[3].item()}, packed grid mismatches {out[4].item()}'


# ---- 2. kernels vs oracle on random inputs ------------------------------------------------------
def _asym(ops, xmin, xmax, n_bits, dev=DEV):
    xmin = torch.as_tensor(xmin, dtype=torch.float32, device=dev).reshape(-1)
    xmax = torch.as_tensor(xmax, dtype=torch.float32, device=dev).reshape(-1)
    d, z = torch.empty_like(xmin), torch.empty_like(xmin)
    ops.set_range_asym(xmin, xmax, n_bits, 1e-8, False, d, z)
    return d, z


@pytest.mark.parametrize('shape', [(32, 128, 768), (8, 12, 128, 128), (1000003,), (5,), (4, 7, 13),
                                   (32, 128, 3072), (32, 12, 128, 128), (64, 128, 512),   # the other BASELINE activation shapes, full size
                                   (8 * 1024 * 1024 + 4096 + 13,),      # bulk-copy staged kernel, ragged
                                   (20 * 1024 * 1024 + 4096 + 13,)])    # one-chunk-per-CTA kernel, ragged
@pytest.mark.parametrize('n_bits', [8, 4])
def test_qdq_tensor_vs_oracle(ops, shape, n_bits):
    rs = np.random.RandomState(1234)
    x = (rs.randn(*shape) * 3).astype(np.float32)
    xt = torch.from_numpy(x).to(DEV)
    d, z = _asym(ops, x.min(), x.max() * 0.7, n_bits)
    spec = ops.spec(d, z, None, n_bits)
    y = ops.qdq(xt, spec)
    yi, yc = ops.quant_int(xt, spec, want_f32=True, want_bf16=True)
    od, oz = O.asym_set_quant_range(x.min(), x.max() * 0.7, n_bits)
    P.assert_same(d, od, 'delta')
    P.assert_same(z, oz, 'zero_float')
    P.assert_same(yi, O.qdq_asym(x, od, oz, n_bits, return_int=True), 'x_int')
    P.assert_same(y, O.qdq_asym(x, od, oz, n_bits), 'x_quant')
    zp = O.asym_zero_point(oz, n_bits)
    P.assert_same(yc.float(), O.qdq_asym(x, od, oz, n_bits, return_int=True) - zp, 'centred bf16 grid')


@pytest.mark.parametrize('signed', [True, False])
def test_qdq_sym_vs_oracle(ops, signed):
    rs = np.random.RandomState(7)
    x = (rs.randn(64, 3072) * 0.02).astype(np.float32)
    if not signed:
        x = np.abs(x)
    xt = torch.from_numpy(x).to(DEV)
    d = torch.empty(1, device=DEV)
    s = torch.empty((), dtype=torch.bool, device=DEV)
    mm = ops.minmax(xt)
    ops.set_range_sym(mm[0:1], mm[1:2], 8, 1e-8, False, d, s)
    od, osg = O.sym_set_quant_range(x.min(), x.max(), 8)
    assert bool(s.item()) == osg == signed
    P.assert_same(d, od, 'delta')
    P.assert_same(ops.qdq(xt, ops.spec(d, None, s, 8)), O.qdq_sym(x, od, osg, 8), 'x_quant')


@pytest.mark.parametrize('shape,axis', [((32, 128, 768), 2), ((16, 128, 3072), 2), ((64, 768), 1),
                                        ((2, 6, 5, 4), 1), ((3, 5, 7), 2), ((768, 3072), 0)])
def test_qdq_and_minmax_axis_vs_oracle(ops, shape, axis):
    rs = np.random.RandomState(99)
    x = (rs.randn(*shape) * 2).astype(np.float32)
    x[..., 0] *= 30
    xt = torch.from_numpy(x).to(DEV)
    C = shape[axis]
    outer = int(np.prod(shape[:axis])) if axis else 1
    inner = int(np.prod(shape[axis + 1:])) if axis + 1 < len(shape) else 1
    mn, mx = ops.minmax_axis(xt, outer, C, inner)
    omn, omx = O.minmax_axis(x, axis)
    P.assert_same(mn, omn, 'axis min')
    P.assert_same(mx, omx, 'axis max')
    d, z = _asym(ops, omn, omx, 8)
    od, oz = O.asym_set_quant_range(omn, omx, 8)
    P.assert_same(d, od, 'delta')
    y = ops.qdq(xt, ops.spec(d, z, None, 8), outer, C, inner)
    if axis == 0:
        ref = O.qdq_asym(x, od, oz, 8, per_channel=True)
    else:
        ref = O.qdq_asym(x, od, oz, 8, axis=axis)
    P.assert_same(y, ref, 'per-axis x_quant')


@pytest.mark.parametrize('n_groups,permute', [(6, False), (6, True), (768, False), (1, True), (3, True)])
def test_group_minmax_vs_oracle(ops, n_groups, permute):
    rs = np.random.RandomState(5)
    mn = -np.abs(rs.randn(768)).astype(np.float32)
    mx = np.abs(rs.randn(768)).astype(np.float32)
    ranges = (mx - mn).astype(np.float32)
    ranges[10] = ranges[500]          # a tie: resolved by index (stable)
    a, b = ops.group_minmax(torch.from_numpy(mn).to(DEV), torch.from_numpy(mx).to(DEV), n_groups,
                            torch.from_numpy(ranges).to(DEV) if permute else None)
    oa, ob = O.group_minmax(mn, mx, n_groups, O.stable_order(ranges) if permute else None)
    P.assert_same(a, oa, 'group min')
    P.assert_same(b, ob, 'group max')


def test_group_minmax_bad_groups(ops):
    v = torch.zeros(768, device=DEV)
    with pytest.raises(tq_native.TQError):
        ops.group_minmax(v, v, 7)


@pytest.mark.parametrize('n', [1, 3, 1000, 4096 * 771 + 3])
def test_minmax_tensor(ops, n):
    rs = np.random.RandomState(n % 1000)
    x = rs.randn(n).astype(np.float32)
    mm = ops.minmax(torch.from_numpy(x).to(DEV))
    P.assert_same(mm, [x.min(), x.max()], 'minmax')
    # workspace is self-cleaning: a second call on different data must not see stale state
    x2 = (x * 0.5).astype(np.float32)
    mm2 = ops.minmax(torch.from_numpy(x2).to(DEV))
    P.assert_same(mm2, [x2.min(), x2.max()], 'minmax second call')


def test_minmax_nan_propagates(ops):
    x = torch.randn(100000, device=DEV)
    x[77777] = float('nan')
    mm = ops.minmax(x)
    assert torch.isnan(mm).all()
    xa = torch.randn(64, 96, device=DEV)
    xa[5, 17] = float('nan')
    mn, mx = ops.minmax_axis(xa, 64, 96, 1)
    nan_cols = torch.isnan(mn).nonzero().flatten().tolist()
    assert nan_cols == [17] and torch.isnan(mx[17])
    # several column blocks and slabs: the per-column-block ticket shares the first flag word of a block with that
    # column's NaN flag (bit 0) -- NaNs in block-leading columns, twice in a row (the workspace must come back clean)
    xb = torch.randn(4096, 768, device=DEV)
    xb[100, 0] = xb[4000, 256] = xb[7, 300] = xb[2222, 767] = float('nan')
    for _ in range(2):
        mn, mx = ops.minmax_axis(xb, 4096, 768, 1)
        assert torch.isnan(mn).nonzero().flatten().tolist() == [0, 256, 300, 767]
        assert torch.isnan(mx).nonzero().flatten().tolist() == [0, 256, 300, 767]
    clean = torch.nan_to_num(xb, nan=0.0)
    mn, mx = ops.minmax_axis(clean, 4096, 768, 1)
    assert torch.equal(mn, clean.min(dim=0).values) and torch.equal(mx, clean.max(dim=0).values)


def test_misaligned_views(ops):
    """tensors whose data pointer is not 16-byte aligned take the scalar kernels."""
    base = torch.randn(4099, device=DEV)
    x = base[1:]                       # 4-byte offset
    d, z = _asym(ops, -2.0, 2.0, 8)
    spec = ops.spec(d, z, None, 8)
    xn = x.cpu().numpy()
    od, oz = O.asym_set_quant_range(-2.0, 2.0, 8)
    P.assert_same(ops.qdq(x, spec), O.qdq_asym(xn, od, oz, 8), 'misaligned qdq')
    P.assert_same(ops.minmax(x), [xn.min(), xn.max()], 'misaligned minmax')


def test_mse_kernel_vs_oracle(ops):
    rs = np.random.RandomState(11)
    x = (rs.randn(70001) * 2).astype(np.float32)     # ragged; spans several CTAs
    thr = [(-0.1 * c, 0.13 * c) for c in range(1, 41)]
    tabs = [O.candidate_qparams(a, b, 8, False) for a, b in thr]
    cand = np.stack([np.array([t[i] for t in tabs], np.float32).reshape(-1) for i in range(4)])
    acc = torch.zeros(len(thr), dtype=torch.float64, device=DEV)
    xt = torch.from_numpy(x).to(DEV)
    ops.mse_sse(xt, torch.from_numpy(cand).to(DEV), len(thr), acc)
    ref = np.array([O.sse(x, *t) for t in tabs])
    np.testing.assert_allclose(acc.cpu().numpy(), ref, rtol=1e-6)
    first = acc.clone()
    ops.mse_sse(xt, torch.from_numpy(cand).to(DEV), len(thr), acc)     # accumulates; deterministic
    assert torch.equal(acc, 2 * first)


# ---- 3. full-size properties ----------------------------------------------------------------------
@pytest.mark.parametrize('shape', [(32, 128, 768), (32, 12, 128, 128), (32, 128, 3072), (64, 128, 512)])
def test_fullsize_properties(ops, shape):
    g = torch.Generator(device='cpu').manual_seed(1234)
    x = (torch.randn(shape, generator=g) * 3).to(DEV)
    mm = ops.minmax(x)
    assert mm[0].item() == x.min().item() and mm[1].item() == x.max().item()
    d, z = torch.empty(1, device=DEV), torch.empty(1, device=DEV)
    ops.set_range_asym(mm[0:1], mm[1:2], 8, 1e-8, False, d, z)
    spec = ops.spec(d, z, None, 8)
    y = ops.qdq(x, spec)
    xi, _ = ops.quant_int(x, spec)
    assert torch.equal(xi, torch.round(xi)) and xi.min().item() >= 0 and xi.max().item() <= 255
    assert xi.min().item() == 0 and xi.max().item() == 255          # range endpoints are reached
    assert torch.equal(ops.qdq(y, spec), y)                          # idempotent
    assert (y - x).abs().max().item() <= 0.5001 * d.item()           # inside the range: <= step/2
    # same tensor through the per-embedding kernel with constant per-dim parameters == per-tensor
    C = shape[-1]
    dv, zv = d.expand(C).contiguous(), z.expand(C).contiguous()
    y_axis = ops.qdq(x, ops.spec(dv, zv, None, 8), x.numel() // C, C, 1)
    assert torch.equal(y_axis, y)


# ---- 4. C-ABI argument errors -----------------------------------------------------------------------
def test_abi_errors(ops):
    lib = ops.lib
    spec = tq_native.QSpec(None, None, None, 8, 0, 1e-8)
    x = torch.zeros(16, device=DEV)
    assert lib.tq_qdq_f32(x.data_ptr(), x.data_ptr(), 16, spec, None) == -1          # no delta
    d = torch.ones(1, device=DEV)
    bad = tq_native.QSpec(d.data_ptr(), d.data_ptr(), None, 17, 0, 1e-8)
    assert lib.tq_qdq_f32(x.data_ptr(), x.data_ptr(), 16, bad, None) == -1           # n_bits
    ok = tq_native.QSpec(d.data_ptr(), d.data_ptr(), None, 8, 0, 1e-8)
    assert lib.tq_qdq_f32(None, x.data_ptr(), 16, ok, None) == -1                     # null x
    assert lib.tq_qdq_f32(x.data_ptr(), x.data_ptr(), 0, ok, None) == 0               # empty is fine
    assert lib.tq_minmax_f32(x.data_ptr(), 16, x.data_ptr(), x.data_ptr(), 4, None) == -3   # workspace
    torch.cuda.synchronize()


def test_uninitialised_quantizer_raises():
    from quantization.quantizers import QMethods, QuantizerNotInitializedError
    q = QMethods.asymmetric_uniform.cls(n_bits=8)
    with pytest.raises(QuantizerNotInitializedError):
        q(torch.zeros(4, device=DEV))
    with pytest.raises(tq_native.TQError):
        q.set_quant_range(-1.0, 1.0)
        q(torch.zeros(4))               # CPU tensor: rejected, no fallback


def test_percentile_on_device_matches_numpy():
    """the percentile ranges are computed on the GPU (torch.sort + numpy's interpolation formula)"""
    from quantization.range_estimators import CurrentMinMaxEstimator
    rs = np.random.RandomState(3)
    for n, q in [(98304, 0.01), (3145728, 0.1), (1000, 50.0), (777, 99.99)]:
        x = (rs.randn(2, n) * 3).astype(np.float32)
        ref = np.percentile(x, (q, 100 - q), axis=-1)
        got = CurrentMinMaxEstimator._percentiles(torch.from_numpy(x).to('cuda'), (q, 100 - q))
        for r, g in zip(ref, got):
            assert g.is_cuda
            assert np.array_equal(torch.Tensor(r).numpy(), g.cpu().numpy())


def test_copy_probe_copies(ops):
    """tq_probe_copy_f32 (bandwidth probe with the library's streaming access pattern): every flag combination
    copies exactly; argument errors are reported"""
    x = torch.randn(1 << 20, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    for flags in range(8):
        y = torch.zeros_like(x)
        assert ops.lib.tq_probe_copy_f32(x.data_ptr(), y.data_ptr(), x.numel(), flags, st) == 0
        assert torch.equal(x, y)
    y = torch.zeros_like(x)
    assert ops.lib.tq_probe_copy_f32(x.data_ptr(), y.data_ptr(), 1002, 0, st) == -1        # n % 4 != 0
    assert ops.lib.tq_probe_copy_f32(None, y.data_ptr(), 1000, 0, st) == -1
