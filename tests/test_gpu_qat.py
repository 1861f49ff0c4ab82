"""GPU parity tests of the training-time path: every call goes
through the C ABI of libtq_b200.so on cuda:0.

1. golden vectors of the reference under torch autograd (tests/golden/qat.npz) through the
   reference-facing quantizer classes: forward and grad_x bit-exact, reduced range gradients within
   2e-6 * sum|terms| (fp32 summation order; see qat_cases.close_sum);
2. tq_qdq_bwd_f32 vs the CPU oracle on seeded random tensors in every kernel variant (per-tensor
   vectorised / ragged / misaligned, per-embedding columns, per-channel rows, generic strides);
3. size-independent properties at BASELINE sizes: determinism (bitwise equal re-runs), linearity of
   all three gradients in grad_y, grad_x == grad_y inside the range / 0 outside, NULL outputs;
4. AdaRound kernels vs goldens and the oracle.
"""
import numpy as np
import pytest
import torch

import tq_native
from oracle import fakequant_oracle as O
from qat_cases import (QAT_MANIFEST, check_adaround_case, check_adaround_layer_case, check_backward_case,
                       check_training_step, close_sum)

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def ops():
    return tq_native.ops()


# ---- 1. goldens ---------------------------------------------------------------------------------
@pytest.mark.parametrize('case', QAT_MANIFEST['backward'], ids=lambda c: c['name'])
def test_backward_golden(case):
    check_backward_case(case, DEV)


@pytest.mark.parametrize('case', QAT_MANIFEST['adaround'], ids=lambda c: c['name'])
def test_adaround_golden(case):
    check_adaround_case(case, DEV)


# ---- 2. kernel vs oracle --------------------------------------------------------------------------
def _params(rs, C, n_bits, asym, dev=DEV):
    delta = (0.01 + rs.rand(C) * 0.03).astype(np.float32)
    zf = (rs.rand(C) * (2 ** n_bits - 1)).astype(np.float32) if asym else None
    return delta, zf


def _run(ops, x, g, delta, zf, signed, n_bits, layout, log=False, **want):
    outer, C, inner = layout
    d = torch.tensor(delta, device=DEV)
    z = torch.tensor(zf, device=DEV) if zf is not None else None
    s = torch.tensor(bool(signed), device=DEV) if zf is None else None
    spec = ops.spec(d, z, s, n_bits, log, 1e-8)
    out = ops.qdq_bwd(x, g, spec, C, outer, C, inner, **want)
    torch.cuda.synchronize()
    return out


def _check_vs_oracle(ops, shape, layout, n_bits, asym, signed=True, misalign=0, seed=0):
    rs = np.random.RandomState(seed)
    C = layout[1]
    n = int(np.prod(shape))
    xh = (rs.randn(n) * 3).astype(np.float32)
    gh = rs.randn(n).astype(np.float32)
    delta, zf = _params(rs, C, n_bits, asym)
    if misalign:                      # 4-byte aligned only: scalar code path
        bx = torch.empty(n + misalign, device=DEV)
        bg = torch.empty(n + misalign, device=DEV)
        x, g = bx[misalign:], bg[misalign:]
        x.copy_(torch.from_numpy(xh))
        g.copy_(torch.from_numpy(gh))
    else:
        x, g = torch.from_numpy(xh).to(DEV), torch.from_numpy(gh).to(DEV)
    gx, gd, gz = _run(ops, x, g, delta, zf, signed, n_bits, layout)
    egx, egd, egz, (mag_s, mag_z) = O.qdq_backward(xh, gh, delta, zf, signed, n_bits, layout=layout)
    assert np.array_equal(gx.cpu().numpy().reshape(-1), egx.reshape(-1)), 'grad_x differs from the oracle'
    close_sum(gd.cpu().numpy(), egd, mag_s, 'grad_delta')
    if asym:
        close_sum(gz.cpu().numpy(), egz, mag_z, 'grad_zero_float')
    else:
        assert gz is None


@pytest.mark.parametrize('n_bits', [8, 4])
@pytest.mark.parametrize('asym', [True, False])
@pytest.mark.parametrize('n', [32 * 128 * 768, 1000003, 5, 1, 4096 + 3, 20 * 1024 * 1024 + 7])   # last: one-chunk-per-CTA kernel
def test_bwd_tensor_vs_oracle(ops, n, asym, n_bits):
    _check_vs_oracle(ops, (n,), (1, 1, n), n_bits, asym, seed=n % 97)


@pytest.mark.parametrize('misalign', [1, 3])
def test_bwd_tensor_misaligned(ops, misalign):
    _check_vs_oracle(ops, (100003,), (1, 1, 100003), 8, True, misalign=misalign, seed=5)


@pytest.mark.parametrize('rows,C', [(4096, 768), (1001, 768), (7, 768), (333, 128), (64, 3072), (9, 4), (4096, 260)])
@pytest.mark.parametrize('asym', [True, False])
def test_bwd_cols_vs_oracle(ops, rows, C, asym):
    """per-embedding / PEG layout [rows, C], parameters per column"""
    _check_vs_oracle(ops, (rows, C), (rows, C, 1), 8, asym, seed=rows + C)


@pytest.mark.parametrize('outer,C,inner', [(1, 768, 3072), (1, 48, 64), (1, 5, 21), (3, 6, 20), (2, 6, 5), (1, 3072, 768),
                                           (40, 30, 1), (1, 2, 4), (1, 3000, 772), (1, 297, 8), (1, 5000, 4)])
@pytest.mark.parametrize('asym', [True, False])
def test_bwd_rows_vs_oracle(ops, outer, C, inner, asym):
    """per-channel weights [C, inner] and generic [outer, C, inner] views"""
    _check_vs_oracle(ops, (outer, C, inner), (outer, C, inner), 4, asym, seed=C + inner)


def test_bwd_cols_misaligned_falls_back(ops):
    _check_vs_oracle(ops, (50, 24), (50, 24, 1), 8, True, misalign=1, seed=11)


def test_bwd_unsigned_symmetric(ops):
    _check_vs_oracle(ops, (70001,), (1, 1, 70001), 8, False, signed=False, seed=3)


def test_bwd_log_domain(ops):
    rs = np.random.RandomState(9)
    n = 50000
    xh, gh = (rs.randn(n) * 3).astype(np.float32), rs.randn(n).astype(np.float32)
    delta, zf = np.log(np.float32([0.03])).astype(np.float32), np.float32([100.4])
    gx, gd, gz = _run(ops, torch.from_numpy(xh).to(DEV), torch.from_numpy(gh).to(DEV), delta, zf, None, 8, (1, 1, n),
                      log=True)
    egx, egd, egz, (mag_s, mag_z) = O.qdq_backward(xh, gh, delta, zf, None, 8, scale_domain='log', layout=(1, 1, n))
    np.testing.assert_allclose(gx.cpu().numpy(), egx, rtol=1e-6, atol=0)
    close_sum(gd.cpu().numpy(), egd, mag_s, 'grad_delta', rtol=4e-6)
    close_sum(gz.cpu().numpy(), egz, mag_z, 'grad_zero_float', rtol=4e-6)


# ---- 3. properties at full size --------------------------------------------------------------------
@pytest.mark.parametrize('layout', [(1, 1, 32 * 128 * 3072), (32 * 128, 768, 1), (1, 3072, 768), (1, 1, 64 * 1024 * 1024)])
def test_bwd_properties_full_size(ops, layout):
    outer, C, inner = layout
    n = outer * C * inner
    gen = torch.Generator(device=DEV).manual_seed(1234)
    x = torch.randn(n, device=DEV, generator=gen) * 3
    g = torch.randn(n, device=DEV, generator=gen)
    rs = np.random.RandomState(1)
    delta, zf = _params(rs, C, 8, True)
    a = _run(ops, x, g, delta, zf, None, 8, layout)
    b = _run(ops, x, g, delta, zf, None, 8, layout)
    for u, v in zip(a, b):
        assert torch.equal(u, v), 'tq_qdq_bwd_f32 is not deterministic'
    # linear in grad_y: 2 * g doubles every output exactly (power-of-two scaling commutes with rounding)
    c = _run(ops, x, g * 2, delta, zf, None, 8, layout)
    assert torch.equal(c[0], a[0] * 2)
    for u, v in zip(c[1:], a[1:]):
        assert torch.allclose(u, v * 2, rtol=1e-6, atol=0)
    # grad_x is grad_y (up to one rounding of g * s / s) inside the range and exactly 0 outside
    d = torch.tensor(delta, device=DEV)
    z = torch.tensor(zf, device=DEV)
    scale = d.clamp(min=1e-8).view(1, C, 1)
    zp = z.round().clamp(0, 255).view(1, C, 1)
    u = torch.round(x.view(outer, C, inner) / scale) + zp
    inside = ((u >= 0) & (u <= 255)).view(-1)
    gx = a[0].view(-1)
    assert (gx[~inside] == 0).all()
    assert torch.allclose(gx[inside], g[inside], rtol=3e-7, atol=0)
    # NULL outputs: only what was asked for is produced, values unchanged
    only_x = _run(ops, x, g, delta, zf, None, 8, layout, want_delta=False, want_zero_float=False)
    assert only_x[1] is None and only_x[2] is None and torch.equal(only_x[0], a[0])
    only_p = _run(ops, x, g, delta, zf, None, 8, layout, want_x=False)
    assert only_p[0] is None and torch.equal(only_p[1], a[1]) and torch.equal(only_p[2], a[2])


def test_bwd_workspace_shared_between_variants(ops):
    """one caller workspace serves every variant on a stream: the per-tensor kernel's partial sums must not
    be mistaken for the column kernel's tickets (and vice versa), in any call order"""
    for _ in range(2):
        _check_vs_oracle(ops, (200003,), (1, 1, 200003), 8, True, seed=21)
        _check_vs_oracle(ops, (64, 3072), (64, 3072, 1), 8, True, seed=22)
        _check_vs_oracle(ops, (4096, 768), (4096, 768, 1), 8, False, seed=23)
        _check_vs_oracle(ops, (1, 400, 64), (1, 400, 64), 4, True, seed=24)       # warp-per-row kernel
        _check_vs_oracle(ops, (32 * 128 * 768,), (1, 1, 32 * 128 * 768), 8, True, seed=25)


def test_bwd_empty_and_errors(ops):
    d = torch.tensor([0.1], device=DEV)
    z = torch.tensor([3.0], device=DEV)
    spec = ops.spec(d, z, None, 8, False, 1e-8)
    x = torch.empty(0, device=DEV)
    gx, gd, gz = ops.qdq_bwd(x, x, spec, 1)
    assert gx.numel() == 0 and gd.item() == 0 and gz.item() == 0
    lib = ops.lib
    bad = ops.spec(d, z, None, 0, False, 1e-8)
    ws = torch.zeros(1 << 16, dtype=torch.uint8, device=DEV)
    y = torch.ones(8, device=DEV)
    assert lib.tq_qdq_bwd_f32(y.data_ptr(), y.data_ptr(), None, None, None, 1, 1, 8, bad, ws.data_ptr(), ws.numel(), None) == -1
    assert lib.tq_qdq_bwd_f32(y.data_ptr(), y.data_ptr(), None, None, None, 1, 1, 8, spec, ws.data_ptr(), 8, None) == -3
    assert lib.tq_qdq_bwd_f32(None, y.data_ptr(), None, None, None, 1, 1, 8, spec, ws.data_ptr(), ws.numel(), None) == -1
    assert lib.tq_qdq_bwd_f32(y.data_ptr(), y.data_ptr(), None, None, None, -1, 1, 8, spec, ws.data_ptr(), ws.numel(), None) == -1


def test_qat_training_step_matches_torch_autograd():
    check_training_step(DEV)


@pytest.mark.parametrize('case', QAT_MANIFEST['adaround_layer'], ids=lambda c: c['name'])
def test_adaround_layer_loop(case):
    check_adaround_layer_case(case, DEV)
