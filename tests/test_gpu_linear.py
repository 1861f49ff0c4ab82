"""tcgen05 fused linear (tq_linear_qdq_bf16) on the GPU.

Integer-grid operands make the GEMM itself EXACT (every partial sum is an integer < 2^24), so the
kernel is compared bit-for-bit against an fp64 integer matmul followed by the same fp32 epilogue
(scale, bias, QDQ) computed with the CPU oracle.  Activation functions (erff / tanhf vs libm) get
a 1-step tolerance on the quantized output.
"""
import numpy as np
import pytest
import torch
from torch import nn

import tq_native
import parity_cases as P
from oracle import fakequant_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _grids(M, N, K, seed, w_bits=8):
    rs = np.random.RandomState(seed)
    a = rs.randint(-255, 256, size=(M, K)).astype(np.float32)
    half = 2 ** (w_bits - 1)
    w = rs.randint(-half, half, size=(N, K)).astype(np.float32)
    return a, w


def _dev_scalar(v):
    return torch.tensor([v], dtype=torch.float32, device=DEV)


@pytest.mark.parametrize('M,N,K', [(128, 64, 64), (128, 128, 128), (256, 192, 768), (4096, 768, 768),
                                   (4096, 3072, 768), (4096, 768, 3072), (32, 768, 768), (100, 72, 128),
                                   (8192, 128, 512), (512, 2304, 768)])
def test_gemm_exact_no_epilogue(M, N, K):
    ops = tq_native.ops()
    a, w = _grids(M, N, K, seed=M + N + K)
    at = torch.from_numpy(a).to(DEV).to(torch.bfloat16)
    wt = torch.from_numpy(w).to(DEV).to(torch.bfloat16)
    y, _ = ops.linear(at, wt, None, M, N, K, 1, None, None, 1, 0, None, 1)
    ref = torch.from_numpy(a).double().to(DEV) @ torch.from_numpy(w).double().to(DEV).T
    assert ref.abs().max().item() < 2 ** 24
    torch.cuda.synchronize()
    assert torch.equal(y.double(), ref), f'max |diff| = {(y.double() - ref).abs().max().item()}'


@pytest.mark.parametrize('ctas', [1, 2])
@pytest.mark.parametrize('bn', [256, 192, 128])
@pytest.mark.parametrize('M,N,K', [(512, 768, 256), (300, 264, 128), (4096, 768, 768), (1000, 2304, 192)])
def test_gemm_exact_forced_tile_shapes(monkeypatch, ctas, bn, M, N, K):
    """every tile width, single CTA and CTA pair (tcgen05 cta_group::2), ragged M / N edges"""
    monkeypatch.setenv('TQ_LINEAR_BN', str(bn))
    monkeypatch.setenv('TQ_LINEAR_CTAS', str(ctas))
    ops = tq_native.ops()
    a, w = _grids(M, N, K, seed=7 * M + N + K + bn + ctas)
    at = torch.from_numpy(a).to(DEV).to(torch.bfloat16)
    wt = torch.from_numpy(w).to(DEV).to(torch.bfloat16)
    y, _ = ops.linear(at, wt, None, M, N, K, 1, None, None, 1, 0, None, 1)
    ref = torch.from_numpy(a).double().to(DEV) @ torch.from_numpy(w).double().to(DEV).T
    torch.cuda.synchronize()
    assert torch.equal(y.double(), ref), f'max |diff| = {(y.double() - ref).abs().max().item()}'


@pytest.mark.parametrize('act', [0, 2])
@pytest.mark.parametrize('per_col', [False, True])
def test_gemm_epilogue_exact(act, per_col):
    """scale, bias, ReLU and the output quantizer (per-tensor and per-column) -- bit-exact."""
    ops = tq_native.ops()
    M, N, K = 384, 192, 256
    a, w = _grids(M, N, K, seed=3)
    rs = np.random.RandomState(4)
    bias = (rs.randn(N) * 0.5).astype(np.float32)
    a_delta, a_zf = O.asym_set_quant_range(-3.1, 2.7, 8)
    w_delta, w_signed = O.sym_set_quant_range(-0.11, 0.09, 8)
    a_scale, w_scale = O.scale_of(a_delta), O.scale_of(w_delta)
    acc = a.astype(np.float64) @ w.astype(np.float64).T                 # exact integers
    # kernel contract: fma(acc, s_a * s_w, bias) -- one rounding (product exact in fp64, then one add)
    pre = (acc * np.float64(np.float32(a_scale * w_scale)) + bias.astype(np.float64)).astype(np.float32)
    if act == 2:
        pre = np.maximum(pre, 0)
    if per_col:
        o_min = pre.min(0) * 0.8
        o_max = pre.max(0) * 0.8
    else:
        o_min, o_max = pre.min() * 0.8, pre.max() * 0.8
    o_delta, o_zf = O.asym_set_quant_range(o_min, o_max, 8)
    ref_int = O.qdq_asym(pre, o_delta, o_zf, 8, axis=1 if per_col else None, return_int=True)
    ref = O.qdq_asym(pre, o_delta, o_zf, 8, axis=1 if per_col else None)

    t = lambda v: torch.from_numpy(np.atleast_1d(np.asarray(v, np.float32))).to(DEV)
    a_d, a_z, w_d, o_d, o_z = t(a_delta), t(a_zf), t(w_delta), t(o_delta), t(o_zf)
    w_s = torch.tensor(bool(w_signed), device=DEV)
    a_spec = ops.spec(a_d, a_z, None, 8)
    w_spec = ops.spec(w_d, None, w_s, 8)
    o_spec = ops.spec(o_d, o_z, None, 8)
    at = torch.from_numpy(a).to(DEV).to(torch.bfloat16)
    wt = torch.from_numpy(w).to(DEV).to(torch.bfloat16)
    y, yc = ops.linear(at, wt, t(bias), M, N, K, 1, a_spec, w_spec, 1, act, o_spec, N if per_col else 1,
                       want_f32=True, want_ctr=True)
    torch.cuda.synchronize()
    P.assert_same(y, ref, 'fused linear + QDQ')
    zp = O.asym_zero_point(o_zf, 8)
    P.assert_same(yc.float(), ref_int - zp.reshape(1, -1), 'centred bf16 output grid')
    # no output quantizer: plain scale + bias (+ReLU)
    y2, _ = ops.linear(at, wt, t(bias), M, N, K, 1, a_spec, w_spec, 1, act, None, 1)
    P.assert_same(y2, pre, 'fused linear without output quantizer')


def test_gemm_split3_fp32_accuracy():
    """arbitrary fp32 activations through the hi|mid|lo split: fp32-GEMM-level accuracy."""
    ops = tq_native.ops()
    M, N, K = 512, 256, 768
    rs = np.random.RandomState(9)
    x = (rs.randn(M, K) * 2).astype(np.float32)
    _, w = _grids(M, N, K, seed=10)
    xt = torch.from_numpy(x).to(DEV)
    a3 = ops.split3(xt)
    rec = a3.float().view(M, 3, K).sum(1)
    assert (rec - xt).abs().max().item() <= 2e-7 * np.abs(x).max()
    wt = torch.from_numpy(w).to(DEV).to(torch.bfloat16)
    y, _ = ops.linear(a3, wt, None, M, N, K, 3, None, None, 1, 0, None, 1)
    ref = xt.double() @ torch.from_numpy(w).double().to(DEV).T
    err = (y.double() - ref).abs().max().item()
    # tensor-core fp32 accumulation over 3K products: ~5e-6 of the output range (fp32 SGEMM: ~1e-6)
    assert err <= 1e-5 * ref.abs().max().item(), err


def test_unsupported_shapes_are_rejected():
    ops = tq_native.ops()
    a = torch.zeros(128, 96, dtype=torch.bfloat16, device=DEV)       # K % 64 != 0
    w = torch.zeros(64, 96, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(tq_native.TQError):
        ops.linear(a, w, None, 128, 64, 96, 1, None, None, 1, 0, None, 1)


@pytest.mark.parametrize('act', [None, nn.GELU, nn.ReLU, nn.Tanh])
@torch.no_grad()                       # the fused kernel is the inference path: it declines while autograd records
def test_quant_linear_fused_vs_unfused(act):
    """QuantLinear through the module API: fused tcgen05 path vs the three-step path (library fp32
    GEMM + activation + QDQ kernel).  <= 1 output step, < 0.2 % of elements differ."""
    from quantization import fused_linear
    from quantization.autoquant_utils import QuantLinear
    from quantization.quantizers import QMethods
    torch.manual_seed(0)
    lin = QuantLinear(768, 3072, bias=True, activation=act() if act else None,
                      method=QMethods.symmetric_uniform, act_method=QMethods.asymmetric_uniform,
                      n_bits=8, n_bits_act=8).to(DEV)
    lin.weight.data.normal_(0, 0.02)
    lin.bias.data.normal_(0, 0.02)
    lin.quantized()
    lin.eval()
    q_in = QMethods.asymmetric_uniform.cls(n_bits=8)
    x_raw = torch.randn(8, 128, 768, device=DEV)
    q_in.set_quant_range(float(x_raw.min()), float(x_raw.max()))
    x = q_in(x_raw)                              # tagged: on an 8-bit grid
    for state in ('estimate', 'fixed'):
        if state == 'fixed':
            lin.fix_ranges()
        fused_linear.ENABLED = True
        y_f = lin(x)
        fused_linear.ENABLED = False
        try:
            y_u = lin(x)
        finally:
            fused_linear.ENABLED = True
        step = float(lin.activation_quantizer.quantizer.scale)
        d = (y_f - y_u).abs()
        assert d.max().item() <= step * 1.001
        assert (d > step * 1e-3).float().mean().item() < 2e-3
    # untagged fp32 input -> split path
    y_s = lin(x_raw)
    fused_linear.ENABLED = False
    try:
        y_su = lin(x_raw)
    finally:
        fused_linear.ENABLED = True
    d = (y_s - y_su).abs()
    assert d.max().item() <= step * 1.001 and (d > step * 1e-3).float().mean().item() < 2e-3


def test_fused_linear_declines_while_autograd_records():
    """eval-mode forward with grad enabled (sensitivity / Fisher passes, eval-mode fine-tuning): the output must
    carry a grad_fn that reaches the weight -- the fused kernel would cut the graph silently."""
    from quantization.autoquant_utils import QuantLinear
    from quantization.quantizers import QMethods
    torch.manual_seed(0)
    lin = QuantLinear(128, 64, bias=True, method=QMethods.symmetric_uniform, act_method=QMethods.asymmetric_uniform,
                      n_bits=8, n_bits_act=8).to(DEV)
    lin.quantized()
    lin.eval()
    x = torch.randn(4, 16, 128, device=DEV)
    with torch.no_grad():
        lin(x)
    lin.fix_ranges()
    lin.cached_params = None
    y = lin(x)
    assert y.grad_fn is not None
    y.sum().backward()
    assert lin.weight.grad is not None and lin.weight.grad.abs().sum().item() > 0
    with torch.no_grad():
        assert lin(x).grad_fn is None


@pytest.mark.parametrize('est_name', ['running_minmax', 'current_minmax', 'allminmax'])
@pytest.mark.parametrize('sym', [False, True])
@torch.no_grad()
def test_calibration_time_fused_gemm(est_name, sym):
    """QuantLinear in estimate_ranges state: min/max out of the GEMM epilogue + tq_calib_finalize_f32 (one launch) must
    give exactly the ranges and outputs of the GEMM -> min/max kernel -> range update -> set_quant_range -> QDQ chain,
    over several calibration batches (EMA / all-time rules included)."""
    from quantization import fused_linear
    from quantization.autoquant_utils import QuantLinear
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators
    act_m = QMethods.symmetric_uniform if sym else QMethods.asymmetric_uniform

    def make():
        torch.manual_seed(0)
        lin = QuantLinear(256, 384, bias=True, method=QMethods.symmetric_uniform, act_method=act_m, n_bits=8, n_bits_act=8,
                          act_range_method=RangeEstimators[est_name]).to(DEV)
        lin.weight.data.normal_(0, 0.05)
        lin.bias.data.normal_(0, 0.05)
        lin.quantized()
        lin.eval()
        return lin

    q_in = QMethods.asymmetric_uniform.cls(n_bits=8)
    g = torch.Generator().manual_seed(3)
    batches = [(torch.randn(4, 64, 256, generator=g) * (1 + i)).to(DEV) for i in range(3)]
    q_in.set_quant_range(-8.0, 8.0)
    a, b = make(), make()
    launches = {}
    for flag, lin in ((True, a), (False, b)):
        fused_linear.CALIBRATION_FUSION = flag
        l0 = tq_native.ops().launches
        try:
            outs = [lin(q_in(x)) for x in batches]
        finally:
            fused_linear.CALIBRATION_FUSION = True
        launches[flag] = tq_native.ops().launches - l0
        lin._outs = outs
    torch.cuda.synchronize()
    qa, qb = a.activation_quantizer.quantizer, b.activation_quantizer.quantizer
    assert torch.equal(qa._delta.reshape(-1), qb._delta.reshape(-1))
    if not sym:
        assert torch.equal(qa._zero_float.reshape(-1), qb._zero_float.reshape(-1))
    else:
        assert bool(qa._signed) == bool(qb._signed)
    ea, eb = a.activation_quantizer.range_estimator, b.activation_quantizer.range_estimator
    assert torch.equal(ea.current_xmin.reshape(-1), eb.current_xmin.reshape(-1))
    assert torch.equal(ea.current_xmax.reshape(-1), eb.current_xmax.reshape(-1))
    for ya, yb in zip(a._outs, b._outs):
        assert torch.equal(ya, yb)
    assert launches[True] < launches[False]          # two launches and one pass over the tensor fewer per batch
