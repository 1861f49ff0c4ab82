"""MobileBERT on the int8 pipeline (BASELINE config 4): tq_linear_nonorm_qdq_i8 (dense -> QDQ [-> + residual -> QDQ] ->
NoNorm -> QDQ in one GEMM epilogue), ReLU / output-stride forms of tq_linear_seg_qdq_i8, tq_attention_pad_qdq_i8 (32-wide
heads in zero-padded slots) and engine/fused_mobilebert.py against the module path."""
import math

import numpy as np
import pytest
import torch

import tq_native
from oracle import fakequant_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def T_(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def asym(ops, lo, hi):
    d, z = O.asym_set_quant_range(lo, hi, 8)
    dt, zt = T_(np.atleast_1d(d)), T_(np.atleast_1d(z))
    return ops.spec(dt, zt, None, 8), (dt, zt), (float(O.scale_of(d)), float(O.asym_zero_point(z, 8)))


@pytest.mark.parametrize('M,N,K,res', [(384, 128, 512, False), (8192, 128, 512, True), (300, 512, 128, True), (256, 128, 256, True)])
def test_linear_nonorm_i8_vs_oracle_chain(M, N, K, res):
    """bit-exact against the oracle's operation chain on the exact integer accumulators"""
    ops = tq_native.ops()
    rs = np.random.RandomState(M + N + K)
    a_int = rs.randint(0, 256, size=(M, K)).astype(np.float32)
    w_int = rs.randint(-8, 8, size=(N, K)).astype(np.float32)                 # 4-bit weight grid
    r_int = rs.randint(0, 256, size=(M, N)).astype(np.float32)
    bias = (rs.randn(N) * 0.3).astype(np.float32)
    gamma = (1 + 0.1 * rs.randn(N)).astype(np.float32)
    beta = (0.5 * rs.randn(N)).astype(np.float32)
    a_sp, k1, (sa, za) = asym(ops, -2.0, 3.0)
    w_d, w_signed = O.sym_set_quant_range(-0.3, 0.28, 4)
    wd_t, ws_t = T_(np.atleast_1d(w_d)), torch.tensor(bool(w_signed), device=DEV)
    w_sp = ops.spec(wd_t, None, ws_t, 4)
    acc = (a_int - za).astype(np.float64) @ w_int.astype(np.float64).T
    cs = np.float32(np.float32(sa) * np.float32(O.scale_of(w_d, scale_domain='linear')))
    pre = (acc * np.float64(cs) + bias.astype(np.float64)).astype(np.float32)
    g_sp, k2, _ = asym(ops, float(pre.min()), float(pre.max()))
    g_d, g_z = k2[0].cpu().numpy(), k2[1].cpu().numpy()
    x = O.qdq_asym(pre, g_d, g_z, 8)
    r_sp = u_sp = None
    keep = []
    if res:
        r_sp, k3, (sr, zr) = asym(ops, -4.0, 4.0)
        resv = (np.float32(sr) * (r_int - np.float32(zr))).astype(np.float32)
        tot = (x + resv).astype(np.float32)
        u_sp, k4, _ = asym(ops, float(tot.min()) * 0.9, float(tot.max()) * 0.9)
        x = O.qdq_asym(tot, k4[0].cpu().numpy(), k4[1].cpu().numpy(), 8)
        keep += [k3, k4]
    y = ((x * gamma).astype(np.float32) + beta).astype(np.float32)
    n_sp, k5, _ = asym(ops, float(y.min()), float(y.max()))
    ref = O.qdq_asym(y, k5[0].cpu().numpy(), k5[1].cpu().numpy(), 8, return_int=True)
    w8 = T_(w_int).to(torch.int8)
    z8 = torch.empty(M, N, dtype=torch.uint8, device=DEV)
    ops.linear_nonorm_i8(T_(a_int).to(torch.uint8), w8, w8.to(torch.int32).sum(dim=1, dtype=torch.int32).contiguous(), T_(bias), M, N, K,
                         a_sp, w_sp, g_sp, T_(r_int).to(torch.uint8) if res else None, r_sp, u_sp, T_(gamma), T_(beta), n_sp, z8)
    torch.cuda.synchronize()
    assert np.array_equal(z8.cpu().numpy().astype(np.float32), ref)


def _model(layers=1, n_bits=4):
    from engine.mobilebert import MobileBertConfig, QuantMobileBertForSequenceClassification
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators
    cfg = MobileBertConfig(vocab_size=1000, hidden_size=256, num_hidden_layers=layers, num_attention_heads=4,
                           intermediate_size=256, embedding_size=64, intra_bottleneck_size=128, num_feedforward_networks=2,
                           max_position_embeddings=128)
    m = QuantMobileBertForSequenceClassification(cfg, method=QMethods.symmetric_uniform, act_method=QMethods.asymmetric_uniform,
                                                 n_bits=n_bits, n_bits_act=8, act_range_method=RangeEstimators.running_minmax)
    m.init_weights(seed=2, std=0.05)
    m.to(DEV).eval()
    m.set_quant_state(True, True)
    return m


@pytest.mark.parametrize('layers,n_bits', [(1, 4), (1, 8), (2, 4)])
def test_mobilebert_engine_vs_module_path(layers, n_bits):
    from engine.fused_mobilebert import FusedMobileBertEngine
    model = _model(layers, n_bits)
    ids = torch.randint(0, 1000, (4, 128), generator=torch.Generator().manual_seed(5)).to(DEV)
    mask = torch.ones_like(ids)
    with torch.no_grad():
        model(ids, mask)
        model.fix_ranges()
        ref_logits = model(ids, mask)
        ref_hidden = model.encode(ids, mask)
        eng = FusedMobileBertEngine(model, 4, 128)
        logits = eng(ids, mask)
        hidden = eng.hidden_states()
    torch.cuda.synchronize()
    hstep = float(model.layers[-1].out_bottleneck.norm.activation_quantizer.quantizer.scale.reshape(-1)[0])
    dh = (hidden - ref_hidden).abs()
    share = float((dh > 0.5 * hstep).float().mean())
    assert float(dh.max()) <= 6 * hstep and share < (0.05 if layers == 1 else 0.25), (float(dh.max()) / hstep, share)
    step = float(model.classifier.activation_quantizer.quantizer.scale.reshape(-1)[0])
    spread = float(ref_logits.max() - ref_logits.min())
    assert float((logits - ref_logits).abs().max()) <= max(3 * step, 0.1 * spread if layers > 1 else 0.0) + 1e-6
