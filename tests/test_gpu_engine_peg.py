"""Per-embedding-group (PEG) activations on the int8 pipeline (BASELINE config 3): tq_linear_peg_qdq_i8,
tq_linear_peg_res_ln_qdq_i8, tq_attention_peg_qdq_i8 and engine/fused_peg.py."""
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


@pytest.mark.parametrize('M,N,K,G,act', [(384, 256, 256, 2, 0), (4096, 2304, 768, 6, 0), (300, 512, 768, 3, 1), (4096, 768, 768, 1, 0)])
def test_linear_peg_i8_vs_numpy(M, N, K, G, act):
    """group-by-group integer accumulation + fp32 combination vs the float64 formula of the reference
    (x_q @ W_q.T + b with x_q = s_g (a - zp_g), W_q = s_w w): integers equal except at rounding boundaries"""
    ops = tq_native.ops()
    rs = np.random.RandomState(M + N + K + G)
    a_int = rs.randint(0, 256, size=(M, K)).astype(np.float32)
    w_int = rs.randint(-128, 128, size=(N, K)).astype(np.float32)
    bias = (rs.randn(N) * 0.3).astype(np.float32)
    gk = K // G
    a_d = (0.02 * (1 + 0.3 * rs.rand(G))).astype(np.float32)
    a_zf = (100 + 40 * rs.rand(G)).astype(np.float32)
    a_zp = np.clip(np.round(a_zf), 0, 255)
    w_d, w_signed = O.sym_set_quant_range(-0.08 * 8 / math.sqrt(K), 0.09 * 8 / math.sqrt(K), 8)
    sw = float(O.scale_of(w_d))
    xq = (a_int - np.repeat(a_zp, gk)[None, :]) * np.repeat(a_d, gk)[None, :].astype(np.float64)
    pre = xq @ (w_int.astype(np.float64) * sw).T + bias.astype(np.float64)
    if act == 1:
        pre = 0.5 * pre * (1.0 + np.vectorize(math.erf)(pre / math.sqrt(2.0)))
    seg = 128
    nseg = N // seg
    o_d = np.zeros(nseg, np.float32)
    o_z = np.zeros(nseg, np.float32)
    for j in range(nseg):
        blk = pre[:, j * seg:(j + 1) * seg]
        d, z = O.asym_set_quant_range(float(blk.min()), float(blk.max()), 8)
        o_d[j], o_z[j] = d, z
    ref = np.empty((M, N), np.float64)
    for j in range(nseg):
        zp = float(O.asym_zero_point(o_z[j], 8))
        ref[:, j * seg:(j + 1) * seg] = np.clip(np.rint(pre[:, j * seg:(j + 1) * seg] / float(O.scale_of(o_d[j]))) + zp, 0, 255)
    ad_t, az_t, od_t, oz_t = T_(a_d), T_(a_zf), T_(o_d), T_(o_z)
    wd_t, ws_t = T_(np.atleast_1d(w_d)), torch.tensor(bool(w_signed), device=DEV)
    a_sp, o_sp, w_sp = ops.spec(ad_t, az_t, None, 8), ops.spec(od_t, oz_t, None, 8), ops.spec(wd_t, None, ws_t, 8)
    w8 = T_(w_int).to(torch.int8)
    grs = T_(w_int).to(torch.int32).view(N, G, gk).sum(dim=2, dtype=torch.int32).t().contiguous()
    y8 = torch.empty(M, N, dtype=torch.uint8, device=DEV)
    ops.linear_peg_i8(T_(a_int).to(torch.uint8), w8, grs, T_(bias), M, N, K, a_sp, G, w_sp, 1, o_sp, nseg, seg, act, out_i8=y8)
    torch.cuda.synchronize()
    d = np.abs(y8.cpu().numpy().astype(np.float64) - ref)
    assert d.max() <= 1.0 and (d > 0).mean() < 2e-3, (d.max(), (d > 0).mean())
    yc = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.linear_peg_i8(T_(a_int).to(torch.uint8), w8, grs, T_(bias), M, N, K, a_sp, G, w_sp, 1, o_sp, nseg, seg, act, out_ctr=yc)
    torch.cuda.synchronize()
    zp_cols = np.repeat(np.array([float(O.asym_zero_point(o_z[j], 8)) for j in range(nseg)]), seg)[None, :]
    assert np.array_equal(yc.float().cpu().numpy() + zp_cols, y8.cpu().numpy().astype(np.float32))


def _peg_model(G, seed=0):
    from engine.bert import BertConfig, QuantBertForSequenceClassification
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators
    cfg = BertConfig(vocab_size=2000, hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512,
                     max_position_embeddings=128)
    m = QuantBertForSequenceClassification(cfg, method=QMethods.symmetric_uniform, act_method=QMethods.asymmetric_uniform,
                                           n_bits=8, n_bits_act=8, act_range_method=RangeEstimators.current_minmax)
    m.init_weights(seed=seed, std=0.05)
    g = torch.Generator().manual_seed(1)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Linear):
            mod.bias.data = torch.randn(mod.bias.shape, generator=g) * 0.05
        elif isinstance(mod, torch.nn.LayerNorm):
            mod.weight.data = 1 + torch.randn(mod.weight.shape, generator=g) * 0.1
            mod.bias.data = torch.randn(mod.bias.shape, generator=g) * 0.05
    m.to(DEV).eval()
    m.set_quant_state(True, True)
    m.set_per_embedding_groups(G, permute=False)
    return m


@pytest.mark.parametrize('G', [2, 1])
def test_peg_engine_vs_module_path(G):
    from engine.fused_peg import FusedBertPegEngine
    model = _peg_model(G)
    ids = torch.randint(0, 2000, (4, 128), generator=torch.Generator().manual_seed(5)).to(DEV)
    mask = torch.ones_like(ids)
    with torch.no_grad():
        model(ids, mask)
        model.fix_ranges()
        ref_logits = model(ids, mask)
        ref_hidden = model.encode(ids, mask)
        eng = FusedBertPegEngine(model, 4, 128)
        logits = eng(ids, mask)
        hidden = eng.hidden_states()
    torch.cuda.synchronize()
    z = model.layers[-1].z.activation_quantizer.quantizer
    hstep = float(z.delta.max())
    dh = (hidden - ref_hidden).abs()
    assert float(dh.max()) <= 6 * hstep and float((dh > 0.5 * hstep).float().mean()) < 0.05
    step = float(model.classifier.activation_quantizer.quantizer.scale.reshape(-1)[0])
    assert float((logits - ref_logits).abs().max()) <= 3 * step + 1e-6


def test_peg_engine_rejects_permuted_groups():
    from engine import configs
    from engine.fused import UnsupportedByEngine
    from engine.fused_peg import FusedBertPegEngine
    model, recipe = configs.build('bert_w8a8_pegp', torch.device(DEV), tiny=False) if False else (None, None)
    model = _peg_model(2)
    model.set_per_embedding_groups(2, permute=True)
    ids = torch.randint(0, 2000, (4, 128), generator=torch.Generator().manual_seed(5)).to(DEV)
    mask = torch.ones_like(ids)
    with torch.no_grad():
        model.full_precision()
        model.set_quant_state(weight_quant=False, act_quant=True)
        model(ids, mask)
        model.set_quant_state(weight_quant=True, act_quant=True)
        from quantization.range_estimators import RangeEstimatorBase
        for m in model.modules():
            if isinstance(m, RangeEstimatorBase):
                m.per_group_range_estimation = False
        model(ids, mask)
        model.fix_ranges()
    with pytest.raises(UnsupportedByEngine):
        FusedBertPegEngine(model, 4, 128)
