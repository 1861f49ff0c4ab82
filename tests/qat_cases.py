"""Shared helpers for the training-path tests (golden file tests/golden/qat.npz)."""
import json
import os

import numpy as np

from oracle import fakequant_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
with open(os.path.join(GOLDEN, 'qat_manifest.json')) as f:
    QAT_MANIFEST = json.load(f)
_QAT = None


def qat_file():
    global _QAT
    if _QAT is None:
        _QAT = np.load(os.path.join(GOLDEN, 'qat.npz'))
    return _QAT


def backward_case_arrays(case):
    """-> dict(x, g, delta, zero_float | None, signed | None, golden grads) of one backward case"""
    g, nm = qat_file(), case['name']
    d = dict(x=g[f'{nm}.x'], g=g[f'{nm}.g'], delta=g[f'{nm}.delta'], y=g[f'{nm}.y'], grad_x=g[f'{nm}.grad_x'],
             grad_delta=g[f'{nm}.grad_delta'], zero_float=None, signed=None, grad_zero_float=None)
    if case['kind'] == 'asym':
        d['zero_float'] = g[f'{nm}.zero_float']
        d['grad_zero_float'] = g[f'{nm}.grad_zero_float']
    else:
        d['signed'] = bool(g[f'{nm}.signed'])
    return d


def oracle_backward(case, a):
    return O.qdq_backward(a['x'], a['g'], a['delta'], a['zero_float'], a['signed'], case['n_bits'],
                          scale_domain=case['scale_domain'], axis=case['axis'], per_channel=case['per_channel'])


def close_sum(got, want, mag, what, rtol=2e-6):
    """|got - want| <= rtol * sum|terms|: the fp32 summation-order error bound of a reduced gradient.
    (`within 1e-5 relative` of the north star, with the magnitude of the summed terms as the base --
    the sums themselves cancel to near zero.)"""
    got, want = np.asarray(got, np.float64).reshape(-1), np.asarray(want, np.float64).reshape(-1)
    tol = rtol * np.asarray(mag, np.float64).reshape(-1) + 1e-30
    bad = np.abs(got - want) > tol
    assert not bad.any(), f'{what}: {bad.sum()} / {got.size} outside tolerance; ' \
                          f'got {got[bad][:4]} want {want[bad][:4]} tol {tol[bad][:4]}'


def adaround_case_arrays(case):
    g, nm = qat_file(), case['name']
    keys = ('y_nearest', 'alpha0', 'y_soft0', 'alpha1', 'y_soft1', 'grad_alpha1', 'y_hard1', 'x_int_hard1', 'delta')
    d = {k: g[f'{nm}.{k}'] for k in keys}
    d['w'], d['g'] = g['ada.w'], g['ada.g']
    d['zero_float'] = g[f'{nm}.zero_float'] if case['kind'] == 'asym' else None
    d['signed'] = bool(g[f'{nm}.signed']) if case['kind'] == 'sym' else None
    return d


def adaround_grid(case, a):
    """(scale, zp, lo, hi) broadcastable against the weight"""
    scale = O.scale_of(a['delta'])
    if case['kind'] == 'asym':
        zp = O.asym_zero_point(a['zero_float'], case['n_bits'])
        lo, hi = 0.0, O.asym_int_max(case['n_bits'])
    else:
        zp = np.zeros_like(scale)
        lo, hi = O.sym_grid(case['n_bits'], a['signed'])
    if case['per_channel']:
        scale, zp = scale.reshape(-1, 1), zp.reshape(-1, 1)
    else:
        scale, zp = scale.reshape(()), zp.reshape(())
    return scale, zp, lo, hi


# ---- the same cases through the reference-facing API (CPU: oracle back-end, GPU: libtq_b200.so) ----
def check_backward_case(case, device):
    """quantizer(x).backward(g) through quantization.quantizers on `device` vs the reference's autograd"""
    import torch
    from quantization.quantizers import QMethods
    a = backward_case_arrays(case)
    cls = QMethods.symmetric_uniform.cls if case['kind'] == 'sym' else QMethods.asymmetric_uniform.cls
    q = cls(n_bits=case['n_bits'], scale_domain=case['scale_domain'], per_channel=case['per_channel'],
            axis=case['axis'])
    delta = torch.tensor(a['delta'], device=device).requires_grad_(True)
    q._delta = delta
    zf = None
    if case['kind'] == 'asym':
        zf = torch.tensor(a['zero_float'], device=device).requires_grad_(True)
        q._zero_float = zf
    else:
        q._signed = torch.tensor(a['signed'], device=device)
    x = torch.tensor(a['x'], device=device).requires_grad_(True)
    y = q(x)
    y.backward(torch.tensor(a['g'], device=device))
    log = case['scale_domain'] == 'log'
    _, _, _, (mag_s, mag_z) = oracle_backward(case, a)
    if log:
        np.testing.assert_allclose(y.detach().cpu().numpy(), a['y'], rtol=3e-7, atol=0)
        np.testing.assert_allclose(x.grad.cpu().numpy(), a['grad_x'], rtol=1e-6, atol=0)
    else:
        assert np.array_equal(y.detach().cpu().numpy(), a['y']), 'forward differs'
        assert np.array_equal(x.grad.cpu().numpy(), a['grad_x']), 'grad_x differs'
    assert delta.grad.shape == delta.shape
    close_sum(delta.grad.cpu().numpy(), a['grad_delta'], mag_s, 'grad_delta')
    if zf is not None:
        close_sum(zf.grad.cpu().numpy(), a['grad_zero_float'], mag_z, 'grad_zero_float')
    # STE only (fixed ranges): same grad_x, no parameter gradients requested
    q2 = cls(n_bits=case['n_bits'], scale_domain=case['scale_domain'], per_channel=case['per_channel'],
             axis=case['axis'])
    q2._delta = torch.tensor(a['delta'], device=device)
    if zf is not None:
        q2._zero_float = torch.tensor(a['zero_float'], device=device)
    else:
        q2._signed = torch.tensor(a['signed'], device=device)
    x2 = torch.tensor(a['x'], device=device).requires_grad_(True)
    q2(x2).backward(torch.tensor(a['g'], device=device))
    assert torch.equal(x2.grad, x.grad)


def check_adaround_case(case, device):
    """AdaRound quantizer classes on `device` vs the reference (alpha init, soft / hard forward, d / d alpha)"""
    import torch
    from quantization.adaround.quantizer import ADAROUND_QUANTIZER_MAP
    from quantization.adaround.utils import AdaRoundMode
    from quantization.quantizers import QMethods
    a = adaround_case_arrays(case)
    base = QMethods.symmetric_uniform.cls if case['kind'] == 'sym' else QMethods.asymmetric_uniform.cls
    q = ADAROUND_QUANTIZER_MAP[base](n_bits=case['n_bits'], per_channel=case['per_channel'])
    q._delta = torch.tensor(a['delta'], device=device)
    if case['kind'] == 'asym':
        q._zero_float = torch.tensor(a['zero_float'], device=device)
    else:
        q._signed = torch.tensor(a['signed'], device=device)
    w = torch.tensor(a['w'], device=device)
    scale, _, lo, hi = adaround_grid(case, a)
    step = float(np.max(scale))
    ytol = 2e-6 * step * max(abs(lo), hi)
    assert np.array_equal(q(w).cpu().numpy(), a['y_nearest']), 'nearest mode differs'
    q.round_mode = AdaRoundMode[case['mode']]
    q.temperature = case['temperature']
    q.soft_targets = True
    y0 = q(w)
    assert isinstance(q.alpha, torch.nn.Parameter) and y0.requires_grad
    np.testing.assert_allclose(q.alpha.detach().cpu().numpy(), a['alpha0'], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(y0.detach().cpu().numpy(), a['y_soft0'], rtol=0, atol=ytol)
    with torch.no_grad():
        q.alpha.copy_(torch.tensor(a['alpha1'], device=device))
    y1 = q(w)
    y1.backward(torch.tensor(a['g'], device=device))
    np.testing.assert_allclose(y1.detach().cpu().numpy(), a['y_soft1'], rtol=0, atol=ytol)
    np.testing.assert_allclose(q.alpha.grad.cpu().numpy(), a['grad_alpha1'], rtol=2e-5, atol=1e-7 * step)
    q.soft_targets = False
    assert np.array_equal(q(w).cpu().numpy(), a['y_hard1']), 'hard targets differ'
    assert np.array_equal(q.to_integer_forward(w).cpu().numpy(), a['x_int_hard1'])


def check_adaround_layer_case(case, device):
    """apply_adaround_to_layer (the whole local-loss loop) on `device` vs the reference run with the same
    seed.  Trajectories differ in the last bits (transcendentals, GEMM order) and Adam normalises the
    noise-level gradients of the first iterations (soft loss ~1e-15), so: alpha within 5e-3 on average and
    0.1 at most, at most 1 % of the hard up / down decisions differ, losses within 2 % (+ 1e-6)."""
    import torch
    from torch import nn
    from quantization.adaround import apply_adaround_to_layer
    from quantization.adaround.config import DEFAULT_ADAROUND_CONFIG, AdaRoundConfig
    from quantization.adaround.utils import AdaRoundMode, AdaRoundTempDecayType
    from quantization.autoquant_utils import QuantLinear
    from quantization.base_quantized_model import QuantizedModel
    from quantization.quantizers import QMethods
    g, nm = qat_file(), case['name']

    class Tiny(QuantizedModel):
        def __init__(self, n_bits):
            super().__init__()
            kw = dict(method=QMethods.symmetric_uniform, act_method=QMethods.asymmetric_uniform, n_bits=n_bits,
                      n_bits_act=8)
            self.fc1 = QuantLinear(32, 48, activation=nn.ReLU(), **kw)
            self.fc2 = QuantLinear(48, 16, **kw)

        def forward(self, x):
            return self.fc2(self.fc1(x))

    model = Tiny(case['n_bits'])
    with torch.no_grad():
        model.fc1.weight.copy_(torch.tensor(g['adal.w1']))
        model.fc1.bias.copy_(torch.tensor(g['adal.b1']))
        model.fc2.weight.copy_(torch.tensor(g['adal.w2']))
        model.fc2.bias.copy_(torch.tensor(g['adal.b2']))
    model.to(device)
    model.eval()
    model.quantized_weights()
    data = torch.tensor(g['adal.data'], device=device)
    with torch.no_grad():
        model(data[:8])
    cfg = AdaRoundConfig(**DEFAULT_ADAROUND_CONFIG)
    cfg.round_mode = AdaRoundMode[case['mode']]
    cfg.decay_type = AdaRoundTempDecayType[case['decay']]
    cfg.iters, cfg.lr, cfg.asym = case['iters'], case['lr'], True
    torch.manual_seed(case['seed'])
    layer = getattr(model, case['layer'])
    res = apply_adaround_to_layer(model, layer, data, batch_size=case['batch_size'], act_quant=False,
                                  adaround_config=cfg, keep_gpu=True)
    qz = layer.weight_quantizer.quantizer
    assert np.array_equal(qz._delta.detach().cpu().numpy().reshape(-1), g[f'{nm}.delta'])
    alpha = qz.alpha.detach().cpu().numpy()
    d = np.abs(alpha - g[f'{nm}.alpha'])
    assert d.mean() < 5e-3 and d.max() < 0.1
    with torch.no_grad():
        w_hard = qz(layer.weight).cpu().numpy()
    assert (w_hard != g[f'{nm}.w_hard']).mean() <= 0.01
    got = np.array([res.loss_soft_before, res.loss_hard_before, res.loss_soft_after, res.loss_hard_after])
    np.testing.assert_allclose(got, g[f'{nm}.losses'], rtol=2e-2, atol=1e-6)
    assert layer.caching and not qz.soft_targets


def check_training_step(device):
    """A QuantLinear in training mode (weights through FakeQuantSTE, learnable ranges) against the same
    computation written with torch ops (the reference's formulation) on `device`."""
    import torch
    from quantization.autoquant_utils import QuantLinear
    from quantization.quantizers import QMethods
    torch.manual_seed(0)
    lin = QuantLinear(256, 192, method=QMethods.symmetric_uniform, act_method=QMethods.asymmetric_uniform, n_bits=4,
                      n_bits_act=8).to(device)
    x = torch.randn(8, 32, 256, device=device)
    coef = torch.linspace(-1, 1, 192, device=device)
    lin.quantized()
    lin.eval()
    with torch.no_grad():
        lin(x)
    lin.learn_ranges()
    lin.train()
    y = lin(x)
    (y * coef).sum().backward()
    wq, aq = lin.weight_quantizer.quantizer, lin.activation_quantizer.quantizer
    assert isinstance(wq._delta, torch.nn.Parameter) and isinstance(aq._zero_float, torch.nn.Parameter)
    # the reference's formulation in torch ops
    w = lin.weight.detach().clone().requires_grad_(True)
    dw = wq._delta.detach().clone().requires_grad_(True)
    da = aq._delta.detach().clone().requires_grad_(True)
    za = aq._zero_float.detach().clone().requires_grad_(True)

    def ste(v):
        return v + (torch.round(v) - v).detach()
    sw = torch.clamp(dw, min=1e-8)
    lo, hi = (-8.0, 7.0) if wq.signed else (0.0, 15.0)
    wqq = sw * torch.clamp(ste(w / sw), lo, hi)
    out = torch.nn.functional.linear(x, wqq, lin.bias.detach())
    sa = torch.clamp(da, min=1e-8)
    zp = torch.clamp(ste(za), 0, 255)
    yq = sa * (torch.clamp(ste(out / sa) + zp, 0, 255) - zp)
    (yq * coef).sum().backward()
    step = float(sa.detach())
    assert (y - yq).abs().max().item() <= step * 1.01          # GEMM summation order: at most one step
    assert (y != yq).float().mean().item() < 5e-3
    for got, want, name in ((lin.weight.grad, w.grad, 'weight'), (wq._delta.grad, dw.grad, 'w delta'),
                            (aq._delta.grad, da.grad, 'a delta'), (aq._zero_float.grad, za.grad, 'a zero_float')):
        err = (got - want.view_as(got)).abs().max().item()
        ref = want.abs().max().item()
        assert err <= 2e-2 * ref + 1e-6, f'{name}: {err} vs {ref}'
