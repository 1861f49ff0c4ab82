"""Host-side logic of the package on CPU: the reference-facing classes run with the CPU oracle
injected as the arithmetic back-end (test-only, see tests/oracle_backend.py) and must reproduce
the golden outputs of the unmodified reference."""
import numpy as np
import pytest
import torch

import tq_native
from conftest import golden_cases
from oracle_backend import OracleOps
import parity_cases as P

CPU = torch.device('cpu')


@pytest.fixture(autouse=True)
def oracle_ops(monkeypatch):
    monkeypatch.setattr(tq_native, '_OPS', OracleOps())
    monkeypatch.setattr(tq_native, 'default_device', lambda: CPU)


@pytest.mark.parametrize('case', golden_cases('quantizers'), ids=lambda c: c['name'])
def test_quantizer_api(case, golden):
    P.check_quantizer_case(case, golden.file('quantizers'), CPU)


@pytest.mark.parametrize('case', golden_cases('estimators'), ids=lambda c: c['name'])
def test_estimator_api(case, golden):
    P.check_estimator_case(case, golden.file('estimators'), CPU)


@pytest.mark.parametrize('case', golden_cases('mse'), ids=lambda c: c['name'])
def test_mse_api(case, golden):
    P.check_mse_case(case, golden.file('mse'), CPU)


@pytest.mark.parametrize('case', golden_cases('manager'), ids=lambda c: c['name'])
def test_manager_api(case, golden):
    P.check_manager_case(case, golden.file('manager'), CPU)


@pytest.mark.parametrize('case', golden_cases('linear'), ids=lambda c: c['name'])
def test_quant_linear_api(case, golden):
    P.check_linear_case(case, golden.file('linear'), CPU, exact_gemm=True)


def test_percentile_estimator_matches_numpy():
    """CurrentMinMaxEstimator(percentile=...) evaluates np.percentile's 'linear' method where the tensor
    lives (sort + numpy's own float32-difference / float64-interpolation formula); the reference calls
    np.percentile on the host (range_estimators.py:121-140).  Bit-identical, NaN slices included."""
    import numpy as np
    import torch
    from quantization.range_estimators import CurrentMinMaxEstimator
    rs = np.random.RandomState(0)
    for trial in range(200):
        n, rows = int(rs.randint(1, 3000)), int(rs.randint(1, 5))
        x = (rs.randn(rows, n) * rs.rand() * 10).astype(np.float32)
        if trial % 9 == 0:
            x[0, rs.randint(0, n)] = np.nan
        q = float(rs.choice([0.0, 0.01, 0.1, 1.0, 5.0, 50.0, 99.9, 99.99, 100.0, rs.rand() * 100]))
        with np.errstate(all='ignore'):
            ref = np.percentile(x, (q, 100 - q), axis=-1)
        got = CurrentMinMaxEstimator._percentiles(torch.from_numpy(x), (q, 100 - q))
        for r, g in zip(ref, got):
            assert np.array_equal(torch.Tensor(r).numpy(), g.numpy(), equal_nan=True), (n, q)


def test_grid_tag_is_bound_to_the_range_that_produced_the_tensor():
    """A fused QuantLinear recovers the integer grid of its input from the tag the producing quantizer attached.
    The quantizer's range buffers are rewritten in place by its next set_quant_range, so a tag must stop being
    valid when that happens (or when the tensor is edited in place)."""
    from quantization import fused_linear
    from quantization.quantizers import QMethods
    for qm in (QMethods.asymmetric_uniform, QMethods.symmetric_uniform):
        q = qm.cls(n_bits=8)
        q.set_quant_range(-1.0, 1.0)
        x = torch.linspace(-1, 1, 64)
        y = q(x)
        tag = fused_linear._valid_tag(y)
        assert tag is not None and tag.quantizer is q
        q.set_quant_range(-2.0, 2.0)                    # new range, same buffers
        assert fused_linear._valid_tag(y) is None
        y2 = q(x)
        assert fused_linear._valid_tag(y2) is not None
        y2 += 1.0                                       # in-place edit (the reference's models do this)
        assert fused_linear._valid_tag(y2) is None
