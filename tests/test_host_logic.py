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
