"""The C-ABI shared library loads without a GPU and exports exactly what include/tq_b200.h declares."""
import os
import re

import pytest

import tq_native
from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'tq_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return set(re.findall(r'\b(tq_[a-z0-9_]+)\s*\(', src))


def test_header_matches_binding():
    assert header_symbols() == set(tq_native.SIGNATURES)


def test_library_exports_every_symbol():
    if not os.path.exists(tq_native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = tq_native.load_library()      # getattr() on every declared symbol; no CUDA call
    assert lib.tq_version() == 1
    assert b'invalid' in lib.tq_error_string(-1)


def test_no_cpu_fallback():
    """CPU tensors must be rejected loudly by the product back-end."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    with pytest.raises(tq_native.TQError):
        tq_native.ops()
