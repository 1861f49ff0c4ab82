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


def test_product_code_never_imports_the_oracle():
    """the CPU oracle is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
    import it -- nothing under the package or tools/ does (a product path routed through it would void parity)"""
    import ast
    import os
    from conftest import PKG, ROOT
    offenders = []
    for top in (PKG, os.path.join(ROOT, 'tools')):
        for dirpath, _, files in os.walk(top):
            for f in files:
                if not f.endswith('.py'):
                    continue
                path = os.path.join(dirpath, f)
                tree = ast.parse(open(path).read())
                for node in ast.walk(tree):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom) and node.module:
                        names = [node.module]
                    if any(n == 'oracle' or n.startswith('oracle.') or n == 'oracle_backend' for n in names):
                        offenders.append(path)
    assert not offenders, offenders
