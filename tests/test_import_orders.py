"""Every public module of the package imports on its own in a fresh interpreter, whatever was imported before it
(the reference's `utils` and `quantization.adaround` packages import each other; a test-suite that always imports
them in one lucky order would hide a cycle)."""
import subprocess
import sys

import pytest

from conftest import PKG

ENTRY_POINTS = [
    'import quantization.adaround',
    'from quantization.adaround.quantizer import ADAROUND_QUANTIZER_MAP',
    'import utils; import quantization.adaround',
    'import utils.quant_options',
    'import utils.adaround_utils',
    'import quantization.autoquant_utils',
    'import engine.configs',
    'import engine.fused',
]


@pytest.mark.parametrize('stmt', ENTRY_POINTS)
def test_module_imports_in_a_fresh_interpreter(stmt):
    r = subprocess.run([sys.executable, '-c', f'import sys; sys.path.insert(0, {PKG!r}); {stmt}'],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
