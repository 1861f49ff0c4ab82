"""pytest configuration: registers the `gpu` marker and exposes shared fixtures.

* `-m "not gpu"`: oracle vs golden vectors, host-side logic (with the oracle injected as the
  arithmetic back-end -- test-only), C-ABI symbol export check.  Runs without a GPU.
* `-m gpu`: parity tests proper -- every call goes through the C-ABI of libtq_b200.so.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'transformer-quantization_b200')
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without a GPU (or without the built library) skips the gpu-marked tests
    instead of failing them; on a GPU box with the library they always run -- and fail loudly if the CUDA
    path is broken."""
    import torch
    lib = os.path.join(PKG, 'lib', 'libtq_b200.so')
    if torch.cuda.is_available() and os.path.exists(os.environ.get('TQ_B200_LIB', lib)):
        return
    why = 'no CUDA device visible' if not torch.cuda.is_available() else f'{lib} not built'
    skip = pytest.mark.skip(reason=f'gpu test: {why}')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


class Golden:
    def __init__(self):
        with open(os.path.join(GOLDEN, 'manifest.json')) as f:
            self.manifest = json.load(f)
        self._files = {}

    def file(self, name):
        if name not in self._files:
            self._files[name] = np.load(os.path.join(GOLDEN, name + '.npz'))
        return self._files[name]

    def cases(self, section):
        return self.manifest[section]


_G = Golden()


@pytest.fixture(scope='session')
def golden():
    return _G


def golden_cases(section):
    return _G.cases(section)
