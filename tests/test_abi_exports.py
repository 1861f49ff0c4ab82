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
    assert lib.tq_version() == tq_native.ABI_VERSION == 4
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


def test_argument_errors_are_reported_before_any_device_work():
    """Every entry point validates its arguments before it touches the device, so the error contract of
    include/tq_b200.h (TQ_EINVAL = -1, TQ_EWORKSPACE = -3, ...) can be checked without a GPU -- no kernel is
    launched by any of these calls."""
    import ctypes
    import tq_native
    lib = tq_native.load_library()
    Q = tq_native.QSpec
    fake = 0x1000                                  # a non-NULL "device pointer" that is never dereferenced
    good = Q(fake, fake, None, 8, 0, 1e-8)
    assert lib.tq_version() == tq_native.ABI_VERSION == 4
    assert lib.tq_error_string(-1) == b'tq: invalid argument' and lib.tq_error_string(0) == b'ok'
    assert lib.tq_error_string(-3) == b'tq: workspace too small'
    # quantize-dequantize
    assert lib.tq_qdq_f32(fake, fake, -1, good, None) == -1
    assert lib.tq_qdq_f32(None, fake, 16, good, None) == -1
    assert lib.tq_qdq_f32(fake, fake, 16, Q(None, fake, None, 8, 0, 1e-8), None) == -1          # no delta
    assert lib.tq_qdq_f32(fake, fake, 16, Q(fake, fake, None, 17, 0, 1e-8), None) == -1         # n_bits > 16
    assert lib.tq_qdq_f32(fake, fake, 16, Q(fake, None, None, 8, 0, 1e-8), None) == -1          # symmetric without `signed`
    assert lib.tq_qdq_axis_f32(fake, fake, 4, 0, 1, good, None) == -1
    assert lib.tq_quant_int_f32(fake, None, None, 1, 1, 16, good, None) == -1                    # no output requested
    # training path
    assert lib.tq_qdq_bwd_f32(fake, fake, None, None, None, -1, 1, 8, good, fake, 1 << 20, None) == -1
    assert lib.tq_qdq_bwd_f32(None, fake, None, None, None, 1, 1, 8, good, fake, 1 << 20, None) == -1
    assert lib.tq_qdq_bwd_f32(fake, fake, None, None, None, 1, 1, 8, good, fake, 8, None) == -3   # workspace too small
    assert lib.tq_qdq_bwd_f32(fake, fake, None, None, None, 1, 1, 8, good, fake + 4, 1 << 20, None) == -2   # misaligned ws
    need = lib.tq_qdq_bwd_workspace_bytes(1, 1, 1 << 20)
    assert need >= 1024 and lib.tq_qdq_bwd_workspace_bytes(4096, 768, 1) > need
    assert lib.tq_qdq_bwd_workspace_bytes(-1, 1, 1) == 0
    assert lib.tq_adaround_fwd_f32(fake, fake, fake, None, 1, 1, 16, good, 3, 1, 0.0, None) == -1         # unknown mode
    assert lib.tq_adaround_fwd_f32(fake, fake, fake, None, 1, 1, 16, good, 2, 1, 0.0, None) == -1         # temp decay needs T > 0
    assert lib.tq_adaround_fwd_f32(fake, None, fake, None, 1, 1, 16, good, 0, 1, 0.0, None) == -1         # no alpha
    assert lib.tq_adaround_bwd_f32(fake, fake, None, fake, 1, 1, 16, good, 0, 0.0, None) == -1            # no grad_y
    # fused linear / attention / LayerNorm: unsupported shapes are refused (callers take the library path), not run
    null, wq = Q(None, None, None, 8, 0, 1e-8), Q(fake, None, fake, 8, 0, 1e-8)

    def linear(a, M, N, K):
        return lib.tq_linear_qdq_bf16(a, fake, None, fake, None, M, N, K, 1, good, wq, 1, 0, null, 1, None, None, 0, None)
    assert linear(fake, 128, 64, 96) == -4                      # K % 64 != 0
    assert linear(fake, 128, 60, 128) == -4                     # N % 8 != 0
    assert linear(None, 128, 64, 128) == -1
    assert linear(fake + 2, 128, 64, 128) == -2                 # operand not 16-byte aligned
    assert lib.tq_attention_qdq_bf16(fake, fake, 2, 64, 4, 64, good, good, good, good, good, good, None, None) == -4
    assert lib.tq_attention_qdq_bf16(fake, fake, 2, 128, 4, 32, good, good, good, good, good, good, None, None) == -4
    assert lib.tq_ln_qdq_bf16(fake, good, 1, fake, fake, 1e-12, good, 1, fake, None, 16, 100, None) == -4
    # range estimation
    assert lib.tq_minmax_f32(fake, 0, fake, fake, 1024, None) == -1
    assert lib.tq_minmax_f32(fake, 16, fake, fake, 4, None) == -3
    assert lib.tq_group_minmax_f32(fake, fake, 10, 3, None, fake, fake, None) == -1   # reference: AssertionError on d % n_groups
    assert lib.tq_mse_sse_f32(fake, 16, fake, 0, fake, fake, 1 << 20, None) == -1
    assert lib.tq_set_range_asym_f32(fake, fake, 1, 0, 1e-8, 0, fake, fake, None) == -1
    # probes
    assert lib.tq_probe_copy_f32(None, fake, 16, 0, None) == -1
    assert lib.tq_probe_copy_f32(fake, fake, 18, 0, None) == -1
    assert lib.tq_selftest_div(1, 0, 1, fake, None) == -1
    # encoder chain plans: argument errors are reported before any CUDA call
    import ctypes
    handle = ctypes.c_void_p()
    stage = tq_native.ChainStage()
    stage.kind, stage.N, stage.K, stage.nseg = 1, 3072, 768, 1
    arr = (tq_native.ChainStage * 1)(stage)
    assert lib.tq_chain_plan_create(None, 1, 128, ctypes.byref(handle)) == -1
    assert lib.tq_chain_plan_create(arr, 0, 128, ctypes.byref(handle)) == -1
    assert lib.tq_chain_plan_create(arr, 1, 0, ctypes.byref(handle)) == -1
    assert lib.tq_chain_plan_create(arr, 1, 128, ctypes.byref(handle)) == -4    # no LayerNorm stage: nothing fixes the cluster size
    stage.kind, stage.N = 2, 1000                                                # LayerNorm stage: N must be a multiple of 192, <= 1536
    arr = (tq_native.ChainStage * 1)(stage)
    assert lib.tq_chain_plan_create(arr, 1, 128, ctypes.byref(handle)) == -4
    assert handle.value is None
    assert lib.tq_chain_plan_run(None, None) == -1
    assert lib.tq_chain_plan_destroy(None) == 0


def test_struct_layouts_match_the_header(tmp_path):
    """struct tq_qspec / tq_chain_stage: size and every field offset of the ctypes mirrors vs the C header (gcc)"""
    import ctypes
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        pytest.skip('gcc not available')
    structs = {'tq_qspec': tq_native.QSpec, 'tq_chain_stage': tq_native.ChainStage}
    lines = ['#include "tq_b200.h"', '#include <stdio.h>', '#include <stddef.h>', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-std=c99', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out[f'{cname}.{fname}']) == getattr(cls, fname).offset, (cname, fname)
