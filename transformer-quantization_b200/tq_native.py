"""ctypes binding of libtq_b200.so (C ABI: include/tq_b200.h) for PyTorch CUDA tensors.

This is the ONLY arithmetic back-end of the package: there is no CPU / eager fallback.  If the
shared library is missing, or a tensor is not on a CUDA device, the call raises immediately.

PyTorch is plumbing here: it owns device memory (caching allocator), the current stream and the
tensors' metadata; the binding only unwraps ``data_ptr()`` / ``cuda_stream`` and forwards them.
Nothing in this module allocates inside a C call or synchronises the host, so every op can be
captured in a CUDA graph (outputs are allocated with ``torch.empty`` before the call).
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('TQ_B200_LIB', os.path.join(_HERE, 'lib', 'libtq_b200.so'))

_c_f32p = ctypes.c_void_p
_i64 = ctypes.c_int64
_i32 = ctypes.c_int32


class QSpec(ctypes.Structure):
    """struct tq_qspec (include/tq_b200.h)."""
    _fields_ = [
        ('delta', ctypes.c_void_p),
        ('zero_float', ctypes.c_void_p),
        ('is_signed', ctypes.c_void_p),
        ('n_bits', ctypes.c_int32),
        ('log_domain', ctypes.c_int32),
        ('eps', ctypes.c_float),
    ]


class ChainStage(ctypes.Structure):
    """struct tq_chain_stage (include/tq_b200.h)."""
    _fields_ = [
        ('a_i8', ctypes.c_void_p), ('w_i8', ctypes.c_void_p), ('w_rowsum', ctypes.c_void_p), ('bias', ctypes.c_void_p),
        ('out', ctypes.c_void_p), ('N', ctypes.c_int64), ('K', ctypes.c_int64),
        ('a_q', QSpec), ('w_q', QSpec), ('out_q', QSpec), ('nseg', ctypes.c_int32), ('kind', ctypes.c_int32),
        ('res_i8', ctypes.c_void_p), ('res_q', QSpec), ('out2_q', QSpec), ('ln_q', QSpec),
        ('ln_gamma_q', ctypes.c_void_p), ('ln_beta', ctypes.c_void_p), ('ln_eps', ctypes.c_float),
    ]


class TQError(RuntimeError):
    pass


class ChainPlan:
    """owner of one tq_chain_plan (device-side stage descriptors); keeps the stage structs alive for inspection"""

    def __init__(self, ops, stages, M):
        self._ops = ops
        self.stages = list(stages)
        self.M = int(M)
        self.flops = sum(self.M * getattr(s, '_flops', 0) for s in self.stages)
        arr = (ChainStage * len(self.stages))(*self.stages)
        h = ctypes.c_void_p()
        ops._check(ops.lib.tq_chain_plan_create(arr, len(self.stages), self.M, ctypes.byref(h)))
        self.handle = h

    def close(self):
        if self.handle is not None and self.handle.value:
            self._ops.lib.tq_chain_plan_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


ACT_FN = {'none': 0, 'gelu': 1, 'relu': 2, 'tanh': 3}

# exported symbol -> (restype, argtypes); also used by the CPU-side export test
SIGNATURES = {
    'tq_version': (ctypes.c_int, []),
    'tq_error_string': (ctypes.c_char_p, [ctypes.c_int]),
    'tq_device_sm_count': (ctypes.c_int, []),
    'tq_selftest_div': (ctypes.c_int, [ctypes.c_uint64, _i32, _i32, ctypes.c_void_p, ctypes.c_void_p]),
    'tq_probe_copy_f32': (ctypes.c_int, [_c_f32p, _c_f32p, _i64, _i32, ctypes.c_void_p]),
    'tq_qdq_f32': (ctypes.c_int, [_c_f32p, _c_f32p, _i64, QSpec, ctypes.c_void_p]),
    'tq_qdq_axis_f32': (ctypes.c_int, [_c_f32p, _c_f32p, _i64, _i64, _i64, QSpec, ctypes.c_void_p]),
    'tq_quant_int_f32': (ctypes.c_int, [_c_f32p, _c_f32p, ctypes.c_void_p, _i64, _i64, _i64, QSpec,
                                        ctypes.c_void_p]),
    'tq_minmax_workspace_bytes': (ctypes.c_size_t, [_i64]),
    'tq_minmax_f32': (ctypes.c_int, [_c_f32p, _i64, _c_f32p, ctypes.c_void_p, ctypes.c_size_t,
                                     ctypes.c_void_p]),
    'tq_minmax_axis_f32': (ctypes.c_int, [_c_f32p, _i64, _i64, _i64, _c_f32p, _c_f32p, ctypes.c_void_p,
                                          ctypes.c_size_t, ctypes.c_void_p]),
    'tq_group_minmax_f32': (ctypes.c_int, [_c_f32p, _c_f32p, _i64, _i32, _c_f32p, _c_f32p, _c_f32p,
                                           ctypes.c_void_p]),
    'tq_dim_ranges_f32': (ctypes.c_int, [_c_f32p, _c_f32p, _i64, _i32, _c_f32p, ctypes.c_void_p]),
    'tq_range_update_f32': (ctypes.c_int, [_c_f32p, _c_f32p, _c_f32p, _c_f32p, _i64, _i32,
                                           ctypes.c_double, _i32, ctypes.c_void_p]),
    'tq_set_range_asym_f32': (ctypes.c_int, [_c_f32p, _c_f32p, _i64, _i32, ctypes.c_float, _i32,
                                             _c_f32p, _c_f32p, ctypes.c_void_p]),
    'tq_set_range_sym_f32': (ctypes.c_int, [_c_f32p, _c_f32p, _i64, _i32, ctypes.c_float, _i32,
                                            _c_f32p, ctypes.c_void_p, ctypes.c_void_p]),
    'tq_calib_finalize_f32': (ctypes.c_int, [ctypes.c_void_p, _c_f32p, _c_f32p, _i32, ctypes.c_double, _i32, _i32, _i32,
                                             ctypes.c_float, _i32, _c_f32p, _c_f32p, ctypes.c_void_p, ctypes.c_void_p]),
    'tq_mse_workspace_bytes': (ctypes.c_size_t, [_i32]),
    'tq_mse_sse_f32': (ctypes.c_int, [_c_f32p, _i64, _c_f32p, _i32, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_size_t, ctypes.c_void_p]),
    'tq_mse_argmin_f64': (ctypes.c_int, [ctypes.c_void_p, _i32, _c_f32p, _c_f32p, _c_f32p, _c_f32p,
                                         ctypes.c_void_p, ctypes.c_void_p]),
    'tq_linear_workspace_bytes': (ctypes.c_size_t, [_i64, _i64, _i64]),
    'tq_linear_qdq_bf16': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _c_f32p, _c_f32p,
                                          ctypes.c_void_p, _i64, _i64, _i64, _i32, QSpec, QSpec, _i64,
                                          _i32, QSpec, _i64, _c_f32p, ctypes.c_void_p,
                                          ctypes.c_size_t, ctypes.c_void_p]),
    'tq_linear_res_qdq_bf16': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _c_f32p, _c_f32p, ctypes.c_void_p,
                                              _i64, _i64, _i64, QSpec, QSpec, _i64, QSpec, _i64, ctypes.c_void_p,
                                              QSpec, QSpec, _i64, ctypes.c_void_p]),
    'tq_linear_res_ln_qdq_bf16': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _c_f32p, _c_f32p, ctypes.c_void_p,
                                                 _i64, _i64, _i64, QSpec, QSpec, _i64, QSpec, ctypes.c_void_p, QSpec,
                                                 QSpec, _c_f32p, _c_f32p, ctypes.c_float, QSpec, ctypes.c_void_p]),
    'tq_attention_qdq_bf16': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _i32, _i32, _i32, _i32, QSpec, QSpec,
                                             QSpec, QSpec, QSpec, QSpec, _c_f32p, ctypes.c_void_p]),
    'tq_ln_qdq_bf16': (ctypes.c_int, [ctypes.c_void_p, QSpec, _i64, _c_f32p, _c_f32p, ctypes.c_float, QSpec, _i64,
                                      ctypes.c_void_p, _c_f32p, _i64, _i32, ctypes.c_void_p]),
    'tq_embed_ln_qdq_bf16': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _i64, _c_f32p, _c_f32p,
                                            _c_f32p, QSpec, _i64, QSpec, _i64, _c_f32p, _c_f32p, ctypes.c_float,
                                            QSpec, _i64, ctypes.c_void_p, _c_f32p, _i64, _i32, ctypes.c_void_p]),
    'tq_split3_bf16': (ctypes.c_int, [_c_f32p, ctypes.c_void_p, _i64, _i64, ctypes.c_void_p]),
    # ---- 8-bit integer operand mode
    'tq_linear_qdq_i8': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _c_f32p, _c_f32p,
                                        ctypes.c_void_p, ctypes.c_void_p, _i64, _i64, _i64, QSpec, QSpec, _i64, _i32,
                                        QSpec, _i64, ctypes.c_void_p]),
    'tq_linear_seg_qdq_i8': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _c_f32p, ctypes.c_void_p,
                                            ctypes.c_void_p, _i64, _i64, _i64, QSpec, QSpec, QSpec, _i32, _i32, _i64,
                                            ctypes.c_void_p]),
    'tq_head_qdq_i8': (ctypes.c_int, [ctypes.c_void_p, _i64, _i32, _i32, _i32, ctypes.c_void_p, ctypes.c_void_p, _c_f32p, QSpec, QSpec, QSpec,
                                      ctypes.c_void_p, ctypes.c_void_p, _c_f32p, QSpec, QSpec, ctypes.c_void_p, _i64, ctypes.c_void_p]),
    'tq_chain_plan_create': (ctypes.c_int, [ctypes.POINTER(ChainStage), _i32, _i64, ctypes.POINTER(ctypes.c_void_p)]),
    'tq_chain_plan_run': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'tq_chain_plan_destroy': (ctypes.c_int, [ctypes.c_void_p]),
    'tq_linear_nonorm_qdq_i8': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _c_f32p, ctypes.c_void_p,
                                               _i64, _i64, _i64, QSpec, QSpec, QSpec, ctypes.c_void_p, QSpec, QSpec, _c_f32p,
                                               _c_f32p, QSpec, _i64, ctypes.c_void_p]),
    'tq_linear_peg_qdq_i8': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _c_f32p, ctypes.c_void_p,
                                            ctypes.c_void_p, _i64, _i64, _i64, QSpec, _i32, QSpec, _i32, QSpec, _i32, _i64,
                                            _i32, ctypes.c_void_p]),
    'tq_linear_peg_res_ln_qdq_i8': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _c_f32p, ctypes.c_void_p,
                                                   ctypes.c_void_p, _i64, _i64, _i64, QSpec, _i32, QSpec, _i32, QSpec, _i32,
                                                   ctypes.c_void_p, QSpec, _i32, QSpec, _i32, _c_f32p, _c_f32p,
                                                   ctypes.c_float, QSpec, _i32, _i64, ctypes.c_void_p]),
    'tq_attention_pad_qdq_i8': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _i32, _i32, _i32, _i32, _i32, QSpec, QSpec,
                                               QSpec, QSpec, QSpec, QSpec, _c_f32p, ctypes.c_void_p]),
    'tq_attention_peg_qdq_i8': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _i32, _i32, _i32, _i32, QSpec, QSpec,
                                               QSpec, _i32, QSpec, QSpec, QSpec, _i32, _c_f32p, ctypes.c_void_p]),
    'tq_linear_qdq_bf16_o8': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _c_f32p, ctypes.c_void_p, _i64, _i64, _i64,
                                             QSpec, QSpec, _i64, _i32, QSpec, _i64, ctypes.c_void_p]),
    'tq_linear_res_ln_qdq_i8': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _c_f32p, _c_f32p,
                                               ctypes.c_void_p, ctypes.c_void_p, _i64, _i64, _i64, QSpec, QSpec, _i64, QSpec,
                                               ctypes.c_void_p, QSpec, QSpec, _c_f32p, _c_f32p, ctypes.c_float, QSpec,
                                               ctypes.c_void_p]),
    'tq_attention_qdq_i8': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _i32, _i32, _i32, _i32, QSpec, QSpec,
                                           QSpec, QSpec, QSpec, QSpec, _c_f32p, ctypes.c_void_p]),
    'tq_embed_ln_qdq_i8': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _i64, _c_f32p, _c_f32p,
                                          _c_f32p, QSpec, _i64, QSpec, _i64, _c_f32p, _c_f32p, ctypes.c_float,
                                          QSpec, _i64, ctypes.c_void_p, _i64, _i32, ctypes.c_void_p]),
    # ---- training-time path (STE backward with learnable ranges, AdaRound)
    'tq_qdq_bwd_workspace_bytes': (ctypes.c_size_t, [_i64, _i64, _i64]),
    'tq_qdq_bwd_f32': (ctypes.c_int, [_c_f32p, _c_f32p, _c_f32p, _c_f32p, _c_f32p, _i64, _i64, _i64, QSpec,
                                      ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    'tq_adaround_init_alpha_f32': (ctypes.c_int, [_c_f32p, _c_f32p, _i64, _i64, _i64, QSpec, _i32, ctypes.c_float,
                                                  ctypes.c_void_p]),
    'tq_adaround_fwd_f32': (ctypes.c_int, [_c_f32p, _c_f32p, _c_f32p, _c_f32p, _i64, _i64, _i64, QSpec, _i32, _i32,
                                           ctypes.c_float, ctypes.c_void_p]),
    'tq_adaround_bwd_f32': (ctypes.c_int, [_c_f32p, _c_f32p, _c_f32p, _c_f32p, _i64, _i64, _i64, QSpec, _i32,
                                           ctypes.c_float, ctypes.c_void_p]),
}

ABI_VERSION = 4          # tq_version() of the library this binding was written against
ADAROUND_MODE = {'learned_sigmoid': 0, 'learned_hard_sigmoid': 1, 'sigmoid_temp_decay': 2}


def load_library(path=None):
    """dlopen the library and attach prototypes.  No CUDA call is made here."""
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise TQError(
            f'tq_b200: {path} not found -- build it with transformer-quantization_b200/csrc/build.sh '
            f'(or __graft_entry__.build()).  There is no CPU fallback.')
    lib = ctypes.CDLL(path)
    if lib.tq_version() < ABI_VERSION:
        raise TQError(f'tq_b200: {path} implements ABI version {lib.tq_version()}, this binding needs {ABI_VERSION} '
                      f'-- rebuild it with transformer-quantization_b200/csrc/build.sh')
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)         # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise TQError('tq_b200: CUDA tensors required (this build has no CPU path); got a '
                          f'{t.device} tensor')


class CudaOps:
    """Tensor-level wrappers around the C ABI.  One instance per process (see ``ops()``)."""

    def __init__(self, lib=None):
        self.lib = lib or load_library()
        self._ws = {}
        self._lock = threading.Lock()
        self.launches = 0            # kernels of this library enqueued so far (bench.py: gpu_launches)
        self.profile = None          # list -> every call appends (name, work, start_event, end_event)

    # -- helpers --------------------------------------------------------------------------------
    def _check(self, code):
        if code != 0:
            raise TQError(f'tq_b200 call failed ({code}): {self.lib.tq_error_string(code).decode()}')

    def _run(self, name, work, kernels, fn, *args):
        """Call one C entry point.  ``work`` = algorithmic bytes (or flops) of the call, ``kernels``
        = how many kernels it enqueues.  With ``self.profile`` set, CUDA events bracket the call on
        the launching stream (used by bench.py's roofline pass; never during graph capture)."""
        if self.profile is None:
            self._check(fn(*args))
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self._check(fn(*args))
            e1.record()
            self.profile.append((name, work, e0, e1))
        self.launches += kernels

    def workspace(self, kind, nbytes, device):
        """Zero-initialised scratch, cached per (kind, device, stream) and grown on demand."""
        key = (kind, device.index, _stream())
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            with self._lock:
                buf = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)
                self._ws[key] = buf
        return buf

    @staticmethod
    def spec(delta, zero_float, is_signed, n_bits, log_domain=False, eps=1e-8):
        return QSpec(_ptr(delta), _ptr(zero_float), _ptr(is_signed), int(n_bits), int(bool(log_domain)),
                     float(eps))

    # -- QDQ ------------------------------------------------------------------------------------
    def qdq(self, x, spec, outer=1, C=1, inner=None, out=None):
        _chk_cuda(x)
        x = x.contiguous()
        if x.dtype != torch.float32:
            raise TQError(f'tq_b200: fp32 tensors only, got {x.dtype}')
        y = torch.empty_like(x) if out is None else out
        n = x.numel()
        if n == 0:
            return y
        if C == 1:
            self._run('qdq_tensor', 8 * n, 1, self.lib.tq_qdq_f32, x.data_ptr(), y.data_ptr(), n, spec, _stream())
        else:
            self._run('qdq_axis', 8 * n, 1, self.lib.tq_qdq_axis_f32, x.data_ptr(), y.data_ptr(), outer, C, inner, spec,
                      _stream())
        return y

    def quant_int(self, x, spec, outer=1, C=1, inner=None, want_f32=True, want_bf16=False):
        _chk_cuda(x)
        x = x.contiguous()
        n = x.numel()
        yi = torch.empty_like(x) if want_f32 else None
        yc = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
        if n == 0:
            return yi, yc
        if C == 1:
            outer, inner = 1, n
        self._run('quant_int', (4 + (4 if want_f32 else 0) + (2 if want_bf16 else 0)) * n, 1, self.lib.tq_quant_int_f32,
                  x.data_ptr(), _ptr(yi), _ptr(yc), outer, C, inner, spec, _stream())
        return yi, yc

    # -- training-time path -----------------------------------------------------------------------
    def qdq_bwd(self, x, grad_y, spec, n_params, outer=1, C=1, inner=None, want_x=True, want_delta=True,
                want_zero_float=True):
        """tq_qdq_bwd_f32 -> (grad_x | None, grad_delta[n_params] | None, grad_zero_float[n_params] | None)"""
        _chk_cuda(x, grad_y)
        x, grad_y = x.contiguous(), grad_y.contiguous()
        if x.dtype != torch.float32 or grad_y.dtype != torch.float32:
            raise TQError(f'tq_b200: fp32 tensors only, got {x.dtype} / {grad_y.dtype}')
        n = x.numel()
        if C == 1:
            outer, inner = 1, n
        gx = torch.empty_like(x) if want_x else None
        gd = torch.empty(n_params, dtype=torch.float32, device=x.device) if want_delta else None
        want_zero_float = want_zero_float and bool(spec.zero_float)      # symmetric quantizer: no zero point
        gz = torch.empty(n_params, dtype=torch.float32, device=x.device) if want_zero_float else None
        nb = self.lib.tq_qdq_bwd_workspace_bytes(outer, C, inner)
        ws = self.workspace('bwd', nb, x.device)
        self._run('qdq_bwd', (12 if want_x else 8) * n, 1, self.lib.tq_qdq_bwd_f32, x.data_ptr(), grad_y.data_ptr(),
                  _ptr(gx), _ptr(gd), _ptr(gz), outer, C, inner, spec, ws.data_ptr(), ws.numel(), _stream())
        return gx, gd, gz

    def adaround_init_alpha(self, w, spec, outer, C, inner, mode, temperature=None):
        _chk_cuda(w)
        w = w.contiguous()
        alpha = torch.empty_like(w)
        self._run('adaround', 8 * w.numel(), 1, self.lib.tq_adaround_init_alpha_f32, w.data_ptr(), alpha.data_ptr(),
                  outer, C, inner, spec, ADAROUND_MODE[mode], float(temperature or 0.0), _stream())
        return alpha

    def adaround_fwd(self, w, alpha, spec, outer, C, inner, mode, soft, temperature=None, want_int=False):
        """-> y (QDQ) or x_int (want_int) of the AdaRound quantizer in a relaxation mode"""
        _chk_cuda(w, alpha)
        w, alpha = w.contiguous(), alpha.contiguous()
        out = torch.empty_like(w)
        self._run('adaround', 12 * w.numel(), 1, self.lib.tq_adaround_fwd_f32, w.data_ptr(), alpha.data_ptr(),
                  None if want_int else out.data_ptr(), out.data_ptr() if want_int else None, outer, C, inner, spec,
                  ADAROUND_MODE[mode], int(bool(soft)), float(temperature or 0.0), _stream())
        return out

    def adaround_bwd(self, w, alpha, grad_y, spec, outer, C, inner, mode, temperature=None):
        _chk_cuda(w, alpha, grad_y)
        w, alpha, grad_y = w.contiguous(), alpha.contiguous(), grad_y.contiguous()
        ga = torch.empty_like(w)
        self._run('adaround', 16 * w.numel(), 1, self.lib.tq_adaround_bwd_f32, w.data_ptr(), alpha.data_ptr(),
                  grad_y.data_ptr(), ga.data_ptr(), outer, C, inner, spec, ADAROUND_MODE[mode],
                  float(temperature or 0.0), _stream())
        return ga

    # -- min / max ------------------------------------------------------------------------------
    def minmax(self, x):
        """-> fp32 tensor [2] = (min, max) on the device."""
        _chk_cuda(x)
        x = x.contiguous()
        if x.numel() == 0:
            raise RuntimeError('min(): Expected reduction dim to be specified for input.numel() == 0')
        out = torch.empty(2, dtype=torch.float32, device=x.device)
        nb = self.lib.tq_minmax_workspace_bytes(1)
        ws = self.workspace('mm', nb, x.device)
        self._run('minmax_tensor', 4 * x.numel(), 1, self.lib.tq_minmax_f32, x.data_ptr(), x.numel(), out.data_ptr(),
                  ws.data_ptr(), ws.numel(), _stream())
        return out

    def minmax_axis(self, x, outer, C, inner):
        """x viewed [outer, C, inner] -> (mn[C], mx[C])."""
        _chk_cuda(x)
        x = x.contiguous()
        out = torch.empty(2, C, dtype=torch.float32, device=x.device)
        nb = self.lib.tq_minmax_workspace_bytes(C)
        ws = self.workspace('mm', nb, x.device)
        self._run('minmax_axis', 4 * x.numel(), 1, self.lib.tq_minmax_axis_f32, x.data_ptr(), outer, C, inner,
                  out[0].data_ptr(), out[1].data_ptr(), ws.data_ptr(), ws.numel(), _stream())
        return out[0], out[1]

    def group_minmax(self, mn, mx, n_groups, ranges=None):
        _chk_cuda(mn, mx, ranges)
        C = mn.numel()
        out = torch.empty(2, C, dtype=torch.float32, device=mn.device)
        self._run('group_minmax', 16 * C, 1, self.lib.tq_group_minmax_f32, mn.data_ptr(), mx.data_ptr(), C,
                  int(n_groups), _ptr(ranges), out[0].data_ptr(), out[1].data_ptr(), _stream())
        return out[0], out[1]

    def dim_ranges(self, mn, mx, first):
        _chk_cuda(mn, mx)
        r = torch.empty_like(mn)
        self._run('dim_ranges', 12 * mn.numel(), 1, self.lib.tq_dim_ranges_f32, mn.data_ptr(), mx.data_ptr(),
                  mn.numel(), int(bool(first)), r.data_ptr(), _stream())
        return r

    def range_update(self, new_min, new_max, cur_min, cur_max, mode, momentum=0.0, first=False):
        """In-place update of (cur_min, cur_max); mode 0 current, 1 running EMA, 2 all-minmax."""
        _chk_cuda(new_min, new_max, cur_min, cur_max)
        self._run('range_update', 16 * new_min.numel(), 1, self.lib.tq_range_update_f32, new_min.data_ptr(),
                  new_max.data_ptr(), cur_min.data_ptr(), cur_max.data_ptr(), new_min.numel(), mode,
                  float(momentum), int(bool(first)), _stream())

    # -- set_quant_range ------------------------------------------------------------------------
    def set_range_asym(self, x_min, x_max, n_bits, eps, log_domain, delta, zero_float):
        _chk_cuda(x_min, x_max, delta, zero_float)
        self._run('set_range', 16 * x_min.numel(), 1, self.lib.tq_set_range_asym_f32, x_min.data_ptr(),
                  x_max.data_ptr(), x_min.numel(), int(n_bits), float(eps), int(bool(log_domain)),
                  delta.data_ptr(), zero_float.data_ptr(), _stream())

    def set_range_sym(self, x_min, x_max, n_bits, eps, log_domain, delta, is_signed):
        _chk_cuda(x_min, x_max, delta, is_signed)
        self._run('set_range', 12 * x_min.numel(), 1, self.lib.tq_set_range_sym_f32, x_min.data_ptr(),
                  x_max.data_ptr(), x_min.numel(), int(n_bits), float(eps), int(bool(log_domain)),
                  delta.data_ptr(), is_signed.data_ptr(), _stream())

    def calib_finalize(self, tile_mm, cur_min, cur_max, mode, momentum, first, symmetric, n_bits, eps, log_domain, delta,
                       zero_float, is_signed):
        """tq_calib_finalize_f32: GEMM-epilogue min/max -> estimator update -> quantizer range, one launch"""
        _chk_cuda(tile_mm, cur_min, cur_max, delta, zero_float, is_signed)
        self._run('set_range', 32, 1, self.lib.tq_calib_finalize_f32, tile_mm.data_ptr(), cur_min.data_ptr(), cur_max.data_ptr(),
                  int(mode), float(momentum), int(bool(first)), int(bool(symmetric)), int(n_bits), float(eps),
                  int(bool(log_domain)), delta.data_ptr(), _ptr(zero_float), _ptr(is_signed), _stream())

    # -- MSE ------------------------------------------------------------------------------------
    def mse_sse(self, x, cand, n_cand, loss_accum):
        """loss_accum[c] += sum((x - QDQ_c(x))^2) for the candidate table ``cand`` [4, n_cand]."""
        _chk_cuda(x, cand, loss_accum)
        x = x.contiguous()
        nb = self.lib.tq_mse_workspace_bytes(n_cand)
        ws = self.workspace('mse', nb, x.device)
        self._run('mse_sse', 4 * x.numel(), 2, self.lib.tq_mse_sse_f32, x.data_ptr(), x.numel(), cand.data_ptr(),
                  n_cand, loss_accum.data_ptr(), ws.data_ptr(), ws.numel(), _stream())

    def mse_argmin(self, loss, cand_xmin, cand_xmax):
        """-> (xmin[1], xmax[1], idx[1]) device tensors; first minimum of the flat loss array."""
        _chk_cuda(loss, cand_xmin, cand_xmax)
        out = torch.empty(2, dtype=torch.float32, device=loss.device)
        idx = torch.empty(1, dtype=torch.int32, device=loss.device)
        self._run('mse_argmin', 8 * loss.numel(), 1, self.lib.tq_mse_argmin_f64, loss.data_ptr(), loss.numel(),
                  cand_xmin.data_ptr(), cand_xmax.data_ptr(), out[0:1].data_ptr(), out[1:2].data_ptr(),
                  idx.data_ptr(), _stream())
        return out[0:1], out[1:2], idx

    # -- fused linear -----------------------------------------------------------------------------
    def split3(self, x2d):
        _chk_cuda(x2d)
        M, K = x2d.shape
        out = torch.empty(M, 3 * K, dtype=torch.bfloat16, device=x2d.device)
        self._run('split3', 10 * M * K, 1, self.lib.tq_split3_bf16, x2d.data_ptr(), out.data_ptr(), M, K, _stream())
        return out

    def linear(self, a_ctr, w_ctr, bias, M, N, K, k_split, a_spec, w_spec, w_params, act_fn,
               out_spec, out_params, want_f32=True, want_ctr=False, tile_minmax=None):
        """tq_linear_qdq_bf16: [M, k_split*K] x [N, K]^T -> (y fp32 [M, N] | None, y_ctr bf16 | None)."""
        _chk_cuda(a_ctr, w_ctr, bias)
        dev = a_ctr.device
        y = torch.empty(M, N, dtype=torch.float32, device=dev) if want_f32 else None
        yc = torch.empty(M, N, dtype=torch.bfloat16, device=dev) if want_ctr else None
        null = QSpec(None, None, None, 8, 0, 1e-8)
        self._run('linear_qdq', 2 * M * N * K * int(k_split), 1, self.lib.tq_linear_qdq_bf16,
                  a_ctr.data_ptr(), w_ctr.data_ptr(), _ptr(bias), _ptr(y), _ptr(yc), M, N, K, int(k_split),
                  a_spec if a_spec is not None else null, w_spec if w_spec is not None else null, int(w_params),
                  int(act_fn), out_spec if out_spec is not None else null, int(out_params), _ptr(tile_minmax),
                  None, 0, _stream())
        return y, yc


    # -- fused encoder blocks -----------------------------------------------------------------------
    def linear_res(self, a_ctr, w_ctr, bias, M, N, K, a_spec, w_spec, w_params, out_spec, out_params,
                   res_ctr, res_spec, out2_spec, out2_params, out_ctr=None, want_f32=False):
        """tq_linear_res_qdq_bf16 -> (y fp32 | None, y_ctr bf16 [M, N])"""
        _chk_cuda(a_ctr, w_ctr, bias, res_ctr)
        dev = a_ctr.device
        y = torch.empty(M, N, dtype=torch.float32, device=dev) if want_f32 else None
        yc = out_ctr if out_ctr is not None else torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        self._run('linear_qdq', 2 * M * N * K, 1, self.lib.tq_linear_res_qdq_bf16, a_ctr.data_ptr(),
                  w_ctr.data_ptr(), _ptr(bias), _ptr(y), yc.data_ptr(), M, N, K, a_spec, w_spec, int(w_params),
                  out_spec, int(out_params), res_ctr.data_ptr(), res_spec, out2_spec, int(out2_params), _stream())
        return y, yc

    def linear_res_ln(self, a_ctr, w_ctr, bias, M, N, K, a_spec, w_spec, w_params, out_spec, res_ctr, res_spec,
                      out2_spec, gamma_q, beta, eps, ln_spec, out_ctr=None, want_f32=False):
        """tq_linear_res_ln_qdq_bf16: residual block with the LayerNorm fused in -> (z fp32 | None, z_ctr bf16).
        Raises TQError(TQ_EUNSUPPORTED) for shapes the cluster kernel does not cover."""
        _chk_cuda(a_ctr, w_ctr, bias, res_ctr, gamma_q, beta)
        dev = a_ctr.device
        z = torch.empty(M, N, dtype=torch.float32, device=dev) if want_f32 else None
        zc = out_ctr if out_ctr is not None else torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        self._run('linear_qdq', 2 * M * N * K, 1, self.lib.tq_linear_res_ln_qdq_bf16, a_ctr.data_ptr(),
                  w_ctr.data_ptr(), _ptr(bias), _ptr(z), zc.data_ptr(), M, N, K, a_spec, w_spec, int(w_params),
                  out_spec, res_ctr.data_ptr(), res_spec, out2_spec, gamma_q.data_ptr(), beta.data_ptr(), float(eps),
                  ln_spec, _stream())
        return z, zc

    def attention(self, qkv_ctr, B, T, H, head_dim, q_spec, k_spec, v_spec, s_spec, p_spec, c_spec, mask=None,
                  out_ctr=None):
        """tq_attention_qdq_bf16: qkv_ctr [B*T, 3*H*hd] -> context grid [B*T, H*hd] (bf16)"""
        _chk_cuda(qkv_ctr, mask)
        c = out_ctr if out_ctr is not None else torch.empty(B * T, H * head_dim, dtype=torch.bfloat16,
                                                            device=qkv_ctr.device)
        work = 4 * B * H * T * T * head_dim
        self._run('attention', work, 1, self.lib.tq_attention_qdq_bf16, qkv_ctr.data_ptr(), c.data_ptr(), B, T, H,
                  head_dim, q_spec, k_spec, v_spec, s_spec, p_spec, c_spec, _ptr(mask), _stream())
        return c

    def ln_qdq(self, x_ctr, in_spec, in_params, gamma_q, beta, eps, out_spec, out_params, out_ctr=None,
               want_f32=False):
        _chk_cuda(x_ctr, gamma_q, beta)
        M, D = x_ctr.shape
        o = out_ctr if out_ctr is not None else torch.empty(M, D, dtype=torch.bfloat16, device=x_ctr.device)
        f = torch.empty(M, D, dtype=torch.float32, device=x_ctr.device) if want_f32 else None
        self._run('ln_qdq', 4 * M * D, 1, self.lib.tq_ln_qdq_bf16, x_ctr.data_ptr(), in_spec, int(in_params),
                  gamma_q.data_ptr(), beta.data_ptr(), float(eps), out_spec, int(out_params), o.data_ptr(), _ptr(f),
                  M, D, _stream())
        return o, f

    def embed_ln_qdq(self, ids, type_ids, pos_ids, T, word_q, type_q, pos_q, e_tok, e_tok_params, e_pos,
                     e_pos_params, gamma_q, beta, eps, out_spec, out_params, out_ctr=None, want_f32=False):
        _chk_cuda(ids, word_q, type_q, pos_q, gamma_q, beta)
        M, D = ids.numel(), word_q.shape[1]
        o = out_ctr if out_ctr is not None else torch.empty(M, D, dtype=torch.bfloat16, device=ids.device)
        f = torch.empty(M, D, dtype=torch.float32, device=ids.device) if want_f32 else None
        self._run('embed_ln_qdq', 14 * M * D, 1, self.lib.tq_embed_ln_qdq_bf16, ids.data_ptr(), _ptr(type_ids),
                  _ptr(pos_ids), int(T), word_q.data_ptr(), type_q.data_ptr(), pos_q.data_ptr(), e_tok,
                  int(e_tok_params), e_pos, int(e_pos_params), gamma_q.data_ptr(), beta.data_ptr(), float(eps),
                  out_spec, int(out_params), o.data_ptr(), _ptr(f), M, D, _stream())
        return o, f

    # -- 8-bit integer operand mode (x_int bytes, int8 tensor cores) ---------------------------------
    def linear_i8(self, a_i8, w_i8, w_rowsum, bias, M, N, K, a_spec, w_spec, w_params, act_fn, out_spec, out_params,
                  want_f32=False, out_ctr=None, out_i8=None):
        """tq_linear_qdq_i8 -> (y fp32 | None); out_ctr (bf16 centred) / out_i8 (x_int bytes) are filled if given"""
        _chk_cuda(a_i8, w_i8, w_rowsum, bias)
        y = torch.empty(M, N, dtype=torch.float32, device=a_i8.device) if want_f32 else None
        null = QSpec(None, None, None, 8, 0, 1e-8)
        self._run('linear_qdq', 2 * M * N * K, 1, self.lib.tq_linear_qdq_i8, a_i8.data_ptr(), w_i8.data_ptr(),
                  w_rowsum.data_ptr(), _ptr(bias), _ptr(y), _ptr(out_ctr), _ptr(out_i8), M, N, K, a_spec, w_spec,
                  int(w_params), int(act_fn), out_spec if out_spec is not None else null, int(out_params), _stream())
        return y

    def linear_seg_i8(self, a_i8, w_i8, w_rowsum, bias, M, N, K, a_spec, w_seg_spec, out_seg_spec, nseg, act_fn,
                      out_ctr=None, out_i8=None, ldc=0):
        """tq_linear_seg_qdq_i8: per-segment weight / output quantizers (nseg slots each); fills out_ctr XOR out_i8
        (``ldc``: output row stride in elements when the output is a column block of a wider buffer)"""
        _chk_cuda(a_i8, w_i8, w_rowsum, bias, out_ctr, out_i8)
        self._run('linear_qdq', 2 * M * N * K, 1, self.lib.tq_linear_seg_qdq_i8, a_i8.data_ptr(), w_i8.data_ptr(),
                  w_rowsum.data_ptr(), _ptr(bias), _ptr(out_ctr), _ptr(out_i8), M, N, K, a_spec, w_seg_spec, out_seg_spec,
                  int(nseg), int(act_fn), int(ldc), _stream())
        return out_i8 if out_i8 is not None else out_ctr

    @staticmethod
    def chain_stage(kind, a_i8, w_i8, w_rowsum, bias, out, N, K, a_spec, w_spec, out_spec, nseg=1, res_i8=None, res_spec=None,
                    out2_spec=None, ln_spec=None, ln_gamma_q=None, ln_beta=None, ln_eps=0.0):
        """one struct tq_chain_stage; kind 0 plain segments (out bf16 centred grid), 1 GELU (bytes), 2 residual + LayerNorm"""
        _chk_cuda(a_i8, w_i8, w_rowsum, bias, out, res_i8, ln_gamma_q, ln_beta)
        st = ChainStage()
        st.a_i8, st.w_i8, st.w_rowsum, st.bias, st.out = a_i8.data_ptr(), w_i8.data_ptr(), w_rowsum.data_ptr(), _ptr(bias), out.data_ptr()
        st.N, st.K, st.nseg, st.kind = int(N), int(K), int(nseg), int(kind)
        st.a_q, st.w_q, st.out_q = a_spec, w_spec, out_spec
        if kind == 2:
            st.res_i8, st.res_q, st.out2_q, st.ln_q = res_i8.data_ptr(), res_spec, out2_spec, ln_spec
            st.ln_gamma_q, st.ln_beta, st.ln_eps = ln_gamma_q.data_ptr(), ln_beta.data_ptr(), float(ln_eps)
        st._flops = 2 * int(N) * int(K)                 # per row
        return st

    @staticmethod
    def chain_attention_stage(qkv_ctr, out_i8, hidden, heads, q_spec, k_spec, v_spec, s_spec, p_spec, c_spec, mask=None):
        """kind 3: attention over the Q | K | V buffer [M, 3 hidden] (bf16 centred grids) -> context bytes [M, hidden]"""
        _chk_cuda(qkv_ctr, out_i8, mask)
        st = ChainStage()
        st.a_i8, st.out, st.bias = qkv_ctr.data_ptr(), out_i8.data_ptr(), _ptr(mask)
        st.N, st.K, st.nseg, st.kind = int(hidden), int(heads), 1, 3
        st.a_q, st.w_q, st.res_q, st.out2_q, st.ln_q, st.out_q = q_spec, k_spec, v_spec, s_spec, p_spec, c_spec
        st._flops = 4 * 128 * int(hidden)               # per row: QK^T and PV over 128 keys
        return st

    def head_i8(self, x_i8, row_stride, B, D, L, wp_i8, wp_rowsum, bp, a_spec, wp_spec, pool_spec, wc_i8, wc_rowsum, bc, wc_spec,
                cls_spec, logits):
        """tq_head_qdq_i8: first token -> pooler (tanh, QDQ) -> classifier (QDQ) in one launch; logits [B, ldl] fp32"""
        _chk_cuda(x_i8, wp_i8, wp_rowsum, bp, wc_i8, wc_rowsum, bc, logits)
        self._run('linear_qdq', 2 * B * D * (D + L), 1, self.lib.tq_head_qdq_i8, x_i8.data_ptr(), int(row_stride), int(B), int(D), int(L),
                  wp_i8.data_ptr(), wp_rowsum.data_ptr(), _ptr(bp), a_spec, wp_spec, pool_spec, wc_i8.data_ptr(), wc_rowsum.data_ptr(),
                  _ptr(bc), wc_spec, cls_spec, logits.data_ptr(), logits.shape[1], _stream())
        return logits

    def chain_plan(self, stages, M):
        """tq_chain_plan_create: descriptors of a stage list in device memory (fixed buffers); run with chain_run"""
        return ChainPlan(self, stages, M)

    def chain_run(self, plan):
        """tq_chain_plan_run: every stage of the plan in ONE launch (a cluster per 128-row panel)"""
        self._run('chain', plan.flops, 1, self.lib.tq_chain_plan_run, plan.handle, _stream())

    def linear_nonorm_i8(self, a_i8, w_i8, w_rowsum, bias, M, N, K, a_spec, w_spec, out_spec, res_i8, res_spec, out2_spec,
                         nn_weight_q, nn_bias_q, nn_spec, out_i8, ldc=0):
        """tq_linear_nonorm_qdq_i8: dense -> QDQ [-> + residual -> QDQ] -> NoNorm -> QDQ in one GEMM epilogue"""
        _chk_cuda(a_i8, w_i8, w_rowsum, bias, res_i8, nn_weight_q, nn_bias_q, out_i8)
        null = QSpec(None, None, None, 8, 0, 1e-8)
        self._run('linear_qdq', 2 * M * N * K, 1, self.lib.tq_linear_nonorm_qdq_i8, a_i8.data_ptr(), w_i8.data_ptr(),
                  w_rowsum.data_ptr(), _ptr(bias), out_i8.data_ptr(), M, N, K, a_spec, w_spec, out_spec, _ptr(res_i8),
                  res_spec if res_spec is not None else null, out2_spec if out2_spec is not None else null,
                  nn_weight_q.data_ptr(), nn_bias_q.data_ptr(), nn_spec, int(ldc), _stream())
        return out_i8

    def linear_peg_i8(self, a_i8, w_i8, w_grp_rowsum, bias, M, N, K, a_spec, a_groups, w_spec, w_params, out_spec, out_params,
                      seg_width, act_fn, out_ctr=None, out_i8=None):
        """tq_linear_peg_qdq_i8: per-group A operand, per-segment output quantizers; fills out_ctr XOR out_i8"""
        _chk_cuda(a_i8, w_i8, w_grp_rowsum, bias, out_ctr, out_i8)
        self._run('linear_qdq', 2 * M * N * K, 1, self.lib.tq_linear_peg_qdq_i8, a_i8.data_ptr(), w_i8.data_ptr(),
                  w_grp_rowsum.data_ptr(), _ptr(bias), _ptr(out_ctr), _ptr(out_i8), M, N, K, a_spec, int(a_groups), w_spec,
                  int(w_params), out_spec, int(out_params), int(seg_width), int(act_fn), _stream())
        return out_i8 if out_i8 is not None else out_ctr

    def linear_peg_res_ln_i8(self, a_i8, w_i8, w_grp_rowsum, bias, M, N, K, a_spec, a_groups, w_spec, w_params, out_spec,
                             out_params, res_i8, res_spec, res_params, out2_spec, out2_params, gamma_q, beta, eps, ln_spec,
                             ln_params, seg_width, out_i8, out_ctr=None):
        _chk_cuda(a_i8, w_i8, w_grp_rowsum, bias, res_i8, gamma_q, beta, out_i8, out_ctr)
        self._run('linear_qdq', 2 * M * N * K, 1, self.lib.tq_linear_peg_res_ln_qdq_i8, a_i8.data_ptr(), w_i8.data_ptr(),
                  w_grp_rowsum.data_ptr(), _ptr(bias), _ptr(out_ctr), out_i8.data_ptr(), M, N, K, a_spec, int(a_groups),
                  w_spec, int(w_params), out_spec, int(out_params), res_i8.data_ptr(), res_spec, int(res_params), out2_spec,
                  int(out2_params), gamma_q.data_ptr(), beta.data_ptr(), float(eps), ln_spec, int(ln_params), int(seg_width),
                  _stream())
        return out_i8

    def attention_pad_i8(self, qkv_ctr, B, T, H, head_dim, true_head_dim, q_spec, k_spec, v_spec, s_spec, p_spec, c_spec, mask,
                         out_i8):
        """tq_attention_pad_qdq_i8: heads of `true_head_dim` < 64 dims in zero-padded 64-column slots"""
        _chk_cuda(qkv_ctr, mask, out_i8)
        self._run('attention', 4 * B * H * T * T * true_head_dim, 1, self.lib.tq_attention_pad_qdq_i8, qkv_ctr.data_ptr(),
                  out_i8.data_ptr(), B, T, H, head_dim, true_head_dim, q_spec, k_spec, v_spec, s_spec, p_spec, c_spec,
                  _ptr(mask), _stream())
        return out_i8

    def attention_peg_i8(self, qkv_ctr, B, T, H, head_dim, q_spec, k_spec, v_spec, qkv_params, s_spec, p_spec, c_spec, c_params,
                         mask, out_i8):
        _chk_cuda(qkv_ctr, mask, out_i8)
        self._run('attention', 4 * B * H * T * T * head_dim, 1, self.lib.tq_attention_peg_qdq_i8, qkv_ctr.data_ptr(),
                  out_i8.data_ptr(), B, T, H, head_dim, q_spec, k_spec, v_spec, int(qkv_params), s_spec, p_spec, c_spec,
                  int(c_params), _ptr(mask), _stream())
        return out_i8

    def linear_res_ln_i8(self, a_i8, w_i8, w_rowsum, bias, M, N, K, a_spec, w_spec, w_params, out_spec, res_i8, res_spec,
                         out2_spec, gamma_q, beta, eps, ln_spec, out_i8, want_f32=False, out_ctr=None):
        _chk_cuda(a_i8, w_i8, w_rowsum, bias, res_i8, gamma_q, beta, out_i8, out_ctr)
        z = torch.empty(M, N, dtype=torch.float32, device=a_i8.device) if want_f32 else None
        self._run('linear_qdq', 2 * M * N * K, 1, self.lib.tq_linear_res_ln_qdq_i8, a_i8.data_ptr(), w_i8.data_ptr(),
                  w_rowsum.data_ptr(), _ptr(bias), _ptr(z), _ptr(out_ctr), out_i8.data_ptr(), M, N, K, a_spec, w_spec, int(w_params),
                  out_spec, res_i8.data_ptr(), res_spec, out2_spec, gamma_q.data_ptr(), beta.data_ptr(), float(eps),
                  ln_spec, _stream())
        return z

    def linear_bf16_o8(self, a_ctr, w_ctr, bias, M, N, K, a_spec, w_spec, w_params, act_fn, out_spec, out_params, out_i8):
        """tq_linear_qdq_bf16_o8: bf16 centred operands, x_int byte output"""
        _chk_cuda(a_ctr, w_ctr, bias, out_i8)
        self._run('linear_qdq', 2 * M * N * K, 1, self.lib.tq_linear_qdq_bf16_o8, a_ctr.data_ptr(), w_ctr.data_ptr(),
                  _ptr(bias), out_i8.data_ptr(), M, N, K, a_spec, w_spec, int(w_params), int(act_fn), out_spec,
                  int(out_params), _stream())
        return out_i8

    def attention_i8(self, qkv_ctr, B, T, H, head_dim, q_spec, k_spec, v_spec, s_spec, p_spec, c_spec, mask, out_i8):
        """tq_attention_qdq_i8: bf16 centred q|k|v in, context x_int bytes out"""
        _chk_cuda(qkv_ctr, mask, out_i8)
        self._run('attention', 4 * B * H * T * T * head_dim, 1, self.lib.tq_attention_qdq_i8, qkv_ctr.data_ptr(),
                  out_i8.data_ptr(), B, T, H, head_dim, q_spec, k_spec, v_spec, s_spec, p_spec, c_spec, _ptr(mask),
                  _stream())
        return out_i8

    def embed_ln_qdq_i8(self, ids, type_ids, pos_ids, T, word_q, type_q, pos_q, e_tok, e_tok_params, e_pos,
                        e_pos_params, gamma_q, beta, eps, out_spec, out_params, out_i8):
        _chk_cuda(ids, word_q, type_q, pos_q, gamma_q, beta, out_i8)
        M, D = ids.numel(), word_q.shape[1]
        self._run('embed_ln_qdq', 14 * M * D, 1, self.lib.tq_embed_ln_qdq_i8, ids.data_ptr(), _ptr(type_ids),
                  _ptr(pos_ids), int(T), word_q.data_ptr(), type_q.data_ptr(), pos_q.data_ptr(), e_tok,
                  int(e_tok_params), e_pos, int(e_pos_params), gamma_q.data_ptr(), beta.data_ptr(), float(eps),
                  out_spec, int(out_params), out_i8.data_ptr(), M, D, _stream())
        return out_i8


_OPS = None


def default_device():
    """Device for quantizer state created from python floats (before any tensor was seen)."""
    return torch.device('cuda', torch.cuda.current_device())


def ops():
    """The process-wide CUDA back-end.  Raises (never falls back) if it cannot be created."""
    global _OPS
    if _OPS is None:
        if not torch.cuda.is_available():
            raise TQError('tq_b200: no CUDA device visible -- the fake-quantization kernels are '
                          'sm_100a only and there is no CPU fallback')
        _OPS = CudaOps()
    return _OPS
