"""N-rank calibration == single-process calibration on the concatenated batch.

world_size 2 over gloo on CPU (127.0.0.1), oracle arithmetic back-end injected in every rank:
exercises quantization/_dist.py -- the package's only collective call site (NCCL on GPUs).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, PKG


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup_backend():
    for p in (ROOT, PKG, os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import tq_native
    from oracle_backend import OracleOps
    tq_native._OPS = OracleOps()
    tq_native.default_device = lambda: torch.device('cpu')


def _batches():
    rs = np.random.RandomState(77)
    return [torch.from_numpy((rs.randn(8, 16, 96) * (1 + i)).astype(np.float32)) for i in range(3)]


def _make_managers():
    from quantization.quantization_manager import QuantizationManager
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators, OptMethod
    A, S, E = QMethods.asymmetric_uniform, QMethods.symmetric_uniform, RangeEstimators
    return {
        'running': QuantizationManager(A, E.running_minmax, qparams=dict(n_bits=8)),
        'current_sym': QuantizationManager(S, E.current_minmax, qparams=dict(n_bits=8)),
        'all': QuantizationManager(A, E.allminmax, qparams=dict(n_bits=8)),
        'peg': QuantizationManager(A, E.running_minmax, axis=2, n_groups=6, qparams=dict(n_bits=8)),
        'mse1d': QuantizationManager(S, E.MSE, qparams=dict(n_bits=8),
                                     init_params=dict(opt_method=OptMethod.grid, num_candidates=20)),
    }


def _calibrate(mgrs, batches, shard=None):
    out = {}
    for name, m in mgrs.items():
        for b in batches:
            x = b if shard is None else b[shard]
            m(x)
        q = m.quantizer
        out[name] = q._delta.detach().numpy().reshape(-1).copy()
    return out


def _worker(rank, world, port, q):
    _setup_backend()
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        B = _batches()
        shard = slice(rank * 4, rank * 4 + 4)           # batch dim 8 split over 2 ranks
        from quantization import _dist
        local = _calibrate(_make_managers(), B, shard)  # outside calibration_sync(): rank-local, no collective
        with _dist.calibration_sync():
            res = _calibrate(_make_managers(), B, shard)
        q.put((rank, res, local, _dist.stats()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_calibration_equals_single_process():
    _setup_backend()
    ref = _calibrate(_make_managers(), _batches())
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    raw = [q.get(timeout=240) for _ in range(2)]
    got = {r[0]: r[1] for r in raw}
    local = {r[0]: r[2] for r in raw}
    # opt-in: without the context the estimators never reduce (the two shards give different ranges) ...
    assert not np.array_equal(local[0]['running'], local[1]['running'])
    # ... and inside it every estimator update issued collectives
    assert all(r[3]['calls'] > 0 and r[3]['bytes'] > 0 for r in raw)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        for name, d in ref.items():
            if name == 'mse1d':       # fp64 loss sums are associative only up to rounding: same argmin expected
                np.testing.assert_allclose(got[rank][name], d, rtol=1e-6)
            else:
                assert np.array_equal(got[rank][name], d), f'rank {rank} {name}'
    import tq_native
    tq_native._OPS = None
