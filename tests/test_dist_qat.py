"""Data-parallel quantization-aware training: 2 ranks over gloo (CPU, oracle arithmetic back-end injected in
every rank), one QuantLinear with learnable ranges under DistributedDataParallel.  The learnable ranges are
ordinary nn.Parameters whose gradients come out of FakeQuantSTE.backward (tq_qdq_bwd_f32 semantics), so DDP's
gradient all-reduce covers them: the averaged 2-rank gradients equal the single-process gradients on the whole
batch (mean loss).  On GPUs the same code runs over NCCL / NVLink; there is no quantizer-specific collective
in training (calibration has one, tests/test_dist_calibration.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, PKG
from test_dist_calibration import _free_port, _setup_backend


def _make_layer():
    from quantization.autoquant_utils import QuantLinear
    from quantization.quantizers import QMethods
    torch.manual_seed(3)
    lin = QuantLinear(64, 48, method=QMethods.symmetric_uniform, act_method=QMethods.asymmetric_uniform, n_bits=4,
                      n_bits_act=8)
    lin.quantized()
    lin.eval()
    rs = np.random.RandomState(5)
    x = torch.from_numpy(rs.randn(16, 8, 64).astype(np.float32))
    with torch.no_grad():
        lin(x)                       # calibrate weight and activation ranges on the whole batch
    lin.learn_ranges()
    lin.train()
    return lin, x


def _grads(module, x):
    module.zero_grad()
    coef = torch.linspace(-1, 1, 48)
    loss = (module(x) * coef).sum(dim=(1, 2)).mean()
    loss.backward()
    return {n: p.grad.detach().numpy().copy() for n, p in module.named_parameters()}


def _worker(rank, world, port, q):
    _setup_backend()
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lin, x = _make_layer()
        ddp = torch.nn.parallel.DistributedDataParallel(lin)
        shard = x[rank * 8:(rank + 1) * 8]
        g = _grads(ddp, shard)
        q.put((rank, {k.replace('module.', ''): v for k, v in g.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_qat_gradients_equal_single_process():
    _setup_backend()
    lin, x = _make_layer()
    ref = _grads(lin, x)
    assert {'weight', 'bias', 'weight_quantizer.quantizer._delta', 'activation_quantizer.quantizer._delta',
            'activation_quantizer.quantizer._zero_float'} <= set(ref)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        assert set(got[rank]) == set(ref)
        for name, g in ref.items():
            scale = np.abs(g).max() + 1e-12
            assert np.abs(got[rank][name] - g).max() <= 1e-5 * scale + 1e-7, f'rank {rank} {name}'
    import tq_native
    tq_native._OPS = None
