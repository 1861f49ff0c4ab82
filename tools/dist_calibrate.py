#!/usr/bin/env python
"""Config 5 smoke on real GPUs: RoBERTa-style encoder, MSE (grid) activation range estimation, calibration
batches sharded over the ranks, statistics all-reduced over NCCL (quantization/_dist.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/dist_calibrate.py

Checks: every rank ends with identical quantizer parameters, and they equal single-process calibration of
the concatenated batch (computed on rank 0 afterwards with the collective disabled) -- exactly for
min/max estimators, to fp64 round-off for the MSE loss sums.  Prints one JSON line from rank 0.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))

from engine.bert import BertConfig, QuantBertForSequenceClassification  # noqa: E402
from quantization.quantizers import QMethods  # noqa: E402
from quantization.range_estimators import RangeEstimators, OptMethod  # noqa: E402


def build(device, est, opts):
    cfg = BertConfig(vocab_size=5000, hidden_size=256, num_hidden_layers=2, num_attention_heads=4,
                     intermediate_size=1024, max_position_embeddings=130, type_vocab_size=1, pad_token_id=1,
                     roberta_positions=True)
    m = QuantBertForSequenceClassification(cfg, method=QMethods.symmetric_uniform,
                                           act_method=QMethods.asymmetric_uniform, n_bits=8, n_bits_act=8,
                                           act_range_method=est, act_range_options=opts)
    m.init_weights(seed=0, std=0.05)
    m.to(device).eval()
    m.set_quant_state(True, True)
    return m


def params(m):
    return torch.cat([q.quantizer._delta.reshape(-1).float() for q in m.act_quantizers()])


def main():
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    g = torch.Generator().manual_seed(7)
    B = 8 * world
    batches = [torch.randint(2, 5000, (B, 128), generator=g) for _ in range(2)]      # global batches
    out = {}
    for name, est, opts in [('running_minmax', RangeEstimators.running_minmax, {}),
                            ('mse_grid', RangeEstimators.MSE, dict(opt_method=OptMethod.grid, num_candidates=100))]:
        os.environ['TQ_DIST_CALIBRATION'] = '1'
        m = build(dev, est, opts)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            for b in batches:
                shard = b[rank * 8:(rank + 1) * 8].to(dev)
                m(shard, torch.ones_like(shard))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        mine = params(m)
        if world > 1:
            gathered = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(gathered, mine)
            same = all(torch.equal(gathered[0], t) for t in gathered)
        else:
            same = True
        res = {'ranks_identical': bool(same), 'calib_s': dt, 'tokens_per_s': 2 * B * 128 / dt}
        if rank == 0:
            os.environ['TQ_DIST_CALIBRATION'] = '0'
            ref = build(dev, est, opts)
            with torch.no_grad():
                for b in batches:
                    full = b.to(dev)
                    ref(full, torch.ones_like(full))
            r = params(ref)
            res['max_rel_diff_vs_single_process'] = float(((mine - r).abs() / r.abs().clamp_min(1e-12)).max())
        out[name] = res
        if world > 1:
            dist.barrier()
    if rank == 0:
        ok = all(v['ranks_identical'] and v['max_rel_diff_vs_single_process'] < (1e-6 if k == 'running_minmax' else 2e-2)
                 for k, v in out.items())
        print(json.dumps({'world': world, 'ok': bool(ok), **out}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
