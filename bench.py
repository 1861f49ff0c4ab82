#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 fake-quantization forward path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): BERT-base, W8 symmetric per-tensor weights (current_minmax),
A8 asymmetric per-tensor activations (running_minmax, one calibration batch, then fix_ranges), eval
forward, batch 32 x seq 128 per GPU, synthetic token ids, random-init weights (seed 0).  A "step" is
one forward over one batch.  N GPUs = N independent replicas (weak scaling, no collective on the
inference path).

Our arm prints ONE JSON line with
  value       tokens/s, inputs resident in HBM, the captured CUDA graph replayed K times
  e2e         tokens/s through the public call with HOST token ids (pinned) -> logits on the host,
              H2D and D2H inside the timed region, one host sync per step
  roofline    the dominant kernel of the step, timed live with CUDA events in an eager pass
  cpu_baseline  oracle port (oracle/bert_oracle.py) on the host cores, bounded sample (rank 0, N=1)
  qdq_standalone  second half of BASELINE's metric: the standalone quant-dequant kernel on a 1 GiB fp32 tensor, GB/s at
              8 B / element, with `cpu_baseline` = the reference's six-pass chain on the host cores (oracle port)
  qat_backward_standalone  tq_qdq_bwd_f32 (straight-through backward + range gradients) on 512 MiB tensors, 12 B / element
The reference arm times the same oracle port -- the reference's algorithm on the host CPU with all
host threads -- on the same config and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'transformer-quantization_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

BATCH, SEQ = 32, 128
METRIC = 'tokens/sec W8A8 BERT-base seq128 b32'
WORKLOAD = ('BERT-base W8A8 per-tensor asymmetric (W8 sym current_minmax / A8 asym running_minmax, '
            'fixed ranges), seq 128 batch 32 per GPU, eval forward')
QDQ_BYTES_PER_TOKEN = 1.3456e6     # algorithmic QDQ traffic per token


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return d['hbm_gbs'], d['bf16_tflops_sustained'], 'measured'
    return 6650.0, 1400.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': mx or None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def dist_env():
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    return rank, world, local


# --------------------------------------------------------------------------------------------------
def build_model(device, seed=0):
    from engine.bert import BertConfig, QuantBertForSequenceClassification
    from quantization.quantizers import QMethods
    from quantization.range_estimators import RangeEstimators
    model = QuantBertForSequenceClassification(
        BertConfig(), method=QMethods.symmetric_uniform, act_method=QMethods.asymmetric_uniform,
        n_bits=8, n_bits_act=8, weight_range_method=RangeEstimators.current_minmax,
        act_range_method=RangeEstimators.running_minmax)
    model.init_weights(seed=seed)
    model.to(device).eval()
    model.set_quant_state(weight_quant=True, act_quant=True)
    return model


def synthetic_ids(seed, n_batches=1):
    g = torch.Generator().manual_seed(seed)
    return [torch.randint(0, 30522, (BATCH, SEQ), generator=g) for _ in range(n_batches)]


def usable_cpus():
    """hardware threads this process may actually use (affinity mask and cgroup CPU quota)"""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    try:
        quota, period = open('/sys/fs/cgroup/cpu.max').read().split()
        if quota != 'max':
            n = max(1, min(n, int(float(quota) / float(period))))
    except (OSError, ValueError):
        pass
    return n


def host_threads():
    """threads the CPU arm uses: every hardware thread the host grants this process"""
    n = usable_cpus()
    torch.set_num_threads(n)
    return n


def cpu_forward_setup(sample_batch=BATCH):
    from oracle.bert_oracle import OracleBert, random_bert_state_dict
    sd = random_bert_state_dict(seed=0)
    m = OracleBert(sd, n_layers=12, n_heads=12, n_bits=8, n_bits_act=8, sym_acts=False)
    ids = synthetic_ids(1234)[0]
    mask = torch.ones_like(ids)
    with torch.no_grad():
        t0 = time.perf_counter()
        m(ids, mask)                 # calibration forward on the full batch (also caches the weights)
        t_cal = time.perf_counter() - t0
        m.fix_ranges()
    return m, ids, mask, t_cal


def cpu_baseline(threads, budget_s=20.0):
    """oracle port of the reference path on the host cores: calibrate on the full batch, fix the
    ranges, then time fixed-range forwards of a bounded sample (the first rows of the same batch)."""
    m, ids, mask, t_cal = cpu_forward_setup()
    rows = BATCH
    while rows > 1 and 2 * t_cal * rows / BATCH > budget_s:
        rows //= 2
    with torch.no_grad():
        ts = []
        for _ in range(2):
            t0 = time.perf_counter()
            logits = m(ids[:rows], mask[:rows])
            ts.append(time.perf_counter() - t0)
    return rows * SEQ / statistics.median(ts), rows, logits


def cpu_qdq_baseline(threads, n=16 * 1024 * 1024, passes=5):
    """the reference's quantize-dequantize chain on the host cores (oracle.bert_oracle.Site: one torch CPU op per
    step of quantizers.py:142-153, 184-185, 209 -- six passes + temporaries), fixed range, on a 64 MiB fp32 tensor:
    algorithmic GB/s at 8 B / element, next to `qdq_standalone`.  None if it cannot run."""
    try:
        from oracle.bert_oracle import Site
        x = torch.randn(n, generator=torch.Generator().manual_seed(1234)) * 3
        site = Site(n_bits=8, symmetric=False, estimator='current_minmax')
        with torch.no_grad():
            site(x)                              # estimates the range
            site.fixed = True
            site(x)                              # warm-up
            ts = []
            for _ in range(passes):
                t0 = time.perf_counter()
                site(x)
                ts.append(time.perf_counter() - t0)
        return {'gbs': 8.0 * n / statistics.median(ts) / 1e9, 'cores': threads, 'kind': 'port',
                'sample': f'{passes} fixed-range passes over {n} fp32 elements (oracle/bert_oracle.py Site: the '
                          'reference op chain on torch CPU)'}
    except Exception as e:                                  # noqa: BLE001
        print(f'bench.py: cpu_qdq_baseline failed: {e!r}', file=sys.stderr)
        return None


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    threads = host_threads()
    m, ids, mask, t_cal = cpu_forward_setup()
    rows = BATCH                     # bounded sample: keep the whole run within a few minutes
    while rows > 1 and (args.steps + args.warmup) * t_cal * rows / BATCH > 150.0:
        rows //= 2
    with torch.no_grad():
        for _ in range(args.warmup):
            m(ids[:rows], mask[:rows])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            m(ids[:rows], mask[:rows])
        dt = time.perf_counter() - t0
    # the CPU path does not shard: N "GPUs" of the reference arm are still one host
    val = args.steps * rows * SEQ / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'tokens/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'global_batch': BATCH, 'seq_len': SEQ, 'parallelism': 'host-cpu'},
        'cpu_baseline': {'value': val, 'unit': 'tokens/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{args.steps} fixed-range forwards of the first {rows} of the 32 sequences '
                                   '(T=128) after one full-batch calibration forward (oracle/bert_oracle.py: '
                                   'the reference op chain on torch CPU)'},
        'e2e': {'value': val, 'unit': 'tokens/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
def profile_graphs(forward, ids_dev, mask_dev, ops, reps=20):
    """Per-kernel-class device time of one step UNDER THE SAME CONDITIONS AS THE TIMED REGION: the
    calls of one forward are recorded (the fused engine's buffers are persistent, so the pointers
    stay valid), every class of kernel is re-captured into its own CUDA graph and that graph is
    replayed ``reps`` times between two CUDA events.  Class times add up to ~ the step time."""
    calls = []
    orig = ops._run

    def record(name, work, kernels, fn, *args):
        calls.append((name, work, fn, args))
        return orig(name, work, kernels, fn, *args)

    ops._run = record
    try:
        with torch.no_grad():
            forward(ids_dev, mask_dev)
    finally:
        ops._run = orig
    torch.cuda.synchronize()
    out = {}
    import tq_native
    for cls in sorted({c[0] for c in calls}):
        mine = [c for c in calls if c[0] == cls]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            st = tq_native._stream()
            for _, _, fn, args in mine:
                rc = fn(*(args[:-1] + (st,)))
                assert rc == 0, rc
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        out[cls] = {'seconds': e0.elapsed_time(e1) * 1e-3 / reps, 'work': float(sum(c[1] for c in mine)),
                    'launches': len(mine)}
    return out


def profile_live(model, ids_dev, mask_dev, ops, reps=10):
    """Per-kernel device time of one step, measured live with CUDA events on the launching stream.
    During one eager forward every call of this library is, right after it ran, re-issued ``reps``
    times back to back between two events (same arguments: the kernels are idempotent and their
    tensors are still alive), so the average is kernel time rather than Python launch overhead.
    (Events cannot be placed inside the captured graph.)"""
    orig_run = ops._run
    agg = {}

    def timed(name, work, kernels, fn, *args):
        orig_run(name, work, kernels, fn, *args)            # the real call of the forward
        for _ in range(2):
            fn(*args)                                       # idempotent re-issue (same inputs/outputs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn(*args)
        e1.record()
        a = agg.setdefault(name, [[], 0.0, 0])
        a[0].append((e0, e1))
        a[1] += work
        a[2] += 1

    ops._run = timed
    try:
        with torch.no_grad():
            model(ids_dev, mask_dev)
        torch.cuda.synchronize()
    finally:
        ops._run = orig_run
    out = {}
    for k, (evs, work, n) in agg.items():
        secs = sum(a.elapsed_time(b) for a, b in evs) * 1e-3 / reps
        out[k] = {'seconds': secs, 'work': work, 'launches': n}
    return out


def qdq_hbm_probe(ops, n=256 * 1024 * 1024, iters=10):
    """standalone quant-dequant on a 1 GiB fp32 tensor (> L2): achieved HBM GB/s (8 B / element)"""
    x = torch.randn(n, device='cuda')
    y = torch.empty_like(x)
    mm = ops.minmax(x)
    d, z = torch.empty(1, device='cuda'), torch.empty(1, device='cuda')
    ops.set_range_asym(mm[0:1], mm[1:2], 8, 1e-8, False, d, z)
    spec = ops.spec(d, z, None, 8)
    for _ in range(3):
        ops.qdq(x, spec, out=y)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.qdq(x, spec, out=y)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    del x, y
    return 8.0 * n / statistics.median(ts) / 1e9


def qdq_bwd_probe(ops, n=128 * 1024 * 1024, iters=10):
    """straight-through backward with learnable-range gradients (tq_qdq_bwd_f32) on 512 MiB fp32 tensors
    (> L2): achieved HBM GB/s at 12 B / element (x read, grad_y read, grad_x written).  None if it cannot run:
    a side measurement must not take the headline line down."""
    try:
        x = torch.randn(n, device='cuda') * 3
        g = torch.randn(n, device='cuda')
        d, z = torch.full((1,), 0.03, device='cuda'), torch.full((1,), 120.3, device='cuda')
        spec = ops.spec(d, z, None, 8)
        for _ in range(3):
            ops.qdq_bwd(x, g, spec, 1)
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.qdq_bwd(x, g, spec, 1)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        del x, g
        return 12.0 * n / statistics.median(ts) / 1e9
    except Exception as e:                                  # noqa: BLE001
        print(f'bench.py: qdq_bwd_probe failed: {e!r}', file=sys.stderr)
        return None


def run_ours(args):
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the product path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        os.environ['NCCL_DEBUG'] = os.environ.get('TQ_NCCL_DEBUG', 'WARN')   # keep stdout to the one JSON line
        dist.init_process_group('nccl', device_id=dev)
    import tq_native
    ops = tq_native.ops()
    hbm_peak, tf_peak, peak_kind = peaks()

    model = build_model(dev)
    ids_host = synthetic_ids(1234 + rank)[0].pin_memory()
    mask_dev = torch.ones(BATCH, SEQ, dtype=torch.int64, device=dev)
    ids_dev = ids_host.to(dev)
    with torch.no_grad():
        os.environ['TQ_DIST_CALIBRATION'] = '0'       # replicas calibrate on their own batch
        model(ids_dev, mask_dev)                      # calibration batch: ranges + weight cache
        model.fix_ranges()
        for _ in range(2):
            model(ids_dev, mask_dev)                  # eager warm-up (function attributes, caches)
        torch.cuda.synchronize()

        # ---- the public forward: fused engine built from the calibrated model (module path if the
        #      configuration is outside the engine's support) ----
        from engine.fused import FusedBertEngine, UnsupportedByEngine
        engine_kind = ('fused (engine/fused.py: 5 kernels per encoder layer -- QKV GEMM, attention, attention-out GEMM + residual + '
                       'LayerNorm, FFN-in GEMM + GELU, FFN-out GEMM + residual + LayerNorm -- bf16 integer-grid carriers)')
        try:
            forward = FusedBertEngine(model, BATCH, SEQ)
        except UnsupportedByEngine as e:
            forward = model
            engine_kind = f'module path (one kernel per quantizer site): {e}'
        if os.environ.get('TQ_BENCH_MODULE_PATH') == '1':
            forward, engine_kind = model, 'module path (forced by TQ_BENCH_MODULE_PATH=1)'
        for _ in range(2):
            forward(ids_dev, mask_dev)
        torch.cuda.synchronize()

        # ---- capture the fixed-range eval forward into a CUDA graph ----
        static_ids = ids_dev.clone()
        l0 = ops.launches
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_logits = forward(static_ids, mask_dev)
        launches_per_step = ops.launches - l0
        graph.replay()
        torch.cuda.synchronize()
        module_logits = model(ids_dev, mask_dev)       # one kernel per site: cross-check of the engine
        engine_vs_module = float((static_logits - module_logits).abs().max())
        logits_host = torch.empty(static_logits.shape, dtype=static_logits.dtype).pin_memory()

        def barrier():
            if world > 1:
                torch.distributed.barrier()
            torch.cuda.synchronize()

        # ---- device-resident throughput ----
        for _ in range(max(args.warmup, 3)):
            graph.replay()
        barrier()
        with ClockSampler(local) as clk:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                graph.replay()
            e1.record()
            barrier()
        t_dev = e0.elapsed_time(e1) * 1e-3

        # ---- end to end: host ids in, host logits out, every step ----
        for _ in range(3):
            static_ids.copy_(ids_host, non_blocking=True)
            graph.replay()
            logits_host.copy_(static_logits, non_blocking=True)
            torch.cuda.synchronize()
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for _ in range(args.steps):
            static_ids.copy_(ids_host, non_blocking=True)
            graph.replay()
            logits_host.copy_(static_logits, non_blocking=True)
            torch.cuda.synchronize()
        e3.record()
        barrier()
        t_e2e = e2.elapsed_time(e3) * 1e-3

        if world > 1:
            t = torch.tensor([t_dev, t_e2e], device=dev, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            t_dev, t_e2e = t.tolist()

        if rank != 0:
            if world > 1:
                torch.distributed.barrier()
                torch.distributed.destroy_process_group()
            return

        # ---- roofline of the dominant kernel (eager pass, events on the launching stream) ----
        torch.cuda.synchronize()
        prof = (profile_graphs if forward is not model else profile_live)(forward, ids_dev, mask_dev, ops)
        qdq_gbs = qdq_hbm_probe(ops)
        bwd_gbs = qdq_bwd_probe(ops)

    tokens = BATCH * SEQ * world
    value = tokens * args.steps / t_dev
    e2e = tokens * args.steps / t_e2e
    top = max(prof.items(), key=lambda kv: kv[1]['seconds'])
    name, st = top
    per_launch_s = st['seconds'] / st['launches']
    flops_kernels = ('linear_qdq', 'attention')
    if name in flops_kernels:
        roof = {'kernel': 'tq_linear_qdq_bf16 / tq_linear_res_ln_qdq_bf16 (tcgen05 GEMM + fused QDQ / residual / LayerNorm epilogue)'
                if name == 'linear_qdq' else 'tq_attention_qdq_bf16', 'bound': 'tensor',
                'achieved': st['work'] / st['seconds'] / 1e12, 'peak': tf_peak, 'unit': 'TFLOP/s'}
    else:
        roof = {'kernel': name, 'bound': 'hbm', 'achieved': st['work'] / st['seconds'] / 1e9,
                'peak': hbm_peak, 'unit': 'GB/s'}
    roof['frac'] = roof['achieved'] / roof['peak']
    roof['traffic'] = None
    try:                      # DRAM bytes per launch of that kernel class from the committed ncu --set full capture
        with open(os.path.join(ROOT, 'profiles', 'r1_ncu_traffic.json')) as f:
            tr = json.load(f)
        if name in tr:
            roof['traffic'] = tr[name]['dram_bytes_per_launch']
            roof['traffic_source'] = tr[name]['source']
    except (OSError, ValueError, KeyError):
        pass
    roof['peak_source'] = f'{peak_kind} (MEASURED_PEAKS.json)' if peak_kind == 'measured' else 'fallback (B200_PROFILING.md)'
    roof['launches_per_step'] = st['launches']
    roof['avg_launch_us'] = per_launch_s * 1e6
    roof['share_of_library_kernel_time'] = st['seconds'] / sum(v['seconds'] for v in prof.values())

    cpu = cpu_qdq = None
    if world == 1:
        threads = host_threads()
        cpu_val, rows, cpu_logits = cpu_baseline(threads)
        cpu = {'value': cpu_val, 'unit': 'tokens/s', 'cores': threads, 'kind': 'port',
               'sample': f'1 full-batch calibration forward + 2 timed fixed-range forwards of the first {rows} of '
                         'the 32 sequences (oracle/bert_oracle.py, torch CPU, reference op chain)',
               'logit_max_abs_diff_vs_gpu': float((cpu_logits - static_logits[:rows].float().cpu()).abs().max())}
        cpu_qdq = cpu_qdq_baseline(threads)

    line = {
        'metric': METRIC, 'value': value, 'unit': 'tokens/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': t_dev / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16 integer-grid operands, fp32 accumulate / fp32 QDQ',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'global_batch': BATCH * world, 'seq_len': SEQ,
                   'parallelism': f'dp{world} (independent replicas)',
                   'l2': 'per-step working set (0.17 GB bf16 weight grids + 0.09 GB fp32 embedding tables, 12 x 63 MB bf16 '
                         'activations) exceeds the 126 MB L2; nothing is flushed between steps',
                   'cuda_graph': True, 'forward': engine_kind,
                   'max_abs_logit_diff_engine_vs_module_path': engine_vs_module},
        'e2e': {'value': e2e, 'unit': 'tokens/s', 'h2d_bytes_per_step': ids_host.numel() * ids_host.element_size(),
                'd2h_bytes_per_step': logits_host.numel() * logits_host.element_size(),
                'ms_per_step': t_e2e / args.steps * 1e3},
        'gpu_launches': launches_per_step * args.steps,
        'gpu_launches_per_step': launches_per_step,
        'clocks': clk.summary(),
        'roofline': roof,
        'cpu_baseline': cpu,
        'memory_roofline': {'qdq_bytes_per_token': QDQ_BYTES_PER_TOKEN,
                            'tokens_per_s_at_peak_per_gpu': hbm_peak * 1e9 / QDQ_BYTES_PER_TOKEN,
                            'frac': value / world / (hbm_peak * 1e9 / QDQ_BYTES_PER_TOKEN)},
        'qdq_standalone': {'gbs': qdq_gbs, 'frac_of_measured_hbm': qdq_gbs / hbm_peak,
                           'shape': '256Mi fp32 (1 GiB in, 1 GiB out)', 'bytes_per_elem': 8, 'cpu_baseline': cpu_qdq},
        'qat_backward_standalone': {'gbs': bwd_gbs, 'frac_of_measured_hbm': bwd_gbs / hbm_peak if bwd_gbs else None,
                                    'shape': '128Mi fp32 (x, grad_y in; grad_x + range gradients out)',
                                    'bytes_per_elem': 12},
        'kernels': {k: {'ms_per_step': v['seconds'] * 1e3, 'launches': v['launches']} for k, v in prof.items()},
    }
    emit(line)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


_JSON_FD = None


def _guard_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line at
    communicator creation whenever NCCL_DEBUG >= VERSION): point fd 1 at stderr for the whole run and keep
    the original stdout for the result line."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
