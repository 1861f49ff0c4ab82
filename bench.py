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
  calibration  BASELINE config 5 at every N: RoBERTa-base W8A8, MSE (grid, 100 candidates) activation ranges, one
              B=32 calibration batch PER RANK, estimator statistics all-reduced over NCCL inside
              quantization/_dist.calibration_sync() -- tokens/s, collective calls and bytes, ranks identical?
  other_configs  (N = 1) BASELINE configs 3 and 4 through tools/run_config.py: BERT-base PEG (K = 6, range-permuted)
              and MobileBERT W4A8 B=64, CUDA-graphed eval forward, tokens/s
The reference arm runs the UNMODIFIED reference (baseline/_ref, tools/install_reference.sh) through its own
public API on the host CPU with all host threads (baseline/reference_arm.py; the oracle port only if that copy is
absent) on the same config and prints the same line with "impl": "reference".

    python bench.py --config bert_w8a8_peg | mobilebert_w4a8 | roberta_w8a8_mse     one of the other configs as the line
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
DTYPE = ('u8 / s8 integer-grid operands with int32 accumulation (QKV, attention-out, FFN-out GEMMs), bf16 integer-grid '
         'operands with fp32 accumulation (FFN-in GEMM, attention products) -- all exact; quantize / dequantize / LayerNorm / '
         'softmax / GELU arithmetic in fp32')
QDQ_BYTES_PER_TOKEN = 1.3456e6     # algorithmic QDQ traffic per token


def peaks():
    """(HBM GB/s, bf16 TFLOP/s burst, bf16 TFLOP/s sustained, source).  The timed region of this benchmark is tens of
    milliseconds at full clocks -- the BURST figure is the denominator for its kernels."""
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return d['hbm_gbs'], d['bf16_tflops'], d['bf16_tflops_sustained'], 'measured'
    return 6650.0, 1590.0, 1400.0, 'fallback'


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


def cpu_forward_setup():
    """The reference's CPU implementation of the workload, calibrated on the benchmark batch with fixed ranges:
    the UNMODIFIED reference from baseline/_ref (kind "reference") when tools/install_reference.sh has installed
    it, else the oracle port (kind "port").  -> (forward(ids, mask), ids, mask, calibration seconds, kind, what)"""
    from oracle.bert_oracle import OracleBert, random_bert_state_dict
    sd = random_bert_state_dict(seed=0)
    ids = synthetic_ids(1234)[0]
    mask = torch.ones_like(ids)
    m, kind, what = None, 'port', 'oracle/bert_oracle.py: the reference op chain on torch CPU'
    try:
        from baseline import reference_arm
        if reference_arm.available():
            m = reference_arm.ReferenceBert(sd)
            kind = 'reference'
            what = ('unmodified reference (baseline/_ref: models/quantized_bert.py + quantization/*, its own public API; '
                    'HF-4.1 container shim tests/hf41_shim.py)')
    except Exception as e:                                  # noqa: BLE001
        print(f'bench.py: reference arm unavailable ({e!r}); timing the oracle port', file=sys.stderr)
        m = None
    if m is None:
        m = OracleBert(sd, n_layers=12, n_heads=12, n_bits=8, n_bits_act=8, sym_acts=False)
    with torch.no_grad():
        t0 = time.perf_counter()
        m(ids, mask)                 # calibration forward on the full batch (also caches the weights)
        t_cal = time.perf_counter() - t0
        m.fix_ranges()
    return m, ids, mask, t_cal, kind, what


def cpu_baseline(threads, n_timed=5):
    """the reference on the host cores: calibrate on the full batch, fix the ranges, then the MEDIAN of
    ``n_timed`` fixed-range forwards of the SAME full batch (B=32, T=128) after one warm-up forward."""
    m, ids, mask, t_cal, kind, what = cpu_forward_setup()
    with torch.no_grad():
        logits = m(ids, mask)        # warm-up
        ts = []
        for _ in range(n_timed):
            t0 = time.perf_counter()
            logits = m(ids, mask)
            ts.append(time.perf_counter() - t0)
    return BATCH * SEQ / statistics.median(ts), n_timed, logits, kind, what


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
    m, ids, mask, t_cal, kind, what = cpu_forward_setup()
    # every step is one forward of the SAME full batch (B=32, T=128) at every N; only if the host is so slow that
    # the whole run would exceed ~5 minutes are the steps cut short (and the line says so)
    steps, warmup = args.steps, args.warmup
    with torch.no_grad():
        t0 = time.perf_counter()
        m(ids, mask)
        t_fwd = time.perf_counter() - t0
        budget = 300.0
        if (steps + warmup) * t_fwd > budget:
            steps = max(5, int(budget / t_fwd) - 1)
            warmup = 1
        for _ in range(max(warmup - 1, 0)):
            m(ids, mask)
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            m(ids, mask)
            ts.append(time.perf_counter() - t0)
    dt = sum(ts)
    # the CPU path does not shard: N "GPUs" of the reference arm are still one host
    val = steps * BATCH * SEQ / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'tokens/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': dt / steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'global_batch': BATCH, 'seq_len': SEQ, 'parallelism': 'host-cpu'},
        'cpu_baseline': {'value': val, 'unit': 'tokens/s', 'cores': threads, 'kind': kind,
                         'sample': f'{steps} fixed-range forwards of the full batch (32 x 128 tokens) after one '
                                   f'full-batch calibration forward; median {statistics.median(ts) * 1e3:.0f} ms per forward; {what}',
                         'steps_requested': args.steps},
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
            keep_alive = forward(ids_dev, mask_dev)      # (its storage is a recorded kernel argument: keep it while the classes replay)
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
    del keep_alive
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


def calibration_leg(dev, rank, world):
    """BASELINE config 5 (every rank calls this): RoBERTa-base W8A8, MSE range estimation on the activations
    (grid search, 100 candidates; the asymmetric two-sided sites take the reference's 2-D grid), ONE calibration
    batch of 32 x 128 tokens PER RANK through ``utils.pass_data_for_range_estimation`` -- the loop that opts into
    quantization/_dist.calibration_sync(): every estimator update all-reduces its statistics over NCCL, so all
    ranks end with the ranges of the global batch.  Timed on the device, max over ranks; a first untimed pass on
    the same model (ranges reset afterwards) takes module loading / allocator warm-up out of the number."""
    from engine import configs
    from quantization import _dist
    from utils.utils import pass_data_for_range_estimation
    import torch.distributed as dist
    model, recipe = configs.build('roberta_w8a8_mse', dev)
    ids = configs.synthetic_batches(model, recipe, 1, seed=4321 + rank)[0]
    loader = [{'input_ids': ids, 'attention_mask': torch.ones_like(ids)}]

    def one_pass():
        with torch.no_grad():
            pass_data_for_range_estimation(loader, model, act_quant=True, weight_quant=True, max_num_batches=1)

    one_pass()                                     # warm-up (also fills the weight caches: weights calibrate once)
    model.reset_act_ranges()
    model.estimate_act_ranges()
    _dist.stats(reset=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    one_pass()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    secs = e0.elapsed_time(e1) * 1e-3
    model.fix_ranges()
    st = _dist.stats()
    deltas = torch.cat([m.quantizer._delta.reshape(-1).float() for m in model.act_quantizers()])
    identical = True
    if world > 1:
        t = torch.tensor([secs, wall], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs, wall = t.tolist()
        gathered = [torch.empty_like(deltas) for _ in range(world)]
        dist.all_gather(gathered, deltas)
        identical = all(torch.equal(gathered[0], g) for g in gathered)
    tokens = recipe.batch * recipe.seq * world
    return {'workload': 'RoBERTa-base W8A8, activation ranges by MSE grid search (100 candidates), weights current_minmax; '
                        'one 32 x 128 calibration batch per rank, statistics all-reduced inside the estimators',
            'tokens_per_s': tokens / secs, 'seconds': secs, 'host_wall_seconds': wall, 'n_gpus': world,
            'global_calibration_tokens': tokens, 'sites': len(model.act_quantizers()),
            'allreduce_calls': st['calls'], 'allreduce_bytes': st['bytes'],
            'collective': 'NCCL all-reduce (MAX on packed [-min, max]; SUM on fp64 loss arrays)' if world > 1 else 'none (one rank)',
            'ranks_identical': bool(identical)}


def other_config_legs(steps, warmup):
    """BASELINE configs 3 and 4 on this GPU (N = 1 only): tools/run_config.py -- build, calibrate on two synthetic
    batches (incl. the FP32 ranges pass for the PEG permutation), fix ranges, CUDA-graphed eval forward."""
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import run_config
    out = {}
    for name in ('bert_w8a8_peg', 'mobilebert_w4a8'):
        try:
            out[name] = run_config.run(name, steps, warmup)
        except Exception as e:                              # noqa: BLE001  (a side leg must not take the headline down)
            out[name] = {'error': repr(e)}
        torch.cuda.empty_cache()
    return out


def run_other_config(args):
    """`--config NAME`: one of BASELINE configs 3 / 4 / 5 as the line of this run (single GPU unless it is the
    calibration config, which shards over the ranks).  Same keys as the headline line where they apply."""
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the product path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    hbm_peak, tf_peak, tf_sustained, peak_kind = peaks()
    if args.config == 'roberta_w8a8_mse':
        if world > 1:
            os.environ['NCCL_DEBUG'] = os.environ.get('TQ_NCCL_DEBUG', 'INFO')
            torch.distributed.init_process_group('nccl', device_id=dev)
        with ClockSampler(local) as clk:
            c = calibration_leg(dev, rank, world)
        if rank == 0:
            emit({'metric': 'calibration tokens/sec RoBERTa-base W8A8 MSE range estimation', 'value': c['tokens_per_s'],
                  'unit': 'tokens/s', 'n_gpus': world, 'steps': 1, 'warmup': 1, 'ms_per_step': c['seconds'] * 1e3,
                  'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 (MSE losses fp64)',
                  'data': 'synthetic', 'config': {'workload': c['workload'], 'global_batch': 32 * world, 'seq_len': SEQ,
                                                  'parallelism': f'dp{world} (calibration batches sharded, statistics all-reduced)'},
                  'clocks': clk.summary(), 'calibration': c})
        if world > 1:
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()
        return
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import run_config
    import tq_native
    ops = tq_native.ops()
    with ClockSampler(local) as clk:
        r = run_config.run(args.config, args.steps, max(args.warmup, 3), profile=True)
    prof = r.pop('kernel_profile', None) or {}
    roof = None
    if prof:
        name, st = max(prof.items(), key=lambda kv: kv[1]['seconds'])
        tensor = name in ('linear_qdq', 'attention', 'chain')
        ach = st['work'] / st['seconds'] / (1e12 if tensor else 1e9)
        roof = {'kernel': name, 'bound': 'tensor' if tensor else 'hbm', 'achieved': ach, 'peak': tf_peak if tensor else hbm_peak,
                'unit': 'TFLOP/s' if tensor else 'GB/s', 'frac': ach / (tf_peak if tensor else hbm_peak), 'traffic': None,
                'launches_per_step': st['launches'], 'avg_launch_us': st['seconds'] / st['launches'] * 1e6,
                'share_of_library_kernel_time': st['seconds'] / sum(v['seconds'] for v in prof.values()),
                'peak_source': f'{peak_kind} (MEASURED_PEAKS.json)'}
    emit({'metric': f'tokens/sec {args.config} seq{r["seq"]} b{r["batch"]}', 'value': r['tokens_per_s'], 'unit': 'tokens/s',
          'n_gpus': 1, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': r['ms_per_step'],
          'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'integer-grid tensor-core GEMMs, fp32 QDQ',
          'data': 'synthetic', 'config': {'workload': args.config, 'global_batch': r['batch'], 'seq_len': r['seq'],
                                          'forward': r['forward'], 'cuda_graph': r['cuda_graph']},
          'gpu_launches': r['library_launches_per_step'] * args.steps, 'clocks': clk.summary(), 'roofline': roof,
          'kernels': {k: {'ms_per_step': v['seconds'] * 1e3, 'launches': v['launches']} for k, v in prof.items()},
          'detail': r})
    del ops


def run_ours(args):
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the product path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        # NCCL's INFO log (communicator size, transports, NVLS) is the evidence that the ranks really talk: leave it
        # on.  It goes to stderr -- fd 1 is redirected for the whole run (_guard_stdout), stdout carries one JSON line.
        os.environ['NCCL_DEBUG'] = os.environ.get('TQ_NCCL_DEBUG', 'INFO')
        dist.init_process_group('nccl', device_id=dev)
    import tq_native
    ops = tq_native.ops()
    hbm_peak, tf_peak, tf_sustained, peak_kind = peaks()

    model = build_model(dev)
    ids_host = synthetic_ids(1234 + rank)[0].pin_memory()
    mask_dev = torch.ones(BATCH, SEQ, dtype=torch.int64, device=dev)
    ids_dev = ids_host.to(dev)
    with torch.no_grad():
        # replicas calibrate on their own batch: no collective (the reduction is opt-in, quantization/_dist.py)
        model(ids_dev, mask_dev)                      # calibration batch: ranges + weight cache
        model.fix_ranges()
        for _ in range(2):
            model(ids_dev, mask_dev)                  # eager warm-up (function attributes, caches)
        torch.cuda.synchronize()

        # ---- the public forward: fused engine built from the calibrated model (module path if the
        #      configuration is outside the engine's support) ----
        from engine.fused import FusedBertEngine, UnsupportedByEngine
        try:
            forward = FusedBertEngine(model, BATCH, SEQ)
            engine_kind = ('fused (engine/fused.py, x_int byte carriers, int8 tensor cores: per encoder layer the attention kernel + '
                           + {0: 'four GEMM kernels (attention-out + residual + LayerNorm, FFN-in + GELU, FFN-out + residual + LayerNorm, '
                                 'next Q|K|V)',
                              1: 'ONE encoder-chain launch (attention-out + residual + LayerNorm -> FFN-in + GELU -> FFN-out + residual + '
                                 'LayerNorm -> next Q|K|V, a 4-CTA cluster per sequence)',
                              2: 'everything in ONE encoder-chain launch for all layers'}[getattr(forward, 'chain', 0)]
                           + ('; one-launch classification head)' if getattr(forward, 'head', False) else ')'))
            if not getattr(forward, 'i8', False):
                engine_kind = ('fused (engine/fused.py: 5 kernels per encoder layer -- QKV GEMM, attention, attention-out GEMM + residual + '
                               'LayerNorm, FFN-in GEMM + GELU, FFN-out GEMM + residual + LayerNorm -- bf16 integer-grid carriers)')
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
        # (cudaProfilerStart / Stop bracket the timed steps: `ncu --profile-from-start off` then lists exactly the
        # launches of the timed region -- profiles/r2_launches_bench.csv.gz; a no-op without a profiler attached)
        torch.cuda.profiler.start()
        with ClockSampler(local) as clk:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                graph.replay()
            e1.record()
            barrier()
        torch.cuda.profiler.stop()
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

    # ---- BASELINE config 5: sharded calibration with the cross-rank all-reduce (every rank takes part) ----
    calib = None
    if os.environ.get('TQ_BENCH_CALIBRATION', '1') != '0':
        try:
            calib = calibration_leg(dev, rank, world)
        except Exception as e:                              # noqa: BLE001
            calib = {'error': repr(e)}
    if rank != 0:
        if world > 1:
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()
        return

    with torch.no_grad():

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
    flops_kernels = ('linear_qdq', 'attention', 'chain')
    if name in flops_kernels:
        kname = {'linear_qdq': 'tq_linear_qdq_i8 / tq_linear_res_ln_qdq_i8 / tq_linear_qdq_bf16_o8 (tcgen05 GEMM + fused QDQ / GELU / residual / '
                               'LayerNorm epilogue)',
                 'chain': 'tq_chain_plan_run (linear_chain_kernel: ' + (
                     'the whole encoder in one launch -- per layer attention, attention-output + LayerNorm, FFN-in + GELU, FFN-out + '
                     'LayerNorm, next Q|K|V; tcgen05 kind::i8 GEMMs + kind::f16 attention products'
                     if getattr(forward, 'chain', 0) == 2 else
                     'one launch per encoder layer, a 4-CTA cluster per 128-token sequence -- attention-output + residual + LayerNorm, '
                     'FFN-in + GELU, FFN-out + residual + LayerNorm, next Q|K|V; tcgen05 kind::i8 GEMMs') + ', fused QDQ epilogues)',
                 'attention': 'tq_attention_qdq_i8'}[name]
        roof = {'kernel': kname, 'bound': 'tensor',
                'achieved': st['work'] / st['seconds'] / 1e12, 'peak': tf_peak, 'unit': 'TFLOP/s',
                'peak_kind': 'bf16 dense BURST (cuBLAS 8192^3, best of 10) -- the timed region is tens of ms at full clocks',
                'frac_of_sustained_bf16_peak': st['work'] / st['seconds'] / 1e12 / tf_sustained}
        if name in ('linear_qdq', 'chain') and getattr(forward, 'i8', False):
            # 2/3 of the GEMM flops of a layer run as kind::i8 (QKV, attention-out, FFN-out), whose pipe rate is 2x
            # bf16 (tools/mainloop_probe_i8.py: 133 cycles per M128 x N256 x K32 MMA alone = K16 bf16): flop-weighted peak
            share_i8 = forward.i8_flop_share(name)
            roof['i8_flop_share'] = share_i8
            roof['peak_i8_weighted'] = tf_peak / (1.0 - share_i8 / 2.0)
            roof['frac_of_i8_weighted_peak'] = roof['achieved'] / roof['peak_i8_weighted']
    else:
        roof = {'kernel': name, 'bound': 'hbm', 'achieved': st['work'] / st['seconds'] / 1e9,
                'peak': hbm_peak, 'unit': 'GB/s'}
    roof['frac'] = roof['achieved'] / roof['peak']
    roof['traffic'] = None
    try:                      # DRAM bytes per launch of that kernel class from the committed ncu --set full capture
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
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

    cpu = cpu_qdq = others = parity = None
    if world == 1:
        others = other_config_legs(10, 3) if os.environ.get('TQ_BENCH_OTHER_CONFIGS', '1') != '0' else None
        threads = host_threads()
        cpu_val, n_timed, cpu_logits, cpu_kind, cpu_what = cpu_baseline(threads)
        cls_step = float(model.classifier.activation_quantizer.quantizer.scale.reshape(-1)[0])
        cpu = {'value': cpu_val, 'unit': 'tokens/s', 'cores': threads, 'kind': cpu_kind,
               'sample': f'1 full-batch calibration forward, 1 warm-up, median of {n_timed} timed fixed-range forwards of the '
                         f'full batch (32 x 128 tokens); {cpu_what}',
               'logit_max_abs_diff_vs_gpu': float((cpu_logits - static_logits.float().cpu()).abs().max()),
               'logit_diff_vs_gpu_in_classifier_steps': float((cpu_logits - static_logits.float().cpu()).abs().max()) / cls_step}
        cpu_qdq = cpu_qdq_baseline(threads)
        # the chaos floor of this workload: the reference's own op chain on cuBLAS fp32 vs the same chain on the host
        # CPU -- same formulas, two GEMM summation orders (tests/test_gpu_fullsize_parity.py)
        try:
            from oracle.bert_oracle import OracleBert, random_bert_state_dict
            sd = {k: v.to(dev) for k, v in random_bert_state_dict(seed=0).items()}
            ob = OracleBert(sd, n_layers=12, n_heads=12, device=dev)
            with torch.no_grad():
                ob(ids_dev, mask_dev)
                ob.fix_ranges()
                floor = float((ob(ids_dev, mask_dev).float().cpu() - cpu_logits).abs().max()) / cls_step
            del ob, sd
        except Exception as e:                              # noqa: BLE001
            print(f'bench.py: parity floor not measured: {e!r}', file=sys.stderr)
            floor = None
        parity = {'classifier_step': cls_step,
                  'engine_vs_module_path_logit_steps': engine_vs_module / cls_step,
                  'engine_vs_reference_cpu_logit_steps': cpu['logit_diff_vs_gpu_in_classifier_steps'],
                  'floor_reference_on_cublas_vs_reference_on_cpu_logit_steps': floor,
                  'note': 'a 12-layer fake-quantized encoder amplifies ONE flipped integer to a saturated difference '
                          '(70-77 % of the last hidden integers, 15-19 classifier steps) within ~8 layers; the reference '
                          'drifts that far against itself on two GEMM libraries (floor).  Per-kernel parity on identical '
                          'inputs at this size: GEMM epilogues bit-exact, LayerNorm < 1e-5, attention < 1e-4 of the integers '
                          'off by one step (tests/test_gpu_fullsize_parity.py, profiles/r2_parity_fullsize.json)'}

    line = {
        'metric': METRIC, 'value': value, 'unit': 'tokens/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': t_dev / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': DTYPE,
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
        'parity': parity,
        'calibration': calib,
        'other_configs': others,
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
    ap.add_argument('--config', default='bert_w8a8_asym',
                    choices=['bert_w8a8_asym', 'bert_w8a8_peg', 'mobilebert_w4a8', 'roberta_w8a8_mse'],
                    help='BASELINE configuration (default: the headline, configs[1]); the others print their own line')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    elif args.config != 'bert_w8a8_asym':
        run_other_config(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
