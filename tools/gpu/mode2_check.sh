#!/usr/bin/env bash
mkdir -p gpurun_out
for c in 2 1; do
TQ_ENGINE_CHAIN=$c TQ_BENCH_OTHER_CONFIGS=0 TQ_BENCH_CALIBRATION=0 timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/m2_bench_$c.json 2> gpurun_out/m2_bench_$c.err
python -c "
import json;p=json.load(open('gpurun_out/m2_bench_$c.json'));print('chain=$c', {k:p.get(k) for k in ('value','ms_per_step','gpu_launches')}, p['kernels'], p['parity']['engine_vs_module_path_logit_steps'])"
done
TRACE_LAYERS=1 timeout 120 python tools/trace_chain.py > gpurun_out/m2_trace.txt 2>&1; grep -E "attention|LN|GELU|QKV|launch" gpurun_out/m2_trace.txt | cut -c1-250
