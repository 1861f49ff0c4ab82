#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/stress_chain.py 300 > gpurun_out/fence_stress.log 2>&1; echo "exit $?" >> gpurun_out/fence_stress.log
tail -3 gpurun_out/fence_stress.log
for f in 0 1; do
TQ_CHAIN_GPU_FENCE=$f TQ_BENCH_OTHER_CONFIGS=0 TQ_BENCH_CALIBRATION=0 timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/fence_bench_$f.json 2> gpurun_out/fence_bench_$f.err
python -c "
import json;p=json.load(open('gpurun_out/fence_bench_$f.json'));print('gpu_fence=$f', {k:p.get(k) for k in ('value','ms_per_step')}, p['kernels']['chain'], p['roofline']['frac'])"
done
TRACE_ATT=0 TRACE_LAYERS=1 timeout 120 python tools/trace_chain.py > gpurun_out/fence_trace.txt 2>&1; grep -E "LN|GELU|QKV|launch" gpurun_out/fence_trace.txt | cut -c1-250
