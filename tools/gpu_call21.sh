#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_engine.py -q --tb=short -p no:cacheprovider -x -k "chain" > gpurun_out/c21_chain.log 2>&1; echo "exit $?" >> gpurun_out/c21_chain.log
tail -25 gpurun_out/c21_chain.log | cut -c1-250
for c in 0 1; do
TQ_ENGINE_CHAIN=$c TQ_BENCH_OTHER_CONFIGS=0 TQ_BENCH_CALIBRATION=0 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/c21_bench_$c.json 2> gpurun_out/c21_bench_$c.err
python -c "
import json;p=json.load(open('gpurun_out/c21_bench_$c.json'));print({k:p.get(k) for k in ('value','ms_per_step','kernels','parity')})"
tail -2 gpurun_out/c21_bench_$c.err
done
