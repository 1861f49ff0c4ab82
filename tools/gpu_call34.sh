#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "attention or engine or fullsize" > gpurun_out/c34_tests.log 2>&1; echo "exit $?" >> gpurun_out/c34_tests.log
tail -5 gpurun_out/c34_tests.log | cut -c1-250
TQ_BENCH_OTHER_CONFIGS=0 TQ_BENCH_CALIBRATION=0 timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/c34_bench.json 2> gpurun_out/c34_bench.err
python -c "
import json;p=json.load(open('gpurun_out/c34_bench.json'));print({k:p.get(k) for k in ('value','ms_per_step')}, p['kernels']['attention'], p['kernels']['chain'])"
tail -2 gpurun_out/c34_bench.err
