#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize_parity.py -q --tb=short -p no:cacheprovider -x > gpurun_out/c35_tests.log 2>&1; echo "exit $?" >> gpurun_out/c35_tests.log
tail -8 gpurun_out/c35_tests.log | cut -c1-250
for h in 1; do
TQ_ENGINE_HEAD=$h TQ_BENCH_OTHER_CONFIGS=0 TQ_BENCH_CALIBRATION=0 timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/c35_bench_$h.json 2> gpurun_out/c35_bench_$h.err
python -c "
import json;p=json.load(open('gpurun_out/c35_bench_$h.json'));print('head=$h', {k:p.get(k) for k in ('value','ms_per_step','gpu_launches')}, p['kernels']['linear_qdq'], p['parity']['engine_vs_module_path_logit_steps'])"
tail -1 gpurun_out/c35_bench_$h.err | cut -c1-200
done
