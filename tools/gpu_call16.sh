#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c16_smoke.log 2>&1; echo "exit $?" >> gpurun_out/c16_smoke.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=8 > gpurun_out/c16_gpu_tests.log 2>&1; echo "exit $?" >> gpurun_out/c16_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err; echo "exit $?" >> gpurun_out/c16_bench.err
tail -4 gpurun_out/c16_smoke.log; tail -25 gpurun_out/c16_gpu_tests.log | cut -c1-250; tail -3 gpurun_out/c16_bench.err
python -c "
import json;p=json.load(open('gpurun_out/c16_bench.json'))
print({k:p.get(k) for k in ('value','ms_per_step','kernels')}); print(p['roofline']['frac'], p['e2e']['value'], p['cpu_baseline']['value'])
print(p['calibration']); print({k:(v.get('tokens_per_s'), v.get('forward')) for k,v in p['other_configs'].items()})"
