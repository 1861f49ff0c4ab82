#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_linear.py tests/test_gpu_fullsize_parity.py -q --tb=short -p no:cacheprovider -x > gpurun_out/c8_tests.log 2>&1; echo "exit $?" >> gpurun_out/c8_tests.log
tail -12 gpurun_out/c8_tests.log
python tools/trace_tiles.py 2>&1 | grep -A4 LEAN > gpurun_out/c8_trace_tiles.log; cat gpurun_out/c8_trace_tiles.log
TQ_BENCH_OTHER_CONFIGS=0 TQ_BENCH_CALIBRATION=0 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err; echo "exit $?" >> gpurun_out/c8_bench.err
python -c "
import json;p=json.load(open('gpurun_out/c8_bench.json'));print({k:p.get(k) for k in ('value','ms_per_step','kernels')}); print(p['roofline']['frac'], p['parity']['engine_vs_module_path_logit_steps'])"
tail -3 gpurun_out/c8_bench.err
