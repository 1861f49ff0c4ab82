#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize_parity.py -q --tb=short -p no:cacheprovider -x > gpurun_out/c6_tests.log 2>&1; echo "exit $?" >> gpurun_out/c6_tests.log
tail -25 gpurun_out/c6_tests.log
TQ_BENCH_OTHER_CONFIGS=0 TQ_BENCH_CALIBRATION=0 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err; echo "exit $?" >> gpurun_out/c6_bench.err
python -c "
import json;p=json.load(open('gpurun_out/c6_bench.json'));print({k:p.get(k) for k in ('value','ms_per_step','kernels','parity')}); print(p['roofline'])"
tail -3 gpurun_out/c6_bench.err
