#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c30_smoke.log 2>&1; echo "exit $?" >> gpurun_out/c30_smoke.log
tail -4 gpurun_out/c30_smoke.log | cut -c1-250
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_linear.py tests/test_gpu_engine_mobilebert.py tests/test_gpu_engine_peg.py tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -x > gpurun_out/c30_tests.log 2>&1; echo "exit $?" >> gpurun_out/c30_tests.log
tail -6 gpurun_out/c30_tests.log | cut -c1-250
TQ_BENCH_CALIBRATION=0 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/c30_bench.json 2> gpurun_out/c30_bench.err
python -c "
import json;p=json.load(open('gpurun_out/c30_bench.json'));print({k:p.get(k) for k in ('value','ms_per_step','kernels')}); print({k:(v['ms_per_step'], v['tokens_per_s']) for k,v in p['other_configs'].items()})"
tail -2 gpurun_out/c30_bench.err
