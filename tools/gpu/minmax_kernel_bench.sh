#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "minmax or range or estimator or manager or golden or fullsize_properties" > gpurun_out/minmax_tests.log 2>&1; echo "exit $?" >> gpurun_out/minmax_tests.log
tail -5 gpurun_out/minmax_tests.log | cut -c1-250
timeout 400 python tools/kernel_bench.py > gpurun_out/minmax_kernel_bench.log 2>&1; echo "exit $?" >> gpurun_out/minmax_kernel_bench.log
grep -E "minmax_axis|minmax_tensor" gpurun_out/minmax_kernel_bench.log | cut -c1-200
ls gpurun_out/*.json | head -30
