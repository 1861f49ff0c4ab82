#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_linear.py -q --tb=short -p no:cacheprovider -x -k "calibration_time or fused_vs_unfused or declines" > gpurun_out/c19_tests.log 2>&1; echo "exit $?" >> gpurun_out/c19_tests.log
tail -15 gpurun_out/c19_tests.log | cut -c1-250
timeout 400 python tools/kernel_bench.py > gpurun_out/c19_kernel_bench.log 2>&1; echo "exit $?" >> gpurun_out/c19_kernel_bench.log
grep -E "minmax_axis|qdq_peg6|minmax_tensor|qdq_tensor|mse_" gpurun_out/c19_kernel_bench.log | cut -c1-260
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -c 12 -f -o gpurun_out/r2_kernels python tools/prof_kernels.py > gpurun_out/c19_ncu.log 2>&1; echo "exit $?" >> gpurun_out/c19_ncu.log
tail -3 gpurun_out/c19_ncu.log
