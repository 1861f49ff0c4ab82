#!/usr/bin/env bash
# One gpurun call: training-path tests, the rest of the GPU suite, QAT kernel timings, ncu capture of the
# backward kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_qat.py -q --tb=short -p no:cacheprovider > gpurun_out/qat_tests.log 2>&1
echo "exit $?" >> gpurun_out/qat_tests.log
timeout "${FULL_TIMEOUT:-330}" python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect tests/test_gpu_qat.py --durations=15 > gpurun_out/gpu_tests.log 2>&1
echo "exit $?" >> gpurun_out/gpu_tests.log
timeout 90 python tools/qat_bench.py > gpurun_out/qat_bench.log 2>&1
echo "exit $?" >> gpurun_out/qat_bench.log
timeout 150 python tools/qdq_variants.py > gpurun_out/qdq_variants.log 2>&1
echo "exit $?" >> gpurun_out/qdq_variants.log
[ -n "${SKIP_NCU:-}" ] || timeout 150 ncu --set full --clock-control none --import-source on -k regex:qdq_bwd -c 4 -f -o gpurun_out/r1_qat_bwd python tools/prof_qat.py > gpurun_out/ncu_qat.log 2>&1
echo "exit $?" >> gpurun_out/ncu_qat.log
tail -4 gpurun_out/qat_tests.log; tail -25 gpurun_out/gpu_tests.log; tail -3 gpurun_out/qat_bench.log; tail -22 gpurun_out/qdq_variants.log; tail -3 gpurun_out/ncu_qat.log
