#!/usr/bin/env bash
# backward span kernel on / off: tests and timings; full GPU suite with the chunk forward kernel as default
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_qat.py -q --tb=short -p no:cacheprovider > gpurun_out/qat_tests.log 2>&1
echo "exit $?" >> gpurun_out/qat_tests.log
TQ_BWD_SPAN_MIN=0 timeout 100 python -m pytest tests/test_gpu_qat.py -q --tb=short -p no:cacheprovider -k "bwd or backward or training" > gpurun_out/qat_tests_span.log 2>&1
echo "exit $?" >> gpurun_out/qat_tests_span.log
timeout 120 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect tests/test_gpu_qat.py > gpurun_out/gpu_tests.log 2>&1
echo "exit $?" >> gpurun_out/gpu_tests.log
timeout 60 python tools/qat_bench.py > gpurun_out/qat_bench.log 2>&1
cp gpurun_out/qat_bench.json gpurun_out/qat_bench_persistent.json
TQ_BWD_SPAN_MIN=0 timeout 60 python tools/qat_bench.py > gpurun_out/qat_bench_span.log 2>&1
cp gpurun_out/qat_bench.json gpurun_out/qat_bench_span.json
tail -3 gpurun_out/qat_tests.log; tail -3 gpurun_out/qat_tests_span.log; tail -3 gpurun_out/gpu_tests.log
grep '"qdq_bwd"' gpurun_out/qat_bench.log | cut -c1-200 | head -4; echo ---; grep '"qdq_bwd"' gpurun_out/qat_bench_span.log | cut -c1-200 | head -4
