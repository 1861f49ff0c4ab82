#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine_mobilebert.py -q --tb=short -p no:cacheprovider -x > gpurun_out/c15_tests.log 2>&1; echo "exit $?" >> gpurun_out/c15_tests.log
tail -30 gpurun_out/c15_tests.log | cut -c1-300
timeout 300 python tools/run_config.py --config mobilebert_w4a8 > gpurun_out/c15_mb.json 2> gpurun_out/c15_mb.err; echo "exit $?" >> gpurun_out/c15_mb.err
cat gpurun_out/c15_mb.json; tail -5 gpurun_out/c15_mb.err
