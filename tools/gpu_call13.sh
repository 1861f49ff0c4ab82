#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine_peg.py -q --tb=short -p no:cacheprovider -x > gpurun_out/c13_tests.log 2>&1; echo "exit $?" >> gpurun_out/c13_tests.log
tail -30 gpurun_out/c13_tests.log
timeout 300 python tools/run_config.py --config bert_w8a8_peg > gpurun_out/c13_peg.json 2> gpurun_out/c13_peg.err; echo "exit $?" >> gpurun_out/c13_peg.err
cat gpurun_out/c13_peg.json; tail -5 gpurun_out/c13_peg.err
