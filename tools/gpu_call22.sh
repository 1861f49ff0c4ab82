#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_engine.py -q --tb=short -p no:cacheprovider -x -k "chain" > gpurun_out/c22_chain.log 2>&1; echo "exit $?" >> gpurun_out/c22_chain.log
tail -12 gpurun_out/c22_chain.log | cut -c1-250
timeout 120 python tools/trace_chain.py > gpurun_out/c22_trace_chain.txt 2>&1; echo "exit $?" >> gpurun_out/c22_trace_chain.txt
cat gpurun_out/c22_trace_chain.txt | cut -c1-300
