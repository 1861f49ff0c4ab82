#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_engine.py -q --tb=short -p no:cacheprovider -x -k "chain" > gpurun_out/c23_chain.log 2>&1; echo "exit $?" >> gpurun_out/c23_chain.log
tail -12 gpurun_out/c23_chain.log | cut -c1-250
timeout 120 python tools/trace_chain.py > gpurun_out/c23_trace_chain.txt 2>&1; echo "exit $?" >> gpurun_out/c23_trace_chain.txt
cat gpurun_out/c23_trace_chain.txt | cut -c1-300
TQ_CHAIN_MCAST=0 timeout 120 python tools/trace_chain.py > gpurun_out/c23_trace_chain_nomc.txt 2>&1
tail -2 gpurun_out/c23_trace_chain_nomc.txt
TQ_BENCH_OTHER_CONFIGS=0 TQ_BENCH_CALIBRATION=0 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/c23_bench.json 2> gpurun_out/c23_bench.err
python -c "
import json;p=json.load(open('gpurun_out/c23_bench.json'));print({k:p.get(k) for k in ('value','ms_per_step','kernels')})"
tail -2 gpurun_out/c23_bench.err
