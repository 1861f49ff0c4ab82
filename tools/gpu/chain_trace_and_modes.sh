#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -q --tb=short -p no:cacheprovider -x -k "chain" > gpurun_out/chain_chain.log 2>&1; echo "exit $?" >> gpurun_out/chain_chain.log
tail -25 gpurun_out/chain_chain.log | cut -c1-250
timeout 120 python tools/trace_chain.py > gpurun_out/chain_trace_chain.txt 2>&1; echo "exit $?" >> gpurun_out/chain_trace_chain.txt; TRACE_ATT=0 TRACE_LAYERS=1 timeout 120 python tools/trace_chain.py > gpurun_out/chain_trace_chain_noatt.txt 2>&1; cat gpurun_out/chain_trace_chain_noatt.txt | cut -c1-300
cat gpurun_out/chain_trace_chain.txt | cut -c1-300
for c in 1 2; do
TQ_ENGINE_CHAIN=$c TQ_BENCH_OTHER_CONFIGS=0 TQ_BENCH_CALIBRATION=0 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/chain_bench_$c.json 2> gpurun_out/chain_bench_$c.err
python -c "
import json;p=json.load(open('gpurun_out/chain_bench_$c.json'));print({k:p.get(k) for k in ('value','ms_per_step','kernels')}); print(p['e2e'], p['roofline']['frac'], p['parity']['engine_vs_module_path_logit_steps'])"
tail -2 gpurun_out/chain_bench_$c.err
done
