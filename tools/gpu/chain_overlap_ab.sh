#!/usr/bin/env bash
# FFN-in -> FFN-out overlap of the chain kernel: tests, determinism stress, per-stage trace, bench A/B (TQ_CHAIN_OVERLAP)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -q --tb=short -p no:cacheprovider -x -k "chain" > gpurun_out/ovl_tests.log 2>&1; echo "exit $?" >> gpurun_out/ovl_tests.log
tail -6 gpurun_out/ovl_tests.log | cut -c1-250
timeout 300 python tools/stress_chain.py 200 > gpurun_out/ovl_stress.log 2>&1; echo "exit $?" >> gpurun_out/ovl_stress.log
tail -2 gpurun_out/ovl_stress.log
for o in 1 0; do
TQ_CHAIN_OVERLAP=$o TRACE_ATT=0 TRACE_LAYERS=1 timeout 120 python tools/trace_chain.py > gpurun_out/ovl_trace_$o.txt 2>&1; echo "overlap=$o"; grep -E "LN|GELU|QKV|launch" gpurun_out/ovl_trace_$o.txt | cut -c1-250
TQ_CHAIN_OVERLAP=$o TQ_BENCH_OTHER_CONFIGS=0 TQ_BENCH_CALIBRATION=0 timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/ovl_bench_$o.json 2> gpurun_out/ovl_bench_$o.err
python -c "
import json;p=json.load(open('gpurun_out/ovl_bench_$o.json'));print('overlap=$o', {k:p.get(k) for k in ('value','ms_per_step')}, p['kernels']['chain'], p['roofline']['frac'], p['parity']['engine_vs_module_path_logit_steps'])"
done
