#!/usr/bin/env bash
# round-2 final validation: the whole GPU suite, smoke, launch list, one ncu --set full of the two kernels of a layer, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=8 > gpurun_out/r2_gpu_tests.log 2>&1; echo "exit $?" >> gpurun_out/r2_gpu_tests.log
tail -14 gpurun_out/r2_gpu_tests.log | cut -c1-220
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1; echo "exit $?" >> gpurun_out/r2_smoke.log
tail -3 gpurun_out/r2_smoke.log
TQ_BENCH_OTHER_CONFIGS=0 TQ_BENCH_CALIBRATION=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2_launches_bench.log 2>&1; echo "exit $?" >> gpurun_out/r2_launches_bench.log
tail -2 gpurun_out/r2_launches_bench.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"linear_chain_kernel|attention_kernel" -c 2 -f -o gpurun_out/r2_chain_layer0 python tools/prof_engine.py > gpurun_out/r2_ncu_chain.log 2>&1; echo "exit $?" >> gpurun_out/r2_ncu_chain.log
tail -3 gpurun_out/r2_ncu_chain.log; ls -la gpurun_out/r2_chain_layer0.ncu-rep
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "exit $?" >> gpurun_out/r2_bench_reference_arm.err
cut -c1-300 gpurun_out/r2_bench_reference_arm.json
timeout 300 python bench.py --config bert_w8a8_peg --steps 20 --warmup 3 > gpurun_out/r2_config3_peg_engine.json 2> gpurun_out/r2_config3.err; echo "exit $?" >> gpurun_out/r2_config3.err
timeout 300 python bench.py --config mobilebert_w4a8 --steps 20 --warmup 3 > gpurun_out/r2_config4_mobilebert_engine.json 2> gpurun_out/r2_config4.err; echo "exit $?" >> gpurun_out/r2_config4.err
python -c "
import json
for f in ('r2_config3_peg_engine','r2_config4_mobilebert_engine'):
    p=json.load(open('gpurun_out/%s.json'%f)); print(f, p['value'], p['ms_per_step'], p['kernels'], p['roofline'] and (p['roofline']['kernel'], round(p['roofline']['frac'],3)))"
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "exit $?" >> gpurun_out/r2_bench_n1.err
python -c "
import json;p=json.load(open('gpurun_out/r2_bench_n1.json'));print({k:p.get(k) for k in ('value','ms_per_step','kernels','e2e','gpu_launches')}); print(p['roofline']); print(p['cpu_baseline']); print(p.get('other_configs')); print(p.get('calibration'))" 2>&1 | cut -c1-1500
tail -3 gpurun_out/r2_bench_n1.err
