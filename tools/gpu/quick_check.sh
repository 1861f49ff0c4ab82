#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -q --tb=short -p no:cacheprovider -x > gpurun_out/qc_tests.log 2>&1; echo "exit $?" >> gpurun_out/qc_tests.log
tail -4 gpurun_out/qc_tests.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1; echo "exit $?" >> gpurun_out/r2_smoke.log
tail -2 gpurun_out/r2_smoke.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "exit $?" >> gpurun_out/r2_bench_n1.err
python -c "
import json;p=json.load(open('gpurun_out/r2_bench_n1.json'));print({k:p.get(k) for k in ('value','ms_per_step','gpu_launches')}, p['e2e']['value'], p['roofline']['frac']); print(p['parity']); print(p['cpu_baseline']['logit_diff_vs_gpu_in_classifier_steps'])" | cut -c1-900
tail -2 gpurun_out/r2_bench_n1.err
