#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -q --tb=short -p no:cacheprovider -x > gpurun_out/c20_tests.log 2>&1; echo "exit $?" >> gpurun_out/c20_tests.log
tail -6 gpurun_out/c20_tests.log | cut -c1-250
timeout 300 python bench.py --config bert_w8a8_peg --steps 20 --warmup 3 > gpurun_out/c20_peg.json 2> gpurun_out/c20_peg.err; echo "exit $?" >> gpurun_out/c20_peg.err
timeout 300 python bench.py --config mobilebert_w4a8 --steps 20 --warmup 3 > gpurun_out/c20_mb.json 2> gpurun_out/c20_mb.err; echo "exit $?" >> gpurun_out/c20_mb.err
python -c "
import json
for f in ('c20_peg','c20_mb'):
    p=json.load(open('gpurun_out/%s.json'%f)); print(f, p['value'], p['ms_per_step'], p['kernels'], p['roofline'] and (p['roofline']['kernel'], round(p['roofline']['frac'],3)))"
tail -3 gpurun_out/c20_peg.err gpurun_out/c20_mb.err
TQ_BENCH_OTHER_CONFIGS=0 TQ_BENCH_CALIBRATION=0 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/c20_bench.json 2> gpurun_out/c20_bench.err
python -c "
import json;p=json.load(open('gpurun_out/c20_bench.json'));print({k:p.get(k) for k in ('value','ms_per_step','kernels')})"
