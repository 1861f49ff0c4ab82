#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize_parity.py tests/test_gpu_engine.py tests/test_gpu_engine_peg.py tests/test_gpu_engine_mobilebert.py -q --tb=short -p no:cacheprovider -x > gpurun_out/c25_tests.log 2>&1; echo "exit $?" >> gpurun_out/c25_tests.log
tail -8 gpurun_out/c25_tests.log | cut -c1-250
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "attention" > gpurun_out/c25_attn.log 2>&1; echo "exit $?" >> gpurun_out/c25_attn.log
tail -5 gpurun_out/c25_attn.log | cut -c1-250
ls profiles/r2_parity_fullsize.json && python - <<'PY'
import json
p=json.load(open('profiles/r2_parity_fullsize.json'))
print({k:(v if not isinstance(v,(dict,list)) else '...') for k,v in p.items()})
PY
TQ_BENCH_OTHER_CONFIGS=0 TQ_BENCH_CALIBRATION=0 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/c25_bench.json 2> gpurun_out/c25_bench.err
python -c "
import json;p=json.load(open('gpurun_out/c25_bench.json'));print({k:p.get(k) for k in ('value','ms_per_step','kernels')}); print(p.get('parity'))"
tail -2 gpurun_out/c25_bench.err
