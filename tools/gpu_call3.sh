#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c3_smoke.log 2>&1; echo "exit $?" >> gpurun_out/c3_smoke.log
timeout 400 python -m pytest tests/test_gpu_fullsize_parity.py -q --tb=short -p no:cacheprovider > gpurun_out/c3_fullsize.log 2>&1; echo "exit $?" >> gpurun_out/c3_fullsize.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; echo "exit $?" >> gpurun_out/c3_bench.err
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x --deselect tests/test_gpu_fullsize_parity.py --durations=8 > gpurun_out/c3_gpu_tests.log 2>&1; echo "exit $?" >> gpurun_out/c3_gpu_tests.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c3_bench_ref.json 2> gpurun_out/c3_bench_ref.err
tail -5 gpurun_out/c3_smoke.log; tail -15 gpurun_out/c3_fullsize.log; tail -12 gpurun_out/c3_gpu_tests.log; tail -5 gpurun_out/c3_bench.err; head -c 600 gpurun_out/c3_bench_ref.json
