#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python tools/trace_tiles.py > gpurun_out/c4_trace_tiles.log 2>&1; echo "exit $?" >> gpurun_out/c4_trace_tiles.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c4_smoke.log 2>&1; echo "exit $?" >> gpurun_out/c4_smoke.log
timeout 400 python -m pytest tests/test_gpu_fullsize_parity.py -q --tb=short -p no:cacheprovider > gpurun_out/c4_fullsize.log 2>&1; echo "exit $?" >> gpurun_out/c4_fullsize.log
cat gpurun_out/c4_trace_tiles.log; tail -4 gpurun_out/c4_smoke.log; tail -6 gpurun_out/c4_fullsize.log
