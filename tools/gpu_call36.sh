#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1; echo "exit $?" >> gpurun_out/r2_smoke.log
tail -4 gpurun_out/r2_smoke.log | cut -c1-300
