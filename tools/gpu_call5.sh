#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"linear_qdq_kernel|linear_lean_kernel|attention_kernel" -c 5 -f -o gpurun_out/r2_engine_layer0 python tools/prof_engine.py > gpurun_out/c5_ncu.log 2>&1; echo "exit $?" >> gpurun_out/c5_ncu.log
tail -5 gpurun_out/c5_ncu.log; ls -la gpurun_out/*.ncu-rep
