#!/usr/bin/env bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 600 python tests/parity_fullsize.py > gpurun_out/c1_parity.log 2>&1; echo "exit $?" >> gpurun_out/c1_parity.log



tail -60 gpurun_out/c1_parity.log 
