#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "exit $?" >> gpurun_out/r2_bench_n2.err
grep -E "NCCL INFO.*(nranks|NVLS|Connected|comm 0x)" gpurun_out/r2_bench_n2.err | head -20 > gpurun_out/r2_bench_n2_nccl_info.txt
python -c "
import json;p=json.load(open('gpurun_out/r2_bench_n2.json'));print({k:p.get(k) for k in ('value','ms_per_step','n_gpus','e2e')}); print(p.get('calibration'))" | cut -c1-900
tail -2 gpurun_out/r2_bench_n2.err | cut -c1-300; wc -l gpurun_out/r2_bench_n2_nccl_info.txt
