#!/usr/bin/env python
"""ncu report (.ncu-rep, read here without a GPU) -> small tracked JSON under profiles/:
    python tools/summarize_ncu.py gpurun_out/r2_kernels.ncu-rep profiles/r2_ncu_kernels.json "<how it was captured>"
Per kernel launch: duration, DRAM bytes read / written, DRAM and L2 throughput %, issue-slot / FMA / ALU / XU / tensor
pipe utilisation, occupancy, registers, grid, and the top warp-stall reasons."""
import csv
import json
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__cluster_dim_x', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max']
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}


def main():
    rep, out, how = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else '')
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units, data = rows[0], rows[1], rows[2:]
    kernels = []
    for r in data:
        d = {'kernel': r[h.index('Kernel Name')]}
        for w in WANT:
            if w not in h:
                continue
            i = h.index(w)
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                continue
            u = units[i]
            if w == 'gpu__time_duration.sum':
                d['duration_us'] = v * UNIT.get(u, 1.0)
            elif 'bytes' in w:
                d[w] = v * UNIT.get(u, 1)
            else:
                d[w] = v
        if 'dram__bytes_read.sum' in d:
            d['dram_bytes_per_launch'] = d['dram__bytes_read.sum'] + d.get('dram__bytes_write.sum', 0.0)
        stalls = {k.split('issue_stalled_')[1].split('_per_')[0]: float(r[h.index(k)]) for k in h
                  if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio') and r[h.index(k)]}
        d['top_stalls_per_issue'] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
        kernels.append(d)
    json.dump({'source': how, 'note': 'ncu replays every launch cold and serialised: compare shares and traffic, not absolute times',
               'kernels': kernels}, open(out, 'w'), indent=1)
    for k in kernels:
        print('%-60s %8.1f us  DRAM %7.1f MB  dram %5.1f%%  issue %5.1f%%  tensor %5.1f%%' % (
            k['kernel'][:60], k.get('duration_us', 0), k.get('dram_bytes_per_launch', 0) / 1e6,
            k.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0), k.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0),
            k.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0)))


if __name__ == '__main__':
    main()
