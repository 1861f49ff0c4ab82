"""Turn the raw ncu outputs of a GPU run (gpurun_out/) into the small tracked summaries under profiles/:
   prof_engine_r1_raw.csv  (ncu --set full, --page raw)       -> r1_ncu_engine_layer0.json, r1_ncu_traffic.json
   launches_r1.csv         (ncu gpu__time_duration launch list) -> r1_launch_shares.json, r1_launches_bench.csv.gz"""
import csv, gzip, json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')


def to_bytes(v, u):
    return float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)


def klass(name):
    for key in ('linear_qdq_kernel', 'attention_kernel', 'ln_qdq_kernel'):
        if key in name:
            return {'linear_qdq_kernel': 'linear_qdq', 'attention_kernel': 'attention', 'ln_qdq_kernel': 'embed_ln_qdq'}[key]
    return 'other (torch glue: mask, first-token gather, casts)'


def ncu_full():
    rows = list(csv.reader(open(os.path.join(OUT, 'prof_engine_r1_raw.csv'))))
    h, units, data = rows[0], rows[1], rows[2:]
    want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size',
            'launch__cluster_dim_x', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
            'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
            'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'lts__t_bytes.sum']
    idx = {w: h.index(w) for w in want if w in h}
    kernels, traffic = [], {}
    for r in data:
        d = {w: {'value': r[i], 'unit': units[i]} for w, i in idx.items()}
        d['dram_bytes_per_launch'] = (to_bytes(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']]) +
                                      to_bytes(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']]))
        kernels.append(d)
        traffic.setdefault(klass(r[idx['Kernel Name']]), []).append(d['dram_bytes_per_launch'])
    src = ('ncu --set full --clock-control none --profile-from-start off -k regex:"linear_qdq_kernel|attention_kernel" -c 5 '
           'python tools/prof_engine.py: encoder layer 0 of one eager engine forward (QKV, attention, attention-out+LN, '
           'FFN-in, FFN-out+LN); cold caches, serialised')
    json.dump({'source': src, 'kernels': kernels}, open(os.path.join(PROF, 'r1_ncu_engine_layer0.json'), 'w'), indent=1)
    json.dump({k: {'dram_bytes_per_launch': sum(v) / len(v), 'launches_profiled': len(v), 'source': src}
               for k, v in traffic.items()}, open(os.path.join(PROF, 'r1_ncu_traffic.json'), 'w'), indent=1)
    for k in kernels:
        print(k['Kernel Name']['value'][:70], k['gpu__time_duration.sum']['value'], 'us, DRAM %.2f MB' % (k['dram_bytes_per_launch'] / 1e6))


def launch_list():
    path = os.path.join(OUT, 'launches_r1.csv')
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(lines))
    h = rows[0]
    iname, ival = h.index('Kernel Name'), h.index('Metric Value')
    seq = [(r[iname], float(r[ival].replace(',', ''))) for r in rows[1:] if len(r) > ival]
    # one engine step = from one embedding kernel (ln_qdq_kernel<1>) to the next
    starts = [i for i, (n, _) in enumerate(seq) if 'ln_qdq_kernel' in n]
    # (the bench's per-class graph replays put embedding kernels back to back: take a pair a whole step apart)
    pairs = [(a, b) for a, b in zip(starts, starts[1:]) if 55 <= b - a <= 80]
    step = seq[pairs[-1][0]:pairs[-1][1]] if pairs else seq
    agg = {}
    for n, ns in step:
        a = agg.setdefault(klass(n), {'launches': 0, 'us_total': 0.0})
        a['launches'] += 1
        a['us_total'] += ns / 1e3
    tot = sum(a['us_total'] for a in agg.values())
    for a in agg.values():
        a['share'] = round(a['us_total'] / tot, 4)
        a['us_total'] = round(a['us_total'], 1)
    json.dump({'command': 'ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file '
                          'gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3',
               'note': 'ncu per-launch times are cold-cache and serialised: compare SHARES with bench.py "kernels", not absolutes',
               'launches_in_list': len(seq), 'one_engine_step': agg},
              open(os.path.join(PROF, 'r1_launch_shares.json'), 'w'), indent=1)
    with open(path, 'rb') as f, gzip.open(os.path.join(PROF, 'r1_launches_bench.csv.gz'), 'wb') as g:
        shutil.copyfileobj(f, g)
    print(json.dumps(agg, indent=1))


if __name__ == '__main__':
    ncu_full()
    launch_list()
