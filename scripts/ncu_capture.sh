#!/bin/bash
# Runs on the GPU box (via gpurun).  Writes only small CSV summaries into gpurun_out/ (the .ncu-rep stays in /tmp).
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
# 1. launch list of the default bench command (graph replay included): per-launch device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3600 -c 1300 --csv \
    --log-file /tmp/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /tmp/ncu_launch.log 2>&1
python - <<'PY' "$TAG"
import csv, sys, collections, json
tag = sys.argv[1]
rows = [r for r in csv.reader(open('/tmp/launches.csv')) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]
kn, mv, mu = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
agg = collections.OrderedDict()
n = 0
for r in rows[hdr + 1:]:
    try:
        v = float(r[mv].replace(',', ''))
    except ValueError:
        continue
    unit = r[mu]
    us = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3)
    name = r[kn].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[:90]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us; n += 1
tot = sum(a[1] for a in agg.values())
out = {'launches': n, 'total_us': tot, 'kernels': [{'kernel': k, 'launches': a[0], 'us': round(a[1], 1), 'share': round(a[1] / tot, 4)}
       for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
json.dump(out, open('gpurun_out/launches_%s.json' % tag, 'w'), indent=1)
print('launch list:', n, 'launches', round(tot / 1e3, 2), 'ms')
PY
# 2. full metric set on a few launches of each hot kernel (eager mode so -k/-s address them directly)
timeout 900 ncu --set full --clock-control none -k regex:"gemm_bf16_|attn_bwd_tc|attn_fwd_tc|relbias_bwd|ln_bwd_kernel|ln_fwd_kernel" \
    -s 760 -c 60 -o /tmp/prof python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > /tmp/ncu_full.log 2>&1
tail -2 /tmp/ncu_full.log | cut -c1-160
ncu -i /tmp/prof.ncu-rep --page raw --csv > /tmp/raw.csv 2>/dev/null
python - <<'PY' "$TAG"
import csv, sys, json
tag = sys.argv[1]
rows = list(csv.reader(open('/tmp/raw.csv')))
H = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu.sum', 'smsp__cycles_active.avg', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_tensor.sum', 'lts__t_bytes.sum']
idx = [(w, H.index(w)) for w in want if w in H]
units = rows[1]
out = []
for r in rows[2:]:
    d = {}
    for w, i in idx:
        d[w] = r[i] if w == 'Kernel Name' else (r[i] + ' ' + units[i]).strip()
    d['Kernel Name'] = d['Kernel Name'].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[:80]
    out.append(d)
json.dump(out, open('gpurun_out/ncu_full_%s.json' % tag, 'w'), indent=1)
print('full capture:', len(out), 'launches')
PY
