#!/bin/bash
# On the GPU box: per-launch device time of the default bench command (graph replays included), aggregated per kernel.
set -u
TAG=${1:-r01b}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${2:-4300} -c ${3:-1100} --csv \
    --log-file /tmp/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /tmp/ncu_launch.log 2>&1
python - <<'PY' "$TAG"
import csv, sys, collections, json
tag = sys.argv[1]
rows = [r for r in csv.reader(open('/tmp/launches.csv')) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]
kn, mv, mu = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
agg = collections.OrderedDict()
n = 0
seq = []
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
    seq.append(name)
tot = sum(a[1] for a in agg.values())
# one step = distance between consecutive rng_advance launches
marks = [i for i, s in enumerate(seq) if 'rng_advance' in s]
per_step = (marks[-1] - marks[0]) / max(1, len(marks) - 1) if len(marks) > 1 else None
out = {'launches': n, 'total_us': tot, 'launches_per_step': per_step, 'steps_covered': (n / per_step) if per_step else None,
       'kernels': [{'kernel': k, 'launches': a[0], 'us': round(a[1], 1), 'share': round(a[1] / tot, 4)}
                   for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
json.dump(out, open('gpurun_out/launches_%s.json' % tag, 'w'), indent=1)
print('launch list:', n, 'launches', round(tot / 1e3, 2), 'ms', 'launches/step', per_step)
PY
