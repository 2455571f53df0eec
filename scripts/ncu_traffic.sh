#!/bin/bash
# On the GPU box: DRAM traffic + duration + tensor-pipe activity of every launch of OUR kernels over one eager step.
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
    --clock-control none -k regex:"gemm_bf16_|attn_|relbias_|ln_fwd|ln_bwd|colsum|cast_|mixed_" -s 1100 -c 340 --csv --log-file /tmp/traffic.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > /tmp/ncu_traffic.log 2>&1
tail -2 /tmp/ncu_traffic.log | cut -c1-160
python - <<'PY'
import csv, json, collections
rows = [r for r in csv.reader(open('/tmp/traffic.csv')) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]
ki, mi, vi, ui, idi = H.index('Kernel Name'), H.index('Metric Name'), H.index('Metric Value'), H.index('Metric Unit'), H.index('ID')
fam_of = [('gemm_bf16_', 'gemm_bf16_tc'), ('attn_bwd', 'attn_bwd'), ('attn_fwd', 'attn_fwd'), ('relbias_mma', 'relbias_mma'), ('relbias_bwd', 'relbias_bwd'),
          ('relbias_fwd', 'relbias_fwd'), ('ln_bwd', 'ln_residual_bwd'), ('ln_fwd', 'ln_residual_fwd'), ('colsum', 'colsum'),
          ('cast_multi', 'cast_multi'), ('cast_kernel', 'cast_f32_to_bf16'), ('mixed_', 'mixed')]
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, '%': 1.0}
per = collections.defaultdict(dict)
for r in rows[hdr + 1:]:
    try:
        v = float(r[vi].replace(',', '')) * scale.get(r[ui], 1.0)
    except ValueError:
        continue
    per[(r[idi], r[ki])][r[mi]] = v
fam = collections.OrderedDict()
for (_, name), m in per.items():
    f = next((b for a, b in fam_of if a in name), None)
    if f is None:
        continue
    d = fam.setdefault(f, {'launches': 0, 'dram_bytes': 0.0, 'l2_bytes': 0.0, 'us': 0.0, 'tensor_pct_x_us': 0.0})
    d['launches'] += 1
    d['dram_bytes'] += m.get('dram__bytes_read.sum', 0.0) + m.get('dram__bytes_write.sum', 0.0)
    d['l2_bytes'] += m.get('lts__t_bytes.sum', 0.0)
    d['us'] += m.get('gpu__time_duration.sum', 0.0)
    d['tensor_pct_x_us'] += m.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0.0) * m.get('gpu__time_duration.sum', 0.0)
out = {}
for f, d in fam.items():
    out[f] = {'launches_per_step': d['launches'], 'dram_bytes_per_launch': d['dram_bytes'] / d['launches'],
              'l2_bytes_per_launch': d['l2_bytes'] / d['launches'], 'us_per_launch_under_ncu': d['us'] / d['launches'],
              'us_per_step_under_ncu': d['us'], 'tensor_pipe_active_pct_time_weighted': d['tensor_pct_x_us'] / max(d['us'], 1e-9)}
json.dump(out, open('gpurun_out/ncu_traffic.json', 'w'), indent=1)
for f, d in out.items():
    print('%-18s n=%3d  dram/launch %8.2f MB  us/launch %7.1f  tensor %5.1f%%' % (f, d['launches_per_step'], d['dram_bytes_per_launch'] / 1e6, d['us_per_launch_under_ncu'], d['tensor_pipe_active_pct_time_weighted']))
PY
