"""Key figures per kernel launch out of an `ncu --set full` report (read here with `ncu -i ... --page raw --csv`):
    python scripts/ncu_extract.py gpurun_out/x.ncu-rep profiles/out.json"""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
H = rows[0]
def col(name):
    return H.index(name) if name in H else None
want = {
    'duration_us': ('gpu__time_duration.sum', 1e-3),
    'tensor_pipe_active_pct': ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 1),
    'tensor_op_hmma_pct': ('sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active', 1),
    'issue_slots_busy_pct': ('sm__inst_issued.avg.pct_of_peak_sustained_active', 1),
    'ipc_active': ('sm__inst_executed.avg.per_cycle_active', 1),
    'dram_read_MB': ('dram__bytes_read.sum', None),
    'dram_write_MB': ('dram__bytes_write.sum', None),
    'dram_throughput_pct': ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 1),
    'l2_hit_pct': ('lts__t_sector_hit_rate.pct', 1),
    'sm_throughput_pct': ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 1),
    'warps_active_pct': ('sm__warps_active.avg.pct_of_peak_sustained_active', 1),
    'registers': ('launch__registers_per_thread', 1),
    'dyn_smem_KB': ('launch__shared_mem_per_block_dynamic', None),
    'cluster': ('launch__cluster_dim_x', 1),
    'smem_bank_conflicts': ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 1),
}
units = rows[1]
res = []
for r in rows[2:]:
    if len(r) < len(H):
        continue
    d = {'kernel': r[col('Kernel Name')][:110], 'grid': r[col('Grid Size')], 'block': r[col('Block Size')]}
    for key, (metric, scale) in want.items():
        c = col(metric)
        if c is None or r[c] == '':
            continue
        v = float(r[c].replace(',', ''))
        u = units[c]
        if scale is None:          # byte-like: normalise by the unit string
            mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(u, 1)
            v = v * mult / (1e6 if key.endswith('_MB') else 1e3)
        elif key == 'duration_us':
            v = v * {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3}.get(u, 1e-3)
        d[key] = round(v, 3)
    res.append(d)
json.dump(res, open(out, 'w'), indent=1)
for d in res:
    print('%-60s %8.1f us  tensor %5s%%  issue %5s%%  dram %6s/%6s MB  regs %s' % (d['kernel'][:60], d.get('duration_us', 0), d.get('tensor_pipe_active_pct'),
          d.get('issue_slots_busy_pct'), d.get('dram_read_MB'), d.get('dram_write_MB'), d.get('registers')))
