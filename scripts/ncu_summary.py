"""Condense `ncu -i rep.ncu-rep --page raw --csv` into the per-kernel summary kept under profiles/:
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv ; python scripts/ncu_summary.py raw.csv > profiles/rNN_ncu_full_summary.json
One entry per captured launch: duration, DRAM bytes, pipe utilisation, registers, grid."""
import csv
import json
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic']

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    e = {'kernel': r[hdr.index('Kernel Name')][:90]}
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            e[k] = f'{r[i]} {units[i]}'.strip()
    out.append(e)
print(json.dumps(out, indent=1))
