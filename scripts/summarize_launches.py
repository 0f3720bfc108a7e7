"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
    python scripts/summarize_launches.py gpurun_out/launches_c2.csv > profiles/r01_launches_c2.md
Only kernels of this library (namespace abopt::) are broken out; everything else is 'torch/other'."""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline='') as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        rows.append((r['Kernel Name'], float(r['Metric Value']), r['Grid Size'], r['Block Size']))
    agg = OrderedDict()
    for name, ns, grid, block in rows:
        m = re.search(r'abopt::(\w+)', name)
        key = m.group(1) if m else 'torch/other'
        a = agg.setdefault(key, dict(n=0, ns=0.0, grid=grid, block=block))
        a['n'] += 1
        a['ns'] += ns
    total = sum(a['ns'] for a in agg.values())
    print(f'| kernel | launches | total us | avg us | share | grid | block |')
    print('|---|---:|---:|---:|---:|---|---|')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['ns']):
        print(f"| {k} | {a['n']} | {a['ns'] / 1e3:.1f} | {a['ns'] / 1e3 / a['n']:.1f} | {100 * a['ns'] / total:.1f}% | {a['grid']} | {a['block']} |")
    print(f'\ntotal {total / 1e6:.3f} ms over {len(rows)} launches (per-launch times are cold-cache and serialised under ncu: compare shares)')


if __name__ == '__main__':
    main(sys.argv[1])
