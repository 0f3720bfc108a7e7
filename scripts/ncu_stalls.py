"""Summarise `ncu --page source --csv` output: total samples per stall reason and the hottest SASS lines.
    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv ; python scripts/ncu_stalls.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[idx['# Samples']].strip().isdigit()]      # several launches: header rows repeat
tot = {c: sum(int(r[idx[c]] or 0) for r in data) for c in stall_cols}
all_s = sum(int(r[idx['# Samples']] or 0) for r in data)
print('total samples', all_s)
for c, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f'  {c:28s} {v:8d} {100 * v / max(all_s, 1):5.1f}%')
print('hottest instructions:')
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']] or 0))[:top]:
    reasons = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"  {int(r[idx['# Samples']]):6d}  {r[idx['Source']].strip()[:70]:70s} " + ' '.join(f'{n}:{v}' for v, n in reasons if v))
