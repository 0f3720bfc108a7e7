"""DRAM traffic of one full-stream pair_stream_kernel launch from an `ncu --page raw --csv` dump -> profiles/pair_kernel_traffic.json
(`roofline.traffic` of bench.py):  python scripts/pair_traffic.py raw.csv out.json [source note]"""
import csv, json, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ik, it, ir, iw = (hdr.index(k) for k in ('Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum'))
def to_bytes(v, u):
    return float(v.replace(',', '')) * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
def to_us(v, u):
    return float(v.replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u, 1.0)
full = [(to_bytes(r[ir], units[ir]), to_bytes(r[iw], units[iw])) for r in rows[2:]
        if len(r) == len(hdr) and 'pair_stream_kernel' in r[ik] and to_us(r[it], units[it]) > 120.0]      # (the generated-rows launches take ~25 us)
assert full, 'no full-stream pair_stream_kernel launch in the capture'
rd = sum(f[0] for f in full) / len(full); wr = sum(f[1] for f in full) / len(full)
B, L = 64, 256
out = {'kernel': 'pair_stream_kernel', 'config': 'C2 B=64 L=256 (one full-stream launch = whole batch, one IPA layer)', 'L': L,
       'dram_bytes_read_per_launch': rd, 'dram_bytes_write_per_launch': wr, 'dram_bytes_per_complex': (rd + wr) / B,
       'source': (sys.argv[3] if len(sys.argv) > 3 else 'ncu --set full --clock-control none') +
                 f' (dram__bytes_read.sum + dram__bytes_write.sum, mean of the {len(full)} full-stream launches captured)',
       'note': 'algorithmic 17.05 MB/complex; the rest is the alpha read (3.1 MB/complex) and the feat write'}
json.dump(out, open(sys.argv[2], 'w'), indent=1)
print(out)
