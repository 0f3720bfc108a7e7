"""Static SASS evidence: counts of the tcgen05 / TMEM / TMA mnemonics per kernel of the built objects.
    python scripts/sass_mnemonics.py > profiles/rNN_final_sass_mnemonics.txt"""
import glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
print('# SASS evidence (cuobjdump -sass of ab_opt_b200/_lib/obj/*.o, sm_100a): tcgen05 / TMEM / TMA mnemonics per kernel')
print('# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = 1-D bulk copy, '
      'UTCBAR = tcgen05.commit, FFMA2 = fma.rn.f32x2, HMMA = legacy mma.sync')
print('# TS-mode MMAs (A operand from tensor memory) show as UTCHMMA with a tmem[...] A operand: counted as UTCHMMA.TS below')
keys = ('UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'FFMA2', 'HMMA')
for obj in sorted(glob.glob(os.path.join(ROOT, 'ab_opt_b200', '_lib', 'obj', '*.o'))):
    out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    name, counts = None, {}
    def flush():
        if name and counts:
            print(os.path.basename(obj), name[:110], ' '.join(f'{k} {v}' for k, v in sorted(counts.items())))
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            flush()
            name, counts = re.sub(r'^_ZN5abopt', '', m.group(1)), {}
            continue
        m = re.search(r'^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);', line)
        if not m:
            continue
        op, args = m.group(1), m.group(2)
        base = op.split('.')[0]
        if base in keys:
            if base == 'UTCHMMA' and re.match(r'\s*tmem\[', args):
                base = 'UTCHMMA.TS'
            counts[base] = counts.get(base, 0) + 1
    flush()
