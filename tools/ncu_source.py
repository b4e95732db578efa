#!/usr/bin/env python
"""Per-SASS-instruction executed counts / samples of one kernel: tools/ncu_source.py rep kernel-regex [min_Minst]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[ix['Instructions Executed']].isdigit()]
ti = sum(int(r[ix['Instructions Executed']]) for r in data); ts = sum(int(r[ix['# Samples']]) for r in data)
print('SASS insts', len(data), 'warp-inst executed', ti, 'samples', ts)
for r in data:
    ie = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
    if ie / 1e6 >= thr:
        print(f"{r[ix['Address']][-4:]} {ie/1e6:8.2f}M thr={float(r[ix['Avg. Threads Executed']]):5.1f} samp={s:6d}  {r[ix['Source']].strip()[:100]}")
