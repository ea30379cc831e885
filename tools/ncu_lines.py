#!/usr/bin/env python
"""Per-source-line executed-instruction and stall-sample breakdown of an .ncu-rep (needs -lineinfo).
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source=cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No'][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0, '']); tot = 0; ts = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    try:
        ln = int(r[0]); ie = int(r[ix['Instructions Executed']]); sm = int(r[ix['# Samples']])
    except ValueError:
        continue
    agg[ln][0] += ie; agg[ln][1] += sm; agg[ln][2] = r[1]; tot += ie; ts += sm
print(f"total warp instructions {tot}, stall samples {ts}")
for ln, (ie, sm, src) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{ln:5d} {ie:12d} {100*ie/max(tot,1):5.1f}%  samples {sm:6d} {100*sm/max(ts,1):5.1f}%  {src.strip()[:95]}")
