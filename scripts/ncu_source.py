#!/usr/bin/env python
"""Per-SASS-instruction view of an ncu report (source page): cumulative executed instructions and stall samples.
    python scripts/ncu_source.py prof.ncu-rep [top N]"""
import csv, io, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
data = [r for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
tot = sum(int(r[iex]) for r in data); tots = sum(int(r[ismp]) for r in data)
print(f"total warp-instructions {tot:,}  samples {tots:,}  sass lines {len(data)}")
mode = sys.argv[2] if len(sys.argv) > 2 else 'all'
cum = 0
for i, r in enumerate(data):
    ex, sm = int(r[iex]), int(r[ismp]); cum += ex
    if mode == 'all' or ex * 200 > tot or sm * 200 > tots:
        print(f"{i:5d} {ex:12,d} {100*ex/tot:5.1f}% {sm:7d} {100*sm/max(tots,1):5.1f}%  {r[isrc].strip()[:90]}")
