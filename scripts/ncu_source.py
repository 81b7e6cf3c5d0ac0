#!/usr/bin/env python
"""Per-SASS-instruction view of an ncu report (source page): executed instructions and stall samples.
    python scripts/ncu_source.py prof.ncu-rep [kernel-regex] [top|all]"""
import csv, io, subprocess, sys
cmd = ['ncu', '-i', sys.argv[1], '--page', 'source', '--csv']
if len(sys.argv) > 2 and sys.argv[2]:
    cmd += ['--kernel-name', 'regex:' + sys.argv[2]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several launches: keep the first table only
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hdr = rows[hdr_idx[0]]
end = hdr_idx[1] - 1 if len(hdr_idx) > 1 else len(rows)
ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[hdr_idx[0] + 1:end] if len(r) > iex and r[iex].isdigit()]
tot = sum(int(r[iex]) for r in data); tots = sum(int(r[ismp]) for r in data)
print(f"total warp-instructions {tot:,}  samples {tots:,}  sass lines {len(data)}")
agg = {h: sum(int(r[i]) for r in data if r[i].isdigit()) for i, h in stall_cols}
print("stall totals:", ", ".join(f"{h[6:]}={v}" for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
mode = sys.argv[3] if len(sys.argv) > 3 else 'top'
for i, r in enumerate(data):
    ex, sm = int(r[iex]), int(r[ismp])
    if mode == 'all' or ex * 150 > tot or sm * 150 > tots:
        top = sorted(((int(r[c]), h[6:]) for c, h in stall_cols if r[c].isdigit() and int(r[c])), reverse=True)[:2]
        print(f"{i:5d} {ex:11,d} {100*ex/tot:5.1f}% {sm:6d} {100*sm/max(tots,1):5.1f}%  {r[isrc].strip()[:70]:70s} {top}")
