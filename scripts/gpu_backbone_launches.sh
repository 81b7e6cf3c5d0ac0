#!/bin/bash
OUT=gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches_backbone.csv \
    python scripts/bench_backbone.py --agents 32 --iters 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_backbone.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
data=rows[hdr+1:]
prev=None;cnt=0;tot=0
for r in data[-41:]+[None]:
    key=(r[ki][:75],r[gi]) if r else None
    if key==prev: cnt+=1; tot+=float(r[vi])
    else:
        if prev: print(f"{prev[0]:77s} {prev[1]:14s} x{cnt:2d} {tot/1e3:8.1f}")
        if r: prev=key;cnt=1;tot=float(r[vi])
PY
