#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches_backbone_r02u.csv \
    python scripts/bench_backbone.py --agents 32 --iters 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_backbone_r02u.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
data=rows[hdr+1:]
# last 41 launches = one backbone call
for r in data[-41:]:
    print(r[ki][:70], r[gi], r[vi])
PY
