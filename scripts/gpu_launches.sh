#!/bin/bash
# launch list of one detector step (ncu serialises launches: shares, not absolutes)
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02z}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_step_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --steps-in-flight 1 --frames-per-step ${FRAMES:-15} > /dev/null 2>&1
python - $OUT/launches_step_$TAG.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
data=rows[hdr+1:]
# the device-resident loop: 3 warm-up + 1 timed step, then e2e 3 + 1; find k_cell_assign occurrences
starts=[i for i,r in enumerate(data) if 'k_cell_assign' in r[ki]]
print(len(data), 'launches;', len(starts), 'steps')
a,b=starts[3],starts[4]
tot=0
for r in data[a:b]:
    print(f"{r[ki][:100]:102s} {r[gi]:16s} {float(r[vi])/1e3:9.1f}")
    tot+=float(r[vi])
print('sum us', tot/1e3, 'launches', b-a)
PY
