#!/bin/bash
for F in 8 12 15 16 24; do
timeout 300 python bench.py --steps 12 --no-cpu-baseline --no-extras --frames-per-step $F > gpurun_out/f$F.json 2>/dev/null
python - gpurun_out/f$F.json $F <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('F', sys.argv[2], round(d['value'],1), 'frames/s', round(d['ms_per_step'],2), 'ms  e2e', round(d['e2e']['value'],1), {k:round(v,2) for k,v in d['stage_ms'].items()})
PY
done
