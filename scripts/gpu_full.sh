#!/bin/bash
# full GPU validation: tests, smoke, bench (N = 1)
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02}
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -2 $OUT/bench_$TAG.err
python - $OUT/bench_$TAG.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])
print(d['stage_ms']); print(d['roofline']); print(d['clocks'])
for k in ('hbm_step','gencomm_sampler','components','other_shape','single_frame'):
    print(k, json.dumps(d.get(k))[:700])
PY
