#!/bin/bash
# end-of-round visit: full GPU suite, smoke, both bench arms
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-final}
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; tail -1 $OUT/bench_ref_$TAG.err; cut -c1-700 $OUT/bench_ref_$TAG.json
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -1 $OUT/bench_$TAG.err
python - $OUT/bench_$TAG.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['clocks'])
print(d['stage_ms']); print(d['roofline']['frac'], d['roofline']['issued_frac'], d['roofline']['traffic'])
print(d['single_frame']); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
