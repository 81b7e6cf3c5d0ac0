#!/bin/bash
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['stage_ms'], d['clocks']['samples'])
PY
}
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-extras > gpurun_out/z1.json 2>/dev/null; show gpurun_out/z1.json
GC_BENCH_CLOCK_MS=25 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-extras > gpurun_out/z2.json 2>/dev/null; show gpurun_out/z2.json
GC_CONV_ROWS=0 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-extras > gpurun_out/z3.json 2>/dev/null; show gpurun_out/z3.json
