#!/bin/bash
# same-box A/B of the step-level switches (sparse planes, noise pre-draw): bench.py headline only, alternating order
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-ab}
run() {  # name, env...
    name=$1; shift
    env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name', 'frames/s %.1f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], d['stage_ms'], 'sm_mhz', d['clocks']['sm_mhz'])
"
}
for rep in 1 2; do
run "planes=sparse predraw@encoder " GC_NOISE_PREDRAW_AT=encoder
run "planes=sparse predraw@backbone" GC_NOISE_PREDRAW_AT=backbone
run "planes=sparse predraw=0       " GC_NOISE_PREDRAW=0
run "planes=dense  predraw=0       " GC_SPARSE_PLANES=0 GC_NOISE_PREDRAW=0
done 2>&1 | tee $OUT/ab_step_$TAG.txt
