#!/bin/bash
timeout 600 python -m pytest tests/test_heter_model_gpu.py tests/test_gencomm_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 300 python scripts/probe/post_anomaly.py 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench_r02aw.json 2> gpurun_out/bench_r02aw.err; tail -2 gpurun_out/bench_r02aw.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02aw.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['stage_ms'])
print(d['single_frame'])
PY
