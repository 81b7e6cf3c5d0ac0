#!/bin/bash
timeout 900 python -m pytest tests/test_backbone_gpu.py tests/test_det_tail_gpu.py tests/test_heter_model_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 20 > gpurun_out/bench_r02y.json 2> gpurun_out/bench_r02y.err; tail -3 gpurun_out/bench_r02y.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02y.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
print(d['stage_ms'])
PY
