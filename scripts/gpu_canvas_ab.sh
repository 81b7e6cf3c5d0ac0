#!/bin/bash
# A/B of the canvas writer on one GPU box: per-tile kernel vs the persistent writer's CTA shapes; bit-exact compare.
OUT=gpurun_out
mkdir -p $OUT
TAG=${1:-cv}
GRID=${2:-square}
: > $OUT/canvas_ab_$TAG.txt
GC_CANVAS_IMPL=tile timeout 300 python scripts/bench_canvas.py --grid $GRID --dump /tmp/cv_tile.pt 2>&1 | tail -1 | tee -a $OUT/canvas_ab_$TAG.txt
for cfg in ${CFGS:-0 1 2 3 4 5}; do
  GC_CANVAS_IMPL=persist GC_CANVAS_CFG=$cfg timeout 60 python scripts/bench_canvas.py --grid $GRID --dump /tmp/cv_p$cfg.pt 2>&1 | tail -1 | tee -a $OUT/canvas_ab_$TAG.txt
  timeout 120 python - <<PY 2>&1 | tee -a $OUT/canvas_ab_$TAG.txt
import torch, os
a = torch.load("/tmp/cv_tile.pt")
p = "/tmp/cv_p$cfg.pt"
if os.path.exists(p):
    b = torch.load(p)
    print("cfg $cfg equal:", torch.equal(a, b), "maxdiff", float((a - b).abs().max()), "mismatch", int((a != b).sum()))
else:
    print("cfg $cfg: no output")
PY
done
