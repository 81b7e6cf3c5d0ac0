#!/bin/bash
# A/B of the warp+fusion configurations.  bash scripts/gpu_fuse_ab.sh tag "cfgs" [mode]
TAG=${1:-ab}
CFGS=${2:-"0 1 2"}
MODE=${3:-all}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_warp_fuse_gpu.py -q -m gpu -x 2>&1 | tail -8 | tee $OUT/pytest_fuse_$TAG.log
for v in $CFGS; do
  echo "== persist cfg $v"
  GC_FUSE_CFG=$v timeout 300 python scripts/bench_fuse.py --quick --mode $MODE 2>&1 | tee $OUT/bench_fuse_${TAG}_p$v.txt
done
