#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for cfg in "8 64 128 2 1" "8 16 24 2 1" "8 64 128 4 2" "8 64 128 1 0"; do
  echo "== $cfg"; timeout 120 python scripts/debug_tma.py $cfg 2>&1 | tail -3
done
echo "== sanitizer"
timeout 300 compute-sanitizer --tool memcheck python scripts/debug_tma.py 4 64 64 2 1 2>&1 | grep -v "^$" | head -60 | tee $OUT/sanitizer.log
