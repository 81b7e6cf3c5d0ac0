#!/bin/bash
# Quick GPU visit: parity tests + bench (+ optional ncu of selected kernels).  bash scripts/gpu_quick.sh tag [ncu-regex]
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -s 2>&1 | tail -60 | tee $OUT/pytest_$TAG.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json
tail -5 $OUT/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_launch_$TAG.log 2>&1
if [ -n "$2" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s 6 -c 3 \
    -f -o $OUT/prof_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_full_$TAG.log 2>&1
fi
