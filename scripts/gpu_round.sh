#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full captures of the two big kernels.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "== pytest -m gpu" 
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -40 | tee $OUT/pytest_$TAG.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 3 2>$OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json
tail -5 $OUT/bench_$TAG.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_launch_$TAG.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_canvas_persist|k_fuse_persist|k_slot_insert" -s 6 -c 3 \
    -f -o $OUT/prof_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT
echo "== fuse microbench"
timeout 600 python scripts/bench_fuse.py --quick 2>&1 | tee $OUT/bench_fuse_$TAG.txt
echo "== component microbenches"
timeout 200 python scripts/bench_detector.py 2>&1 | tail -1 | tee $OUT/bench_detector_$TAG.json
timeout 200 python scripts/bench_backbone.py --torch 2>&1 | tail -1 | tee $OUT/bench_backbone_$TAG.json
timeout 100 python scripts/bench_postprocess.py 2>&1 | tail -1 | tee $OUT/bench_postprocess_$TAG.json
timeout 100 python scripts/bench_det_tail.py 2>&1 | tail -1 | tee $OUT/bench_det_tail_$TAG.json
timeout 100 python scripts/bench_me.py 2>&1 | tail -1 | tee $OUT/bench_me_$TAG.json
timeout 100 python scripts/bench_enh.py 2>&1 | tail -1 | tee $OUT/bench_enh_$TAG.json
echo "== ncu: backbone conv + postprocess kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_me_conv" -s 60 -c 6 \
    -f -o $OUT/prof_det_$TAG python scripts/bench_detector.py --frames 2 --iters 1 > $OUT/ncu_det_$TAG.log 2>&1
