#!/bin/bash
# r02a: first run of the cluster-resident UNet middle
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu_r02a.txt 2>&1
timeout 600 python -m pytest tests/test_gencomm_gpu.py -q -m gpu -p no:cacheprovider -s -x -k "cluster" 2>&1 | tail -40 | tee $OUT/pytest_r02a.log
timeout 300 python scripts/bench_sampler.py --iters 20 2>&1 | tee $OUT/bench_sampler_r02a.txt
timeout 300 python scripts/bench_sampler.py --iters 20 --frames 1 --precision cluster 2>&1 | tee -a $OUT/bench_sampler_r02a.txt
timeout 300 python scripts/bench_sampler.py --iters 20 --frames 1 --precision tc 2>&1 | tee -a $OUT/bench_sampler_r02a.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_sampler_r02a.csv \
    python scripts/bench_sampler.py --iters 1 --precision cluster > $OUT/ncu_sampler_r02a.log 2>&1
