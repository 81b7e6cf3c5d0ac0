#!/bin/bash
# r02k: cluster kernel v3 (MMA warp, row barriers, st.async hand-over)
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gencomm_gpu.py -q -m gpu -p no:cacheprovider -s -x -k "cluster" 2>&1 | tail -12 | tee $OUT/pytest_r02k.log
timeout 120 python scripts/bench_sampler.py --iters 20 --precision cluster 2>&1 | tee $OUT/bench_sampler_r02k.txt
timeout 120 python scripts/bench_sampler.py --iters 20 --precision cluster --frames 1 2>&1 | tee -a $OUT/bench_sampler_r02k.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file $OUT/launches_sampler_r02k.csv \
    python scripts/bench_sampler.py --iters 1 --precision cluster > /dev/null 2>&1
grep -E "k_unet_middle|k_conv_in|k_conv_out|q_sample" $OUT/launches_sampler_r02k.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tail -7
GC_CL_DEBUG=32 timeout 120 python scripts/bench_sampler.py --iters 1 --precision cluster --frames 1 2>&1 | grep "trace layer" | tail -26 | tee $OUT/trace_r02k.txt
