#!/bin/bash
# r02i: conv_in v2 parity + timing; cluster kernel timeline trace
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gencomm_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -4 | tee $OUT/pytest_r02i.log
timeout 300 python scripts/bench_sampler.py --iters 20 --precision cluster 2>&1 | tee $OUT/bench_sampler_r02i.txt
GC_CONV_IN_V1=1 timeout 300 python scripts/bench_sampler.py --iters 20 --precision cluster 2>&1 | tee -a $OUT/bench_sampler_r02i.txt
timeout 300 python scripts/bench_sampler.py --iters 20 --precision cluster --frames 8 --agents 5 --C 256 2>&1 | tee -a $OUT/bench_sampler_r02i.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file $OUT/launches_sampler_r02i.csv \
    python scripts/bench_sampler.py --iters 1 --precision cluster > /dev/null 2>&1
grep -E "k_unet_middle|k_conv_in|k_conv_out|q_sample" $OUT/launches_sampler_r02i.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tail -7
GC_CL_DEBUG=32 timeout 120 python scripts/bench_sampler.py --iters 1 --precision cluster --frames 1 2>&1 | grep "trace layer" | tail -26 | tee $OUT/trace_r02i.txt
