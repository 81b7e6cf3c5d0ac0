#!/bin/bash
# r02h: 512-thread cluster kernel; detection-level parity numbers
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gencomm_gpu.py tests/test_heter_model_gpu.py -q -m gpu -p no:cacheprovider -s -k "cluster or detections_match" 2>&1 | grep -v "^$" | tail -25 | tee $OUT/pytest_r02h.log
timeout 300 python scripts/bench_sampler.py --iters 20 --precision cluster 2>&1 | tee $OUT/bench_sampler_r02h.txt
timeout 300 python scripts/bench_sampler.py --iters 20 --precision cluster --frames 1 2>&1 | tee -a $OUT/bench_sampler_r02h.txt
timeout 300 python scripts/bench_sampler.py --iters 20 --precision cluster --frames 8 --agents 5 --C 256 2>&1 | tee -a $OUT/bench_sampler_r02h.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file $OUT/launches_sampler_r02h.csv \
    python scripts/bench_sampler.py --iters 1 --precision cluster > /dev/null 2>&1
grep -E "k_unet_middle|k_conv_in|k_conv_out|q_sample" $OUT/launches_sampler_r02h.csv | awk -F'","' '{print $5, $NF}' | tail -8
