#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gencomm_gpu.py -q -m gpu -p no:cacheprovider -s -x -k "cluster" 2>&1 | tail -8 | tee $OUT/pytest_r02o.log
for d in 0; do
echo "GC_CL_DEBUG=$d" | tee -a $OUT/bench_sampler_r02o.txt
GC_CL_DEBUG=$d timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_unet_middle -c 3 --csv \
   python scripts/bench_sampler.py --iters 1 --precision cluster 2>&1 | grep k_unet_middle | awk -F'","' '{print $NF}' | tee -a $OUT/bench_sampler_r02o.txt
done
GC_CL_DEBUG=32 timeout 120 python scripts/bench_sampler.py --iters 1 --precision cluster --frames 1 2>&1 | grep "trace layer" | tail -26 | tee $OUT/trace_r02o.txt
