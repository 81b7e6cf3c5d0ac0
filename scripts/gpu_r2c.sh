#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gencomm_gpu.py -q -m gpu -p no:cacheprovider -s -x -k "cluster" 2>&1 | tail -12 | tee $OUT/pytest_r02c.log
for d in 0 1 2 4 8 15; do
  echo "GC_CL_DEBUG=$d" | tee -a $OUT/bisect_r02c.txt
  GC_CL_DEBUG=$d timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_unet_middle -c 2 --csv \
     python scripts/bench_sampler.py --iters 1 --precision cluster 2>&1 | grep k_unet_middle | awk -F'","' '{print $NF}' | tee -a $OUT/bisect_r02c.txt
done
timeout 300 python scripts/bench_sampler.py --iters 20 --precision cluster 2>&1 | tee $OUT/bench_sampler_r02c.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_unet_middle -s 3 -c 1 -f -o $OUT/prof_cluster_r02c \
    python scripts/bench_sampler.py --iters 1 --precision cluster > $OUT/ncu_full_r02c.log 2>&1
