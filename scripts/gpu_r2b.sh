#!/bin/bash
# r02b: where does the cluster middle kernel spend its time?  (timing bisection + ncu full with source)
OUT=gpurun_out; mkdir -p $OUT
for d in 0 1 2 4 8 3 7 15; do
  echo "GC_CL_DEBUG=$d" | tee -a $OUT/bisect_r02b.txt
  GC_CL_DEBUG=$d timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_unet_middle -c 3 --csv \
     python scripts/bench_sampler.py --iters 1 --precision cluster 2>&1 | grep k_unet_middle | awk -F'","' '{print $NF}' | tee -a $OUT/bisect_r02b.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_unet_middle -s 3 -c 1 -f -o $OUT/prof_cluster_r02b \
    python scripts/bench_sampler.py --iters 1 --precision cluster > $OUT/ncu_full_r02b.log 2>&1
ls -la $OUT | tail -5
