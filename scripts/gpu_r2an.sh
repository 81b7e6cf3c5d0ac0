#!/bin/bash
OUT=gpurun_out
for k in k_canvas_planes k_enh_dw_cl k_enh_ln k_slot_insert; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o $OUT/${k}_r02an -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
done
ls -la $OUT/*_r02an.ncu-rep
