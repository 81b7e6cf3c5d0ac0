#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_rows -s 2 -c 1 -o $OUT/conv_rows64_r02w -f \
    python scripts/bench_backbone.py --agents 32 --iters 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_me_conv -s 14 -c 1 -o $OUT/me_conv256_r02w -f \
    python scripts/bench_backbone.py --agents 32 --iters 1 > /dev/null 2>&1
ls -la $OUT/*.ncu-rep
