#!/bin/bash
GC_CONV_PERSIST=1 timeout 600 python -m pytest tests/test_backbone_gpu.py tests/test_det_tail_gpu.py -m gpu -x -q 2>&1 | tail -4
GC_CONV_MT2=0 GC_CONV_PERSIST=1 timeout 300 python scripts/bench_backbone.py --agents 32 2>&1 | tail -1
GC_CONV_MT2=0 timeout 300 python scripts/bench_backbone.py --agents 32 2>&1 | tail -1
export GC_CONV_PERSIST=1 GC_CONV_MT2=0
bash scripts/gpu_r2ac.sh
