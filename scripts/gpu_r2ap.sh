#!/bin/bash
timeout 120 python scripts/bench_backbone.py --agents 32 2>&1 | tail -1
GC_CONV_ROWS=0 timeout 120 python scripts/bench_backbone.py --agents 32 2>&1 | tail -1
export GC_CONV_ROWS=0
timeout 200 bash scripts/gpu_r2ac.sh
