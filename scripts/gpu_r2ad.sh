#!/bin/bash
timeout 300 python scripts/bench_backbone.py --agents 32 2>&1 | tail -1
GC_CONV_MT2=0 timeout 300 python scripts/bench_backbone.py --agents 32 2>&1 | tail -1
