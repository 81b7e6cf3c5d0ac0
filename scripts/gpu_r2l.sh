#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
GC_CL_DEBUG=32 timeout 120 python scripts/bench_sampler.py --iters 1 --precision cluster --frames 1 2>&1 | grep "trace layer" | tail -26 | tee $OUT/trace_r02l.txt
