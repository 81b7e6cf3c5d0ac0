#!/bin/bash
timeout 600 python -m pytest tests/test_backbone_gpu.py tests/test_det_tail_gpu.py tests/test_enhancer_gpu.py tests/test_heter_model_gpu.py -m gpu -x -q 2>&1 | tail -6
echo "--- default (SC 64 for N<=64)"; timeout 120 scripts/probe/conv_tma_trace.bin 2>&1 | cut -c1-150
echo "--- SC=64 forced for N<=128"; GC_CONV_SC=64 timeout 120 scripts/probe/conv_tma_trace.bin 2>&1 | cut -c1-150
timeout 120 python scripts/bench_backbone.py --agents 32 2>&1 | tail -1
GC_CONV_SC=64 timeout 120 python scripts/bench_backbone.py --agents 32 2>&1 | tail -1
GC_CONV_SC=32 timeout 120 python scripts/bench_backbone.py --agents 32 2>&1 | tail -1
