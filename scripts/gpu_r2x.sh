#!/bin/bash
timeout 900 python -m pytest tests/test_backbone_gpu.py tests/test_det_tail_gpu.py tests/test_heter_model_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 120 scripts/probe/conv_rows_trace.bin
echo "--- rows on"; timeout 300 python scripts/bench_backbone.py --agents 32 2>&1 | tail -3
