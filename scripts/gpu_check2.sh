#!/bin/bash
timeout 600 python -m pytest tests/test_backbone_gpu.py tests/test_det_tail_gpu.py tests/test_enhancer_gpu.py tests/test_heter_model_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 120 scripts/probe/conv_tma_trace.bin 2>&1 | cut -c1-210
timeout 120 python scripts/bench_backbone.py --agents 32 2>&1 | tail -1
timeout 300 python scripts/probe/post_anomaly.py 2>&1 | tail -4 | head -2
