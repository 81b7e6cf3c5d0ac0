#!/bin/bash
timeout 300 python -m pytest tests/test_backbone_gpu.py tests/test_det_tail_gpu.py tests/test_heter_model_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 120 scripts/probe/conv_tma_trace.bin
timeout 300 python scripts/probe/post_anomaly.py 2>&1 | tail -4 | head -2
