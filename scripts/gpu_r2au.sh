#!/bin/bash
timeout 600 python -m pytest tests/test_backbone_gpu.py tests/test_det_tail_gpu.py tests/test_enhancer_gpu.py tests/test_heter_model_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 120 python scripts/bench_backbone.py --agents 32 2>&1 | tail -1
GC_CONV_TMAR=0 timeout 120 python scripts/bench_backbone.py --agents 32 2>&1 | tail -1
timeout 200 bash scripts/gpu_r2ac.sh
timeout 300 python scripts/probe/post_anomaly.py 2>&1 | tail -4 | head -2
