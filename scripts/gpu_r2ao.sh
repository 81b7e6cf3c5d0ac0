#!/bin/bash
timeout 900 python -m pytest tests/test_pillars_gpu.py tests/test_enhancer_gpu.py tests/test_heter_model_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 300 python scripts/probe/post_anomaly.py 2>&1 | tail -4 | head -2
