#!/bin/bash
timeout 600 python -m pytest tests/test_gencomm_gpu.py tests/test_heter_model_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 300 python scripts/bench_sampler.py 2>&1 | tail -6 | cut -c1-400
timeout 300 python scripts/probe/post_anomaly.py 2>&1 | tail -4 | head -2
