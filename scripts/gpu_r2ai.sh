#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/probe/post_anomaly.py 2>&1 | tail -4 | head -2
