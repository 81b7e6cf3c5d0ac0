#!/bin/bash
# r02g: full GPU test suite incl. the new stage-2 / detection-level / deterministic-LSS tests
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -60 | tee $OUT/pytest_r02g.log
