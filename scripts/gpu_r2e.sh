#!/bin/bash
# r02e: AttFusion changes (parity + sweep incl. variant A/B), GPU-eager reference bar
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_warp_fuse_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -5 | tee $OUT/pytest_fuse_r02e.log
timeout 300 python scripts/bench_fuse.py --quick --mode att 2>&1 | tee $OUT/bench_fuse_r02e.txt
echo "--- GC_FUSE_CFG=5 (16x16, two-pass for C>64)" | tee -a $OUT/bench_fuse_r02e.txt
GC_FUSE_CFG=5 timeout 300 python scripts/bench_fuse.py --quick --mode att 2>&1 | tee -a $OUT/bench_fuse_r02e.txt
echo "--- GC_FUSE_CFG=6 (16x8)" | tee -a $OUT/bench_fuse_r02e.txt
GC_FUSE_CFG=6 timeout 300 python scripts/bench_fuse.py --quick --mode att 2>&1 | tee -a $OUT/bench_fuse_r02e.txt
timeout 300 python scripts/bench_fuse.py --quick --mode max 2>&1 | tee -a $OUT/bench_fuse_r02e.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>$OUT/bench_ref_r02e.err | tee $OUT/bench_ref_r02e.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps(d['gpu_eager_baseline'])[:1500])"
tail -3 $OUT/bench_ref_r02e.err
