#!/bin/bash
# r02s: full GPU suite, smoke, bench (N=1), sampler bench with graph timing, ncu of the sampler kernels
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -6 | tee $OUT/pytest_r02s.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_r02s.log
timeout 300 python scripts/bench_sampler.py --iters 20 2>&1 | tee $OUT/bench_sampler_r02s.txt
timeout 300 python scripts/bench_sampler.py --iters 20 --frames 1 --precision cluster 2>&1 | tee -a $OUT/bench_sampler_r02s.txt
timeout 300 python scripts/bench_sampler.py --iters 20 --frames 8 --agents 5 --C 256 --precision cluster 2>&1 | tee -a $OUT/bench_sampler_r02s.txt
timeout 900 python bench.py --steps 20 --warmup 3 2>$OUT/bench_r02s.err | tee $OUT/bench_r02s.json | cut -c1-600
tail -3 $OUT/bench_r02s.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_unet_middle|k_conv_in_tc2|k_conv_out_tc" -s 3 -c 3 -f -o $OUT/prof_sampler_r02s \
    python scripts/bench_sampler.py --iters 1 --precision cluster > $OUT/ncu_sampler_r02s.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_sampler_r02s.csv \
    python scripts/bench_sampler.py --iters 1 --precision cluster > /dev/null 2>&1
