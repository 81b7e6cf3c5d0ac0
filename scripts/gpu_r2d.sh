#!/bin/bash
# r02d: new bench (GenComm frame headline) at N=1, reference arm, backbone vs cuDNN TF32/bf16
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python bench.py --steps 20 --warmup 3 2>$OUT/bench_r02d.err | tee $OUT/bench_r02d.json | cut -c1-3000
tail -5 $OUT/bench_r02d.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 --cpu-protocol 2>$OUT/bench_ref_r02d.err | tee $OUT/bench_ref_r02d.json | cut -c1-3000
tail -3 $OUT/bench_ref_r02d.err
timeout 300 python scripts/bench_backbone.py --torch --agents 4 2>&1 | tail -1 | tee $OUT/bench_backbone_r02d.json
timeout 300 python scripts/bench_backbone.py --torch --agents 32 2>&1 | tail -1 | tee -a $OUT/bench_backbone_r02d.json
