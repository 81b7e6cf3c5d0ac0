#!/bin/bash
# BaseBEVBackbone against torch / cuDNN (TF32 default, strict fp32, bf16 channels_last) at 4 and 32 agents -> one JSON per line
OUT=gpurun_out; mkdir -p $OUT
for a in 4 32; do timeout 600 python scripts/bench_backbone.py --agents $a --torch 2>/dev/null | tail -1; done | tee $OUT/backbone_vs_cudnn.jsonl | cut -c1-900
