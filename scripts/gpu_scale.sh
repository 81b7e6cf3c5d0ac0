#!/bin/bash
# usage: bash scripts/gpu_scale.sh N tag   (under gpurun --gpus N)
N=$1; TAG=$2; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 2>$OUT/bench${N}_$TAG.err | tee $OUT/bench${N}_$TAG.json | cut -c1-1200
tail -5 $OUT/bench${N}_$TAG.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 3 --shape v2xreal 2>$OUT/bench${N}_v2x_$TAG.err | tee $OUT/bench${N}_v2x_$TAG.json | cut -c1-1200
tail -3 $OUT/bench${N}_v2x_$TAG.err
