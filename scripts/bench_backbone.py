#!/usr/bin/env python
"""Microbench of the tcgen05 BaseBEVBackbone (GenComm m1 config) on 256x512x64 canvases.

    python scripts/bench_backbone.py [--agents 4] [--iters 5]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gencomm_b200 import BaseBEVBackbone  # noqa: E402

CFG = {"layer_nums": [3, 5, 8], "layer_strides": [2, 2, 2], "num_filters": [64, 128, 256],
       "upsample_strides": [1, 2, 4], "num_upsample_filter": [128, 128, 128]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=4)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--torch", action="store_true", help="also time the same layers through torch/cuDNN (fp32, TF32 off)")
    args = ap.parse_args()
    torch.manual_seed(0)
    m = BaseBEVBackbone(CFG, 64).cuda().eval()
    x = torch.randn(args.agents, 64, 256, 512, device="cuda")
    d = {"spatial_features": x}
    for _ in range(2):
        m(d)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(args.iters):
        m(d)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / args.iters
    px = [128 * 256, 64 * 128, 32 * 64]
    fl = 0.0
    cin = 64
    for n, f, p in zip(CFG["layer_nums"], CFG["num_filters"], px):
        fl += 2.0 * p * 9 * (cin * f + n * f * f)
        cin = f
    fl += 2.0 * (px[0] * 64 * 128 + px[1] * 128 * 128 * 4 + px[2] * 256 * 128 * 16)
    out = {"workload": f"BaseBEVBackbone m1, {args.agents} agents, 64x256x512 -> 384x128x256", "ms_per_call": ms,
           "agents_per_s": args.agents / ms * 1e3, "tflops": fl * args.agents / ms / 1e9, "launches_per_call": 1 + 19 + 21}
    if args.torch:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

        @torch.no_grad()
        def f():   # the module tree holds the reference's own torch layers (blocks / deblocks): this is the cuDNN path
            y, ups = x, []
            for blk, deb in zip(m.blocks, m.deblocks):
                y = blk(y)
                ups.append(deb(y))
            return torch.cat(ups, dim=1)
        for _ in range(2):
            f()
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(args.iters):
            f()
        ev[1].record()
        torch.cuda.synchronize()
        out["torch_cudnn_fp32_ms"] = ev[0].elapsed_time(ev[1]) / args.iters
    print(json.dumps(out))


if __name__ == "__main__":
    main()
