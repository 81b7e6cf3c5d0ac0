#!/usr/bin/env python
"""Microbench of the LSS voxel pooling kernel at the OPV2V camera shape (B agents x 4 cameras x 48 depth bins x 48 x 64
feature pixels, C = 64) against its HBM roofline (features read once + grid written once)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gencomm_b200 as G  # noqa: E402
from gencomm_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=4)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    B, N, D, H, W, C = args.agents, 4, 48, 48, 64, 64
    geom, x = synth.lss_frustum(5, B=B, N=N, D=D, H=H, W=W, C=C)
    pool = G.VoxelPooling(synth.LSS_GRID_CONF).cuda()
    geom, x = geom.cuda(), x.cuda()
    for _ in range(3):
        out = pool(geom, x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.iters):
        out = pool(geom, x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    alg = x.numel() * 4 + geom.numel() * 4 + out.numel() * 4
    print(json.dumps({"workload": f"LSS voxel pooling, {B} agents x {N} cams x {D} x {H} x {W} points, C={C} -> {list(out.shape)}",
                      "ms_per_call": ms, "algorithmic_MB": alg / 1e6, "GBps": alg / ms / 1e6}))


if __name__ == "__main__":
    main()
