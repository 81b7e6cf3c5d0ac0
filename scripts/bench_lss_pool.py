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
    ap.add_argument("--coarse", type=float, default=0.0, help="fraction of points snapped onto a 3 m lattice (contention)")
    args = ap.parse_args()
    B, N, D, H, W, C = args.agents, 4, 48, 48, 64, 64
    geom, x = synth.lss_frustum(5, B=B, N=N, D=D, H=H, W=W, C=C, coarse_frac=args.coarse)
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
    from gencomm_b200 import ops
    dx, bx, nx = pool._host
    for _ in range(2):
        ops.lss_voxel_pooling(geom, x, dx, bx, nx, vector=False)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.iters):
        ops.lss_voxel_pooling(geom, x, dx, bx, nx, vector=False)
    e1.record()
    torch.cuda.synchronize()
    ms_scalar = e0.elapsed_time(e1) / args.iters
    print(json.dumps({"scalar_red_ms_per_call": ms_scalar, "workload": f"LSS voxel pooling, {B} agents x {N} cams x {D} x {H} x {W} points, C={C} -> {list(out.shape)}",
                      "coarse_frac": args.coarse, "points_per_occupied_cell": float(x.numel() // C) / max(int((out.abs().sum(1) != 0).sum()), 1), "ms_per_call": ms, "algorithmic_MB": alg / 1e6, "GBps": alg / ms / 1e6}))


if __name__ == "__main__":
    main()
