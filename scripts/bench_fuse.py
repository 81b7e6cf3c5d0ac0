#!/usr/bin/env python
"""Microbench of gc_warp_fuse (BASELINE config[4] sweep shapes): time, achieved algorithmic GB/s, and a
max-abs comparison of the tiled TMA path against the gather kernels on the same inputs.

    python scripts/bench_fuse.py [--quick] [--mode att|max|warp] [--frames F]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gencomm_b200 import ops, synth  # noqa: E402

MODES = {"warp": ops.FUSE_WARP_ONLY, "max": ops.FUSE_MAX, "att": ops.FUSE_ATT}


def inputs(F, N, C, H, W, L=5, seed=0):
    dev = "cuda"
    feat = torch.randn(F * N, C, H, W, device=dev)
    pw = np.stack([synth.pairwise_t_matrix(seed + f, N, max(L, N), spread=(0.3 * W * 0.4, 0.3 * H * 0.4)) for f in range(F)])
    theta = ops.normalize_pairwise_tfm(torch.from_numpy(pw).to(dev), H * 0.4, W * 0.4, 1.0)
    off = torch.arange(0, F * N + 1, N, dtype=torch.int32, device=dev)
    return feat, off, theta


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(iters):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--mode", default="all")
    ap.add_argument("--frames", type=int, default=8)
    args = ap.parse_args()
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
    shapes = [(4, 64, 256, 256), (4, 128, 64, 128), (5, 256, 64, 128)]
    if not args.quick:
        shapes += [(2, 64, 256, 256), (8, 64, 256, 256), (4, 64, 256, 512), (4, 128, 256, 256), (4, 256, 256, 256), (5, 64, 512, 512),
                   (8, 128, 128, 256), (8, 256, 512, 512)]
    modes = list(MODES) if args.mode == "all" else [args.mode]
    print(f"{'mode':5s} {'N':>2s} {'C':>4s} {'H':>4s} {'W':>4s} {'F':>3s} {'tile ms':>9s} {'GB/s':>8s} {'frac':>6s} {'gather ms':>10s} {'maxdiff':>9s}")
    for N, C, H, W in shapes:
        F = max(1, min(args.frames, int(2.0e9 // (N * C * H * W * 4))))
        feat, off, theta = inputs(F, N, C, H, W)
        for m in modes:
            mode = MODES[m]
            lead = F * N if mode == ops.FUSE_WARP_ONLY else F
            out = torch.empty(lead, C, H, W, device="cuda")
            os.environ["GC_WARP_FUSE_GATHER"] = "0"
            t_tile = timeit(lambda: ops.warp_fuse(feat, off, theta, mode, out=out, max_agents=N))
            res_tile = out.clone()
            os.environ["GC_WARP_FUSE_GATHER"] = "1"
            t_gather = timeit(lambda: ops.warp_fuse(feat, off, theta, mode, out=out, max_agents=N), iters=3, warm=1)
            diff = (res_tile - out).abs().max().item()
            os.environ["GC_WARP_FUSE_GATHER"] = "0"
            nbytes = 4 * F * N * C * H * W + 4 * lead * C * H * W
            gbs = nbytes / (t_tile * 1e-3) / 1e9
            print(f"{m:5s} {N:2d} {C:4d} {H:4d} {W:4d} {F:3d} {t_tile:9.4f} {gbs:8.1f} {gbs / peak:6.3f} {t_gather:10.4f} {diff:9.2e}",
                  flush=True)


if __name__ == "__main__":
    main()
