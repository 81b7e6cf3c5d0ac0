#!/usr/bin/env python
"""A/B of the two voxelizers behind gc_voxelize (GC_VOXELIZE_IMPL=legacy: k_cell_assign + k_pillar_count +
k_pillar_assign + k_slot_insert; default: k_cell_assign2 + k_pillar_build) on the bench workload: voxelizer alone and
the whole front end (voxelize + PFN + scatter), CUDA events, 3 input sets cycled, bit-exact workspace comparison.

    python scripts/bench_voxelize.py [--frames F] [--grid square|opv2v]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gencomm_b200 import pipeline, synth  # noqa: E402


def timed(fn, iters):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for k in range(3):
        fn(k)
    torch.cuda.synchronize()
    for k in range(iters):
        ev[k][0].record()
        fn(k)
        ev[k][1].record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--agents", type=int, default=4)
    ap.add_argument("--points", type=int, default=100_000)
    ap.add_argument("--grid", default="square")
    ap.add_argument("--iters", type=int, default=30)
    args = ap.parse_args()
    rng = [-51.2, -51.2, -3, 51.2, 51.2, 1] if args.grid == "square" else [-102.4, -51.2, -3, 102.4, 51.2, 1]
    pipe = pipeline.FramePipeline(args.frames, args.agents, args.points, rng, [0.4, 0.4, 4], fusion="max",
                                  pfn=synth.pfn_weights(0))
    sets = []
    for s in range(3):
        pts, _ = pipeline.synthetic_step_inputs(s, args.frames, args.agents, args.points, rng)
        sets.append(torch.from_numpy(pts).cuda())
    peak_path = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
    peak = json.load(open(peak_path))["hbm_gbs"] if os.path.exists(peak_path) else 6548.0
    out = {"grid": [pipe.nx, pipe.ny], "frames": args.frames, "front_end_bytes": pipe.scatter_bytes()}
    sums = {}
    for impl in ("legacy", "fused"):
        os.environ["GC_VOXELIZE_IMPL"] = impl
        vox = timed(lambda k: pipe.pre.voxelize_device(sets[k % 3], pipe.point_offsets, pipe.P), args.iters)
        full = timed(lambda k: pipe.encode(sets[k % 3]), args.iters)
        canvas = pipe.encode(sets[0])
        sums[impl] = (float(canvas.double().sum()), float(canvas.double().abs().sum()),
                      pipe.pre._ws.n_pillars.cpu().tolist())
        out[impl] = {"voxelize_ms": vox, "front_end_ms": full, "front_end_gbs": pipe.scatter_bytes() / full / 1e6,
                     "front_end_frac": pipe.scatter_bytes() / full / 1e6 / peak}
    out["identical"] = sums["legacy"] == sums["fused"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
