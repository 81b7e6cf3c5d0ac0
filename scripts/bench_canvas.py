#!/usr/bin/env python
"""Microbench of the fused PFN + scatter canvas writer (gc_pillar_canvas) on the bench workload
(F frames x 4 agents x 100k LiDAR-like points, 256x256x64 canvas): time, algorithmic GB/s, and a
bit-exact comparison of the persistent writer against the per-tile kernel on the same workspace.

    python scripts/bench_canvas.py [--frames F] [--grid opv2v|square] [--uniform]

GC_CANVAS_IMPL=tile selects the per-tile kernel, GC_CANVAS_CFG=k a CTA shape of the persistent writer; both are
read once per process, so A/B runs are separate processes (scripts/gpu_canvas_ab.sh).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gencomm_b200 import ops, pipeline, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--agents", type=int, default=4)
    ap.add_argument("--points", type=int, default=100_000)
    ap.add_argument("--grid", default="square")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--dump", default="")
    args = ap.parse_args()
    rng = [-51.2, -51.2, -3, 51.2, 51.2, 1] if args.grid == "square" else [-102.4, -51.2, -3, 102.4, 51.2, 1]
    pipe = pipeline.FramePipeline(args.frames, args.agents, args.points, rng, [0.4, 0.4, 4], fusion="max",
                                  pfn=synth.pfn_weights(0))
    sets = []
    for s in range(3):
        pts, _ = pipeline.synthetic_step_inputs(s, args.frames, args.agents, args.points, rng)
        sets.append(torch.from_numpy(pts).cuda())
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
    for k in range(3):
        pipe.encode(sets[k % 3])
    torch.cuda.synchronize()
    for k in range(args.iters):
        pipe.encode(sets[k % 3], events=ev[k])
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
    nbytes = pipe.scatter_bytes()
    peak_path = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
    peak = json.load(open(peak_path))["hbm_gbs"] if os.path.exists(peak_path) else 6548.0
    canvas = pipe.encode(sets[0]).clone()
    torch.cuda.synchronize()
    occ = float((canvas.abs().sum(1) > 0).float().mean())
    out = {"impl": os.environ.get("GC_CANVAS_IMPL", "persist"), "cfg": os.environ.get("GC_CANVAS_CFG", "0"),
           "grid": [pipe.nx, pipe.ny], "frames": args.frames, "ms": ms, "gbs": nbytes / ms / 1e6,
           "frac": nbytes / ms / 1e6 / peak, "occupied_cells": occ,
           "checksum": float(canvas.double().sum()), "abs_checksum": float(canvas.double().abs().sum())}
    if args.dump:
        torch.save(canvas.cpu(), args.dump)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
