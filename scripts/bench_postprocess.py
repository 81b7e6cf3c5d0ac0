#!/usr/bin/env python
"""Microbench of the GPU decode + rotated NMS (gc_postprocess) on synthetic head maps.

    python scripts/bench_postprocess.py [--frames 8] [--bias -3.0] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gencomm_b200 import VoxelPostprocessor, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--bias", type=float, default=-3.0)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    pp = VoxelPostprocessor(synth.postprocess_params(), train=False)
    anchors = torch.from_numpy(pp.generate_anchor_box()).float().cuda()
    heads = [synth.head_outputs(100 + f, bias=args.bias) for f in range(args.frames)]
    cls, reg, dr = (torch.cat([h[i] for h in heads]).cuda() for i in range(3))
    for _ in range(3):
        boxes, scores, counts = pp.post_process_batch(cls, reg, dr, anchors)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        boxes, scores, counts = pp.post_process_batch(cls, reg, dr, anchors)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    out = {"workload": f"decode + rotated NMS, {args.frames} frames, 64x128x2 anchors, bias {args.bias}",
           "candidates_per_frame": [int((torch.sigmoid(h[0]) > 0.2).sum()) for h in heads][:4],
           "kept_per_frame": counts.tolist()[:4], "ms_per_call": ms, "frames_per_s": args.frames / ms * 1e3}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
