#!/usr/bin/env python
"""Microbench of gc_enhancer (Enhancer) at the GenComm shapes.

    python scripts/bench_enh.py [--frames 8] [--agents 4] [--C 128] [--H 64] [--W 128]
Prints time per call, the dense FLOP rate and the mandatory HBM traffic rate.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gencomm_b200 import Enhancer, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--agents", type=int, default=4)
    ap.add_argument("--C", type=int, default=128)
    ap.add_argument("--H", type=int, default=64)
    ap.add_argument("--W", type=int, default=128)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    A = args.frames * args.agents
    torch.manual_seed(0)
    model = Enhancer(args.C, [8, 8], 4).cuda().eval()
    xs = [torch.randn(A, args.C, args.H, args.W, device="cuda") for _ in range(3)]
    for k in range(3):
        model(xs[k])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for k in range(args.iters):
        model(xs[k % 3])
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / args.iters
    hw = args.H * args.W
    C = args.C
    flops = 2.0 * A * hw * (9 * (C // 4) ** 2 + C * 4 * C + 2 * C * C + 2 * C * 9)
    nbytes = 4.0 * A * hw * 2 * C   # x read once, out written once
    print(json.dumps({"workload": f"Enhancer {A} agents, C={C}, {args.H}x{args.W}", "ms_per_call": ms,
                      "frames_per_s": args.frames / ms * 1e3, "tflops": flops / ms / 1e9,
                      "mandatory_gbs": nbytes / ms / 1e6, "launches_per_call": 11 + C // 64 + C // 128}))


if __name__ == "__main__":
    main()
