#!/usr/bin/env python
"""Microbench of gc_double_conv (GenComm stage-1 shrink header: 384 -> 128, stride 2, 128x256 -> 64x128) and gc_det_heads.

    python scripts/bench_det_tail.py [--frames 8] [--agents 4]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gencomm_b200 import DetectionHeads, DownsampleConv  # noqa: E402


def timeit(fn, iters):
    for _ in range(3):
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
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--agents", type=int, default=4)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    A = args.frames * args.agents
    torch.manual_seed(0)
    shrink = DownsampleConv({"kernal_size": [3], "stride": [2], "padding": [1], "dim": [128], "input_dim": 384}).cuda().eval()
    heads = DetectionHeads(128, 2).cuda().eval()
    x = torch.randn(A, 384, 128, 256, device="cuda")
    f = torch.randn(args.frames, 128, 64, 128, device="cuda")
    ms_s = timeit(lambda: shrink(x), args.iters)
    ms_h = timeit(lambda: heads(f), args.iters)
    fl_s = 2.0 * A * 64 * 128 * 128 * 9 * (384 + 128)
    fl_h = 2.0 * args.frames * 64 * 128 * 128 * 20
    print(json.dumps({"shrink_header": {"workload": f"{A} agents, 384x128x256 -> 128x64x128, stride 2", "ms_per_call": ms_s,
                                        "tflops": fl_s / ms_s / 1e9, "input_gbs": x.numel() * 4 / ms_s / 1e6},
                      "det_heads": {"workload": f"{args.frames} frames, 128x64x128 -> 20 maps", "ms_per_call": ms_h,
                                    "tflops": fl_h / ms_h / 1e9}}))


if __name__ == "__main__":
    main()
