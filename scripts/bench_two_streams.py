#!/usr/bin/env python
"""Throughput of the detector step with one step in flight (bench.py's device-resident loop) against two steps in flight:
two DetectorPipeline instances (own weights, workspaces, planes) on two CUDA streams, consecutive steps alternate.
The tails of one step (NMS: 15 - 128 CTAs, the sampler's 120-CTA cluster launches, every persistent kernel's last wave)
can then be filled by the other step's kernels.  Same box, same process, alternating order.

    python scripts/bench_two_streams.py [--frames 15] [--steps 30]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gencomm_b200 import pipeline, shard  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=15)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--shape", default="opv2v_h")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    F, K = args.frames, args.steps
    N = 5 if args.shape == "v2xreal" else 4
    pipes = [pipeline.DetectorPipeline(F, N, 100_000, shape=args.shape, fusion="att", device=dev) for _ in range(2)]
    sets = []
    for s in range(3):
        p, pw = pipeline.synthetic_detector_inputs(1 + s, F, N, 100_000, pipes[0].lidar_range)
        sets.append((torch.from_numpy(p).to(dev), torch.from_numpy(pw).to(dev)))
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    main_s = torch.cuda.current_stream()

    def serial(k0, n):
        for k in range(k0, k0 + n):
            out = pipes[0].step(*sets[k % 3])
        return out

    def dual(k0, n):
        outs = [None, None]
        for s in streams:
            s.wait_stream(main_s)
        for k in range(k0, k0 + n):
            i = k % 2
            with torch.cuda.stream(streams[i]):
                outs[i] = pipes[i].step(*sets[k % 3])
        for s in streams:
            main_s.wait_stream(s)
        return outs[(k0 + n - 1) % 2]

    def timed(fn):
        fn(0, 4)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn(0, K)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / K, out

    res = {"frames_per_step": F, "steps": K, "shape": args.shape}
    ref = None
    for rep in range(2):
        for name, fn in (("serial", serial), ("two_streams", dual)):
            ms, out = timed(fn)
            res.setdefault(name, []).append({"ms_per_step": ms, "frames_per_s": F / ms * 1e3})
            packed = shard.pack_detections(*out)
            if ref is None:
                ref = packed.clone()
            res.setdefault("same_detections", []).append(bool(torch.equal(packed[:, 0], ref[:, 0])))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
