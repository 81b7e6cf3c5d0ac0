#!/usr/bin/env python
"""GenComm sampler microbench (BASELINE configs[2]/[3] shapes): ms per call for the bf16 tensor-core and fp32 paths.

    python scripts/bench_sampler.py [--frames 8] [--agents 4] [--C 128] [--iters 10] [--precision both|bf16|fp32]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gencomm_b200 as G  # noqa: E402
from gencomm_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--agents", type=int, default=4)
    ap.add_argument("--C", type=int, default=128)
    ap.add_argument("--H", type=int, default=64)
    ap.add_argument("--W", type=int, default=128)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--precision", default="both")
    a = ap.parse_args()
    C, H, W = a.C, a.H, a.W
    torch.manual_seed(0)
    m = G.GenComm({"model": {"embed_dim": C + 2, "in_channels": C, "out_ch": C, "ch": 8, "ch_mult": [1, 1],
                             "num_res_blocks": 2, "attn_resolutions": [16], "dropout": 0.0, "resamp_with_conv": True},
                   "diffusion": {"beta_schedule": "linear", "beta_start": 0.0005, "beta_end": 0.02,
                                 "num_diffusion_timesteps": 3}}).cuda().eval()
    A = a.frames * a.agents
    feat = synth.bev_features(40, A, C, H, W).cuda()
    cond = synth.bev_features(40, A, 2, H, W, salt=4).cuda()
    n0, steps = synth.sampler_noise(40, A, C, H, W, T=3)
    noise = (n0.cuda(), torch.stack(steps).cuda())
    rl = torch.full((a.frames,), a.agents, dtype=torch.int64).cuda()
    for name in (("cluster", "tc", "bf16", "fp32") if a.precision == "both" else (a.precision,)):
        m.precision = name
        for _ in range(3):
            m(feat, cond, rl, noise=noise)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.iters):
            m(feat, cond, rl, noise=noise)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        # the same call replayed from a CUDA graph: device time without the host-side launch path
        gms = float("nan")
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                m(feat, cond, rl, noise=noise)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                m(feat, cond, rl, noise=noise)
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(a.iters):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            gms = e0.elapsed_time(e1) / a.iters
        except Exception as exc:
            print("graph capture failed:", repr(exc))
        flop = 486.8e6 if C == 128 else 788.8e6
        print(f"sampler {name}: {a.frames} frames x {a.agents} agents C={C} {H}x{W}: {ms:.3f} ms/call eager, {gms:.3f} ms/call "
              f"CUDA graph, {a.frames / gms * 1e3:.0f} frames/s, {A * 3 * flop / (gms * 1e-3) / 1e12:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
