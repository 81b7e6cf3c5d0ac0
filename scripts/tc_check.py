#!/usr/bin/env python
"""tcgen05 denoiser check: conv_in / conv_out tensor-core kernels against the fp32 CUDA-core path on the
same inputs (per layer selection + the validation 'materialize' operand mode), then sampler timings.

    python scripts/tc_check.py [--agents 4] [--C 128]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gencomm_b200 as G  # noqa: E402
from gencomm_b200 import ops, synth  # noqa: E402


def cfg(C):
    return {"model": {"embed_dim": C + 2, "in_channels": C, "out_ch": C, "ch": 8, "ch_mult": [1, 1],
                      "num_res_blocks": 2, "attn_resolutions": [16], "dropout": 0.0, "resamp_with_conv": True},
            "diffusion": {"beta_schedule": "linear", "beta_start": 0.0005, "beta_end": 0.02,
                          "num_diffusion_timesteps": 3}}


def rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30), \
        ((a - b).abs().mean() / b.abs().mean().clamp_min(1e-30)).item()


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=4)
    ap.add_argument("--C", type=int, default=128)
    ap.add_argument("--H", type=int, default=64)
    ap.add_argument("--W", type=int, default=128)
    args = ap.parse_args()
    A, C, H, W = args.agents, args.C, args.H, args.W
    torch.manual_seed(0)
    m = G.GenComm(cfg(C))
    with torch.no_grad():
        for name, p in m.named_parameters():
            if "norm" in name:
                p.add_(0.2 * torch.randn_like(p))
            elif name.endswith(".bias"):
                p.add_(0.05 * torch.randn_like(p))
    m = m.cuda().eval()
    feat = synth.bev_features(30, A, C, H, W).cuda()
    cond = synth.bev_features(30, A, 2, H, W, salt=4).cuda()
    n0, steps = synth.sampler_noise(30, A, C, H, W, T=3)
    noise = (n0.cuda(), torch.stack(steps).cuda())
    rl = torch.tensor([A], dtype=torch.int64, device="cuda")
    x = torch.cat([cond, feat], dim=1)
    t = torch.full((A,), 1, device="cuda")

    m.denoiser.precision = ops.PREC_F32
    ref_u = m.denoiser(x, t)
    ref_s = m(feat, cond, rl, noise=noise)["pred_feature"]
    torch.cuda.synchronize()
    names = {1: "conv_in tc", 2: "conv_out tc", 2 | 4: "conv_out tc (materialized A)", 3: "both tc", 8: "middle tf32 tc", 11: "all tc"}
    for prec, name in names.items():
        m.denoiser.precision = prec
        u = m.denoiser(x, t)
        torch.cuda.synchronize()
        print(f"unet   {name:32s} max-rel {rel(u, ref_u)[0]:.3e} mean-rel {rel(u, ref_u)[1]:.3e} finite={bool(torch.isfinite(u).all())}", flush=True)
    m.denoiser.precision = ops.PREC_BF16_TC
    s = m(feat, cond, rl, noise=noise)["pred_feature"]
    torch.cuda.synchronize()
    print(f"sampler bf16-tc vs fp32: max-rel {rel(s, ref_s)[0]:.3e} mean-rel {rel(s, ref_s)[1]:.3e}", flush=True)
    m.denoiser.precision = ops.PREC_TC_ALL
    s = m(feat, cond, rl, noise=noise)["pred_feature"]
    torch.cuda.synchronize()
    print(f"sampler all-tc vs fp32: max-rel {rel(s, ref_s)[0]:.3e} mean-rel {rel(s, ref_s)[1]:.3e}", flush=True)
    for prec, name in ((0, "fp32"), (3, "bf16 tc"), (11, "all tc")):
        m.denoiser.precision = prec
        ms = timeit(lambda: m(feat, cond, rl, noise=noise))
        print(f"sampler {name:8s} A={A} C={C} {H}x{W}: {ms:.3f} ms / call ({ms / A:.3f} ms per agent)", flush=True)


if __name__ == "__main__":
    main()
