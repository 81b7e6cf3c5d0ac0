#!/usr/bin/env python
"""Microbench of the tcgen05 BaseBEVBackbone (GenComm m1 config) on 256x512x64 canvases.

    python scripts/bench_backbone.py [--agents 4] [--iters 5]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gencomm_b200 import BaseBEVBackbone  # noqa: E402

CFG = {"layer_nums": [3, 5, 8], "layer_strides": [2, 2, 2], "num_filters": [64, 128, 256],
       "upsample_strides": [1, 2, 4], "num_upsample_filter": [128, 128, 128]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=4)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--torch", action="store_true", help="also time the same layers through torch/cuDNN: TF32 (PyTorch default), strict fp32, bf16 channels_last")
    args = ap.parse_args()
    torch.manual_seed(0)
    m = BaseBEVBackbone(CFG, 64).cuda().eval()
    x = torch.randn(args.agents, 64, 256, 512, device="cuda")
    d = {"spatial_features": x}
    for _ in range(2):
        m(d)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(args.iters):
        m(d)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / args.iters
    px = [128 * 256, 64 * 128, 32 * 64]
    fl = 0.0
    cin = 64
    for n, f, p in zip(CFG["layer_nums"], CFG["num_filters"], px):
        fl += 2.0 * p * 9 * (cin * f + n * f * f)
        cin = f
    fl += 2.0 * (px[0] * 64 * 128 + px[1] * 128 * 128 * 4 + px[2] * 256 * 128 * 16)
    out = {"workload": f"BaseBEVBackbone m1, {args.agents} agents, 64x256x512 -> 384x128x256", "ms_per_call": ms,
           "agents_per_s": args.agents / ms * 1e3, "tflops": fl * args.agents / ms / 1e9, "launches_per_call": 1 + 19 + 3}
    if args.torch:
        @torch.no_grad()
        def f(inp, blocks, deblocks):   # the module tree holds the reference's own torch layers (blocks / deblocks): the cuDNN path
            y, ups = inp, []
            for blk, deb in zip(blocks, deblocks):
                y = blk(y)
                ups.append(deb(y))
            return torch.cat(ups, dim=1)

        def timed(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(args.iters):
                fn()
            ev[1].record()
            torch.cuda.synchronize()
            return ev[0].elapsed_time(ev[1]) / args.iters

        torch.backends.cudnn.benchmark = True
        ref32 = None
        # (1) PyTorch defaults = what the reference runs (it never touches allow_tf32: cuDNN convolutions use TF32)
        torch.backends.cudnn.allow_tf32 = True
        out["torch_cudnn_tf32_default_ms"] = timed(lambda: f(x, m.blocks, m.deblocks))
        y_tf32 = f(x, m.blocks, m.deblocks)
        # (2) strict fp32 (TF32 off)
        torch.backends.cudnn.allow_tf32 = False
        out["torch_cudnn_fp32_ms"] = timed(lambda: f(x, m.blocks, m.deblocks))
        ref32 = f(x, m.blocks, m.deblocks)
        torch.backends.cudnn.allow_tf32 = True
        # (3) channels_last bf16 autocast-free: weights + activations in bf16 NHWC (the fastest library configuration)
        import copy
        b16 = copy.deepcopy(m.blocks).to(dtype=torch.bfloat16, memory_format=torch.channels_last)
        d16 = copy.deepcopy(m.deblocks).to(dtype=torch.bfloat16, memory_format=torch.channels_last)
        x16 = x.to(dtype=torch.bfloat16, memory_format=torch.channels_last)
        out["torch_cudnn_bf16_channels_last_ms"] = timed(lambda: f(x16, b16, d16))
        y16 = f(x16, b16, d16).float()
        ours = m(d)["spatial_features_2d"]
        den = ref32.abs().max().item()
        out["max_err_vs_cudnn_fp32"] = {"ours_bf16x3": (ours - ref32).abs().max().item() / den,
                                         "cudnn_tf32": (y_tf32 - ref32).abs().max().item() / den,
                                         "cudnn_bf16": (y16 - ref32).abs().max().item() / den}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
