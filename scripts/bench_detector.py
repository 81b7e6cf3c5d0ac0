#!/usr/bin/env python
"""Whole-detector microbench: HeterModelBaselineWGenComm (stage-1 GenComm detector, m1_att.yaml model args) from raw
LiDAR points to decoded, NMS-filtered boxes, F frames of N agents per call, per-stage CUDA-event breakdown.

    python scripts/bench_detector.py [--frames 8] [--agents 4] [--points 100000] [--iters 5] [--fusion att]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gencomm_b200 as G  # noqa: E402
from gencomm_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--agents", type=int, default=4)
    ap.add_argument("--points", type=int, default=100_000)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--fusion", default="att")
    ap.add_argument("--precision", default="tc")
    args = ap.parse_args()
    F, N = args.frames, args.agents
    m = G.HeterModelBaselineWGenComm(synth.gencomm_stage1_args(args.fusion))
    m.load_state_dict(synth.fill_state_dict(m.state_dict(), 11))
    m = m.cuda().eval()
    m.gencomm.precision = args.precision
    clouds, pairwise = synth.heter_frames(7000, [N] * F, args.points)
    off = np.concatenate([[0], np.cumsum([len(c) for c in clouds])]).astype(np.int32)
    data = {"inputs_m1": {"points": torch.from_numpy(np.concatenate(clouds)).cuda(),
                          "point_offsets": torch.from_numpy(off).cuda(), "max_agent_points": args.points},
            "agent_modality_list": ["m1"] * (F * N), "pairwise_t_matrix": torch.from_numpy(pairwise).cuda(),
            "record_len": torch.full((F,), N, dtype=torch.int64).cuda()}

    # per-stage timing through forward hooks on the sub-modules (events on the current stream)
    stages = {"encoder_m1": "pillars", "backbone_m1": "backbone", "shrinker_m1": "shrink",
              "message_extractor_m1": "message_extractor", "gencomm": "sampler", "enhancer": "enhancer",
              "fusion_net": "warp_fuse"}
    marks = []
    for name, label in stages.items():
        mod = getattr(m, name)
        mod.register_forward_pre_hook(lambda _m, _i, label=label: marks.append((label, 0, _ev())))
        mod.register_forward_hook(lambda _m, _i, _o, label=label: marks.append((label, 1, _ev())))
    pp = G.VoxelPostprocessor(synth.postprocess_params(score_threshold=0.6), train=False)
    anchors = torch.from_numpy(pp.generate_anchor_box()).float().cuda()

    def run():
        out = m(dict(data))
        marks.append(("postprocess", 0, _ev()))
        det = pp.post_process_batch(out["cls_preds"], out["reg_preds"], out["dir_preds"], anchors)
        marks.append(("postprocess", 1, _ev()))
        return out, det

    for _ in range(2):
        out, det = run()
    torch.cuda.synchronize()
    marks.clear()
    e0, e1 = _ev(sync=False), None
    for _ in range(args.iters):
        out, det = run()
    e1 = _ev(sync=False)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    per = {}
    for (l0, k0, a), (l1, k1, b) in zip(marks[0::2], marks[1::2]):
        assert l0 == l1 and k0 == 0 and k1 == 1
        per[l0] = per.get(l0, 0.0) + a.elapsed_time(b) / args.iters
    per["heads_and_glue"] = ms - sum(per.values())
    print(json.dumps({"workload": f"HeterModelBaselineWGenComm m1_att, {F} frames x {N} agents x {args.points} points, "
                                  f"OPV2V-H grid, {args.fusion} fusion, sampler {args.precision}",
                      "ms_per_call": ms, "frames_per_s": F / ms * 1e3, "stage_ms": {k: round(v, 3) for k, v in per.items()},
                      "cls_preds": list(out["cls_preds"].shape), "detections_per_frame": det[2].tolist()}))


def _ev(sync=False):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


if __name__ == "__main__":
    main()
