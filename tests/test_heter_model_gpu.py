"""GPU parity of the full stage-1 / stage-2 GenComm detectors (``HeterModelBaselineWGenComm`` /
``HeterModelBaselineWDiffCommStage2``) against the golden outputs of the UNMODIFIED reference class and the oracle, on
the shipped configuration (m1_att.yaml model args, full OPV2V-H grid, 2 + 1 agents).

Tolerances, stated per output (the path mixes fp32-grade bf16x3 tensor-core GEMMs -- backbone, shrink header, Enhancer,
heads -- with the plain-bf16 deformable layer of MessageExtractorv2, whose own stated bound is 2e-2 / 1e-2,
tests/test_message_extractor_gpu.py, and which conditions the sampler):
  gt_feature (encoder + backbone + shrink):  max|d| <= 1e-3 * max|ref|, mean|d| <= 3e-4 * mean|ref|  (measured 8e-5 / 8e-5)
  message, pred_feature, heads; fp32 sampler: max|d| <= 2e-2 * max|ref|, mean|d| <= 1e-2 * mean|ref|  (measured: message
                                              6.6e-3 / 6.8e-3, pred_feature 4.0e-3 / 3.2e-3, heads 4.8e-3 / 3.5e-3)
  same, sampler on tensor cores ('tc'):       max|d| <= 5e-2 * max|ref|, mean|d| <= 3e-2 * mean|ref|  (bf16 conv_in / conv_out,
                                              tf32 middle layers; measured heads 1.5e-2 / 1.1e-2)
The pillar canvas under it is bit-exact (tests/test_pillars_gpu.py).
"""
import pytest
import torch

import gencomm_b200 as G
from gencomm_b200 import synth
from oracle import gen_golden
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
T = torch.from_numpy
DEV = "cuda"


def _err(got, ref):
    assert got.shape == ref.shape, (got.shape, ref.shape)
    d = (got - ref).abs()
    return float(d.max()) / float(ref.abs().max()), float(d.mean()) / float(ref.abs().mean())


def _close(got, ref, what, tmax, tmean):
    emax, emean = _err(got, ref)
    print(f"{what}: max {emax:.2e} mean {emean:.2e}")
    assert emax <= tmax and emean <= tmean, (what, emax, emean)


def _model(args=None, cls=None, seed=gen_golden.HETER_WSEED):
    m = (cls or G.HeterModelBaselineWGenComm)(args or synth.gencomm_stage1_args("att"))
    m.load_state_dict(synth.fill_state_dict(m.state_dict(), seed))
    return m.to(DEV).eval()


def _data(heter_inputs):
    voxels, pairwise, record_len, n0, steps = heter_inputs
    return {"inputs_m1": {k: v.to(DEV) for k, v in voxels.items()},
            "agent_modality_list": ["m1"] * int(record_len.sum()), "pairwise_t_matrix": pairwise.to(DEV),
            "record_len": record_len.to(DEV), "gencomm_noise": (n0.to(DEV), torch.stack(list(steps)).to(DEV))}


def _check(out, g, tmax, tmean):
    errs = {k: _err(out[k].cpu(), T(g[k])) for k in ("cls_preds", "reg_preds", "dir_preds", "message")}
    errs.update({k: _err(out[k][:, ::8].cpu(), T(g[k + "_c8"])) for k in ("gt_feature", "pred_feature")})
    print({k: (f"{a:.2e}", f"{b:.2e}") for k, (a, b) in errs.items()})
    bad = {k: v for k, v in errs.items() if v[0] > tmax or v[1] > tmean}
    if errs["gt_feature"][0] > 1e-3 or errs["gt_feature"][1] > 3e-4:
        bad["gt_feature"] = errs["gt_feature"]
    assert not bad, bad


def test_stage1_detector_matches_reference_fp32_sampler(golden_heter_model, heter_inputs):
    m = _model()
    m.gencomm.precision = "fp32"
    out = m(_data(heter_inputs))
    assert set(out) == {"cls_preds", "reg_preds", "dir_preds", "gt_feature", "pred_feature", "message"}
    assert out["cls_preds"].shape == (2, 2, 64, 128) and out["reg_preds"].shape == (2, 14, 64, 128)
    _check(out, golden_heter_model, 2e-2, 1e-2)


def test_stage1_detector_matches_reference_default_precision(golden_heter_model, heter_inputs):
    out = _model()(_data(heter_inputs))
    _check(out, golden_heter_model, 5e-2, 3e-2)


def test_raw_point_inputs_equal_voxel_inputs(heter_inputs):
    """Extension: inputs_m1 = raw points + offsets (fused voxelize -> PFN -> canvas kernels) gives the same detector
    output as the spconv-style voxel tensors (the canvas is bit-exact either way)."""
    import numpy as np
    m = _model()
    m.gencomm.precision = "fp32"
    data = _data(heter_inputs)
    ref = m(dict(data))
    clouds, _ = synth.heter_frames(gen_golden.HETER_SEED, gen_golden.HETER_RECORD_LEN, gen_golden.HETER_POINTS)
    off = np.concatenate([[0], np.cumsum([len(c) for c in clouds])]).astype(np.int32)
    data["inputs_m1"] = {"points": T(np.concatenate(clouds)).to(DEV), "point_offsets": T(off).to(DEV),
                         "max_agent_points": max(len(c) for c in clouds)}
    out = m(data)
    assert torch.equal(out["gt_feature"], ref["gt_feature"])
    assert torch.equal(out["cls_preds"], ref["cls_preds"])


def test_stage2_detector_max_fusion_trick_matches_oracle(heter_inputs):
    """Stage-2 class (args['diffcomm'], trick mask) with max fusion, against the oracle composition."""
    args = synth.gencomm_stage1_args("max")
    args["diffcomm"] = args.pop("gencomm")
    args["trick"] = True
    m = _model(args, G.HeterModelBaselineWDiffCommStage2, seed=23)
    m.gencomm.precision = "fp32"
    out = m(_data(heter_inputs))
    voxels, pairwise, record_len, n0, steps = heter_inputs
    oargs = synth.gencomm_stage1_args("max")
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    # the oracle composition with the stage-2 mask (…_stage2.py:284-285,293-294) applied between sampler and enhancer
    ref = R.heter_gencomm_forward(sd, oargs, voxels, pairwise, record_len, n0, steps,
                                  mask_generated=True)
    for k in ("cls_preds", "reg_preds", "dir_preds"):
        _close(out[k].cpu(), ref[k], "stage2 " + k, 2e-2, 1e-2)
    _close(out["pred_feature"].cpu(), ref["pred_feature"], "stage2 pred_feature", 2e-2, 1e-2)


def test_detector_outputs_through_gpu_postprocess_match_oracle_postprocess(heter_inputs):
    """points -> detector -> decode + NMS entirely on the GPU; the oracle post-processes the same head maps on the host
    (both frames of the batch in one gc_postprocess call)."""
    m = _model()
    out = m(_data(heter_inputs))
    params = synth.postprocess_params(score_threshold=0.6)
    pp = G.VoxelPostprocessor(params, train=False)
    anchors = pp.generate_anchor_box()
    boxes, scores, counts = pp.post_process_batch(out["cls_preds"], out["reg_preds"], out["dir_preds"], anchors)
    for f in range(2):
        ref_b, ref_s = R.post_process(out["cls_preds"][f:f + 1].cpu(), out["reg_preds"][f:f + 1].cpu(),
                                      out["dir_preds"][f:f + 1].cpu(), T(anchors), torch.eye(4), params)
        k = int(counts[f])
        assert ref_b is not None and k == ref_b.shape[0] and k > 10, (f, k)
        assert float((boxes[f, :k].cpu() - ref_b).abs().max()) <= 1e-4
        assert float((scores[f, :k].cpu() - ref_s).abs().max()) <= 1e-6


def test_detector_is_inference_only():
    m = _model()
    m.train()
    with pytest.raises(RuntimeError, match="inference-only"):
        m({})


def _data2(heter2_inputs):
    voxels, bev, pairwise, record_len, n0, steps = heter2_inputs
    return {"inputs_m1": {k: v.to(DEV) for k, v in voxels.items()}, "inputs_m2": {"bev_feature": bev.to(DEV)},
            "agent_modality_list": list(gen_golden.HETER2_MODALITIES), "pairwise_t_matrix": pairwise.to(DEV),
            "record_len": record_len.to(DEV), "gencomm_noise": (n0.to(DEV), torch.stack(list(steps)).to(DEV))}


@pytest.mark.parametrize("precision,tmax,tmean", [("fp32", 2e-2, 1e-2), (None, 5e-2, 3e-2)])
def test_stage2_lidar_camera_detector_matches_unmodified_reference(golden_heter_model_stage2, heter2_inputs, precision, tmax, tmean):
    """HeterModelBaselineWDiffCommStage2 with LiDAR (m1) + camera (m2) agents, 3 + 2 agents with interleaved modalities,
    against the UNMODIFIED reference class (only its EfficientNet image encoder stubbed to return the injected BEV
    feature): camera backbone (inplanes 128) + shrink + message extractor, CenterCrop zero-padding of the camera
    features and messages to 64 x 128, per-agent re-assembly, sampler, Enhancer, AttFusion, heads."""
    m = _model(synth.gencomm_stage2_hetero_args("att"), G.HeterModelBaselineWDiffCommStage2, seed=gen_golden.HETER2_WSEED)
    if precision is not None:
        m.gencomm.precision = precision
    out = m(_data2(heter2_inputs))
    g = golden_heter_model_stage2
    assert out["gt_feature"].shape == (5, 128, 64, 128) and out["message"].shape == (5, 2, 64, 128)
    # the camera agents (1 and 4) carry exact zeros outside the central 64 columns, features and messages alike
    for a in (1, 4):
        assert float(out["gt_feature"][a][:, :, :32].abs().max()) == 0.0 and float(out["gt_feature"][a][:, :, 96:].abs().max()) == 0.0
        assert float(out["message"][a][:, :, :32].abs().max()) == 0.0 and float(out["message"][a][:, :, 96:].abs().max()) == 0.0
    _check(out, g, tmax, tmean)


def _match_detections(boxes_a, scores_a, boxes_b, scores_b):
    """Greedy one-to-one matching of two detection sets by bottom-face centre distance; returns (matched pairs,
    max centre distance, max |score difference|) over the pairs closer than 0.25 m."""
    ca, cb = boxes_a[:, :4, :2].mean(1), boxes_b[:, :4, :2].mean(1)
    d = torch.cdist(ca, cb)
    pairs, used = [], set()
    for i in torch.argsort(scores_a, descending=True).tolist():
        j = int(torch.argmin(d[i]))
        if float(d[i, j]) < 0.25 and j not in used:
            used.add(j)
            pairs.append((i, j))
    if not pairs:
        return pairs, 0.0, 0.0
    ia, ib = torch.tensor([p[0] for p in pairs]), torch.tensor([p[1] for p in pairs])
    return pairs, float(d[ia, ib].max()), float((scores_a[ia] - scores_b[ib]).abs().max())


@pytest.mark.parametrize("precision,min_match,max_dist,max_score", [("fp32", 0.98, 0.15, 0.01), (None, 0.96, 0.25, 0.02)])
def test_detections_match_reference_fp32_heads(golden_heter_model, heter_inputs, precision, min_match, max_dist, max_score):
    """Detection-level parity: the head maps of the UNMODIFIED fp32 reference model (golden) and the GPU detector's own head
    maps go through the same post-processor at the shipped thresholds (score 0.2, NMS 0.15).  The synthetic-weight scene is
    dense clutter (~750 kept boxes per frame, many overlapping near the NMS threshold), i.e. a worst case for greedy-NMS
    stability.  Stated bound: kept counts within 2 %; a one-to-one matched GPU detection within ``max_dist`` m of the centre
    and ``max_score`` of the score for at least ``min_match`` of the reference detections (the rest are NMS decision flips
    between overlapping neighbours).  Measured on the B200:
      fp32 sampler            746 / 752 and 734 / 739 matched (99.2 %), centre <= 0.08 m, score <= 0.0023
      default (tensor cores)  733 / 752 and 721 / 739 matched (97.5 %), centre <= 0.17 m, score <= 0.0073"""
    g = golden_heter_model
    m = _model()
    if precision is not None:
        m.gencomm.precision = precision
    out = m(_data(heter_inputs))
    params = synth.postprocess_params(score_threshold=0.2)
    pp = G.VoxelPostprocessor(params, train=False)
    anchors = pp.generate_anchor_box()
    ref = pp.post_process_batch(T(g["cls_preds"]).to(DEV), T(g["reg_preds"]).to(DEV), T(g["dir_preds"]).to(DEV), anchors)
    got = pp.post_process_batch(out["cls_preds"], out["reg_preds"], out["dir_preds"], anchors)
    for f in range(2):
        kr, kg = int(ref[2][f]), int(got[2][f])
        rb, rs = ref[0][f, :kr].cpu(), ref[1][f, :kr].cpu()
        gb, gs = got[0][f, :kg].cpu(), got[1][f, :kg].cpu()
        pairs, dmax, smax = _match_detections(rb, rs, gb, gs)
        print(f"precision {precision or 'default'} frame {f}: reference {kr} detections, gpu {kg}; matched {len(pairs)} "
              f"({100.0 * len(pairs) / kr:.1f} %), max centre distance {dmax:.3f} m, max score diff {smax:.4f}")
        assert kr > 20 and abs(kr - kg) <= max(3, 0.02 * kr)
        assert len(pairs) >= min_match * kr
        assert dmax <= max_dist and smax <= max_score
