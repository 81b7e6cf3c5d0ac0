"""Pins the oracle (oracle/ref_ops.py + oracle/pillar_ref.c) against the golden vectors produced by
the unmodified reference classes (oracle/gen_golden.py).  CPU only."""
import numpy as np
import torch

from oracle import ref_ops as R

T = torch.from_numpy


def _split_points(g):
    return [g["points0"], g["points1"], g["points2"]]


def test_voxelizer_restatement_is_self_consistent(golden_pillars):
    """spconv is absent (parity unpinned): check the restatement against an independent pure-Python
    transcription of the sequential algorithm on the golden clouds, incl. the max_voxels cap."""
    g = golden_pillars
    rng, vs, cap = g["lidar_range"].tolist(), g["voxel_size"].tolist(), int(g["max_voxels"])
    grid = R.grid_size(rng, vs)
    outs = []
    for pts in _split_points(g):
        out = R.voxelize(pts, rng, vs, 32, cap)
        outs.append(out)
        vmap, coords, vox = {}, [], []
        for p in pts:
            c = [int(np.floor((np.float32(p[j]) - np.float32(rng[j])) / np.float32(vs[j]))) for j in range(3)]
            if any(c[j] < 0 or c[j] >= grid[j] for j in range(3)):
                continue
            key = (c[2], c[1], c[0])
            if key not in vmap:
                if len(coords) >= cap:
                    continue
                vmap[key] = len(coords)
                coords.append(key)
                vox.append([])
            if len(vox[vmap[key]]) < 32:
                vox[vmap[key]].append(p)
        assert out["voxel_coords"].tolist() == [list(k) for k in coords]
        assert out["voxel_num_points"].tolist() == [len(v) for v in vox]
        for m, v in enumerate(vox):
            assert np.array_equal(out["voxel_features"][m, :len(v)], np.stack(v))
            assert not out["voxel_features"][m, len(v):].any()
    batch = R.collate_voxels(outs)
    assert np.array_equal(batch["voxel_features"].numpy(), g["voxel_features"])
    assert np.array_equal(batch["voxel_coords"].numpy(), g["voxel_coords"])
    assert np.array_equal(batch["voxel_num_points"].numpy(), g["voxel_num_points"])
    assert outs[0]["voxel_features"].shape[0] == cap and outs[1]["voxel_features"].shape[0] < cap


def _pfn(g):
    return {k[4:]: T(v) for k, v in g.items() if k.startswith("pfn_")}


def test_pillar_vfe_torch_order_matches_reference(golden_pillars):
    g, w = golden_pillars, _pfn(golden_pillars)
    out = R.pillar_vfe(T(g["voxel_features"]), T(g["voxel_num_points"]), T(g["voxel_coords"]), w["weight"],
                       w["bn_weight"], w["bn_bias"], w["bn_mean"], w["bn_var"], g["voxel_size"].tolist(),
                       g["lidar_range"].tolist())
    assert torch.equal(out, T(g["ref_pillar_features"]))


def test_pillar_vfe_kernel_order_close_to_reference(golden_pillars):
    """The fixed evaluation order shared with the CUDA kernel is a regrouping of the same sum:
    |delta| <= 1e-5 * max|ref| (the reference's own fp32 rounding noise is of that size)."""
    g, w = golden_pillars, _pfn(golden_pillars)
    sc, sh = R.fold_bn(w["bn_weight"], w["bn_bias"], w["bn_mean"], w["bn_var"])
    out = R.pillar_vfe_kernel_order(T(g["voxel_features"]), T(g["voxel_num_points"]), T(g["voxel_coords"]),
                                    w["weight"], sc, sh, g["voxel_size"].tolist(), g["lidar_range"].tolist())
    ref = T(g["ref_pillar_features"])
    assert (out - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


def test_scatter_matches_reference_bit_exact(golden_pillars):
    g = golden_pillars
    grid = R.grid_size(g["lidar_range"].tolist(), g["voxel_size"].tolist())
    canvas = R.scatter(T(g["ref_pillar_features"]), T(g["voxel_coords"]), int(grid[0]), int(grid[1]))
    assert torch.equal(canvas, T(g["ref_canvas"]))


def test_normalize_pairwise_tfm_bit_exact(golden_warp):
    g = golden_warp
    theta = R.normalize_pairwise_tfm(T(g["pairwise"]), float(g["Hm"]), float(g["Wm"]), 1)
    assert theta.dtype == torch.float64 and torch.equal(theta, T(g["ref_theta"]))


def test_warp_and_fusion_match_reference(golden_warp):
    g = golden_warp
    feat, rl, theta = T(g["feat"]), T(g["record_len"]), T(g["ref_theta"])
    assert torch.equal(R.warp_only(feat, rl, theta), T(g["ref_warped"]))
    assert torch.equal(R.max_fusion(feat, rl, theta), T(g["ref_max"]))
    assert torch.equal(R.att_fusion(feat, rl, theta), T(g["ref_att"]))


def _sd(g):
    return {k[len("sd/denoiser."):]: T(v) for k, v in g.items() if k.startswith("sd/denoiser.")}


def test_unet_matches_reference(golden_gencomm):
    g = golden_gencomm
    x = torch.cat([T(g["cond"]), T(g["feat"])], dim=1)
    out = R.unet_forward(x, T(g["unet_t"]), _sd(g))
    assert torch.allclose(out, T(g["ref_unet"]), rtol=0, atol=1e-6)


def test_gencomm_schedule_matches_reference_buffers(golden_gencomm):
    g = golden_gencomm
    sch = R.gencomm_schedule(3)
    for k, v in sch.items():
        assert torch.equal(v, T(g["sd/" + k])), k


def test_gencomm_sampler_matches_reference(golden_gencomm):
    g = golden_gencomm
    out = R.gencomm_sample(T(g["feat"]), T(g["cond"]), T(g["record_len"]), _sd(g), T(g["noise0"]),
                           [T(s) for s in g["step_noises"]])
    assert torch.allclose(out, T(g["ref_pred"]), rtol=0, atol=2e-6)


def test_message_extractor_restatement_matches_reference(golden_message_extractor):
    """oracle/ref_ops.message_extractor_v2 (incl. the restated torchvision deform_conv2d) vs the golden outputs of the
    reference MessageExtractorv2 class (oracle/gen_golden.py::gen_message_extractor)."""
    g = golden_message_extractor
    sd = {k[3:]: T(v) for k, v in g.items() if k.startswith("sd/")}
    out, offset, b1 = R.message_extractor_v2(T(g["x"]), sd)
    assert torch.allclose(offset, T(g["ref_offset"]), rtol=0, atol=1e-5)
    assert torch.allclose(b1, T(g["ref_b1"]), rtol=0, atol=1e-5)
    assert torch.allclose(out, T(g["ref_out"]), rtol=0, atol=1e-6)
    assert float(T(g["ref_offset"]).abs().max()) > 2.0   # the fixture really moves the sampling taps


def test_enhancer_restatement_matches_reference(golden_enhancer):
    """oracle/ref_ops.enhancer vs the golden output of the reference Enhancer class (gen_golden.py::gen_enhancer)."""
    g = golden_enhancer
    sd = {k[3:]: T(v) for k, v in g.items() if k.startswith("sd/")}
    out = R.enhancer(T(g["x"]), sd)
    assert torch.allclose(out, T(g["ref_out"]), rtol=0, atol=5e-6)


def test_downsample_conv_restatement_matches_reference(golden_det_tail):
    g = golden_det_tail
    sd = {k[3:]: T(v) for k, v in g.items() if k.startswith("sd/")}
    out = R.downsample_conv(T(g["x"]), sd, [2])
    assert torch.allclose(out, T(g["ref_out"]), rtol=0, atol=1e-6)
    heads = R.det_heads(out, *[T(g[f"head{i}/{k}"]) for i in range(3) for k in ("weight", "bias")])
    for i, h in enumerate(heads):
        assert torch.allclose(h, T(g[f"head{i}/out"]), rtol=0, atol=1e-6)


def test_bev_backbone_restatement_matches_reference(golden_backbone):
    from conftest import BACKBONE_CFG, backbone_input
    g = golden_backbone
    x = backbone_input()
    assert abs(float(x.double().sum()) - float(g["x_checksum"])) < 1e-6
    sd = {k[3:]: T(v) for k, v in g.items() if k.startswith("sd/")}
    out = R.bev_backbone(x, sd, BACKBONE_CFG["layer_nums"], BACKBONE_CFG["layer_strides"], BACKBONE_CFG["upsample_strides"])
    assert torch.allclose(out[:, ::4], T(g["ref_out_c4"]), rtol=0, atol=1e-6)


def test_heter_model_restatement_matches_reference(golden_heter_model, heter_inputs):
    """The composed oracle (voxels -> ... -> heads) against the UNMODIFIED HeterModelBaselineWGenComm on the full
    OPV2V-H grid (2 + 1 agents); the weights are regenerated from the reference's own state_dict key list."""
    from gencomm_b200 import HeterModelBaselineWGenComm, synth
    from oracle import gen_golden
    g = golden_heter_model
    model = HeterModelBaselineWGenComm(synth.gencomm_stage1_args("att"))
    assert sorted(model.state_dict().keys()) == list(g["state_dict_keys"])      # a reference checkpoint loads as is
    sd = synth.fill_state_dict(model.state_dict(), gen_golden.HETER_WSEED)
    voxels, pairwise, record_len, n0, steps = heter_inputs
    assert voxels["voxel_coords"].shape[0] == int(g["n_pillars"])
    out = R.heter_gencomm_forward(sd, synth.gencomm_stage1_args("att"), voxels, pairwise, record_len, n0, steps)
    for k in ("cls_preds", "reg_preds", "dir_preds", "message"):
        ref = T(g[k])
        assert float((out[k] - ref).abs().max()) <= 1e-5 * float(ref.abs().max()), k
    for k in ("gt_feature", "pred_feature"):
        ref = T(g[k + "_c8"])
        assert float((out[k][:, ::8] - ref).abs().max()) <= 1e-5 * float(ref.abs().max()), k


def test_create_model_resolves_the_gencomm_detectors():
    """train_utils.create_model (tools/train_utils.py:255-288) over gencomm_b200, incl. the yaml / class-name quirk."""
    import pytest
    import gencomm_b200 as G
    from gencomm_b200 import synth
    m1 = G.create_model({"model": {"core_method": "heter_model_baseline_w_gencomm_stage1",
                                   "args": synth.gencomm_stage1_args("max")}})
    assert type(m1).__name__ == "HeterModelBaselineWGenComm" and isinstance(m1.fusion_net, G.MaxFusion)
    a2 = synth.gencomm_stage1_args("att")
    a2["diffcomm"] = a2.pop("gencomm")
    a2["trick"] = True
    m2 = G.create_model({"model": {"core_method": "heter_model_baseline_w_gencomm_stage2", "args": a2}})
    assert type(m2).__name__ == "HeterModelBaselineWDiffCommStage2" and m2.trick
    assert set(m2.state_dict().keys()) == set(m1.state_dict().keys())
    with pytest.raises(NotImplementedError):
        G.create_model({"model": {"core_method": "point_pillar_v2vnet", "args": {}}})
    # camera (lift_splat_shoot) modalities enter at the encoder boundary (BEV feature input); other encoders are refused
    cam = G.HeterModelBaselineWDiffCommStage2(synth.gencomm_stage2_hetero_args("att"))
    assert isinstance(cam.encoder_m2, G.BEVFeatureInput) and cam.crop_ratio_W_m2 == 2.0 and cam.crop_ratio_H_m2 == 1.0
    sec = synth.gencomm_stage1_args("att")
    sec["m3"] = dict(sec["m1"], core_method="second")
    with pytest.raises(NotImplementedError, match="LiDAR"):
        G.HeterModelBaselineWGenComm(sec)
    bad = synth.gencomm_stage1_args("att")
    bad["fusion_method"] = "v2xvit"
    with pytest.raises(NotImplementedError, match="fusion_method"):
        G.HeterModelBaselineWGenComm(bad)


def test_postprocess_restatement_matches_reference(golden_postprocess):
    """Decode + rotated NMS: the oracle against the UNMODIFIED VoxelPostprocessor.post_process
    (bit-exact boxes / scores / order), incl. the top-1000 cap, a rigid cav -> ego transform and the empty frame."""
    from gencomm_b200 import VoxelPostprocessor, synth
    from oracle import gen_golden
    g = golden_postprocess
    params = synth.postprocess_params()
    anchors = R.generate_anchor_box(params["anchor_args"], params["order"])
    assert np.array_equal(anchors, g["anchors"])
    assert np.array_equal(VoxelPostprocessor(params, train=False).generate_anchor_box(), g["anchors"])   # host mirror
    for name, (seed, bias, moved) in gen_golden.POSTPROCESS_CASES.items():
        cls, reg, dr = synth.head_outputs(seed, bias=bias)
        boxes, scores = R.post_process(cls, reg, dr, T(anchors), gen_golden.postprocess_transform(moved), params)
        if int(g[f"{name}/count"]) == 0:
            assert boxes is None and scores is None
            continue
        assert torch.equal(boxes, T(g[f"{name}/boxes"])), name
        assert torch.equal(scores, T(g[f"{name}/scores"])), name


def test_nms_c_restatement_matches_python_restatement():
    """oracle/nms_ref.c against the numpy / Python restatement of nms_rotated + polygon IoU on a small clustered case."""
    g = torch.Generator().manual_seed(5)
    n = 120
    centre = torch.rand(n, 1, 2, generator=g) * 12.0
    yaw = torch.rand(n, generator=g) * 3.1416
    half = torch.tensor([[1.95, -0.8], [1.95, 0.8], [-1.95, 0.8], [-1.95, -0.8]])
    rot = torch.stack([torch.stack([yaw.cos(), -yaw.sin()], -1), torch.stack([yaw.sin(), yaw.cos()], -1)], -2)
    quad = torch.einsum("nij,kj->nki", rot, half) + centre
    boxes = torch.cat([quad, torch.zeros(n, 4, 1)], dim=-1)
    scores = torch.rand(n, generator=g)
    scores[7] = scores[3]                                            # a tie: larger index first
    a = R.nms_rotated(boxes, scores, 0.15)
    b = R.nms_rotated_py(boxes, scores, 0.15)
    assert np.array_equal(a, b) and 5 < len(a) < n
    assert abs(float(R.polygon_iou(quad[0].numpy(), quad[0].numpy())) - 1.0) < 1e-6


def test_lss_voxel_pooling_restatement_matches_reference(golden_lss_pool):
    """the restated LiftSplatShoot.voxel_pooling against the unmodified reference method (bit-exact:
    same sort, same fp32 cumsum differences), and the float64 per-voxel sums it approximates."""
    from conftest import LSS_CASES, lss_case
    g = golden_lss_pool
    for name in LSS_CASES:
        geom, x, conf = lss_case(name)
        dx, bx, nx = R.gen_dx_bx(conf["xbound"], conf["ybound"], conf["zbound"])
        out = R.lss_voxel_pooling(geom, x, dx, bx, nx)
        assert list(out.shape) == g[f"{name}/shape"].tolist()
        cells = T(g[f"{name}/cells"]).long()
        assert torch.equal(out[cells[:, 0], :, cells[:, 1], cells[:, 2]], T(g[f"{name}/values"])), name
        assert int((out.abs().sum(1) != 0).sum()) == cells.shape[0]          # nothing outside the recorded cells
        exact = R.lss_voxel_pooling_exact(geom, x, dx, bx, nx)
        # the reference's cumsum differences carry cancellation noise ~1e-7 * |running sum|
        assert float((out.double() - exact).abs().max()) <= 2e-4, name


def test_quad_iou_properties():
    """Domain properties of the restated polygon IoU (oracle/nms_ref.c): symmetry, identity, disjointness, containment,
    orientation independence and a closed-form case (two axis-aligned boxes shifted by half their length)."""
    import ctypes
    lib = R._lib()
    def iou(p, q):
        p = np.ascontiguousarray(p, np.float64); q = np.ascontiguousarray(q, np.float64)
        return float(lib.gc_ref_quad_iou(R._p(p), R._p(q)))
    rng = np.random.default_rng(3)
    def box(cx, cy, l, w, yaw):
        c, s = np.cos(yaw), np.sin(yaw)
        pts = np.array([[l / 2, -w / 2], [l / 2, w / 2], [-l / 2, w / 2], [-l / 2, -w / 2]])
        return pts @ np.array([[c, s], [-s, c]]) + [cx, cy]
    a = box(0, 0, 4, 2, 0.0)
    assert abs(iou(a, box(2, 0, 4, 2, 0.0)) - (4.0 / 12.0)) < 1e-6          # overlap 2x2 = 4, union 8 + 8 - 4
    assert iou(a, box(10, 0, 4, 2, 0.3)) == 0.0                              # disjoint
    assert abs(iou(a, box(0, 0, 2, 1, 0.0)) - 0.25) < 1e-6                   # contained: 2 / 8
    for _ in range(200):
        p = box(*rng.uniform(-3, 3, 2), *rng.uniform(1, 5, 2), rng.uniform(-3.2, 3.2))
        q = box(*rng.uniform(-3, 3, 2), *rng.uniform(1, 5, 2), rng.uniform(-3.2, 3.2))
        v = iou(p, q)
        assert 0.0 <= v <= 1.0 + 1e-6
        assert abs(v - iou(q, p)) < 1e-6                                      # symmetric
        assert abs(v - iou(p[::-1], q)) < 1e-6                                # orientation of the corner order
        assert abs(v - iou(p + 7.5, q + 7.5)) < 1e-5                          # translation
        assert abs(iou(p, p) - 1.0) < 1e-6
        assert abs(v - float(R.polygon_iou(p, q))) < 1e-6                     # C vs Python restatement


def test_lss_both_cumsum_variants_are_pinned(golden_lss_pool):
    """heter_encoders.py:202-207 picks cumsum_trick or the QuickCumsum autograd function by ``use_quickcumsum``: the
    fixture holds BOTH results of the unmodified method; their forward arithmetic is the same, and the restatement
    reproduces them bit for bit."""
    from conftest import LSS_CASES, lss_case
    g = golden_lss_pool
    for name in LSS_CASES:
        assert np.array_equal(g[f"{name}/values"], g[f"{name}/values_quickcumsum"]), name
        geom, x, conf = lss_case(name)
        dx, bx, nx = R.gen_dx_bx(conf["xbound"], conf["ybound"], conf["zbound"])
        out = R.lss_voxel_pooling(geom, x, dx, bx, nx)
        cells = T(g[f"{name}/cells"]).long()
        assert torch.equal(out[cells[:, 0], :, cells[:, 1], cells[:, 2]], T(g[f"{name}/values_quickcumsum"])), name


def test_collate_matches_reference():
    """SpVoxelPreprocessor.collate_batch (list and dict form, sp_voxel_preprocessor.py:87-174), run unbound from the
    unmodified class (tests/golden/collate.npz): the agent-index column, concatenation order and dtypes of the collated
    voxel tensors -- for the oracle's collate_voxels and for the drop-in's host-side collate."""
    import os
    from gencomm_b200 import SpVoxelPreprocessor, synth
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "collate.npz"))
    clouds = [synth.lidar_points(77, a, 6000 + 500 * a) for a in range(3)]
    per_agent = [R.voxelize(c, synth.OPV2V_H_RANGE, [0.4, 0.4, 4.0]) for c in clouds]
    ours = R.collate_voxels(per_agent)
    as_np = [{k: np.asarray(v) for k, v in d.items()} for d in per_agent]
    drop_list = SpVoxelPreprocessor.collate_batch_list(as_np)
    drop_dict = SpVoxelPreprocessor.collate_batch_dict({k: [d[k] for d in as_np] for k in as_np[0]})
    for res in (ours, drop_list, drop_dict):
        assert np.array_equal(np.asarray(res["voxel_coords"]), g["voxel_coords"])
        assert np.array_equal(np.asarray(res["voxel_num_points"]), g["voxel_num_points"])
        assert float(torch.as_tensor(res["voxel_features"]).double().sum()) == float(g["voxel_features_checksum"])
        assert [str(torch.as_tensor(res[k]).dtype) for k in ("voxel_features", "voxel_coords", "voxel_num_points")] == list(g["dtypes"])
    assert g["voxel_coords"][:, 0].max() == 2 and g["voxel_coords"].dtype == np.int32


def test_stage2_hetero_state_dict_keys_match_reference():
    """The LiDAR + camera stage-2 drop-in exposes exactly the parameters of the unmodified reference model with its image
    encoder stubbed out (tests/golden/heter_model_stage2.npz): a reference checkpoint loads with strict=False, the only
    keys left over being the encoder_m2.* (EfficientNet / LSS) weights that stay with the reference."""
    import os
    from gencomm_b200 import HeterModelBaselineWDiffCommStage2, synth
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "heter_model_stage2.npz"))
    m = HeterModelBaselineWDiffCommStage2(synth.gencomm_stage2_hetero_args("att"))
    assert sorted(m.state_dict().keys()) == list(g["state_dict_keys"])
    # the yaml spelling (``gencomm:`` instead of ``diffcomm:``) is accepted too
    a = synth.gencomm_stage2_hetero_args("att")
    a["gencomm"] = a.pop("diffcomm")
    assert sorted(HeterModelBaselineWDiffCommStage2(a).state_dict().keys()) == list(g["state_dict_keys"])


def test_center_crop_matches_torchvision():
    import torchvision
    from gencomm_b200.modules import center_crop
    x = torch.randn(2, 3, 7, 10)
    for th, tw in ((7, 20), (7, 10), (4, 6), (9, 5), (14, 21), (6, 11)):
        assert torch.equal(center_crop(x, th, tw), torchvision.transforms.CenterCrop((th, tw))(x)), (th, tw)
