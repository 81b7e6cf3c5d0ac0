"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz from the UNMODIFIED reference classes.

Run in the build container (``/root/reference`` mounted):  ``python oracle/gen_golden.py``

The reference ships no tests, golden vectors or fixtures for this path, so
these fixtures are created here by importing the reference's own modules (oracle/ref_import.py)
and executing them on CPU with seeded inputs.  They pin the oracle restatement
(oracle/ref_ops.py, oracle/pillar_ref.c) in ``tests/test_oracle_cpu.py`` and are the committed
ground truth for the ``-m gpu`` parity tests (``/root/reference`` does not exist on the GPU box).

Voxelizer inputs come from the oracle's own spconv restatement (spconv is absent and unpinned:
"parity unpinned"); everything downstream of the voxel tensors is produced by reference code.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gencomm_b200 import synth  # noqa: E402
from oracle import ref_import, ref_ops  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

SMALL_RANGE = [-12.8, -6.4, -3.0, 12.8, 6.4, 1.0]   # 64 x 32 x 1 grid at 0.4 m
VOXEL = [0.4, 0.4, 4.0]

GENCOMM_CFG = {
    "model": {"embed_dim": 18, "in_channels": 16, "out_ch": 16, "ch": 8, "ch_mult": [1, 1],
              "num_res_blocks": 2, "attn_resolutions": [16], "dropout": 0.0, "resamp_with_conv": True},
    "diffusion": {"beta_schedule": "linear", "beta_start": 0.0005, "beta_end": 0.02,
                  "num_diffusion_timesteps": 3},
}


def small_cloud(agent, n):
    pts = synth.lidar_points(frame=900, agent=agent, n_points=n, lidar_range=SMALL_RANGE)
    pts[:, :2] *= 0.12   # pull the 120 m cloud into the 25 m box (some points stay outside)
    return np.ascontiguousarray(pts.astype(np.float32))


def gen_pillars(ns):
    torch.manual_seed(1)
    clouds = [small_cloud(0, 3000), small_cloud(1, 600), small_cloud(2, 2000)]
    vox = [ref_ops.voxelize(c, SMALL_RANGE, VOXEL, 32, 400) for c in clouds]   # cap 400 is hit
    batch = ref_ops.collate_voxels(vox)
    w = synth.pfn_weights(3)
    grid = ref_ops.grid_size(SMALL_RANGE, VOXEL)
    vfe = ns.PillarVFE({"use_norm": True, "with_distance": False, "use_absolute_xyz": True, "num_filters": [64]},
                       num_point_features=4, voxel_size=VOXEL, point_cloud_range=SMALL_RANGE).eval()
    with torch.no_grad():
        vfe.pfn_layers[0].linear.weight.copy_(w["weight"])
        vfe.pfn_layers[0].norm.weight.copy_(w["bn_weight"])
        vfe.pfn_layers[0].norm.bias.copy_(w["bn_bias"])
        vfe.pfn_layers[0].norm.running_mean.copy_(w["bn_mean"])
        vfe.pfn_layers[0].norm.running_var.copy_(w["bn_var"])
        bd = {k: v.clone() for k, v in batch.items()}
        bd = vfe(bd)
        sc = ns.PointPillarScatter({"num_features": 64, "grid_size": grid})
        bd = sc(bd)
    np.savez_compressed(
        os.path.join(OUT, "pillars.npz"),
        lidar_range=np.array(SMALL_RANGE), voxel_size=np.array(VOXEL), max_voxels=400,
        points0=clouds[0], points1=clouds[1], points2=clouds[2],
        voxel_features=batch["voxel_features"].numpy(), voxel_coords=batch["voxel_coords"].numpy(),
        voxel_num_points=batch["voxel_num_points"].numpy(),
        **{"pfn_" + k: v.numpy() for k, v in w.items()},
        ref_pillar_features=bd["pillar_features"].numpy(), ref_canvas=bd["spatial_features"].numpy())
    print("pillars.npz: M =", batch["voxel_features"].shape[0], "per agent", [v["voxel_features"].shape[0] for v in vox])


def gen_warp(ns):
    C, H, W = 8, 16, 24
    record_len = torch.tensor([3, 2], dtype=torch.int64)
    feat = synth.bev_features(901, 5, C, H, W)
    pw = np.stack([synth.pairwise_t_matrix(901, 3, 5, spread=(6.0, 3.0)),
                   synth.pairwise_t_matrix(902, 2, 5, spread=(6.0, 3.0))])
    pw_t = torch.from_numpy(pw)
    Hm, Wm = 16 * 0.8, 24 * 0.8   # metres covered by the map (heter_model_baseline.py:87-89 convention)
    theta = ns.normalize_pairwise_tfm(pw_t.clone(), Hm, Wm, 1)
    with torch.no_grad():
        warped = torch.cat([ns.warp_affine_simple(x, theta[b][0, :int(record_len[b])], (H, W))
                            for b, x in enumerate(ns.regroup(feat, record_len))])
        mx = ns.MaxFusion()(feat, record_len, theta)
        att = ns.AttFusion(C)(feat, record_len, theta)
    np.savez_compressed(os.path.join(OUT, "warp_fuse.npz"), feat=feat.numpy(), record_len=record_len.numpy(),
                        pairwise=pw, Hm=Hm, Wm=Wm, ref_theta=theta.numpy(), ref_warped=warped.numpy(),
                        ref_max=mx.numpy(), ref_att=att.numpy())
    print("warp_fuse.npz:", tuple(feat.shape), "theta", tuple(theta.shape))


def gen_gencomm(ns):
    import opencood.models.gencomm_modules.cond_diff as cd
    torch.manual_seed(7)
    C, H, W = 16, 16, 24
    model = ns.GenComm(GENCOMM_CFG).eval()
    with torch.no_grad():   # default GroupNorm affine is (1,0): randomise so it is exercised
        for name, p in model.named_parameters():
            if "norm" in name:
                p.add_(0.2 * torch.randn_like(p))
            elif name.endswith(".bias"):
                p.add_(0.05 * torch.randn_like(p))
    record_len = torch.tensor([2, 1], dtype=torch.int64)
    feat = synth.bev_features(903, 3, C, H, W)
    cond = synth.bev_features(903, 3, 2, H, W, salt=4)
    n0, steps = synth.sampler_noise(903, 3, C, H, W, T=3)

    # inject the pre-drawn noise in the reference's draw order (SURVEY.md App. A.6 RNG)
    queue = [n0, torch.zeros(1, C, H, W), torch.zeros(1, C, H, W)]
    step_queue = list(steps)
    orig_randn_like, orig_noise_like = torch.randn_like, cd.noise_like
    torch.randn_like = lambda t, *a, **k: queue.pop(0).to(t)
    cd.noise_like = lambda shape, device, repeat=False: step_queue.pop(0)
    try:
        with torch.no_grad():
            out = model(feat, cond, record_len)
    finally:
        torch.randn_like, cd.noise_like = orig_randn_like, orig_noise_like
    assert not queue and not step_queue
    with torch.no_grad():
        x_in = torch.cat([cond, feat], dim=1)
        t = torch.tensor([2.0, 1.0, 0.0])
        unet_out = model.denoiser(x_in, t)
    sd = {k: v.numpy() for k, v in model.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "gencomm.npz"), feat=feat.numpy(), cond=cond.numpy(),
                        record_len=record_len.numpy(), noise0=n0.numpy(),
                        step_noises=np.stack([s.numpy() for s in steps]),
                        ref_pred=out["pred_feature"].numpy(), unet_t=t.numpy(), ref_unet=unet_out.numpy(),
                        **{"sd/" + k: v for k, v in sd.items()})
    print("gencomm.npz: pred", tuple(out["pred_feature"].shape), "params",
          sum(p.numel() for p in model.parameters()))


def gen_message_extractor(ns):
    """MessageExtractorv2 through the reference class (torchvision DeformConv2d on CPU)."""
    torch.manual_seed(11)
    C, H, W, N = 64, 4, 64, 3
    model = ns.MessageExtractorv2(C, 2).eval()
    with torch.no_grad():   # default inits give ~0.3-pixel offsets; scale them up so the bilinear taps really move
        model.bev_extractor.offset1.weight.mul_(3.0)
        model.bev_extractor.offset1.bias.add_(0.3 * torch.randn(18))
    x = synth.bev_features(1201, N, C, H, W)
    x[2, :, :, 40:] = 0.0   # a camera-like agent: empty outside its field of view
    with torch.no_grad():
        out = model(x)
        offset = model.bev_extractor.offset1(x)
        b1 = model.bev_extractor.dcn1(x, offset)
    sd = {k: v.numpy() for k, v in model.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "message_extractor.npz"), x=x.numpy(), ref_out=out.numpy(),
                        ref_offset=offset.numpy(), ref_b1=b1.numpy(), **{"sd/" + k: v for k, v in sd.items()})
    print("message_extractor.npz: out", tuple(out.shape), "|offset| max", float(offset.abs().max()))


def gen_enhancer(ns):
    """Enhancer through the reference class."""
    torch.manual_seed(13)
    C, H, W = 128, 4, 64
    model = ns.Enhancer(C, [8, 8], 4).eval()
    with torch.no_grad():   # LayerNorm affines default to (1, 0), biases to small values: randomise so they matter
        for name, p in model.named_parameters():
            if "norm" in name or "bn1" in name:
                p.add_(0.2 * torch.randn_like(p))
            elif name.endswith(".bias"):
                p.add_(0.1 * torch.randn_like(p))
    x = synth.bev_features(1301, 3, C, H, W)
    x[2, :, :, 40:] = 0.0
    record_len = torch.tensor([2, 1], dtype=torch.int64)
    affine = torch.randn(2, 5, 5, 2, 3)
    with torch.no_grad():
        out = model(x, affine, record_len)
    import json
    sd = {k: v.numpy() for k, v in model.state_dict().items()}
    # tensors of the evaluated parameters; names + shapes only for the ones the reference declares but never uses
    used = {k: v for k, v in sd.items() if (k.startswith("block_1.") and ".attn." not in k) or k.startswith("split_attn.")}
    shapes = np.frombuffer(json.dumps({k: list(v.shape) for k, v in sd.items()}).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "enhancer.npz"), x=x.numpy(), record_len=record_len.numpy(),
                        affine=affine.numpy(), ref_out=out.numpy(), sd_shapes_json=shapes,
                        **{"sd/" + k: v for k, v in used.items()})
    print("enhancer.npz: out", tuple(out.shape), "keys", len(sd), "stored", len(used))


def gen_det_tail(ns):
    """DownsampleConv (shrink header, stride 2 like the GenComm stage-1 yaml) through the reference class, and the three
    1x1 heads (plain nn.Conv2d in heter_model_baseline.py:130-135)."""
    torch.manual_seed(17)
    cfg = {"kernal_size": [3], "stride": [2], "padding": [1], "dim": [64], "input_dim": 64}
    model = ns.DownsampleConv(cfg).eval()
    x = synth.bev_features(1401, 2, 64, 16, 32)
    with torch.no_grad():
        out = model(x)
    heads = [torch.nn.Conv2d(64, n, 1) for n in (2, 14, 4)]
    with torch.no_grad():
        hout = [h(out) for h in heads]
    sd = {k: v.numpy() for k, v in model.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "det_tail.npz"), x=x.numpy(), ref_out=out.numpy(),
                        **{"sd/" + k: v for k, v in sd.items()},
                        **{f"head{i}/weight": h.weight.detach().numpy() for i, h in enumerate(heads)},
                        **{f"head{i}/bias": h.bias.detach().numpy() for i, h in enumerate(heads)},
                        **{f"head{i}/out": o.numpy() for i, o in enumerate(hout)})
    print("det_tail.npz: out", tuple(out.shape))


BACKBONE_CFG = {"layer_nums": [1, 1, 2], "layer_strides": [2, 2, 2], "num_filters": [64, 64, 128],
                "upsample_strides": [1, 2, 4], "num_upsample_filter": [64, 64, 64]}


def backbone_input():
    return synth.bev_features(1501, 1, 64, 64, 128, sparsity=0.7)


def gen_backbone(ns):
    """BaseBEVBackbone through the reference class (eval mode, randomised BatchNorm statistics)."""
    torch.manual_seed(19)
    model = ns.BaseBEVBackbone(BACKBONE_CFG, 64).eval()
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
                m.weight.add_(0.2 * torch.randn_like(m.weight)); m.bias.add_(0.1 * torch.randn_like(m.bias))
    x = backbone_input()
    with torch.no_grad():
        out = model({"spatial_features": x})["spatial_features_2d"]
    sd = {k: v.numpy() for k, v in model.state_dict().items()}
    # the input is regenerated from its seed (backbone_input); every 4th output channel is stored to keep the fixture small
    np.savez_compressed(os.path.join(OUT, "backbone.npz"), ref_out_c4=out[:, ::4].numpy(), x_checksum=np.float64(x.double().sum()),
                        **{"sd/" + k: v for k, v in sd.items()})
    print("backbone.npz: out", tuple(out.shape), "keys", len(sd))


HETER_RECORD_LEN = [2, 1]
HETER_SEED, HETER_POINTS, HETER_WSEED = 4100, 30_000, 11


def heter_inputs():
    """Inputs of tests/golden/heter_model.npz, regenerated from seeds (also used by the tests)."""
    clouds, pairwise = synth.heter_frames(HETER_SEED, HETER_RECORD_LEN, HETER_POINTS)
    voxels = ref_ops.collate_voxels([ref_ops.voxelize(c, synth.OPV2V_H_RANGE, [0.4, 0.4, 4.0]) for c in clouds])
    n = len(clouds)
    noise0, steps = synth.sampler_noise(HETER_SEED, n, 128, 64, 128, T=3)
    return voxels, torch.from_numpy(pairwise), torch.tensor(HETER_RECORD_LEN, dtype=torch.int64), noise0, steps


def gen_heter_model(ns):
    """The UNMODIFIED HeterModelBaselineWGenComm (stage-1 detector, m1_att.yaml model args) on two frames (2 + 1 agents)
    of the full OPV2V-H grid.  Weights: synth.fill_state_dict (regenerated in the tests from the shared state_dict
    keys, so the key list itself is pinned); sampler noise injected in the reference's draw order."""
    import opencood.models.gencomm_modules.cond_diff as cd
    from opencood.models.heter_model_baseline_w_gencomm_stage1 import HeterModelBaselineWGenComm
    args = synth.gencomm_stage1_args("att")
    model = HeterModelBaselineWGenComm(args).eval()
    model.load_state_dict(synth.fill_state_dict(model.state_dict(), HETER_WSEED))
    voxels, pairwise, record_len, n0, steps = heter_inputs()
    data = {"inputs_m1": voxels, "agent_modality_list": ["m1"] * int(record_len.sum()), "pairwise_t_matrix": pairwise,
            "record_len": record_len}
    C, H, W = 128, 64, 128
    queue = [n0, torch.zeros(1, C, H, W), torch.zeros(1, C, H, W)]
    step_queue = list(steps)
    orig_randn_like, orig_noise_like = torch.randn_like, cd.noise_like
    torch.randn_like = lambda t, *a, **k: queue.pop(0).to(t)
    cd.noise_like = lambda shape, device, repeat=False: step_queue.pop(0)
    try:
        with torch.no_grad():
            out = model(data)
    finally:
        torch.randn_like, cd.noise_like = orig_randn_like, orig_noise_like
    assert not queue and not step_queue
    keys = sorted(model.state_dict().keys())
    np.savez_compressed(os.path.join(OUT, "heter_model.npz"), state_dict_keys=np.array(keys),
                        n_pillars=np.int64(voxels["voxel_coords"].shape[0]),
                        cls_preds=out["cls_preds"].numpy(), reg_preds=out["reg_preds"].numpy(),
                        dir_preds=out["dir_preds"].numpy(), message=out["message"].numpy(),
                        gt_feature_c8=out["gt_feature"][:, ::8].numpy(), pred_feature_c8=out["pred_feature"][:, ::8].numpy())
    print("heter_model.npz: cls", tuple(out["cls_preds"].shape), "pillars", voxels["voxel_coords"].shape[0],
          "|cls| max", float(out["cls_preds"].abs().max()), "|gt| max", float(out["gt_feature"].abs().max()),
          "|pred| max", float(out["pred_feature"].abs().max()))


HETER2_RECORD_LEN = [3, 2]
HETER2_MODALITIES = ["m1", "m2", "m1", "m1", "m2"]     # per-agent order; the ego of each frame is LiDAR (ego_modality m1)
HETER2_SEED, HETER2_WSEED = 4300, 17


def heter2_inputs():
    """Inputs of tests/golden/heter_model_stage2.npz, regenerated from seeds (also used by the tests): LiDAR voxels of the
    three m1 agents, the synthetic LSS BEV feature [2,128,256,256] of the two m2 (camera) agents (75 % empty cells, like
    a splatted frustum), poses, record_len, sampler noise."""
    n = sum(HETER2_RECORD_LEN)
    clouds_all, pairwise = synth.heter_frames(HETER2_SEED, HETER2_RECORD_LEN, HETER_POINTS)
    lidar = [c for c, m in zip(clouds_all, HETER2_MODALITIES) if m == "m1"]
    voxels = ref_ops.collate_voxels([ref_ops.voxelize(c, synth.OPV2V_H_RANGE, [0.4, 0.4, 4.0]) for c in lidar])
    bev = synth.bev_features(HETER2_SEED, HETER2_MODALITIES.count("m2"), 128, 256, 256, sparsity=0.75)
    noise0, steps = synth.sampler_noise(HETER2_SEED, n, 128, 64, 128, T=3)
    return voxels, bev, torch.from_numpy(pairwise), torch.tensor(HETER2_RECORD_LEN, dtype=torch.int64), noise0, steps


def gen_heter_model_stage2(ns):
    """The UNMODIFIED HeterModelBaselineWDiffCommStage2 (m1m2_att.yaml model args: LiDAR m1 + camera m2) on two frames
    (3 + 2 agents, modalities interleaved).  Only the third-party-dependent image encoder is replaced: heter_encoders'
    LiftSplatShoot (EfficientNet) is swapped for a stub that returns the injected BEV feature, so the camera branch
    (backbone with inplanes 128, shrink header, message extractor), the CenterCrop zero-padding of features AND messages
    (:223-236), the per-agent re-assembly (:245-256), sampler, Enhancer, AttFusion and heads are all reference code."""
    import torch.nn as nn
    import opencood.models.gencomm_modules.cond_diff as cd
    import opencood.models.heter_encoders as he
    import opencood.models.heter_model_baseline_w_gencomm_stage2 as st2

    class LiftSplatShoot(nn.Module):   # stand-in for the image encoder only
        def __init__(self, args):
            super().__init__()

        def forward(self, data_dict, modality_name):
            return data_dict[f"inputs_{modality_name}"]["bev_feature"]

    orig = he.LiftSplatShoot
    he.LiftSplatShoot = LiftSplatShoot
    try:
        model = st2.HeterModelBaselineWDiffCommStage2(synth.gencomm_stage2_hetero_args("att")).eval()
    finally:
        he.LiftSplatShoot = orig
    model.load_state_dict(synth.fill_state_dict(model.state_dict(), HETER2_WSEED))
    voxels, bev, pairwise, record_len, n0, steps = heter2_inputs()
    data = {"inputs_m1": voxels, "inputs_m2": {"bev_feature": bev}, "agent_modality_list": list(HETER2_MODALITIES),
            "pairwise_t_matrix": pairwise, "record_len": record_len}
    C, H, W = 128, 64, 128
    queue = [n0, torch.zeros(1, C, H, W), torch.zeros(1, C, H, W)]
    step_queue = list(steps)
    orig_randn_like, orig_noise_like = torch.randn_like, cd.noise_like
    torch.randn_like = lambda t, *a, **k: queue.pop(0).to(t)
    cd.noise_like = lambda shape, device, repeat=False: step_queue.pop(0)
    try:
        with torch.no_grad():
            out = model(data)
    finally:
        torch.randn_like, cd.noise_like = orig_randn_like, orig_noise_like
    assert not queue and not step_queue
    keys = sorted(model.state_dict().keys())
    np.savez_compressed(os.path.join(OUT, "heter_model_stage2.npz"), state_dict_keys=np.array(keys),
                        cls_preds=out["cls_preds"].numpy(), reg_preds=out["reg_preds"].numpy(),
                        dir_preds=out["dir_preds"].numpy(), message=out["message"].numpy(),
                        gt_feature_c8=out["gt_feature"][:, ::8].numpy(), pred_feature_c8=out["pred_feature"][:, ::8].numpy())
    print("heter_model_stage2.npz: cls", tuple(out["cls_preds"].shape), "gt", tuple(out["gt_feature"].shape),
          "camera feature nonzero columns", int((out["gt_feature"][1].abs().sum((0, 1)) > 0).sum()),
          "|cls| max", float(out["cls_preds"].abs().max()))


def gen_collate(ns):
    """The UNMODIFIED SpVoxelPreprocessor.collate_batch_dict / collate_batch_list (sp_voxel_preprocessor.py:87-174), called
    unbound (the constructor needs spconv): pins the agent-index column and the concatenation order of the collated
    voxel tensors that feed PillarVFE."""
    import importlib
    import types
    import sys
    for name in ("open3d", "pypcd"):    # pcd_utils.py:10-12 (point-cloud file IO; never called here)
        ref_import._stub(name, pypcd=ref_import._Anything())
    for name in ("spconv", "spconv.utils", "spconv.pytorch", "spconv.pytorch.utils", "cumm", "cumm.tensorview"):
        ref_import._stub(name, VoxelGenerator=ref_import._Anything, VoxelGeneratorV2=ref_import._Anything,
                         Point2VoxelCPU3d=ref_import._Anything, tv=ref_import._Anything())
    mod = importlib.import_module("opencood.data_utils.pre_processor.sp_voxel_preprocessor")
    cls = mod.SpVoxelPreprocessor
    clouds = [synth.lidar_points(77, a, 6000 + 500 * a) for a in range(3)]
    per_agent = [ref_ops.voxelize(c, synth.OPV2V_H_RANGE, [0.4, 0.4, 4.0]) for c in clouds]
    as_np = [{k: np.asarray(v) for k, v in d.items()} for d in per_agent]
    me = types.SimpleNamespace(collate_batch_list=cls.collate_batch_list, collate_batch_dict=cls.collate_batch_dict)
    res_list = cls.collate_batch(me, as_np)                                   # list form (:87-110, :112-142)
    batch = {k: [d[k] for d in as_np] for k in as_np[0]}
    res_dict = cls.collate_batch(me, batch)                                   # dict form (:144-174)
    for k in res_list:
        assert torch.equal(res_list[k], res_dict[k]), k
    np.savez_compressed(os.path.join(OUT, "collate.npz"), voxel_coords=res_dict["voxel_coords"].numpy(),
                        voxel_num_points=res_dict["voxel_num_points"].numpy(),
                        voxel_features_checksum=np.float64(res_dict["voxel_features"].double().sum()),
                        dtypes=np.array([str(res_dict[k].dtype) for k in ("voxel_features", "voxel_coords", "voxel_num_points")]))
    print("collate.npz:", {k: (tuple(v.shape), str(v.dtype)) for k, v in res_dict.items()})


POSTPROCESS_CASES = {"mid": (1, -3.0, False), "cap": (2, -1.0, False), "few": (4, -4.5, True), "none": (3, -9.0, False)}


def postprocess_transform(moved):
    """ego -> ego is the identity (intermediate fusion); ``moved`` exercises project_box3d with a rigid transform."""
    return torch.from_numpy(synth.pose_matrix(1.5, -0.75, 3.0)).float() if moved else torch.eye(4)


def gen_postprocess(ns):
    """The UNMODIFIED VoxelPostprocessor.generate_anchor_box / post_process (decode, direction fix, corners, projection,
    size / z filters, rotated NMS, range mask) on synthetic head maps.  The polygon area inside nms_rotated comes from
    the convex-clipping stand-in for shapely (oracle/ref_import.py::_ConvexPolygon): everything else is reference code."""
    from opencood.data_utils.post_processor.voxel_postprocessor import VoxelPostprocessor
    params = synth.postprocess_params()
    pp = VoxelPostprocessor(params, train=False)
    anchors = pp.generate_anchor_box()
    out = {"anchors": anchors}
    for name, (seed, bias, moved) in POSTPROCESS_CASES.items():
        cls, reg, dr = synth.head_outputs(seed, bias=bias)
        data = {"ego": {"transformation_matrix": postprocess_transform(moved), "anchor_box": torch.from_numpy(anchors)}}
        res = {"ego": {"cls_preds": cls.clone(), "reg_preds": reg.clone(), "dir_preds": dr.clone()}}
        boxes, scores = pp.post_process(data, res)
        n = 0 if boxes is None else boxes.shape[0]
        out[f"{name}/count"] = np.int64(n)
        out[f"{name}/boxes"] = np.zeros((0, 8, 3), np.float32) if boxes is None else boxes.numpy()
        out[f"{name}/scores"] = np.zeros((0,), np.float32) if boxes is None else scores.numpy()
        print("postprocess", name, "candidates", int((torch.sigmoid(cls) > 0.2).sum()), "kept", n)
    np.savez_compressed(os.path.join(OUT, "postprocess.npz"), **out)


def gen_lss_pool(ns):
    """The UNMODIFIED LiftSplatShoot.voxel_pooling (heter_encoders.py:161-217), called unbound on a stand-in ``self`` that
    only carries dx / bx / nx / use_quickcumsum (building the class needs the EfficientNet image encoder)."""
    import types
    from opencood.models.heter_encoders import LiftSplatShoot
    out = {}
    for name, conf, kw in (("z1", synth.LSS_GRID_CONF, {}),
                           ("z2", dict(synth.LSS_GRID_CONF, zbound=[-10, 10, 10.0], xbound=[-20.0, 20.0, 0.8]), {"B": 1, "N": 3})):
        dx, bx, nx = ref_ops.gen_dx_bx(conf["xbound"], conf["ybound"], conf["zbound"])
        geom, x = synth.lss_frustum(31, grid_conf=conf, **kw)
        both = {}
        for quick in (False, True):   # cumsum_trick (camera_utils.py:209-217) and the QuickCumsum autograd function (:220-246)
            me = types.SimpleNamespace(dx=dx, bx=bx, nx=nx, use_quickcumsum=quick)
            both[quick] = LiftSplatShoot.voxel_pooling(me, geom, x)
        res = both[False]
        nz = torch.nonzero(res.abs().sum(1))          # occupied cells only: the grid is sparse
        nzq = torch.nonzero(both[True].abs().sum(1))
        assert torch.equal(nz, nzq), "the two cumsum variants disagree on the occupied cells"
        out[f"{name}/shape"] = np.array(res.shape)
        out[f"{name}/cells"] = nz.numpy().astype(np.int32)
        out[f"{name}/values"] = res[nz[:, 0], :, nz[:, 1], nz[:, 2]].numpy()
        out[f"{name}/values_quickcumsum"] = both[True][nz[:, 0], :, nz[:, 1], nz[:, 2]].numpy()
        print("lss_pool", name, tuple(res.shape), "occupied cells", nz.shape[0], "max |cumsum_trick - QuickCumsum|",
              float((both[False] - both[True]).abs().max()))
    np.savez_compressed(os.path.join(OUT, "lss_pool.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_import.load()
    gen_pillars(ns)
    gen_warp(ns)
    gen_gencomm(ns)
    gen_message_extractor(ns)
    gen_enhancer(ns)
    gen_det_tail(ns)
    gen_backbone(ns)
    gen_heter_model(ns)
    gen_heter_model_stage2(ns)
    gen_collate(ns)
    gen_postprocess(ns)
    gen_lss_pool(ns)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
