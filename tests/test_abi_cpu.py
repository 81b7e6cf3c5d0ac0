"""The C-ABI library loads and exports every symbol include/gencomm_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest
import torch

from gencomm_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gencomm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound():
    syms = _declared_symbols()
    assert len(syms) >= 10
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes binding table out of sync with the header"
    assert _lib.load().gc_version() >= 100


def test_argument_errors_are_reported_without_a_gpu():
    lib = _lib.load()
    geom = ops.make_geom([-1, -1, -1, 1, 1, 1], [0.5, 0.5, 2], 100, max_points=16)   # unsupported max_points
    rc = lib.gc_voxelize(None, None, 1, 0, 0, ctypes.byref(geom), None, None, None)
    assert rc == -2 and b"32" in lib.gc_last_error()
    rc = lib.gc_warp_fuse(None, None, 1, 1, 0, None, 5, 8, 4, 4, 7, None, None)
    assert rc == -1 and b"mode" in lib.gc_last_error()
    with pytest.raises(RuntimeError, match="argument error"):
        _lib.check(rc, "gc_warp_fuse")


def test_ops_refuse_cpu_tensors():
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.warp_fuse(torch.zeros(1, 1, 2, 2), torch.zeros(2, dtype=torch.int32),
                      torch.zeros(1, 1, 1, 2, 3, dtype=torch.float64), ops.FUSE_MAX)


def test_geometry_and_pfn_packing_host_logic():
    g = ops.make_geom([-102.4, -51.2, -3, 102.4, 51.2, 1], [0.4, 0.4, 4], 70000)
    assert list(g.grid) == [512, 256, 1]
    w = torch.arange(640, dtype=torch.float32).reshape(64, 10) / 100
    gamma, beta, mean, var = torch.full((64,), 2.0), torch.full((64,), -0.5), torch.full((64,), 0.25), torch.ones(64)
    t = ops.pack_pfn(w, gamma, beta, mean, var, eps=0.0)     # scale = 2, shift = -0.5 - 0.25*2 = -1
    assert t.shape == (64, 16)
    assert torch.equal(t[:, 0], ((w[:, 0] + w[:, 4]) + w[:, 7]) * 2) and torch.equal(t[:, 7], -w[:, 4] * 2)
    assert torch.equal(t[:, 3], w[:, 3] * 2) and torch.equal(t[:, 5], w[:, 1] * 2)
    assert torch.equal(t[:, 10], torch.full((64,), -1.0)) and not t[:, 11:].any()


def test_modules_keep_reference_state_dict_keys():
    from gencomm_b200 import PillarVFE
    vfe = PillarVFE({"use_norm": True, "with_distance": False, "use_absolute_xyz": True, "num_filters": [64]},
                    4, [0.4, 0.4, 4], [-102.4, -51.2, -3, 102.4, 51.2, 1])
    keys = set(vfe.state_dict())
    assert {"pfn_layers.0.linear.weight", "pfn_layers.0.norm.weight", "pfn_layers.0.norm.bias",
            "pfn_layers.0.norm.running_mean", "pfn_layers.0.norm.running_var"} <= keys
    with pytest.raises(NotImplementedError):
        PillarVFE({"use_norm": True, "with_distance": True, "use_absolute_xyz": True, "num_filters": [64]},
                  4, [0.4, 0.4, 4], [-102.4, -51.2, -3, 102.4, 51.2, 1])


def test_message_extractor_keeps_reference_state_dict_keys(golden_message_extractor):
    from gencomm_b200 import MessageExtractorv2
    m = MessageExtractorv2(64, 2)
    ref_keys = {k[3:] for k in golden_message_extractor if k.startswith("sd/")}
    assert set(m.state_dict()) == ref_keys
    m.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in golden_message_extractor.items() if k.startswith("sd/")})
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 64, 4, 32))
    with pytest.raises(NotImplementedError):
        MessageExtractorv2(64, 3)


def test_enhancer_keeps_reference_state_dict_keys(golden_enhancer):
    import json
    from gencomm_b200 import Enhancer
    shapes = json.loads(bytes(golden_enhancer["sd_shapes_json"]).decode())
    m = Enhancer(128, [8, 8], 4)
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == shapes
    m.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in golden_enhancer.items() if k.startswith("sd/")}, strict=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 128, 4, 32))


def test_downsample_conv_keeps_reference_state_dict_keys(golden_det_tail):
    from gencomm_b200 import DetectionHeads, DownsampleConv
    m = DownsampleConv({"kernal_size": [3], "stride": [2], "padding": [1], "dim": [64], "input_dim": 64})
    assert set(m.state_dict()) == {k[3:] for k in golden_det_tail if k.startswith("sd/")}
    assert set(DetectionHeads(64, 2).state_dict()) == {"cls_head.weight", "cls_head.bias", "reg_head.weight",
                                                       "reg_head.bias", "dir_head.weight", "dir_head.bias"}
    with pytest.raises(NotImplementedError):
        DownsampleConv({"kernal_size": [5], "stride": [1], "padding": [2], "dim": [64], "input_dim": 64})


def test_bev_backbone_keeps_reference_state_dict_keys(golden_backbone):
    from conftest import BACKBONE_CFG
    from gencomm_b200 import BaseBEVBackbone
    m = BaseBEVBackbone(BACKBONE_CFG, 64)
    assert set(m.state_dict()) == {k[3:] for k in golden_backbone if k.startswith("sd/")}
    assert m.num_bev_features == 192
    with pytest.raises(NotImplementedError):
        BaseBEVBackbone({**BACKBONE_CFG, "upsample_strides": [1, 2, 0.5]}, 64)


def test_postprocessor_and_lss_host_logic():
    """Host-side mirrors that need no GPU: argument errors, anchor generation shape, gen_dx_bx."""
    import pytest
    import torch
    import gencomm_b200 as G
    from gencomm_b200 import ops, synth
    params = synth.postprocess_params()
    pp = G.VoxelPostprocessor(params, train=False)
    assert pp.generate_anchor_box().shape == (64, 128, 2, 7)
    with pytest.raises(NotImplementedError, match="one cav"):
        pp.post_process({"a": {}, "b": {}}, {"a": {}, "b": {}})
    with pytest.raises(ValueError, match="order"):
        ops.make_post_params(0.2, 0.15, 0.78, 2, "xyz", params["gt_range"])
    bad = dict(params, order="whl")
    with pytest.raises(ValueError, match="order"):
        G.VoxelPostprocessor(bad, train=False).generate_anchor_box()
    dx, bx, nx = G.gen_dx_bx([-51.2, 51.2, 0.4], [-51.2, 51.2, 0.4], [-10, 10, 20.0])
    assert nx.tolist() == [256, 256, 1] and torch.allclose(bx, torch.tensor([-51.0, -51.0, 0.0]))
    with pytest.raises(RuntimeError, match="CUDA tensor"):          # no CPU path
        G.voxel_pooling(torch.zeros(1, 1, 1, 1, 1, 3), torch.zeros(1, 1, 1, 1, 1, 4), dx, bx, nx)
