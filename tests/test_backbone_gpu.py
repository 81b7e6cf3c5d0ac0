"""GPU parity of BaseBEVBackbone against the golden output of the reference class and the oracle.

Tolerance: every layer is a bf16x3 tensor-core GEMM with fp32 accumulation and BatchNorm folded into the weights; over
the 4 - 17 layer deep stacks: max|d| <= 2e-3 * max|ref|, mean|d| <= 3e-4 * mean|ref|.
"""
import pytest
import torch

from conftest import BACKBONE_CFG, backbone_input
from gencomm_b200 import BaseBEVBackbone, synth
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
T = torch.from_numpy
DEV = "cuda"


def _close(got, ref, what):
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    d = (got - ref).abs()
    assert float(d.max()) <= 2e-3 * float(ref.abs().max()), (what, float(d.max()), float(ref.abs().max()))
    assert float(d.mean()) <= 3e-4 * float(ref.abs().mean()), (what, float(d.mean()), float(ref.abs().mean()))


def _randomise_bn(model):
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
                m.weight.add_(0.2 * torch.randn_like(m.weight)); m.bias.add_(0.1 * torch.randn_like(m.bias))


def test_backbone_matches_golden(golden_backbone):
    g = golden_backbone
    m = BaseBEVBackbone(BACKBONE_CFG, 64)
    m.load_state_dict({k[3:]: T(v) for k, v in g.items() if k.startswith("sd/")})
    out = m.to(DEV).eval()({"spatial_features": backbone_input().to(DEV)})["spatial_features_2d"].cpu()
    _close(out[:, ::4], T(g["ref_out_c4"]), "backbone")


@pytest.mark.parametrize("rows", [False, True])
def test_backbone_shipped_config_matches_oracle(rows, monkeypatch):
    """The GenComm m1 backbone (layer_nums [3,5,8], filters [64,128,256], up-sampling [1,2,4] -> 384 channels) on one
    256 x 512 canvas; ``rows``: the 64- / 128-channel stride-1 layers through the row-staged kernel (conv_rows.cu, opt-in)
    instead of the TMA-fed one."""
    monkeypatch.setenv("GC_CONV_ROWS", "1" if rows else "0")
    torch.manual_seed(5)
    cfg = {"layer_nums": [3, 5, 8], "layer_strides": [2, 2, 2], "num_filters": [64, 128, 256],
           "upsample_strides": [1, 2, 4], "num_upsample_filter": [128, 128, 128]}
    m = BaseBEVBackbone(cfg, 64).eval()
    _randomise_bn(m)
    x = synth.bev_features(77, 1, 64, 256, 512, sparsity=0.7)
    ref = R.bev_backbone(x, {k: v.detach() for k, v in m.state_dict().items()}, cfg["layer_nums"], cfg["layer_strides"],
                         cfg["upsample_strides"])
    out = m.to(DEV)({"spatial_features": x.to(DEV)})["spatial_features_2d"].cpu()
    assert out.shape == (1, 384, 128, 256)
    _close(out, ref, "backbone m1")


def test_backbone_refuses_training_mode():
    m = BaseBEVBackbone(BACKBONE_CFG, 64).to(DEV).train()
    with pytest.raises(RuntimeError, match="inference-only"):
        m({"spatial_features": torch.zeros(1, 64, 64, 128, device=DEV)})


def test_plane_handover_to_shrink_header_is_bit_identical():
    """BaseBEVBackbone(emit_planes) -> DownsampleConv over the channel-last planes (deblock phases written pixel-shuffled into
    the planes, DoubleConv without the NCHW fp32 round trip) returns exactly what the NCHW hand-over returns: the value /
    residual split of an fp32 activation is the same function in the layout-conversion kernel and in the conv epilogue."""
    from gencomm_b200 import DownsampleConv, ops
    torch.manual_seed(9)
    cfg = {"layer_nums": [3, 5, 8], "layer_strides": [2, 2, 2], "num_filters": [64, 128, 256],
           "upsample_strides": [1, 2, 4], "num_upsample_filter": [128, 128, 128]}
    bb = BaseBEVBackbone(cfg, 64).eval()
    _randomise_bn(bb)
    bb = bb.to(DEV)
    x = synth.bev_features(78, 2, 64, 128, 256, sparsity=0.7).to(DEV)
    for stride in (2, 1):
        sh = DownsampleConv({"kernal_size": [3], "stride": [stride], "padding": [1], "dim": [128], "input_dim": 384}).to(DEV).eval()
        nchw = bb({"spatial_features": x})["spatial_features_2d"]
        want = sh(nchw)
        bb.emit_planes = True
        try:
            feat = bb({"spatial_features": x})["spatial_features_2d"]
        finally:
            bb.emit_planes = False
        assert isinstance(feat, ops.PlaneFeature) and feat.shape == tuple(nchw.shape)
        got = sh(feat)
        assert got.shape == want.shape and torch.equal(got, want), float((got - want).abs().max())
