"""GPU parity: normalize_pairwise_tfm (bit-exact), warp_affine_simple / MaxFusion / AttFusion within
1e-5 relative (max|delta| / max|ref|) of the fp32 reference, the tolerance north_star states."""
import numpy as np
import pytest
import torch

import gencomm_b200 as G
from gencomm_b200 import ops, synth
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
T = torch.from_numpy
DEV = "cuda"
TOL = 1e-5


@pytest.fixture(params=["persist", "persist16x8", "tile", "gather"], autouse=True)
def fuse_path(request, monkeypatch):
    """Every test runs through every implementation of gc_warp_fuse: the persistent TMA-staged default
    (csrc/warp_fuse_persist.cu: 16x16 tiles, AttFusion park in tensor memory where it fits; also its 16x8 / 2-group
    configuration with the shared-memory park), the round-1b per-tile kernel (csrc/warp_fuse_tile.cu) and the gather
    kernels (csrc/warp_fuse.cu).  Shapes the TMA paths do not take (W % 4 != 0) fall through to the gather kernels."""
    monkeypatch.setenv("GC_WARP_FUSE_GATHER", "1" if request.param == "gather" else "0")
    if request.param == "tile":
        monkeypatch.setenv("GC_FUSE_IMPL", "tile")
    else:
        monkeypatch.delenv("GC_FUSE_IMPL", raising=False)
    if request.param == "persist16x8":
        monkeypatch.setenv("GC_FUSE_CFG", "2")
    else:
        monkeypatch.delenv("GC_FUSE_CFG", raising=False)
    return request.param


def rel_err(out, ref):
    return (out - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)


def test_normalize_pairwise_tfm_bit_exact(golden_warp):
    g = golden_warp
    pw = T(g["pairwise"]).to(DEV)
    theta = G.normalize_pairwise_tfm(pw, float(g["Hm"]), float(g["Wm"]), 1)
    assert torch.equal(theta.cpu(), T(g["ref_theta"]))
    assert torch.equal(pw.cpu(), T(g["pairwise"]))   # input untouched, like the reference


def test_warp_and_fusion_match_golden(golden_warp):
    g = golden_warp
    feat, rl, theta = T(g["feat"]).to(DEV), T(g["record_len"]).to(DEV), T(g["ref_theta"]).to(DEV)
    assert rel_err(G.warp_feature(feat, rl, theta).cpu(), T(g["ref_warped"])) <= TOL
    assert rel_err(G.MaxFusion()(feat, rl, theta).cpu(), T(g["ref_max"])) <= TOL
    assert rel_err(G.AttFusion(feat.shape[1])(feat, rl, theta).cpu(), T(g["ref_att"])) <= TOL
    # warp_affine_simple drop-in on the first frame (3 agents), float64 theta row [0, :3]
    w = G.warp_affine_simple(feat[:3], theta[0, 0, :3], feat.shape[2:])
    assert rel_err(w.cpu(), T(g["ref_warped"])[:3]) <= TOL
    for j, piece in enumerate(G.regroup(feat, rl)):
        assert piece.shape[0] == int(g["record_len"][j])


def _frames(seed, record_len, C, H, W, L=5, sparsity=0.0):
    n = int(sum(record_len))
    feat = synth.bev_features(seed, n, C, H, W, sparsity=sparsity)
    pw = np.stack([synth.pairwise_t_matrix(seed + b, int(k), L, spread=(0.3 * W * 0.4, 0.3 * H * 0.4))
                   for b, k in enumerate(record_len)])
    theta = R.normalize_pairwise_tfm(T(pw), H * 0.4, W * 0.4, 1)
    return feat, torch.tensor(record_len, dtype=torch.int64), theta


@pytest.mark.parametrize("record_len,C,H,W,L", [
    ([4], 64, 64, 64, 5), ([5, 3, 1], 32, 32, 64, 5), ([2, 2], 128, 64, 128, 5), ([3], 16, 20, 50, 5),
    ([8], 24, 33, 31, 8), ([1], 8, 16, 16, 5), ([3, 4], 8, 100, 72, 5), ([2], 4, 40, 260, 5),
    ([8, 6], 16, 48, 64, 8), ([7], 64, 32, 128, 8), ([4, 3], 64, 48, 80, 5), ([5], 12, 40, 36, 5)])
def test_fusion_matches_oracle(record_len, C, H, W, L):
    feat, rl, theta = _frames(10 + C, record_len, C, H, W, L, sparsity=0.3)
    fd, rd, td = feat.to(DEV), rl.to(DEV), theta.to(DEV)
    assert rel_err(G.MaxFusion()(fd, rd, td).cpu(), R.max_fusion(feat, rl, theta)) <= TOL
    assert rel_err(G.AttFusion(C)(fd, rd, td).cpu(), R.att_fusion(feat, rl, theta)) <= TOL
    assert rel_err(G.warp_feature(fd, rd, td).cpu(), R.warp_only(feat, rl, theta)) <= TOL
    # record_len may also be a host list (sync-free offsets)
    assert torch.equal(G.MaxFusion()(fd, record_len, td), G.MaxFusion()(fd, rd, td))


def test_full_size_properties():
    """BASELINE config[1] size (4 agents, C=64, 256x256): size-independent properties."""
    C, H, W = 64, 256, 256
    feat, rl, theta = _frames(77, [4], C, H, W)
    fd, rd, td = feat.to(DEV), rl.to(DEV), theta.to(DEV)
    warped = G.warp_feature(fd, rd, td)
    # identity transform (ego row) passes the ego feature through bit for bit (App. A.5)
    assert torch.equal(warped[0], fd[0])
    # linearity of the warp in the features
    other = synth.bev_features(78, 4, C, H, W).to(DEV)
    lin = G.warp_feature(fd + other, rd, td)
    assert rel_err(lin, warped + G.warp_feature(other, rd, td)) <= 1e-6
    # max fusion == max over the materialised warp; >= ego everywhere
    mx = G.MaxFusion()(fd, rd, td)
    assert torch.equal(mx[0], warped.max(dim=0)[0])
    # fusing N copies of one agent under identity transforms is the identity for max and att
    same = fd[:1].repeat(4, 1, 1, 1).contiguous()
    eye = R.normalize_pairwise_tfm(T(np.tile(np.eye(4), (1, 5, 5, 1, 1))), H * 0.4, W * 0.4, 1).to(DEV)
    assert torch.equal(G.MaxFusion()(same, rd, eye)[0], fd[0])
    assert rel_err(G.AttFusion(C)(same, rd, eye)[0], fd[0]) <= 1e-6
    # full-size oracle comparison for max (CPU reference takes ~0.1 s)
    assert rel_err(mx.cpu(), R.max_fusion(feat, rl, theta)) <= TOL
    assert rel_err(G.AttFusion(C)(fd, rd, td).cpu(), R.att_fusion(feat, rl, theta)) <= TOL


def test_far_away_agent_contributes_zeros():
    C, H, W = 8, 32, 32
    feat = synth.bev_features(5, 2, C, H, W)
    pw = np.tile(np.eye(4), (1, 5, 5, 1, 1))
    pw[0, 0, 1] = synth.pose_matrix(1e4, -1e4, 33.0)
    theta = R.normalize_pairwise_tfm(T(pw), H * 0.4, W * 0.4, 1)
    rl = torch.tensor([2])
    out = G.warp_feature(feat.to(DEV), rl.to(DEV), theta.to(DEV)).cpu()
    assert not out[1].any() and torch.equal(out[0], feat[0])
    assert rel_err(G.AttFusion(C)(feat.to(DEV), rl.to(DEV), theta.to(DEV)).cpu(), R.att_fusion(feat, rl, theta)) <= TOL


def test_non_isometric_affine_overflows_the_box_and_still_matches():
    """A zooming affine map makes the source footprint of a 32x32 tile larger than the 48x48 TMA box:
    the kernel must detect it per tile and sample that agent from global memory."""
    C, H, W = 6, 96, 128
    feat = synth.bev_features(9, 3, C, H, W)
    theta = torch.zeros(1, 5, 5, 2, 3, dtype=torch.float64)
    theta[..., 0, 0] = 1.0; theta[..., 1, 1] = 1.0
    theta[0, 0, 1] = torch.tensor([[2.5, 0.3, 0.1], [-0.2, 1.9, -0.05]])     # zoom out x2.5 / x1.9
    theta[0, 0, 2] = torch.tensor([[0.2, 0.0, 0.4], [0.0, 0.25, -0.3]])      # zoom in x5 / x4
    rl = torch.tensor([3])
    fd, rd, td = feat.to(DEV), rl.to(DEV), theta.to(DEV)
    assert rel_err(G.warp_feature(fd, rd, td).cpu(), R.warp_only(feat, rl, theta)) <= TOL
    assert rel_err(G.MaxFusion()(fd, rd, td).cpu(), R.max_fusion(feat, rl, theta)) <= TOL
    assert rel_err(G.AttFusion(C)(fd, rd, td).cpu(), R.att_fusion(feat, rl, theta)) <= TOL
