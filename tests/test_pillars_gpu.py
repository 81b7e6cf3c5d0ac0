"""GPU parity: voxelize / PillarVFE / PointPillarScatter / fused front end vs the oracle and the
golden vectors.  Integer outputs and the scatter are bit-exact; PFN outputs are bit-exact against
the oracle's kernel-order restatement and within 1e-5*max|ref| of the reference's torch order."""
import numpy as np
import pytest
import torch

from gencomm_b200 import PillarVFE, PointPillar, PointPillarScatter, SpVoxelPreprocessor, ops, synth
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
T = torch.from_numpy
DEV = "cuda"


def _params(rng, vs, cap):
    return {"cav_lidar_range": rng, "args": {"voxel_size": vs, "max_points_per_voxel": 32,
                                             "max_voxel_train": cap, "max_voxel_test": cap}}


def _pfn(g):
    return {k[4:]: T(v) for k, v in g.items() if k.startswith("pfn_")}


def _oracle_batch(clouds, rng, vs, cap):
    return R.collate_voxels([R.voxelize(c, rng, vs, 32, cap) for c in clouds])


def _check_voxels(dev, ref):
    assert torch.equal(dev["voxel_coords"].cpu(), ref["voxel_coords"])
    assert torch.equal(dev["voxel_num_points"].cpu(), ref["voxel_num_points"])
    assert torch.equal(dev["voxel_features"].cpu(), ref["voxel_features"])


@pytest.fixture(params=["fused", "legacy"])
def voxelizer(request, monkeypatch):
    """Both voxelizers of gc_voxelize: k_cell_assign2 + k_pillar_build (default) and the round-1 four-kernel chain."""
    monkeypatch.setenv("GC_VOXELIZE_IMPL", request.param)
    return request.param


def test_voxelize_matches_golden(golden_pillars, voxelizer):
    g = golden_pillars
    rng, vs, cap = g["lidar_range"].tolist(), g["voxel_size"].tolist(), int(g["max_voxels"])
    pre = SpVoxelPreprocessor(_params(rng, vs, cap), train=False)
    d = pre.preprocess_batch([g["points0"], g["points1"], g["points2"]])
    _check_voxels(d, {k: T(g[k]) for k in ("voxel_features", "voxel_coords", "voxel_num_points")})
    # reference contract of preprocess(): numpy dict, coords [M,3] (z,y,x)
    single = pre.preprocess(g["points1"])
    ref1 = R.voxelize(g["points1"], rng, vs, 32, cap)
    for k in ref1:
        assert isinstance(single[k], np.ndarray) and np.array_equal(single[k], ref1[k]), k


@pytest.mark.parametrize("cap", [200, 57])
def test_voxelize_crowded_cells_across_blocks(cap, voxelizer):
    """~250 points per cell spread over ~49 look-back blocks per agent: every pillar overflows its 32 slots, first points
    and their followers sit in different blocks (the cross-block wait of k_pillar_build), run repeatedly on one workspace."""
    rng, vs = [-4.0, -2.0, -3.0, 4.0, 2.0, 1.0], [0.4, 0.4, 4.0]
    g = np.random.default_rng(7)
    clouds = []
    for n in (50_000, 1025, 30_000):
        clouds.append(np.c_[g.uniform(-4.2, 4.2, n), g.uniform(-2.1, 2.1, n), g.uniform(-3, 1, n),
                            g.random(n)].astype(np.float32))
    pre = SpVoxelPreprocessor(_params(rng, vs, cap), train=False)
    ref = _oracle_batch(clouds, rng, vs, cap)
    for _ in range(3):
        _check_voxels(pre.preprocess_batch(clouds), ref)


@pytest.mark.parametrize("uniform,cap", [(False, 70000), (True, 70000), (True, 20000)])
def test_voxelize_full_size_matches_oracle(uniform, cap, voxelizer):
    clouds = [synth.lidar_points(1, a, 100_000, uniform=uniform) for a in range(4)]
    pre = SpVoxelPreprocessor(_params(synth.OPV2V_H_RANGE, synth.VOXEL_SIZE, cap), train=False)
    d = pre.preprocess_batch(clouds)
    ref = _oracle_batch(clouds, synth.OPV2V_H_RANGE, synth.VOXEL_SIZE, cap)
    _check_voxels(d, ref)
    if uniform and cap == 20000:
        assert d["voxel_features"].shape[0] == 4 * cap   # every agent hit the cap


def test_voxelize_edge_cases(voxelizer):
    rng, vs = [-4.0, -2.0, -3.0, 4.0, 2.0, 1.0], [0.4, 0.4, 4.0]   # 20 x 10 grid (nx % 4 == 0, < one tile)
    g = np.random.default_rng(0)
    inside = np.c_[g.uniform(-4, 4, 500), g.uniform(-2, 2, 500), g.uniform(-3, 1, 500), g.random(500)].astype(np.float32)
    outside = inside.copy(); outside[:, 0] += 100
    dense = np.tile(np.array([[0.1, 0.1, 0.0, 0.5]], np.float32), (100, 1)); dense[:, 3] = np.arange(100)
    edge = np.array([[-4.0, -2.0, -3.0, 1], [4.0, 0, 0, 2], [3.9999, 1.9999, 0.9999, 3], [0, 0, 1.0, 4],
                     [np.nan, 0, 0, 5], [0, np.inf, 0, 6], [-4.0000005, 0, 0, 7]], np.float32)
    clouds = [inside, np.zeros((0, 4), np.float32), outside, dense, edge, inside[:1]]
    pre = SpVoxelPreprocessor(_params(rng, vs, 150), train=False)
    d = pre.preprocess_batch(clouds)
    ref = _oracle_batch(clouds, rng, vs, 150)
    _check_voxels(d, ref)
    counts = [R.voxelize(c, rng, vs, 32, 150)["voxel_features"].shape[0] for c in clouds]
    assert counts == [150, 0, 0, 1, 2, 1]
    assert pre._ws.n_pillars.cpu().tolist() == counts


def test_pillar_vfe_and_scatter_modules_match_golden(golden_pillars):
    g, w = golden_pillars, _pfn(golden_pillars)
    rng, vs = g["lidar_range"].tolist(), g["voxel_size"].tolist()
    grid = ops.grid_size(rng, vs)
    vfe = PillarVFE({"use_norm": True, "with_distance": False, "use_absolute_xyz": True, "num_filters": [64]},
                    4, vs, rng)
    vfe.load_state_dict({"pfn_layers.0.linear.weight": w["weight"], "pfn_layers.0.norm.weight": w["bn_weight"],
                         "pfn_layers.0.norm.bias": w["bn_bias"], "pfn_layers.0.norm.running_mean": w["bn_mean"],
                         "pfn_layers.0.norm.running_var": w["bn_var"]}, strict=False)
    vfe = vfe.to(DEV).eval()
    bd = {k: T(g[k]).to(DEV) for k in ("voxel_features", "voxel_coords", "voxel_num_points")}
    bd = vfe(bd)
    out = bd["pillar_features"].cpu()
    sc, sh = R.fold_bn(w["bn_weight"], w["bn_bias"], w["bn_mean"], w["bn_var"])
    ko = R.pillar_vfe_kernel_order(T(g["voxel_features"]), T(g["voxel_num_points"]), T(g["voxel_coords"]),
                                   w["weight"], sc, sh, vs, rng)
    assert torch.equal(out, ko), "kernel differs from its own fixed-order oracle"
    ref = T(g["ref_pillar_features"])
    assert (out - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()

    scat = PointPillarScatter({"num_features": 64, "grid_size": grid})
    canvas = scat(dict(bd))["spatial_features"].cpu()
    assert torch.equal(canvas, R.scatter(out, T(g["voxel_coords"]), int(grid[0]), int(grid[1])))
    # scatter of the reference's own pillar features reproduces the reference canvas bit for bit
    bd2 = {"pillar_features": ref.to(DEV), "voxel_coords": bd["voxel_coords"], "batch_size": 3}
    assert torch.equal(scat(bd2)["spatial_features"].cpu(), T(g["ref_canvas"]))
    with pytest.raises(RuntimeError, match="inference-only"):
        vfe.train()(dict(bd))


@pytest.mark.parametrize("C,nx,ny", [(64, 50, 7), (96, 130, 3), (5, 257, 2)])
def test_scatter_odd_shapes(C, nx, ny):
    g = torch.Generator().manual_seed(C)
    n_batch, per = 3, min(40, nx * ny // 2)
    coords = []
    for b in (0, 2):   # batch element 1 stays empty -> all-zero canvas
        cells = torch.randperm(nx * ny, generator=g)[:per]
        coords.append(torch.stack([torch.full((per,), b), torch.zeros(per, dtype=torch.long), cells // nx, cells % nx], 1))
    coords = torch.cat(coords).to(torch.int32)
    feat = torch.randn(coords.shape[0], C, generator=g)
    out = ops.scatter_canvas(feat.to(DEV), coords.to(DEV), nx, ny, n_batch).cpu()
    assert torch.equal(out, R.scatter(feat, coords, nx, ny, n_batch))
    assert not out[1].any()


@pytest.mark.parametrize("writer", ["tile", "persist"])
@pytest.mark.parametrize("rng,uniform", [(synth.OPV2V_H_RANGE, False), (synth.SQUARE_RANGE, False),
                                         (synth.OPV2V_H_RANGE, True)])
def test_fused_front_end_full_size(rng, uniform, writer, monkeypatch):
    """points -> canvas in one pipeline == oracle voxelize -> kernel-order PFN -> scatter, bit-exact, through both
    canvas writers (per-tile kernel and persistent TMA-store writer)."""
    monkeypatch.setenv("GC_CANVAS_IMPL", writer)
    vs, cap = synth.VOXEL_SIZE, 70000
    clouds = [synth.lidar_points(2, a, 100_000, lidar_range=rng, uniform=uniform) for a in range(4)]
    clouds[2] = clouds[2][:777]   # ragged
    w = synth.pfn_weights(1)
    enc = PointPillar({"lidar_range": rng, "voxel_size": vs, "max_voxels": cap,
                       "pillar_vfe": {"use_norm": True, "with_distance": False, "use_absolute_xyz": True,
                                      "num_filters": [64]},
                       "point_pillar_scatter": {"num_features": 64}})
    layer = enc.pillar_vfe.pfn_layers[0]
    with torch.no_grad():
        layer.linear.weight.copy_(w["weight"]); layer.norm.weight.copy_(w["bn_weight"])
        layer.norm.bias.copy_(w["bn_bias"]); layer.norm.running_mean.copy_(w["bn_mean"])
        layer.norm.running_var.copy_(w["bn_var"])
    enc = enc.to(DEV).eval()
    sizes = [c.shape[0] for c in clouds]
    pts = T(np.concatenate(clouds)).to(DEV)
    off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32, device=DEV)
    canvas = enc({"inputs_m1": {"points": pts, "point_offsets": off, "max_agent_points": max(sizes)}}, "m1").cpu()

    batch = _oracle_batch(clouds, rng, vs, cap)
    sc, sh = R.fold_bn(w["bn_weight"], w["bn_bias"], w["bn_mean"], w["bn_var"])
    feats = R.pillar_vfe_kernel_order(batch["voxel_features"], batch["voxel_num_points"], batch["voxel_coords"],
                                      w["weight"], sc, sh, vs, rng)
    grid = ops.grid_size(rng, vs)
    ref = R.scatter(feats, batch["voxel_coords"], int(grid[0]), int(grid[1]), 4)
    assert torch.equal(canvas, ref)
    # the same through the reference-shaped three-step API (voxel tensors -> PillarVFE -> Scatter)
    pre = SpVoxelPreprocessor(_params(rng, vs, cap), train=False)
    d = pre.preprocess_batch(clouds)
    d["batch_size"] = 4
    canvas2 = enc({"inputs_m1": d}, "m1").cpu()
    assert torch.equal(canvas2, ref)
    # size-independent property: occupied canvas cells == voxel coordinates
    occ = (canvas != 0).any(1).nonzero()
    assert occ.shape[0] <= batch["voxel_coords"].shape[0]
    # the canvas written directly as the backbone's operand planes == the layout conversion of the fp32 canvas, bit for bit
    if writer == "tile":
        enc.emit_planes = True
        feat = enc({"inputs_m1": {"points": pts, "point_offsets": off, "max_agent_points": max(sizes)}}, "m1")
        enc.emit_planes = False
        assert isinstance(feat, ops.PlaneFeature) and feat.shape == tuple(ref.shape)
        xh, xl = ops.to_planes(ref.to(DEV))
        assert torch.equal(feat.xh, xh) and torch.equal(feat.xl, xl)
        # that was the sparse path (persistent all-zero planes, occupied cells written, cleared on release); frames in a row,
        # a frame that is never released (full re-zero on the next one), and the dense writer give the same planes
        assert enc.sparse_planes and feat.on_consumed is not None and enc._planes_dirty
        inp = {"inputs_m1": {"points": pts, "point_offsets": off, "max_agent_points": max(sizes)}}
        enc.emit_planes = True
        try:
            monkeypatch.setenv("GC_SPARSE_WRITER", "tile")   # the tile writer's sparse form (default: pillar-centric kernel)
            feat2 = enc(inp, "m1")                      # previous frame not released -> buffers re-zeroed
            monkeypatch.delenv("GC_SPARSE_WRITER")
            assert torch.equal(feat2.xh, xh) and torch.equal(feat2.xl, xl)
            feat2.consumed()
            assert not enc._planes_dirty and feat2.on_consumed is None
            assert int(feat2.xh.count_nonzero()) == 0 and int(feat2.xl.count_nonzero()) == 0
            half = pts[: sizes[0]]                       # a different frame (first agent only, the others empty)
            off2 = torch.tensor([0] + [sizes[0]] * len(sizes), dtype=torch.int32, device=DEV)
            inp2 = {"inputs_m1": {"points": half, "point_offsets": off2, "max_agent_points": sizes[0]}}
            f3 = enc(inp2, "m1")
            a3, b3 = f3.xh.clone(), f3.xl.clone()
            f3.consumed()
            f4 = enc(inp, "m1")
            assert torch.equal(f4.xh, xh) and torch.equal(f4.xl, xl)
            f4.consumed()
            enc.sparse_planes = False
            d3 = enc(inp2, "m1")
            assert d3.on_consumed is None and torch.equal(d3.xh, a3) and torch.equal(d3.xl, b3)
        finally:
            enc.emit_planes = False
            enc.sparse_planes = True


@pytest.mark.parametrize("writer", ["tile", "persist"])
@pytest.mark.parametrize("nx,ny", [(20, 10), (132, 3), (256, 1), (516, 7)])
def test_canvas_writers_small_and_ragged_grids(nx, ny, writer, monkeypatch):
    """Partial tiles (nx not a multiple of 128), single rows, empty agents, crowded cells (> 32 points), one agent:
    both canvas writers against the oracle, bit-exact."""
    monkeypatch.setenv("GC_CANVAS_IMPL", writer)
    vs = [0.4, 0.4, 4.0]
    rng = [0.0, 0.0, -3.0, 0.4 * nx, 0.4 * ny, 1.0]
    g = np.random.default_rng(nx * 1000 + ny)
    def cloud(n):
        return np.c_[g.uniform(-1, 0.4 * nx + 1, n), g.uniform(-1, 0.4 * ny + 1, n), g.uniform(-3, 1, n),
                     g.random(n)].astype(np.float32)
    crowded = cloud(400); crowded[:300, 0] = 0.4 * (nx - 1) + 0.2; crowded[:300, 1] = 0.2   # 300 points in the last cell of row 0
    clouds = [cloud(3000), np.zeros((0, 4), np.float32), crowded, cloud(1), cloud(5000)]
    for sub in (clouds, clouds[:1]):
        cap = 2000
        w = synth.pfn_weights(3)
        enc = PointPillar({"lidar_range": rng, "voxel_size": vs, "max_voxels": cap,
                           "pillar_vfe": {"use_norm": True, "with_distance": False, "use_absolute_xyz": True,
                                          "num_filters": [64]},
                           "point_pillar_scatter": {"num_features": 64}})
        layer = enc.pillar_vfe.pfn_layers[0]
        with torch.no_grad():
            layer.linear.weight.copy_(w["weight"]); layer.norm.weight.copy_(w["bn_weight"])
            layer.norm.bias.copy_(w["bn_bias"]); layer.norm.running_mean.copy_(w["bn_mean"])
            layer.norm.running_var.copy_(w["bn_var"])
        enc = enc.to(DEV).eval()
        sizes = [c.shape[0] for c in sub]
        pts = T(np.concatenate(sub)).to(DEV)
        off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32, device=DEV)
        canvas = enc({"inputs_m1": {"points": pts, "point_offsets": off, "max_agent_points": max(sizes)}}, "m1").cpu()
        batch = _oracle_batch(sub, rng, vs, cap)
        sc, sh = R.fold_bn(w["bn_weight"], w["bn_bias"], w["bn_mean"], w["bn_var"])
        feats = R.pillar_vfe_kernel_order(batch["voxel_features"], batch["voxel_num_points"], batch["voxel_coords"],
                                          w["weight"], sc, sh, vs, rng)
        ref = R.scatter(feats, batch["voxel_coords"], nx, ny, len(sub))
        assert canvas.shape == ref.shape
        assert torch.equal(canvas, ref)
