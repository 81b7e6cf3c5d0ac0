"""GPU parity of the LSS voxel pooling (``LiftSplatShoot.voxel_pooling``).

Bar: the voxel assignment is integer work -> the set of occupied cells is identical to the reference's; the per-voxel
sums are fp32 reductions in a different order than the reference's cumsum differences: |d| <= 1e-5 * max|ref| against the
float64 per-voxel sums (the reference itself is only within ~1e-4 of those), <= 2e-4 absolute against the reference.
"""
import pytest
import torch

import gencomm_b200 as G
from conftest import LSS_CASES, lss_case
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
T = torch.from_numpy
DEV = "cuda"


@pytest.mark.parametrize("name", list(LSS_CASES))
def test_voxel_pooling_matches_reference(golden_lss_pool, name):
    g = golden_lss_pool
    geom, x, conf = lss_case(name)
    pool = G.VoxelPooling(conf).to(DEV)
    out = pool(geom.to(DEV), x.to(DEV)).cpu()
    assert list(out.shape) == g[f"{name}/shape"].tolist()
    cells = T(g[f"{name}/cells"]).long()
    occ = torch.nonzero(out.abs().sum(1))
    assert torch.equal(occ, cells), name                                   # identical occupied cells
    ref = T(g[f"{name}/values"])
    assert float((out[cells[:, 0], :, cells[:, 1], cells[:, 2]] - ref).abs().max()) <= 2e-4
    dx, bx, nx = R.gen_dx_bx(conf["xbound"], conf["ybound"], conf["zbound"])
    exact = R.lss_voxel_pooling_exact(geom, x, dx, bx, nx)
    assert float((out.double() - exact).abs().max()) <= 1e-5 * float(exact.abs().max())


def test_voxel_pooling_function_form_and_empty_input():
    geom, x, conf = lss_case("z1")
    dx, bx, nx = G.gen_dx_bx(conf["xbound"], conf["ybound"], conf["zbound"])
    a = G.voxel_pooling(geom.to(DEV), x.to(DEV), dx, bx, nx)
    b = G.VoxelPooling(conf).to(DEV)(geom.to(DEV), x.to(DEV))
    assert a.shape == b.shape and float((a - b).abs().max()) <= 1e-5       # atomics: run-to-run order may differ
    # scalar-reduction path (no workspace; also taken when C % 4 != 0) against the 128-bit-reduction path
    from gencomm_b200 import ops
    c = ops.lss_voxel_pooling(geom.to(DEV), x.to(DEV), dx, bx, nx, vector=False)
    assert float((a - c).abs().max()) <= 1e-5
    x3 = x[..., :3].contiguous()
    d3 = G.voxel_pooling(geom.to(DEV), x3.to(DEV), dx, bx, nx)
    assert float((d3 - a[:, :3]).abs().max()) <= 1e-5
    far = torch.full_like(geom, 1.0e4)
    assert float(G.voxel_pooling(far.to(DEV), x.to(DEV), dx, bx, nx).abs().max()) == 0.0


@pytest.mark.parametrize("name", list(LSS_CASES))
def test_deterministic_voxel_pooling(golden_lss_pool, name):
    """deterministic=True: fixed-point 64-bit integer reductions -> bit-identical on every run (the reference's sort +
    cumsum is deterministic too), same occupied cells, same accuracy bound as the fp32-reduction path."""
    g = golden_lss_pool
    geom, x, conf = lss_case(name)
    pool = G.VoxelPooling(conf, deterministic=True).to(DEV)
    a = pool(geom.to(DEV), x.to(DEV))
    for _ in range(3):
        assert torch.equal(pool(geom.to(DEV), x.to(DEV)), a)
    out = a.cpu()
    cells = T(g[f"{name}/cells"]).long()
    assert torch.equal(torch.nonzero(out.abs().sum(1)), cells)
    assert float((out[cells[:, 0], :, cells[:, 1], cells[:, 2]] - T(g[f"{name}/values"])).abs().max()) <= 2e-4
    dx, bx, nx = R.gen_dx_bx(conf["xbound"], conf["ybound"], conf["zbound"])
    exact = R.lss_voxel_pooling_exact(geom, x, dx, bx, nx)
    assert float((out.double() - exact).abs().max()) <= 1e-5 * float(exact.abs().max())
