"""Host-side state machines added in round 2, on the CPU (no kernel is called: the ops entry points are stubbed)."""
import torch

from gencomm_b200 import PointPillar, ops


def _pillar():
    return PointPillar({"lidar_range": [-4.0, -2.0, -3.0, 4.0, 2.0, 1.0], "voxel_size": [0.4, 0.4, 4.0], "max_voxels": 100,
                        "pillar_vfe": {"use_norm": True, "with_distance": False, "use_absolute_xyz": True, "num_filters": [64]},
                        "point_pillar_scatter": {"num_features": 64}}).eval()


def test_plane_feature_consumed_runs_once():
    calls = []
    f = ops.PlaneFeature(torch.zeros(4, dtype=torch.uint8), torch.zeros(4, dtype=torch.uint8), (1, 64, 1, 1))
    f.consumed()                      # nothing registered: no-op
    f.on_consumed = lambda: calls.append(1)
    f.consumed()
    f.consumed()
    assert calls == [1] and f.on_consumed is None


def test_sparse_planes_buffer_state_machine(monkeypatch):
    """The persistent planes are all-zero between frames: allocated as zeros, re-zeroed in full when a frame was never
    released, handed back clean by _planes_release (which clears through the C ABI: stubbed here)."""
    enc = _pillar()
    cleared = []
    monkeypatch.setattr(ops, "plane_bytes", lambda ws: 64)
    monkeypatch.setattr(ops, "planes_clear_occupied", lambda ws, xh, xl: (cleared.append(ws), xh.zero_(), xl.zero_()))
    dev = torch.device("cpu")
    xh, xl = enc._planes_out("ws0", dev)
    assert enc._planes_dirty and int(xh.count_nonzero()) == 0 and xh.numel() == 64
    xh[3] = 7; xl[5] = 9              # the writer's occupied cells
    enc._planes_release("ws0")
    assert cleared == ["ws0"] and not enc._planes_dirty and int(xh.count_nonzero()) == 0 and int(xl.count_nonzero()) == 0
    a, b = enc._planes_out("ws1", dev)
    assert a is xh and b is xl        # same buffers, no re-zero needed
    a[1] = 1
    c, d = enc._planes_out("ws2", dev)   # frame ws1 never released: full re-zero
    assert c is xh and int(c.count_nonzero()) == 0 and enc._planes_dirty
    monkeypatch.setattr(ops, "plane_bytes", lambda ws: 128)
    e, _ = enc._planes_out("ws3", dev)   # another canvas size: new zeroed buffers
    assert e is not xh and e.numel() == 128 and int(e.count_nonzero()) == 0
