"""Per-frame collaborative-perception pipeline over the in-scope operators, batched over frames.

    points[F*N*P,4] --voxelize--> workspace --PFN+scatter--> canvas[F*N,64,ny,nx]
    pairwise[F,L,L,4,4] f64 --normalize_pairwise_tfm--> theta[F,L,L,2,3]
    (canvas, theta, record_len) --warp + regroup + Att/Max fusion--> fused[F,64,ny,nx]

This is BASELINE.json config[1] ("PointPillars + AttFusion, 4 agents, 256x256x64 BEV"): the scatter
canvas feeds the fusion directly; the cuDNN backbone/shrink convs that sit between them in the
full detector are out of scope and are not part of the measured step.

All buffers are preallocated, every launch is asynchronous on the current stream and nothing
synchronises with the host: the sequence can be enqueued back to back or captured in a CUDA graph.
"""
import numpy as np
import torch

from . import ops
from .modules import PointPillar, SpVoxelPreprocessor

KERNELS_PER_STEP = 7   # cell_assign, pillar_count, pillar_assign, slot_insert, canvas, normalize_tfm, warp_fuse


class FramePipeline:
    def __init__(self, n_frames, n_agents, points_per_agent, lidar_range, voxel_size, max_voxels=70000,
                 fusion="att", max_cav=5, device="cuda", pfn=None):
        self.F, self.N, self.P, self.L = int(n_frames), int(n_agents), int(points_per_agent), int(max_cav)
        self.device = torch.device(device)
        self.lidar_range, self.voxel_size = list(lidar_range), list(voxel_size)
        self.mode = {"att": ops.FUSE_ATT, "max": ops.FUSE_MAX}[fusion]
        self.encoder = PointPillar({"lidar_range": self.lidar_range, "voxel_size": self.voxel_size,
                                    "max_voxels": max_voxels,
                                    "pillar_vfe": {"use_norm": True, "with_distance": False,
                                                   "use_absolute_xyz": True, "num_filters": [64]},
                                    "point_pillar_scatter": {"num_features": 64}})
        if pfn is not None:
            layer = self.encoder.pillar_vfe.pfn_layers[0]
            with torch.no_grad():
                layer.linear.weight.copy_(pfn["weight"]); layer.norm.weight.copy_(pfn["bn_weight"])
                layer.norm.bias.copy_(pfn["bn_bias"]); layer.norm.running_mean.copy_(pfn["bn_mean"])
                layer.norm.running_var.copy_(pfn["bn_var"])
        self.encoder = self.encoder.to(self.device).eval()
        self.pre = SpVoxelPreprocessor(
            {"cav_lidar_range": self.lidar_range,
             "args": {"voxel_size": self.voxel_size, "max_points_per_voxel": 32,
                      "max_voxel_train": max_voxels, "max_voxel_test": max_voxels}}, train=False, device=self.device)
        self.encoder._pre = self.pre
        g = ops.grid_size(self.lidar_range, self.voxel_size)
        self.nx, self.ny = int(g[0]), int(g[1])
        # metric extents, heter_model_baseline.py:87-89 convention (H = y extent, W = x extent, ratio 1)
        self.Hm = self.lidar_range[4] - self.lidar_range[1]
        self.Wm = self.lidar_range[3] - self.lidar_range[0]
        A = self.F * self.N
        self.canvas = torch.empty(A, 64, self.ny, self.nx, dtype=torch.float32, device=self.device)
        self.fused = torch.empty(self.F, 64, self.ny, self.nx, dtype=torch.float32, device=self.device)
        self.agent_offsets = torch.arange(0, A + 1, self.N, dtype=torch.int32, device=self.device)
        self.point_offsets = torch.arange(0, A * self.P + 1, self.P, dtype=torch.int32, device=self.device)

    # algorithmic bytes, per step of F frames
    def scatter_bytes(self):
        return self.F * self.N * (16 * self.P + 4 * 64 * self.ny * self.nx)

    def fuse_bytes(self):
        return self.F * (4 * self.N * 64 * self.ny * self.nx + 4 * 64 * self.ny * self.nx + 48 * self.N)

    def encode(self, points, point_offsets=None, events=None):
        """points [F*N*P,4] f32 (device) -> canvas [F*N,64,ny,nx].  events: optional (start, end) pair
        recorded around the canvas writer kernel."""
        off = self.point_offsets if point_offsets is None else point_offsets
        ws = self.pre.voxelize_device(points, off, self.P)
        v = self.encoder.pillar_vfe
        if events:
            events[0].record()
        ops.pillar_canvas(points, off, ws, v.pfn_table(points.device), (v.x_offset, v.y_offset, v.z_offset),
                          out=self.canvas)
        if events:
            events[1].record()
        return self.canvas

    def fuse(self, feat, pairwise, events=None):
        theta = ops.normalize_pairwise_tfm(pairwise, self.Hm, self.Wm, 1.0)
        if events:
            events[0].record()
        ops.warp_fuse(feat, self.agent_offsets, theta, self.mode, out=self.fused, max_agents=self.N)
        if events:
            events[1].record()
        return self.fused

    def step(self, points, pairwise, ev_canvas=None, ev_fuse=None):
        return self.fuse(self.encode(points, events=ev_canvas), pairwise, events=ev_fuse)


def synthetic_step_inputs(step_seed, n_frames, n_agents, points_per_agent, lidar_range, max_cav=5):
    """Host-side (numpy) inputs of one step: points [F*N*P,4] f32 and pairwise [F,L,L,4,4] f64."""
    from . import synth
    clouds, pws = [], []
    half_w = 0.3 * (lidar_range[3] - lidar_range[0])
    half_h = 0.3 * (lidar_range[4] - lidar_range[1])
    for f in range(n_frames):
        frame = step_seed * 1000 + f
        for a in range(n_agents):
            clouds.append(synth.lidar_points(frame, a, points_per_agent, lidar_range=lidar_range))
        pws.append(synth.pairwise_t_matrix(frame, n_agents, max_cav, spread=(half_w, half_h)))
    return np.concatenate(clouds), np.stack(pws)
