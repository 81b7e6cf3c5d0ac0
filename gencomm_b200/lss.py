"""LSS voxel pooling ("splat") for the camera agents.

Mirrors ``LiftSplatShoot.voxel_pooling`` (``opencood/models/heter_encoders.py:161-217``) and ``gen_dx_bx``
(``opencood/utils/camera_utils.py:129-134``).  Drop-in for a reference model instance:

    LiftSplatShoot.voxel_pooling = lambda self, geom_feats, x: gencomm_b200.voxel_pooling(geom_feats, x, self.dx, self.bx, self.nx)

The image encoder (EfficientNet / ResNet ``CamEncode``) and the frustum geometry stay with the reference: they are library
convolutions and tiny matrix products; the pooling is the HBM-bound scatter this package is about.  No CPU path.
"""
import torch
import torch.nn as nn

from . import ops


def gen_dx_bx(xbound, ybound, zbound):
    dx = torch.Tensor([row[2] for row in [xbound, ybound, zbound]])
    bx = torch.Tensor([row[0] + row[2] / 2.0 for row in [xbound, ybound, zbound]])
    nx = torch.LongTensor([(row[1] - row[0]) / row[2] for row in [xbound, ybound, zbound]])
    return dx, bx, nx


def voxel_pooling(geom_feats, x, dx, bx, nx, deterministic=False):
    """geom_feats [B,N,D,H,W,3], x [B,N,D,H,W,C] -> [B, nz*C, ny, nx] (the reference's return value).
    ``deterministic=True``: fixed-point integer reductions, bit-identical run to run like the reference's sort + cumsum
    (the default fp32 reductions differ by ~1e-7 relative between runs)."""
    return ops.lss_voxel_pooling(geom_feats.float().contiguous(), x.float().contiguous(), dx, bx, nx,
                                 deterministic=deterministic)


class VoxelPooling(nn.Module):
    """``VoxelPooling(grid_conf)(geom_feats, x)``; grid_conf = {'xbound','ybound','zbound'} as in the LSS yaml blocks."""

    def __init__(self, grid_conf, deterministic=False):
        super().__init__()
        self.deterministic = deterministic
        dx, bx, nx = gen_dx_bx(grid_conf['xbound'], grid_conf['ybound'], grid_conf['zbound'])
        self.dx = nn.Parameter(dx, requires_grad=False)      # heter_encoders.py:99-101 keeps them as frozen parameters
        self.bx = nn.Parameter(bx, requires_grad=False)
        self.nx = nn.Parameter(nx, requires_grad=False)
        self._host = (dx.clone(), bx.clone(), nx.clone())    # host copies for the launch (no device -> host sync per call)

    @torch.no_grad()
    def forward(self, geom_feats, x):
        return voxel_pooling(geom_feats, x, *self._host, deterministic=self.deterministic)
