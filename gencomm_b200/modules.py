"""Drop-in operators with the reference's names, constructor arguments, forward signatures and
state_dict keys, backed by the sm_100a kernels in ``csrc/``.

Reference counterparts (paths relative to /root/reference/opencood):

=====================  ==========================================================================
``SpVoxelPreprocessor``  data_utils/pre_processor/sp_voxel_preprocessor.py:18-174
``PFNLayer``/``PillarVFE``  models/sub_modules/pillar_vfe.py:10-155
``PointPillarScatter``   models/sub_modules/point_pillar_scatter.py:9-76
``PointPillar``          models/heter_encoders.py:22-50
``warp_affine_simple``   models/sub_modules/torch_transformation_utils.py:323-332
``normalize_pairwise_tfm``  utils/transformation_utils.py:68-92
``regroup`` ``warp_feature`` ``MaxFusion`` ``AttFusion``  models/fuse_modules/fusion_in_one.py:48-151
=====================  ==========================================================================

All modules are inference-only (the hot path is ``torch.no_grad()`` inference,
tools/inference.py:135); calling them in training mode raises.  There is no CPU fallback: inputs
must be CUDA tensors and the CUDA library must be built.
"""
import sys

import numpy as np
import os

import torch
import torch.nn as nn

from . import ops


def _as_offsets(record_len, device):
    """record_len (tensor on any device, list or ndarray) -> [B+1] int32 offsets on `device`, sync-free."""
    if isinstance(record_len, torch.Tensor):
        return ops.agent_offsets_from_record_len(record_len.to(device))
    rl = torch.as_tensor(np.asarray(record_len), dtype=torch.int32)
    off = torch.zeros(rl.numel() + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(rl, 0)
    return off.to(device)


# ------------------------------------------------------------------------------------------------
# voxelizer
# ------------------------------------------------------------------------------------------------
class SpVoxelPreprocessor:
    """``SpVoxelPreprocessor(preprocess_params, train).preprocess(pcd_np)`` on the GPU.

    Reproduces spconv ``Point2VoxelCPU3d.point_to_voxel`` (sequential first-come semantics,
    ``max_num_points_per_voxel`` and ``max_num_voxels`` caps) with an order-preserving parallel
    formulation (csrc/pillars.cu).  ``preprocess`` keeps the reference contract (numpy in, numpy
    dict out); ``preprocess_batch`` keeps everything on the device for the fused front end.
    """

    def __init__(self, preprocess_params, train, device="cuda"):
        self.params = preprocess_params
        self.train = train
        self.device = torch.device(device)
        self.lidar_range = self.params['cav_lidar_range']
        self.voxel_size = self.params['args']['voxel_size']
        self.max_points_per_voxel = self.params['args']['max_points_per_voxel']
        self.max_voxels = self.params['args']['max_voxel_train' if train else 'max_voxel_test']
        self.grid_size = ops.grid_size(self.lidar_range, self.voxel_size)
        self.geom = ops.make_geom(self.lidar_range, self.voxel_size, self.max_voxels, self.max_points_per_voxel)
        self._ws = None

    def _workspace(self, n_agents, total_points):
        ws = self._ws
        if ws is None or ws.n_agents != n_agents or ws.total_points != total_points:
            ws = self._ws = ops.VoxelWorkspace(self.geom, n_agents, total_points, self.device)
        return ws

    def voxelize_device(self, points, point_offsets, max_agent_points=0):
        """points [sumP,4] f32 cuda, point_offsets [A+1] i32 cuda -> workspace (sync-free)."""
        ws = self._workspace(point_offsets.numel() - 1, points.shape[0])
        ops.voxelize(points, point_offsets, ws, max_agent_points)
        return ws

    def preprocess_batch(self, pcd_list):
        """list of [P_a,4] numpy clouds -> collated device dict (voxel_features, voxel_coords [M,4], voxel_num_points)."""
        sizes = [int(p.shape[0]) for p in pcd_list]
        host = np.concatenate([np.ascontiguousarray(p, np.float32).reshape(-1, 4) for p in pcd_list]) \
            if sum(sizes) else np.zeros((0, 4), np.float32)
        pts = torch.from_numpy(host).to(self.device)
        off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32, device=self.device)
        ws = self.voxelize_device(pts, off, max(sizes) if sizes else 0)
        voxels, coords, npts = ops.voxel_gather(pts, off, ws)
        return {'voxel_features': voxels, 'voxel_coords': coords, 'voxel_num_points': npts}

    def preprocess(self, pcd_np):
        """Reference contract: dict of numpy arrays, coords [M,3] (z,y,x) int32."""
        d = self.preprocess_batch([pcd_np])
        return {'voxel_features': d['voxel_features'].cpu().numpy(),
                'voxel_coords': d['voxel_coords'][:, 1:].contiguous().cpu().numpy(),
                'voxel_num_points': d['voxel_num_points'].cpu().numpy()}

    # host-side collation, same contract as the reference (:87-174)
    def collate_batch(self, batch):
        if isinstance(batch, list):
            return self.collate_batch_list(batch)
        if isinstance(batch, dict):
            return self.collate_batch_dict(batch)
        sys.exit('Batch has too be a list or a dictionarn')

    @staticmethod
    def collate_batch_list(batch):
        return SpVoxelPreprocessor.collate_batch_dict({
            'voxel_features': [b['voxel_features'] for b in batch],
            'voxel_num_points': [b['voxel_num_points'] for b in batch],
            'voxel_coords': [b['voxel_coords'] for b in batch]})

    @staticmethod
    def collate_batch_dict(batch):
        coords = [np.pad(c, ((0, 0), (1, 0)), mode='constant', constant_values=i)
                  for i, c in enumerate(batch['voxel_coords'])]
        return {'voxel_features': torch.from_numpy(np.concatenate(batch['voxel_features'])),
                'voxel_coords': torch.from_numpy(np.concatenate(coords)),
                'voxel_num_points': torch.from_numpy(np.concatenate(batch['voxel_num_points']))}


# ------------------------------------------------------------------------------------------------
# PillarVFE / PointPillarScatter / PointPillar encoder
# ------------------------------------------------------------------------------------------------
class PFNLayer(nn.Module):
    """Parameter container with the reference's state_dict keys (``linear.weight``, ``norm.*``)."""

    def __init__(self, in_channels, out_channels, use_norm=True, last_layer=False):
        super().__init__()
        if not (use_norm and last_layer):
            raise NotImplementedError("gencomm_b200 PillarVFE implements use_norm=True with a single PFN layer "
                                      "(num_filters=[64]), the configuration of every hot-path yaml")
        self.last_vfe, self.use_norm = last_layer, use_norm
        self.linear = nn.Linear(in_channels, out_channels, bias=False)
        self.norm = nn.BatchNorm1d(out_channels, eps=1e-3, momentum=0.01)


class PillarVFE(nn.Module):
    def __init__(self, model_cfg, num_point_features, voxel_size, point_cloud_range):
        super().__init__()
        self.model_cfg = model_cfg
        self.use_norm = model_cfg['use_norm']
        self.with_distance = model_cfg['with_distance']
        self.use_absolute_xyz = model_cfg['use_absolute_xyz']
        self.num_filters = list(model_cfg['num_filters'])
        if not self.use_absolute_xyz or self.with_distance or self.num_filters != [64] or num_point_features != 4:
            raise NotImplementedError("gencomm_b200 PillarVFE: only use_absolute_xyz=True, with_distance=False, "
                                      "num_filters=[64], 4 point features are implemented")
        self.pfn_layers = nn.ModuleList([PFNLayer(10, 64, self.use_norm, last_layer=True)])
        self.voxel_x, self.voxel_y, self.voxel_z = voxel_size[0], voxel_size[1], voxel_size[2]
        self.x_offset = self.voxel_x / 2 + point_cloud_range[0]
        self.y_offset = self.voxel_y / 2 + point_cloud_range[1]
        self.z_offset = self.voxel_z / 2 + point_cloud_range[2]
        self._pfn_cache = None

    def get_output_feature_dim(self):
        return self.num_filters[-1]

    def invalidate(self):
        """Drop the packed PFN table (needed after in-place writes through ``p.data``, which the cache key cannot see)."""
        self._pfn_cache = None

    def _load_from_state_dict(self, *args, **kwargs):
        self._pfn_cache = None
        return super()._load_from_state_dict(*args, **kwargs)

    def pfn_table(self, device):
        """Packed [64,16] PFN table on `device`, rebuilt whenever a parameter/buffer changes."""
        layer = self.pfn_layers[0]
        ts = (layer.linear.weight, layer.norm.weight, layer.norm.bias, layer.norm.running_mean,
              layer.norm.running_var)
        key = tuple((t.data_ptr(), t._version) for t in ts) + (str(device),)
        if self._pfn_cache is None or self._pfn_cache[0] != key:
            table = ops.pack_pfn(*ts, eps=layer.norm.eps).to(device)
            self._pfn_cache = (key, table)
        return self._pfn_cache[1]

    def forward(self, batch_dict):
        if self.training:
            raise RuntimeError("gencomm_b200 PillarVFE is inference-only: call .eval()")
        vf, npts, coords = batch_dict['voxel_features'], batch_dict['voxel_num_points'], batch_dict['voxel_coords']
        feats = ops.pillar_vfe(vf.contiguous(), npts.to(torch.int32).contiguous(),
                               coords.to(torch.int32).contiguous(), self.pfn_table(vf.device),
                               (self.voxel_x, self.voxel_y, self.voxel_z),
                               (self.x_offset, self.y_offset, self.z_offset))
        batch_dict['pillar_features'] = feats.squeeze()   # reference squeezes (pillar_vfe.py:152)
        return batch_dict


class PointPillarScatter(nn.Module):
    def __init__(self, model_cfg):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_bev_features = model_cfg['num_features']
        self.nx, self.ny, self.nz = (int(v) for v in model_cfg['grid_size'])
        assert self.nz == 1

    def forward(self, batch_dict):
        pf, coords = batch_dict['pillar_features'], batch_dict['voxel_coords']
        if pf.dim() == 1:   # M == 1 after the reference's squeeze()
            pf = pf.unsqueeze(0)
        # The reference derives the batch size from the data with a host sync
        # (point_pillar_scatter.py:45).  Pass batch_dict['batch_size'] to stay sync-free.
        batch_size = batch_dict.get('batch_size')
        if batch_size is None:
            batch_size = int(coords[:, 0].max().item()) + 1
        batch_dict['spatial_features'] = ops.scatter_canvas(
            pf.contiguous(), coords.to(torch.int32).contiguous(), self.nx, self.ny * self.nz, int(batch_size))
        return batch_dict


class PointPillar(nn.Module):
    """heter_encoders.PointPillar: PillarVFE + PointPillarScatter behind ``forward(data_dict, modality_name)``.

    Extension: when ``data_dict['inputs_<m>']`` carries raw ``points`` [sumP,4] + ``point_offsets`` [A+1]
    (device tensors) instead of spconv voxels, the fused voxelize -> PFN -> canvas kernels run and no
    voxel tensor is ever materialised.
    """

    def __init__(self, args):
        super().__init__()
        grid_size = ops.grid_size(args['lidar_range'], args['voxel_size'])
        args['point_pillar_scatter']['grid_size'] = grid_size
        self.pillar_vfe = PillarVFE(args['pillar_vfe'], num_point_features=4, voxel_size=args['voxel_size'],
                                    point_cloud_range=args['lidar_range'])
        self.scatter = PointPillarScatter(args['point_pillar_scatter'])
        self.lidar_range, self.voxel_size = args['lidar_range'], args['voxel_size']
        self.max_voxels = int(args.get('max_voxels', 70000))
        self._pre = None
        # True: the raw-point path returns an ops.PlaneFeature (the backbone's operand planes) instead of the fp32 canvas --
        # set by the model when this package's BaseBEVBackbone consumes the canvas directly
        self.emit_planes = False
        # the planes live in a persistent buffer that is all-zero between frames: a frame writes its occupied cells only
        # (16 % of the 512 x 256 grid at 100 k points) and zeroes them again once the backbone's first convolution has
        # been enqueued (ops.PlaneFeature.consumed) -- 0.66 GB of stores per 60 agents instead of 2.0 GB
        self.sparse_planes = os.environ.get("GC_SPARSE_PLANES", "1") != "0"     # A/B switch
        self._planes = None
        self._planes_dirty = False

    def _planes_out(self, ws, device):
        n = ops.plane_bytes(ws)
        if self._planes is None or self._planes[0].numel() != n or self._planes[0].device != device:
            self._planes = (torch.zeros(n, dtype=torch.uint8, device=device), torch.zeros(n, dtype=torch.uint8, device=device))
        elif self._planes_dirty:     # the previous frame's planes were never released by a consumer
            self._planes[0].zero_()
            self._planes[1].zero_()
        self._planes_dirty = True
        return self._planes

    def _planes_release(self, ws):
        ops.planes_clear_occupied(ws, *self._planes)
        self._planes_dirty = False

    def forward(self, data_dict, modality_name):
        inp = data_dict[f'inputs_{modality_name}']
        if 'points' in inp:
            return self.forward_points(inp['points'], inp['point_offsets'], inp.get('max_agent_points', 0))
        batch_dict = {'voxel_features': inp['voxel_features'], 'voxel_coords': inp['voxel_coords'],
                      'voxel_num_points': inp['voxel_num_points']}
        if 'batch_size' in inp:
            batch_dict['batch_size'] = inp['batch_size']
        batch_dict = self.pillar_vfe(batch_dict)
        batch_dict = self.scatter(batch_dict)
        return batch_dict['spatial_features']

    def forward_points(self, points, point_offsets, max_agent_points=0, out=None):
        if self.training:
            raise RuntimeError("gencomm_b200 PointPillar is inference-only: call .eval()")
        if self._pre is None:
            params = {'cav_lidar_range': self.lidar_range,
                      'args': {'voxel_size': self.voxel_size, 'max_points_per_voxel': 32,
                               'max_voxel_train': self.max_voxels, 'max_voxel_test': self.max_voxels}}
            self._pre = SpVoxelPreprocessor(params, train=False, device=points.device)
        ws = self._pre.voxelize_device(points, point_offsets, max_agent_points)
        v = self.pillar_vfe
        if self.emit_planes and out is None:
            if not self.sparse_planes:
                return ops.pillar_canvas_planes(points, point_offsets, ws, v.pfn_table(points.device),
                                                (v.x_offset, v.y_offset, v.z_offset))
            feat = ops.pillar_canvas_planes(points, point_offsets, ws, v.pfn_table(points.device),
                                            (v.x_offset, v.y_offset, v.z_offset), out=self._planes_out(ws, points.device),
                                            sparse=True)
            feat.on_consumed = lambda: self._planes_release(ws)
            return feat
        return ops.pillar_canvas(points, point_offsets, ws, v.pfn_table(points.device),
                                 (v.x_offset, v.y_offset, v.z_offset), out=out)


class BEVFeatureInput(nn.Module):
    """Encoder boundary of a camera modality: returns the BEV feature [n, C, Hc, Wc] the reference's ``LiftSplatShoot``
    encoder produced for the n agents of this modality (``data_dict['inputs_<m>']['bev_feature']``, device f32)."""

    def __init__(self, args=None):
        super().__init__()
        self.args = args

    def forward(self, data_dict, modality_name):
        inp = data_dict[f'inputs_{modality_name}']
        if 'bev_feature' not in inp:
            raise KeyError(f"gencomm_b200: camera modality {modality_name} needs data_dict['inputs_{modality_name}']"
                           "['bev_feature'] (the LSS encoder output) or an encoder plugged in with set_encoder()")
        return inp['bev_feature'].contiguous()


def center_crop(x, out_h, out_w):
    """torchvision.transforms.CenterCrop((out_h, out_w)) on a [..., H, W] tensor: a target larger than the input is
    zero-PADDED (left/top get floor, right/bottom ceil of the difference), then the centre window is taken."""
    h, w = x.shape[-2:]
    if out_w > w or out_h > h:
        pl = (out_w - w) // 2 if out_w > w else 0
        pt = (out_h - h) // 2 if out_h > h else 0
        pr = (out_w - w + 1) // 2 if out_w > w else 0
        pb = (out_h - h + 1) // 2 if out_h > h else 0
        x = torch.nn.functional.pad(x, (pl, pr, pt, pb))
        h, w = x.shape[-2:]
        if (h, w) == (out_h, out_w):
            return x
    top, left = int(round((h - out_h) / 2.0)), int(round((w - out_w) / 2.0))
    return x[..., top:top + out_h, left:left + out_w]


# ------------------------------------------------------------------------------------------------
# pose normalisation, warp, regroup, fusion
# ------------------------------------------------------------------------------------------------
def normalize_pairwise_tfm(pairwise_t_matrix, H, W, discrete_ratio, downsample_rate=1):
    """[B,L,L,4,4] -> [B,L,L,2,3]; returns a new tensor (the reference also leaves its input untouched)."""
    if pairwise_t_matrix.dtype != torch.float64:
        pairwise_t_matrix = pairwise_t_matrix.double()
    return ops.normalize_pairwise_tfm(pairwise_t_matrix.contiguous(), H, W, discrete_ratio, downsample_rate)


def regroup(x, record_len):
    """Split by record_len.  Needs the lengths on the host (tensor_split), like the reference's ``.cpu()``
    (fusion_in_one.py:50); pass a list / CPU tensor to avoid the device sync."""
    rl = record_len.cpu() if isinstance(record_len, torch.Tensor) else torch.as_tensor(np.asarray(record_len))
    return torch.tensor_split(x, torch.cumsum(rl, dim=0)[:-1])


def warp_affine_simple(src, M, dsize, mode='bilinear', padding_mode='zeros', align_corners=False):
    """``mode`` / ``padding_mode`` are accepted and ignored, exactly like the reference (App. B.3)."""
    B, C, H, W = src.size()
    if tuple(dsize) != (H, W):
        raise NotImplementedError("gencomm_b200 warp_affine_simple: dsize must equal the source size "
                                  "(every call site on the hot path passes (H, W))")
    if align_corners:
        raise NotImplementedError("gencomm_b200 warp_affine_simple: align_corners=False only")
    theta = M.to(torch.float64).reshape(B, 1, 1, 2, 3).contiguous()
    off = torch.arange(B + 1, dtype=torch.int32, device=src.device)
    return ops.warp_fuse(src.contiguous(), off, theta, ops.FUSE_WARP_ONLY)


def _max_agents(record_len):
    """Host-side bound on the agents per frame when record_len lives on the host (no device read-back)."""
    if isinstance(record_len, torch.Tensor):
        return int(record_len.max()) if record_len.device.type == "cpu" and record_len.numel() else 0
    rl = np.asarray(record_len)
    return int(rl.max()) if rl.size else 0


def _fuse(x, record_len, affine_matrix, mode):
    off = _as_offsets(record_len, x.device)
    theta = affine_matrix if affine_matrix.dtype == torch.float64 else affine_matrix.double()
    return ops.warp_fuse(x.contiguous(), off, theta.contiguous(), mode, max_agents=_max_agents(record_len))


def warp_feature(x, record_len, affine_matrix):
    return _fuse(x, record_len, affine_matrix, ops.FUSE_WARP_ONLY)


class MaxFusion(nn.Module):
    def forward(self, x, record_len, affine_matrix):
        return _fuse(x, record_len, affine_matrix, ops.FUSE_MAX)


class AttFusion(nn.Module):
    def __init__(self, feature_dims):
        super().__init__()
        self.feature_dims = feature_dims   # sqrt(dim) scaling is derived from the tensor's C, as in the reference configs

    def forward(self, xx, record_len, affine_matrix):
        if xx.shape[1] != self.feature_dims:
            raise ValueError(f"AttFusion(feature_dims={self.feature_dims}) got {xx.shape[1]} channels")
        return _fuse(xx, record_len, affine_matrix, ops.FUSE_ATT)
