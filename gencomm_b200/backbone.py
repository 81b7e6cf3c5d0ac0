"""Drop-in ``BaseBEVBackbone``.

Mirrors ``opencood/models/sub_modules/base_bev_backbone.py:6-124``: same constructor, sub-module tree and ``state_dict``
keys (``blocks.i.{1,2,4,5,...}``, ``deblocks.i.{0,1}``), same ``forward(data_dict)`` contract
(``data_dict['spatial_features']`` -> ``data_dict['spatial_features_2d']``).  Inference only (BatchNorm in eval mode is
folded into the convolution weights and bias when the operands are packed).  Every layer is a tcgen05 implicit GEMM over
channel-last bf16 value + residual planes (``gc_conv_planes``): activations stay in that layout between the layers, the
up-sampling ``ConvTranspose2d`` (kernel == stride) is evaluated phase by phase as 1x1 GEMMs with a pixel-shuffle store
straight into the concatenated output.  No CPU path.
"""
import torch
import torch.nn as nn

from . import ops


def _fold(conv_w, bn, transposed=False):
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    shift = bn.bias - bn.running_mean * scale
    w = conv_w * (scale.view(1, -1, 1, 1) if transposed else scale.view(-1, 1, 1, 1))
    return w, shift


class BaseBEVBackbone(nn.Module):
    def __init__(self, model_cfg, input_channels):
        super().__init__()
        self.model_cfg = model_cfg
        layer_nums = model_cfg.get('layer_nums', [])
        layer_strides = model_cfg.get('layer_strides', [])
        num_filters = model_cfg.get('num_filters', [])
        upsample_strides = model_cfg.get('upsample_strides', [])
        num_upsample_filters = model_cfg.get('num_upsample_filter', [])
        assert len(layer_nums) == len(layer_strides) == len(num_filters)
        assert len(upsample_strides) == len(num_upsample_filters)
        if len(upsample_strides) != len(layer_nums) or any(int(s) != s or s < 1 for s in upsample_strides):
            raise NotImplementedError("gencomm_b200 BaseBEVBackbone: one ConvTranspose2d deblock (stride >= 1) per level")
        self.num_levels = len(layer_nums)
        c_in_list = [input_channels, *num_filters[:-1]]
        self.blocks = nn.ModuleList()
        self.deblocks = nn.ModuleList()
        for idx in range(self.num_levels):
            cur = [nn.ZeroPad2d(1),
                   nn.Conv2d(c_in_list[idx], num_filters[idx], kernel_size=3, stride=layer_strides[idx], padding=0, bias=False),
                   nn.BatchNorm2d(num_filters[idx], eps=1e-3, momentum=0.01), nn.ReLU()]
            for _ in range(layer_nums[idx]):
                cur.extend([nn.Conv2d(num_filters[idx], num_filters[idx], kernel_size=3, padding=1, bias=False),
                            nn.BatchNorm2d(num_filters[idx], eps=1e-3, momentum=0.01), nn.ReLU()])
            self.blocks.append(nn.Sequential(*cur))
            s = int(upsample_strides[idx])
            self.deblocks.append(nn.Sequential(
                nn.ConvTranspose2d(num_filters[idx], num_upsample_filters[idx], s, stride=s, bias=False),
                nn.BatchNorm2d(num_upsample_filters[idx], eps=1e-3, momentum=0.01), nn.ReLU()))
        self.num_bev_features = sum(num_upsample_filters)
        self._strides = [int(s) for s in layer_strides]
        self._ups = [int(s) for s in upsample_strides]
        self._key, self._packed = None, None
        # True: 'spatial_features_2d' is an ops.PlaneFeature (channel-last bf16 value + residual planes) for this package's
        # DownsampleConv instead of an NCHW fp32 tensor -- the model sets it when the shrink header follows directly
        self.emit_planes = False

    def invalidate(self):
        """Drop the packed-weight cache.  The cache key is (data_ptr, _version) of every parameter, which in-place writes
        through ``p.data`` do not change: call this after such an update (``load_state_dict`` does it by itself)."""
        self._key = None

    def _load_from_state_dict(self, *args, **kwargs):
        self._key = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _pack(self):
        key = tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))
        if key == self._key:
            return self._packed
        levels = []
        for blk, deb, up in zip(self.blocks, self.deblocks, self._ups):
            mods = list(blk)
            convs = []
            for i, m in enumerate(mods):
                if isinstance(m, nn.Conv2d):
                    w, b = _fold(m.weight.detach(), mods[i + 1])
                    convs.append((ops.conv_pack(w), b.float().contiguous(), m.out_channels, m.in_channels, m.stride[0]))
            wt, bt = _fold(deb[0].weight.detach(), deb[1], transposed=True)      # [c_in, c_out, up, up]
            phases = []      # phase (dy, dx) = a 1x1 GEMM; packed back to back in phase order dy * up + dx (gc_conv_planes, up_dy = -1)
            for dy in range(up):
                for dx in range(up):
                    w1 = wt[:, :, dy, dx].t().contiguous().view(wt.shape[1], wt.shape[0], 1, 1)   # -> [c_out, c_in, 1, 1]
                    phases.append(ops.conv_pack(w1))
            levels.append((convs, torch.cat(phases), bt.float().contiguous(), wt.shape[1], wt.shape[0]))
        self._key, self._packed = key, levels
        return levels

    @torch.no_grad()
    def forward(self, data_dict):
        if self.training:
            raise RuntimeError("gencomm_b200 BaseBEVBackbone is inference-only (BatchNorm is folded): call .eval()")
        x = data_dict['spatial_features']
        levels = self._pack()
        if isinstance(x, ops.PlaneFeature):      # handed over by this package's PointPillar (emit_planes)
            (A, C, H, W), planes = x.shape, (x.xh, x.xl)
        else:
            x = x.contiguous()
            A, C, H, W = x.shape
            planes = ops.to_planes(x)
        h, w = H, W
        out, ch_off, C_out = None, 0, self.num_bev_features
        for (convs, phases, up_bias, up_out, up_in), up in zip(levels, self._ups):
            c_in = None
            for packed, bias, n_out, c_in, stride in convs:
                planes, h, w = ops.conv_planes(planes, A, c_in, h, w, packed, bias, n_out, 9, stride=stride)
                if isinstance(x, ops.PlaneFeature):
                    x.consumed()     # the input planes have been read for the last time (PointPillar's sparse planes)
            if out is None:
                Hu, Wu = h * up, w * up
                if self.emit_planes:
                    oh = torch.empty(A * Hu * Wu * C_out * 2, dtype=torch.uint8, device=x.device)
                    out = (oh, torch.empty_like(oh))
                else:
                    out = torch.empty(A, C_out, Hu, Wu, dtype=torch.float32, device=x.device)
            elif (h * up, w * up) != (Hu, Wu):
                raise ValueError("BaseBEVBackbone: the deblock outputs do not share one resolution")
            # all up * up phases of the ConvTranspose2d in one call (one launch when the TMA kernel applies)
            if self.emit_planes:
                ops.conv_planes(planes, A, up_in, h, w, phases, up_bias, up_out, 1, out_planes=out, out_ch_total=C_out,
                                out_ch_off=ch_off, up=up, up_dy=-1)
            else:
                ops.conv_planes(planes, A, up_in, h, w, phases, up_bias, up_out, 1, out_nchw=out, out_ch_off=ch_off, up=up,
                                up_dy=-1)
            ch_off += up_out
        if isinstance(x, ops.PlaneFeature):
            x.consumed()
        if self.emit_planes:
            out = ops.PlaneFeature(out[0], out[1], (A, C_out, Hu, Wu))
        data_dict['spatial_features_2d'] = out
        return data_dict
