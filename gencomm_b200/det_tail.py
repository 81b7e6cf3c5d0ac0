"""Drop-in ``DownsampleConv`` / ``DoubleConv`` (shrink header) and the shared detection heads.

Mirrors ``opencood/models/sub_modules/downsample_conv.py:7-50`` (same class names, config keys and ``state_dict`` keys
``layers.N.double_conv.{0,2}.{weight,bias}``) and the three ``nn.Conv2d`` heads of ``heter_model_baseline.py:130-135``.
``forward`` runs tcgen05 implicit GEMMs through the C ABI (``gc_double_conv`` / ``gc_det_heads``); no CPU path.
"""
import torch
import torch.nn as nn

from . import ops


class DoubleConv(nn.Module):
    """downsample_conv.py:7-27."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding):
        super().__init__()
        if kernel_size != 3 or padding != 1 or stride not in (1, 2):
            raise NotImplementedError("gencomm_b200 DoubleConv: kernel 3, padding 1, stride 1 or 2 (the shipped configs)")
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding), nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1), nn.ReLU(inplace=True))
        self.stride, self.out_channels = stride, out_channels
        self._key, self._blobs = None, None

    def invalidate(self):
        """Drop the packed-weight cache.  The cache key is (data_ptr, _version) of every parameter, which in-place writes
        through ``p.data`` do not change: call this after such an update (``load_state_dict`` does it by itself)."""
        self._key = None

    def _load_from_state_dict(self, *args, **kwargs):
        self._key = None
        return super()._load_from_state_dict(*args, **kwargs)

    @torch.no_grad()
    def forward(self, x):
        c0, c2 = self.double_conv[0], self.double_conv[2]
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if key != self._key:
            self._blobs = ops.double_conv_pack(c0.weight, c0.bias, c2.weight, c2.bias)
            self._key = key
        if isinstance(x, ops.PlaneFeature):      # handed over by this package's BaseBEVBackbone (emit_planes)
            return ops.double_conv_planes(x, self._blobs[0], self._blobs[1], self.out_channels, self.stride)
        return ops.double_conv(x.contiguous(), self._blobs[0], self._blobs[1], self.out_channels, self.stride)


class DownsampleConv(nn.Module):
    """downsample_conv.py:30-50."""

    def __init__(self, config):
        super().__init__()
        self.layers = nn.ModuleList([])
        input_dim = config['input_dim']
        for (ksize, dim, stride, padding) in zip(config['kernal_size'], config['dim'], config['stride'], config['padding']):
            self.layers.append(DoubleConv(input_dim, dim, kernel_size=ksize, stride=stride, padding=padding))
            input_dim = dim

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


class DetectionHeads(nn.Module):
    """cls_head / reg_head / dir_head of heter_model_baseline.py:130-135 as one GEMM.  ``forward(fused)`` returns
    ``(cls_preds, reg_preds, dir_preds)`` like the three separate calls at :165-167."""

    def __init__(self, in_head, anchor_number, num_class=1, dir_bins=2):
        super().__init__()
        self.cls_head = nn.Conv2d(in_head, anchor_number * num_class * num_class, kernel_size=1)
        self.reg_head = nn.Conv2d(in_head, 7 * anchor_number * num_class, kernel_size=1)
        self.dir_head = nn.Conv2d(in_head, dir_bins * anchor_number, kernel_size=1)
        self._key, self._blobs = None, None

    def invalidate(self):
        """Drop the packed-weight cache.  The cache key is (data_ptr, _version) of every parameter, which in-place writes
        through ``p.data`` do not change: call this after such an update (``load_state_dict`` does it by itself)."""
        self._key = None

    def _load_from_state_dict(self, *args, **kwargs):
        self._key = None
        return super()._load_from_state_dict(*args, **kwargs)

    @torch.no_grad()
    def forward(self, x):
        heads = (self.cls_head, self.reg_head, self.dir_head)
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if key != self._key:
            self._blobs = ops.det_heads_pack([h.weight for h in heads], [h.bias for h in heads])
            self._key = key
        packed, bias, splits = self._blobs
        out = ops.det_heads(x.contiguous(), packed, bias)
        return torch.split(out, splits, dim=1)
