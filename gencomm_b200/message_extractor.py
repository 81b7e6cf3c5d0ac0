"""Drop-in ``MessageExtractorv2``.

Mirrors ``opencood/models/gencomm_modules/message_extractor_v2.py:70-120``: same class names, constructor
arguments, ``forward`` signature and ``state_dict`` keys (``bev_extractor.offset1/dcn1/fuse.{0,2}/attn.{1,3}``), so a
reference checkpoint loads unchanged.  ``forward`` runs the sm_100a kernels of ``csrc/message_extractor.cu`` through the
C ABI (``gc_message_extractor``): the two 3x3 layers (plain + deformable) are tcgen05 implicit GEMMs with bf16 operands
and fp32 accumulation.  There is no CPU path.
"""
import math

import torch
import torch.nn as nn

from . import ops


class DeformConv2d(nn.Module):
    """Parameter container with torchvision.ops.DeformConv2d's parameter names, shapes and initialisation
    (weight [out, in, kh, kw], bias [out]; kaiming_uniform(a=sqrt(5)) / uniform(+-1/sqrt(fan_in)))."""

    def __init__(self, in_channels, out_channels, kernel_size=3, padding=1):
        super().__init__()
        if kernel_size != 3 or padding != 1:
            raise NotImplementedError("gencomm_b200 DeformConv2d: 3x3, padding 1 (message_extractor_v2.py:78)")
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, 3, 3))
        self.bias = nn.Parameter(torch.empty(out_channels))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(in_channels * 9)
        nn.init.uniform_(self.bias, -bound, bound)


class BEVDeformableExtractor(nn.Module):
    """message_extractor_v2.py:70-112."""

    def __init__(self, in_channels=128, out_channels=2):
        super().__init__()
        if out_channels != 2:
            raise NotImplementedError("gencomm_b200 MessageExtractorv2: out_channels must be 2 (the GenComm condition)")
        self.offset1 = nn.Conv2d(in_channels, 18, kernel_size=3, padding=1)
        self.dcn1 = DeformConv2d(in_channels, 64, kernel_size=3, padding=1)
        self.fuse = nn.Sequential(nn.Conv2d(64, 64, kernel_size=1), nn.ReLU(), nn.Conv2d(64, out_channels, kernel_size=1))
        self.attn = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(64, 32, kernel_size=1), nn.ReLU(),
                                  nn.Conv2d(32, 64, kernel_size=1), nn.Sigmoid())
        self._key = None
        self._packed = None
        self._params = None

    def invalidate(self):
        """Drop the packed-weight cache.  The cache key is (data_ptr, _version) of every parameter, which in-place writes
        through ``p.data`` do not change: call this after such an update (``load_state_dict`` does it by itself)."""
        self._key = None

    def _load_from_state_dict(self, *args, **kwargs):
        self._key = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _blobs(self):
        ps = list(self.parameters())
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if key != self._key:
            sd = {"x." + k: v for k, v in self.state_dict().items()}
            self._packed = ops.me_pack_weights(self.offset1.weight.detach().contiguous(), self.dcn1.weight.detach().contiguous())
            self._params = ops.me_pack_params(sd, prefix="x.")
            self._key = key
        return self._packed, self._params

    @torch.no_grad()
    def forward(self, x):
        packed, params = self._blobs()
        return ops.message_extractor(x.contiguous(), packed, params)


class MessageExtractorv2(nn.Module):
    """message_extractor_v2.py:114-120."""

    def __init__(self, in_channels=128, out_channels=2):
        super().__init__()
        self.bev_extractor = BEVDeformableExtractor(in_channels, out_channels)

    def forward(self, bev_feature):
        return self.bev_extractor(bev_feature)
