"""Drop-in ``Enhancer``.

Mirrors ``opencood/models/gencomm_modules/enhancer.py:316-383``: same class names, constructor arguments, ``forward``
signature and ``state_dict`` keys, so a reference checkpoint loads unchanged -- including the parameters the reference
declares but never evaluates (``block_2``, ``block_3`` and every ``attn`` sub-module: the attention call is commented
out at enhancer.py:326 and only ``block_1`` feeds ``split_attn``, :369-374).  ``forward`` runs the sm_100a kernels of
``csrc/enhancer.cu`` through the C ABI (``gc_enhancer``); there is no CPU path.
"""
import torch
import torch.nn as nn

from . import ops


class LinearProjection(nn.Module):          # enhancer.py:43-62 (parameters only)
    def __init__(self, dim, heads=8, dim_head=64, bias=True):
        super().__init__()
        inner = dim_head * heads
        self.to_q = nn.Linear(dim, inner, bias=bias)
        self.to_kv = nn.Linear(dim, inner * 2, bias=bias)


class Attention(nn.Module):                 # enhancer.py:92-114 (parameters only; never evaluated by the reference)
    def __init__(self, dim, num_heads):
        super().__init__()
        self.angle_bias_table = nn.Parameter(torch.ones(5, num_heads))
        self.qkv = LinearProjection(dim, num_heads, dim // num_heads, bias=True)
        self.proj = nn.Linear(dim, dim)


class FRFN(nn.Module):                      # enhancer.py:205-222
    def __init__(self, dim, hidden_dim):
        super().__init__()
        self.linear1 = nn.Sequential(nn.Linear(dim, hidden_dim * 2), nn.GELU())
        self.dwconv = nn.Sequential(nn.Conv2d(hidden_dim, hidden_dim, groups=hidden_dim, kernel_size=3, stride=1, padding=1),
                                    nn.GELU())
        self.linear2 = nn.Sequential(nn.Linear(hidden_dim, dim))
        self.partial_conv3 = nn.Conv2d(dim // 4, dim // 4, 3, 1, 1, bias=False)


class Enhancer_block(nn.Module):            # enhancer.py:316-324
    def __init__(self, C, win_size, num_heads):
        super().__init__()
        self.attn = Attention(C, num_heads)
        self.mlp = FRFN(C, C * 2)
        self.norm1 = nn.LayerNorm(C)
        self.norm2 = nn.LayerNorm(C)


class SplitAttn(nn.Module):                 # enhancer.py:286-298
    def __init__(self, input_dim):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, input_dim, bias=False)
        self.bn1 = nn.LayerNorm(input_dim)
        self.fc2 = nn.Linear(input_dim, input_dim, bias=False)


class Enhancer(nn.Module):
    """enhancer.py:335-383.  ``forward(x, affine_matrix=None, record_len=None)``: x [sumN,C,H,W] -> [sumN,C,H,W];
    ``affine_matrix`` and ``record_len`` are accepted for signature parity (the reference only uses them to split the
    batch per frame and for the disabled attention)."""

    def __init__(self, C, win_size, num_heads):
        super().__init__()
        self.block_1 = Enhancer_block(C, [4, 4], num_heads)
        self.block_2 = Enhancer_block(C, win_size, num_heads)
        self.block_3 = Enhancer_block(C, [16, 16], num_heads)
        self.split_attn = SplitAttn(C)
        self._key = None
        self._blobs = None

    def invalidate(self):
        """Drop the packed-weight cache.  The cache key is (data_ptr, _version) of every parameter, which in-place writes
        through ``p.data`` do not change: call this after such an update (``load_state_dict`` does it by itself)."""
        self._key = None

    def _load_from_state_dict(self, *args, **kwargs):
        self._key = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _packed(self):
        used = [p for n, p in self.named_parameters() if n.startswith("block_1.") or n.startswith("split_attn.")]
        key = tuple((p.data_ptr(), p._version) for p in used)
        if key != self._key:
            self._blobs = ops.enhancer_pack(self.state_dict())
            self._key = key
        return self._blobs

    @torch.no_grad()
    def forward(self, x, affine_matrix=None, record_len=None):
        packed, params = self._packed()
        return ops.enhancer(x.contiguous(), packed, params)
