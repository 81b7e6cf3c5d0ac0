"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference modules from /root/reference.

Used exclusively by ``oracle/gen_golden.py`` (run in the build container, where
``/root/reference`` is mounted read-only) to pin the oracle restatement in
``oracle/ref_ops.py`` and to write the golden fixtures under ``tests/golden/``.
Nothing in ``gencomm_b200/``, ``bench.py``, ``smoke()`` or the ``-m gpu`` tests imports
this file: ``/root/reference`` does not exist on the GPU box.

The reference imports a handful of packages at module top level that are absent from this
image and irrelevant to the hot path.  They are replaced by inert
``sys.modules`` stubs:

* ``icecream.ic``                      (debug print; torch_transformation_utils.py:11)
* ``matplotlib`` / ``matplotlib.pyplot`` (plotting;    torch_transformation_utils.py:10)
* ``pyquaternion.Quaternion``          (transformation_utils.py:13)
* ``shapely.geometry.Polygon``         (utils/common_utils.py:12; a functional convex-quad stand-in, see _ConvexPolygon,
  so that the reference's own post_process / nms_rotated can run for the decode + NMS golden)
* ``opencood.utils.box_overlaps`` (Cython extension, label generation only) and ``opencood.visualization.vis_utils``
  (open3d / matplotlib) -- imported by voxel_postprocessor.py:20-21, never called by post_process
* ``timm.models.layers``               (cond_diff.py:20, DropPath & friends; unused on the eval path)
* ``efficientnet_pytorch.EfficientNet`` (lss_submodule.py:7, camera encoder; only needed to import heter_encoders.py
  for the full-detector golden, the LiDAR path never touches it); ``termcolor`` and ``spconv`` (sparse-conv classes
  of the SECOND encoder, sparse_backbone_3d.py:2-9) are stubbed for the same import and never called
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GENCOMM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "opencood"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return None

    def __getattr__(self, item):
        return _Anything()


class _ConvexPolygon:
    """Functional stand-in for ``shapely.geometry.Polygon`` restricted to what ``box_utils.nms_rotated`` uses
    (``common_utils.convert_format`` / ``compute_iou``: ``a.intersection(b).area``, ``a.union(b).area``) on convex quads.
    shapely (GEOS) is absent from this image, so the reference's NMS runs on the oracle's own polygon clipping
    (oracle/ref_ops.py::convex_intersection_area): everything in post_process EXCEPT the polygon area is reference code."""

    def __init__(self, pts=None, area=None):
        import numpy as np
        self.pts = None if pts is None else np.asarray([(float(x), float(y)) for x, y in pts], dtype=np.float64)
        self._area = area

    @property
    def area(self):
        from oracle import ref_ops
        return self._area if self.pts is None else abs(ref_ops._quad_area(self.pts))

    def intersection(self, other):
        from oracle import ref_ops
        return _ConvexPolygon(area=ref_ops.convex_intersection_area(self.pts, other.pts))

    def union(self, other):
        return _ConvexPolygon(area=self.area + other.area - self.intersection(other).area)


def install_stubs():
    import torch.nn as nn

    _stub("icecream", ic=lambda *a, **k: None)
    plt = _stub("matplotlib.pyplot")
    plt.__dict__.setdefault("figure", _Anything())
    mpl = _stub("matplotlib", pyplot=plt, use=lambda *a, **k: None)
    mpl.__dict__.setdefault("cm", _Anything())
    _stub("matplotlib.cm")
    _stub("matplotlib.colors")
    _stub("pyquaternion", Quaternion=_Anything)
    geo = _stub("shapely.geometry", Polygon=_ConvexPolygon, Point=_Anything, MultiPoint=_Anything)
    _stub("shapely", geometry=geo)

    class DropPath(nn.Module):  # identity at eval; never instantiated by GenComm's eval path
        def __init__(self, p=0.0):
            super().__init__()

        def forward(self, x):
            return x

    def to_2tuple(x):
        return (x, x) if not isinstance(x, (tuple, list)) else tuple(x)

    def trunc_normal_(t, std=0.02, **k):
        return nn.init.trunc_normal_(t, std=std)

    layers = _stub("timm.models.layers", DropPath=DropPath, to_2tuple=to_2tuple,
                   trunc_normal_=trunc_normal_, lecun_normal_=trunc_normal_, Mlp=_Anything,
                   PatchEmbed=_Anything, to_ntuple=lambda n: to_2tuple)
    models = _stub("timm.models", layers=layers)
    _stub("timm", models=models)
    _stub("termcolor", colored=lambda s, *a, **k: s)       # sparse_backbone_3d.py:2 (SECOND encoder, unused here)
    _stub("spconv", **{n: _Anything for n in ("SparseSequential", "SubMConv3d", "SparseConv3d", "SparseInverseConv3d",
                                              "SparseConvTensor")})   # sparse_backbone_3d.py:3-9, import-time only
    _stub("efficientnet_pytorch", EfficientNet=_Anything)
    _stub("opencood.utils.box_overlaps", bbox_overlaps=_Anything())
    _stub("opencood.visualization.vis_utils")   # camera encoder import of heter_encoders.py (lss_submodule.py:7)


def load():
    """Returns a namespace of the reference callables on the hot path."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    from opencood.models.sub_modules.pillar_vfe import PillarVFE
    from opencood.models.sub_modules.point_pillar_scatter import PointPillarScatter
    from opencood.models.sub_modules.torch_transformation_utils import warp_affine_simple
    from opencood.models.fuse_modules.fusion_in_one import MaxFusion, AttFusion, regroup
    from opencood.utils.transformation_utils import normalize_pairwise_tfm
    from opencood.models.gencomm_modules.unet import DiffusionUNet
    from opencood.models.gencomm_modules.cond_diff import GenComm, Config
    from opencood.models.gencomm_modules.message_extractor_v2 import MessageExtractorv2
    from opencood.models.gencomm_modules.enhancer import Enhancer
    from opencood.models.sub_modules.downsample_conv import DownsampleConv
    from opencood.models.sub_modules.base_bev_backbone import BaseBEVBackbone
    ns.PillarVFE = PillarVFE
    ns.PointPillarScatter = PointPillarScatter
    ns.warp_affine_simple = warp_affine_simple
    ns.MaxFusion = MaxFusion
    ns.AttFusion = AttFusion
    ns.regroup = regroup
    ns.normalize_pairwise_tfm = normalize_pairwise_tfm
    ns.DiffusionUNet = DiffusionUNet
    ns.GenComm = GenComm
    ns.Config = Config
    ns.MessageExtractorv2 = MessageExtractorv2
    ns.Enhancer = Enhancer
    ns.DownsampleConv = DownsampleConv
    ns.BaseBEVBackbone = BaseBEVBackbone
    return ns
