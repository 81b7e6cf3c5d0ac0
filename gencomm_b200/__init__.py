"""gencomm_b200 -- B200-native (sm_100a) implementation of GenComm's per-frame hot path.

    points -> voxelize -> PillarVFE -> PointPillarScatter -> BEV canvas
    BEV features -> warp to ego (affine_grid + grid_sample) + regroup + Max/Att fusion
    GenComm conditional-diffusion sampler (DiffusionUNet denoiser)

Hand-written CUDA behind a C ABI (include/gencomm_b200.h, csrc/), loaded through ctypes; the
classes in ``gencomm_b200.modules`` carry the reference's operator names and signatures.
"""
from . import _lib, ops, synth  # noqa: F401
from .modules import (AttFusion, MaxFusion, PFNLayer, PillarVFE, PointPillar, PointPillarScatter,  # noqa: F401
                      SpVoxelPreprocessor, normalize_pairwise_tfm, regroup, warp_affine_simple, warp_feature)

from .gencomm import Config, DiffusionUNet, GenComm  # noqa: F401
from .message_extractor import BEVDeformableExtractor, MessageExtractorv2  # noqa: F401
from .enhancer import Enhancer  # noqa: F401
from .det_tail import DetectionHeads, DoubleConv, DownsampleConv  # noqa: F401
from .backbone import BaseBEVBackbone  # noqa: F401

__version__ = "0.1.0"
