import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_artifacts():
    """Oracle C library (test infrastructure) and, if stale or missing, the CUDA library."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "libgc_oracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    from gencomm_b200 import build
    if os.path.exists("/usr/local/cuda/bin/nvcc"):
        build.build()


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def golden_pillars():
    return load_golden("pillars.npz")


@pytest.fixture(scope="session")
def golden_warp():
    return load_golden("warp_fuse.npz")


@pytest.fixture(scope="session")
def golden_gencomm():
    return load_golden("gencomm.npz")


@pytest.fixture(scope="session")
def golden_message_extractor():
    return load_golden("message_extractor.npz")


@pytest.fixture(scope="session")
def golden_enhancer():
    return load_golden("enhancer.npz")


@pytest.fixture(scope="session")
def golden_det_tail():
    return load_golden("det_tail.npz")


@pytest.fixture(scope="session")
def golden_backbone():
    return load_golden("backbone.npz")


BACKBONE_CFG = {"layer_nums": [1, 1, 2], "layer_strides": [2, 2, 2], "num_filters": [64, 64, 128],
                "upsample_strides": [1, 2, 4], "num_upsample_filter": [64, 64, 64]}


def backbone_input():
    """The input of tests/golden/backbone.npz (oracle/gen_golden.py::backbone_input), regenerated from its seed."""
    from gencomm_b200 import synth
    return synth.bev_features(1501, 1, 64, 64, 128, sparsity=0.7)


@pytest.fixture(scope="session")
def golden_heter_model():
    return load_golden("heter_model.npz")


@pytest.fixture(scope="session")
def heter_inputs():
    """(voxels, pairwise, record_len, noise0, step_noises) of tests/golden/heter_model.npz, regenerated from seeds."""
    from oracle import gen_golden
    return gen_golden.heter_inputs()


@pytest.fixture(scope="session")
def golden_postprocess():
    return load_golden("postprocess.npz")


@pytest.fixture(scope="session")
def golden_lss_pool():
    return load_golden("lss_pool.npz")


LSS_CASES = {"z1": (None, {}), "z2": ({"zbound": [-10, 10, 10.0], "xbound": [-20.0, 20.0, 0.8]}, {"B": 1, "N": 3})}


def lss_case(name):
    """(geom_feats, x, grid_conf) of tests/golden/lss_pool.npz case `name`, regenerated from its seed."""
    from gencomm_b200 import synth
    over, kw = LSS_CASES[name]
    conf = dict(synth.LSS_GRID_CONF, **(over or {}))
    geom, x = synth.lss_frustum(31, grid_conf=conf, **kw)
    return geom, x, conf


@pytest.fixture(scope="session")
def golden_heter_model_stage2():
    return load_golden("heter_model_stage2.npz")


@pytest.fixture(scope="session")
def heter2_inputs():
    """(voxels, bev_feature, pairwise, record_len, noise0, step_noises) of tests/golden/heter_model_stage2.npz."""
    from oracle import gen_golden
    return gen_golden.heter2_inputs()
