import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
SAMPLE = os.path.join(ROOT, "data", "sample_ankle_dyn_mri.nii.gz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


HASH_CASES = ["v1_d4", "v1_d3", "v1_d2_f4", "v1_d3_f1", "v2_d3", "v2_d4_f1"]
SIREN_CASES = ["small", "d4_h64", "d2_out3"]


def oracle_levels(fx):
    from oracle import hashgrid
    return [hashgrid.Level(tuple(int(v) for v in r), int(t)) for r, t in zip(fx["resolutions"], fx["rows"])]


def build_encoder(fx, device=None):
    """Product encoder configured like a golden hash fixture, tables loaded from the fixture."""
    from mri_interpolation_b200 import encoding
    dim, L, F = int(fx["dim"]), int(fx["n_levels"]), int(fx["n_features"])
    if bool(fx["anisotropic"]):
        enc = encoding.MultiResHashGridV2(dim, L, F, int(fx["log2_hashmap_size"]), tuple(int(v) for v in fx["base"]),
                                          tuple(int(v) for v in fx["finest"]))
    else:
        enc = encoding.MultiResHashGrid(dim, L, F, int(fx["log2_hashmap_size"]), int(fx["base"]), int(fx["finest"]))
    with torch.no_grad():
        for li, lv in enumerate(enc.levels):
            lv.embedding.weight.copy_(torch.from_numpy(fx[f"table{li}"]))
    return enc.to(device) if device is not None else enc
