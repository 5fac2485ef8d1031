"""Import the UNMODIFIED reference modules (encoding.py, models.py) from baseline/_ref/ for bench.py's reference arm.

baseline/_ref/ is git-ignored (reference sources never enter the history) but travels to the GPU box with the repo
snapshot; `python baseline/install_reference.py` (also run by __graft_entry__.build()) fills it from /root/reference.
The reference has no setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` cannot work -
the two source files of the hot path are copied verbatim instead.  Their third-party imports that are absent from the
image (pytorch_lightning, commentjson, rff, the reference's own utils.py -> nibabel/torchio/matplotlib) take no part in
the arithmetic of the path and are replaced by inert stand-ins for the duration of the import.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "encoding.py")) and os.path.isfile(os.path.join(REF_DIR, "models.py"))


class _LightningModuleStandIn(torch.nn.Module):
    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def log(self, *a, **k):
        pass


def load():
    """(encoding, models) modules of the reference, imported from baseline/_ref/."""
    if not available():
        raise RuntimeError(f"no reference copy under {REF_DIR}: run python baseline/install_reference.py where /root/reference exists")
    if "_mri_baseline_models" in sys.modules:
        return sys.modules["_mri_baseline_encoding"], sys.modules["_mri_baseline_models"]
    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = _LightningModuleStandIn
    pl.LightningDataModule = object
    pl_util = types.ModuleType("pytorch_lightning.utilities")
    pl_types = types.ModuleType("pytorch_lightning.utilities.types")
    pl_types.STEP_OUTPUT = object
    cj = types.ModuleType("commentjson")
    rff = types.ModuleType("rff")
    rff.layers = types.SimpleNamespace(GaussianEncoding=None)
    utils = types.ModuleType("utils")
    utils.create_mgrid = lambda shape: None
    names = ("pytorch_lightning", "pytorch_lightning.utilities", "pytorch_lightning.utilities.types", "commentjson", "rff",
             "utils", "encoding", "models")
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update({"pytorch_lightning": pl, "pytorch_lightning.utilities": pl_util,
                        "pytorch_lightning.utilities.types": pl_types, "commentjson": cj, "rff": rff, "utils": utils})
    sys.modules.pop("encoding", None)
    sys.modules.pop("models", None)
    sys.path.insert(0, REF_DIR)
    try:
        import encoding as ref_encoding  # noqa
        import models as ref_models  # noqa
    finally:
        sys.path.remove(REF_DIR)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    sys.modules["_mri_baseline_encoding"] = ref_encoding
    sys.modules["_mri_baseline_models"] = ref_models
    return ref_encoding, ref_models
