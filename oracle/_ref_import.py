"""Import the UNMODIFIED reference modules from /root/reference with stub third-party packages.

Only usable in the build container (the GPU box has no /root/reference); used by
oracle/make_golden.py to pin the oracle and generate tests/golden/*.npz.
The reference needs pytorch_lightning, commentjson, rff and its own utils.py
(-> nibabel/torchio/matplotlib), none of which are installed; none of them takes
part in the arithmetic of encoding.py / models.py, so inert stand-ins suffice.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("MRI_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "encoding.py"))


class _StubLightningModule(torch.nn.Module):
    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def log(self, *a, **k):
        pass


def load():
    """Returns (encoding, models) reference modules."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if "_mri_ref_models" in sys.modules:
        return sys.modules["_mri_ref_encoding"], sys.modules["_mri_ref_models"]

    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = _StubLightningModule
    pl.LightningDataModule = object
    pl_util = types.ModuleType("pytorch_lightning.utilities")
    pl_types = types.ModuleType("pytorch_lightning.utilities.types")
    pl_types.STEP_OUTPUT = object
    cj = types.ModuleType("commentjson")
    rff = types.ModuleType("rff")
    rff.layers = types.SimpleNamespace(GaussianEncoding=None)
    utils = types.ModuleType("utils")
    utils.create_mgrid = lambda shape: None

    saved = {k: sys.modules.get(k) for k in
             ("pytorch_lightning", "pytorch_lightning.utilities", "pytorch_lightning.utilities.types",
              "commentjson", "rff", "utils", "encoding", "models")}
    sys.modules.update({"pytorch_lightning": pl, "pytorch_lightning.utilities": pl_util,
                        "pytorch_lightning.utilities.types": pl_types, "commentjson": cj, "rff": rff,
                        "utils": utils})
    sys.modules.pop("encoding", None)
    sys.modules.pop("models", None)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import encoding as ref_encoding  # noqa
        import models as ref_models  # noqa
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    sys.modules["_mri_ref_encoding"] = ref_encoding
    sys.modules["_mri_ref_models"] = ref_models
    return ref_encoding, ref_models
