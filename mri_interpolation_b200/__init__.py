"""B200-native coordinate-network hot path of Benjamin-Fouquet/mri_interpolation.

Same nn.Module / LightningModule surface as the reference's ``encoding.py`` and ``models.py``,
backed by hand-written sm_100a CUDA kernels behind a plain C ABI (include/mri_b200.h,
mri_interpolation_b200/libmri_b200.so).  CUDA only - there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import MriB200Error  # noqa: F401

__all__ = ["encoding", "models", "datamodules", "sweep", "optim", "functional", "metrics", "nifti", "config",
           "MriB200Error"]
__version__ = "0.1.0"
