"""Configuration dataclasses - the field names are the launcher's config keys (config/base.py:17-89).

Differences from the reference, all forced by its as-shipped breakage:
  * ``image_shape`` / ``dim_in`` are filled in ``__post_init__`` from the NIfTI header instead of by a
    class-level ``nib.load`` at import time;
  * ``HashConfig`` ships 4-axis resolutions for the 4-D sample (the reference's 3-tuples were tuned for
    the 2-D+t slice and make ``_HashGridV2.forward`` fail on the bundled volume);
  * ``interp_shapes`` entries must have ``dim_in`` axes.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Any, Optional, Tuple

from . import nifti
from .datamodules import MriDataModule
from .models import HashMLP, SirenNet

_DEFAULT_IMAGE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data",
                              "sample_ankle_dyn_mri.nii.gz")


@dataclass
class BaseConfig:
    checkpoint_path: Optional[str] = None
    image_path: str = _DEFAULT_IMAGE
    image_shape: Tuple[int, ...] = ()
    batch_size: int = 4096
    epochs: int = 1
    num_workers: int = os.cpu_count()
    accumulate_grad_batches: Any = None
    # Network parameters
    dim_in: int = 0
    dim_hidden: int = 128
    dim_out: int = 1
    n_layers: int = 6
    n_sample: int = 3
    w0: float = 30.0
    w0_initial: float = 30.0
    use_bias: bool = True
    final_activation: Any = None
    lr: float = 1e-4
    datamodule: Any = MriDataModule
    model_cls: Any = SirenNet
    interp_shapes: Any = ()
    # filled by the launcher; kept for parity with the reference's superset of ctor kwargs
    encoder_type: str = "hash"
    n_levels: int = 16
    n_features_per_level: int = 2
    log2_hashmap_size: int = 19
    base_resolution: Any = 16
    finest_resolution: Any = 512
    per_level_scale: float = 1.2
    interpolation: str = "Linear"
    dropout: float = 0.0

    def __post_init__(self):
        if not self.image_shape and self.image_path and os.path.isfile(self.image_path):
            self.image_shape = nifti.load(self.image_path).shape
        if not self.dim_in:
            self.dim_in = len(self.image_shape)

    def export_to_txt(self, file_path: str = "") -> None:
        with open(file_path + "config.txt", "w") as f:
            for key in self.__dict__:
                f.write(str(key) + " : " + str(self.__dict__[key]) + "\n")


@dataclass
class HashConfig(BaseConfig):
    batch_size: int = 10000
    encoder_type: str = "hash"
    n_levels: int = 4
    n_features_per_level: int = 1
    log2_hashmap_size: int = 23
    base_resolution: Any = (64, 64, 5, 5)
    finest_resolution: Any = (352, 352, 6, 15)
    per_level_scale: float = 1.2
    interpolation: str = "Linear"
    dim_hidden: int = 64
    dim_out: int = 1
    n_layers: int = 2
    lr: float = 5e-3
    dropout: float = 0.0
    model_cls: Any = HashMLP
    interp_shapes: Any = ((352, 352, 6, 29),)


def g4_hash_kwargs() -> dict:
    """config/hash_config.json (L=16, F=2, log2T=19, base 16, per_level_scale 1.4) expressed through the
    python API: finest = round(16 * 1.4**15) = 2489 (valid because base-1 == L-1 == 15)."""
    return dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16, finest_resolution=2489)
