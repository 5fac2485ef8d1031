"""Image-quality metrics and the linear-in-time baseline, computed on the GPU (csrc/metrics.cu).

The reference imports ``skimage.metrics`` but only calls it from a commented block
(legacy_code/hash_experimentation.py:445-453: mean_squared_error, peak_signal_noise_ratio,
structural_similarity); ``interp.py:35-52`` is the linear-in-time baseline the networks are compared with.
Definitions follow skimage's published algorithms (data_range = 1 for [0,1]-normalised images, 7x7 uniform SSIM
window, K1 = 0.01, K2 = 0.03, sample covariance, borders cropped); SSIM is applied slice by slice over the first
two axes.  Inputs may be numpy arrays or tensors on any device; they are moved to the CUDA device and reduced there
in float64.  There is no host implementation in the product: without a CUDA device the calls raise (the numpy
restatement lives in oracle/sweep.py and is what the tests compare against).
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import MriB200Error


def _device(*xs) -> torch.device:
    for x in xs:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    if not torch.cuda.is_available():
        raise MriB200Error("metrics run on the GPU (csrc/metrics.cu); no CUDA device is available and there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _f32(x, dev: torch.device) -> torch.Tensor:
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _pair(truth, test):
    dev = _device(truth, test)
    a, b = _f32(truth, dev), _f32(test, dev)
    if a.shape != b.shape:
        raise MriB200Error(f"metrics: shapes differ, {tuple(a.shape)} vs {tuple(b.shape)}")
    return a, b


def mean_squared_error(a, b) -> float:
    a, b = _pair(a, b)
    acc = torch.zeros(1, dtype=torch.float64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.call("mri_sq_err_sum", a.data_ptr(), b.data_ptr(), a.numel(), acc.data_ptr(), _lib.stream())
    return float(acc) / max(a.numel(), 1)


def peak_signal_noise_ratio(truth, test, data_range: float = 1.0) -> float:
    err = mean_squared_error(truth, test)
    return float("inf") if err == 0.0 else float(10.0 * math.log10((data_range ** 2) / err))


def structural_similarity(truth, test, data_range: float = 1.0, win: int = 7) -> float:
    a, b = _pair(truth, test)
    if a.dim() < 2:
        raise MriB200Error("structural_similarity needs at least two axes")
    nx, ny = int(a.shape[0]), int(a.shape[1])
    planes = a.numel() // (nx * ny)
    acc = torch.zeros(1, dtype=torch.float64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.call("mri_ssim_sum", a.data_ptr(), b.data_ptr(), nx, ny, planes, int(win), float(data_range), acc.data_ptr(),
                  _lib.stream())
    return float(acc) / ((nx - win + 1) * (ny - win + 1) * planes)


def linear_time_baseline(data, device: Optional[torch.device] = None) -> torch.Tensor:
    """interp.py:35-50: keep frames ::2, linear interpolation at continuous index t/2 along the last axis (clamped at
    the last kept frame).  Returns a CUDA tensor of data's shape."""
    dev = device if device is not None else _device(data)
    d = _f32(data, torch.device(dev))
    out = torch.empty_like(d)
    t = int(d.shape[-1])
    with torch.cuda.device(d.device):
        _lib.call("mri_linear_time_interp", d.data_ptr(), d.numel() // t, t, out.data_ptr(), _lib.stream())
    return out


def write_scores(path: str, truth, test, extra: dict = None) -> dict:
    """The scores.txt of the reference's commented block, made live."""
    scores = {"MSE": mean_squared_error(truth, test), "PSNR": peak_signal_noise_ratio(truth, test),
              "SSIM": structural_similarity(truth, test)}
    scores.update(extra or {})
    with open(path, "w") as f:
        for k, v in scores.items():
            f.write(f"{k} : {v}\n")
    return scores
