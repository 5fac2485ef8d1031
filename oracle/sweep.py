"""Oracle: coordinate grids, the dense-grid query sweep, image metrics (CPU).

Restates:
  * utils.create_mgrid                         utils.py:14-23
  * MriImage coords / intensity normalisation  datamodules.py:130-166
  * MriDataModule.upsampling                   datamodules.py:229-252
  * launcher dense sweep                       launcher.py:191-222
  * PSNR (skimage.metrics.peak_signal_noise_ratio, commented call sites
    legacy_code/hash_experimentation.py:447-450); SSIM restated from the
    published skimage algorithm - skimage is absent: SSIM parity is UNPINNED.
Test infrastructure only - never imported by the product package.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np
import torch


def axis_values(n: int, norm_siren: bool = False) -> torch.Tensor:
    """torch.linspace(0,1,n) (or (-1,1,n) for norm_siren) - utils.py:19-20, datamodules.py:141-147."""
    return torch.linspace(-1, 1, n) if norm_siren else torch.linspace(0, 1, n)


def grid_coords(shape: Sequence[int], norm_siren: bool = False) -> torch.Tensor:
    """(prod(shape), D) f32 coordinates, 'ij' meshgrid, C-order flatten (last axis fastest)."""
    axes = [axis_values(s, norm_siren) for s in shape]
    return torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(-1, len(shape))


def normalise_intensities(volume: torch.Tensor, norm_siren: bool = False) -> torch.Tensor:
    """datamodules.py:150-160: min-max to [0,1] (or [-1,1]), flattened C-order, shape (M,1)."""
    p = volume.flatten()
    p = (p - torch.min(p)) / (torch.max(p) - torch.min(p))
    if norm_siren:
        p = p * 2 - 1
    return p.unsqueeze(-1)


def dense_sweep(model_fn: Callable[[torch.Tensor], torch.Tensor], shape: Sequence[int], batch_size: int,
                norm_siren: bool = False) -> torch.Tensor:
    """launcher.py:191-217: query the model on the dense grid in order, concat, reshape(shape)."""
    coords = grid_coords(shape, norm_siren)
    outs = [model_fn(coords[i : i + batch_size]) for i in range(0, coords.shape[0], batch_size)]
    return torch.cat(outs).reshape(tuple(shape))


def linear_time_interpolation(frames: np.ndarray, n_out: int) -> np.ndarray:
    """Linear-in-time baseline of interp.py:35-50 (ITK LinearInterpolateImageFunction on the last axis).

    ``frames``: (..., T_in); returns (..., n_out) sampled at continuous index t * (T_in-1)/(n_out-1).
    """
    t_in = frames.shape[-1]
    pos = np.linspace(0.0, t_in - 1, n_out)
    lo = np.clip(np.floor(pos).astype(np.int64), 0, t_in - 1)
    hi = np.clip(lo + 1, 0, t_in - 1)
    a = (pos - lo).astype(frames.dtype)
    return frames[..., lo] * (1 - a) + frames[..., hi] * a


def linear_time_baseline(data: np.ndarray) -> np.ndarray:
    """interp.py:35-50 as the reference runs it: keep frames ::2 and evaluate the ITK linear interpolator at the
    continuous index t/2 along the last axis for every output frame t (clamped at the last kept frame)."""
    values = data[..., ::2]
    t_in = values.shape[-1]
    pos = np.minimum(np.arange(data.shape[-1]) / 2.0, t_in - 1)
    lo = np.floor(pos).astype(np.int64)
    hi = np.minimum(lo + 1, t_in - 1)
    a = (pos - lo).astype(data.dtype)
    return values[..., lo] * (1 - a) + values[..., hi] * a


def psnr(truth: np.ndarray, test: np.ndarray, data_range: float = 1.0) -> float:
    """10 log10(data_range^2 / MSE), float64 accumulation (skimage's definition)."""
    err = np.mean((truth.astype(np.float64) - test.astype(np.float64)) ** 2)
    return float(10.0 * np.log10((data_range**2) / err))


def ssim2d(a: np.ndarray, b: np.ndarray, data_range: float = 1.0, win: int = 7) -> float:
    """Mean SSIM of two 2-D images, skimage defaults (uniform 7x7 window, K1=.01, K2=.03,
    sample covariance, borders cropped by (win-1)//2).  UNPINNED (skimage absent)."""
    from scipy.ndimage import uniform_filter

    a = a.astype(np.float64)
    b = b.astype(np.float64)
    npix = win * win
    cov_norm = npix / (npix - 1)
    ua, ub = uniform_filter(a, win), uniform_filter(b, win)
    uaa, ubb, uab = uniform_filter(a * a, win), uniform_filter(b * b, win), uniform_filter(a * b, win)
    va, vb, vab = cov_norm * (uaa - ua * ua), cov_norm * (ubb - ub * ub), cov_norm * (uab - ua * ub)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    s = ((2 * ua * ub + c1) * (2 * vab + c2)) / ((ua**2 + ub**2 + c1) * (va + vb + c2))
    pad = (win - 1) // 2
    return float(s[pad:-pad, pad:-pad].mean())


def ssim_slices(truth: np.ndarray, test: np.ndarray, data_range: float = 1.0) -> float:
    """Mean of 2-D SSIM over all leading-plane slices (axes 0,1 are the in-plane axes)."""
    t = truth.reshape(truth.shape[0], truth.shape[1], -1)
    p = test.reshape(test.shape[0], test.shape[1], -1)
    return float(np.mean([ssim2d(t[..., k], p[..., k], data_range) for k in range(t.shape[-1])]))
