"""Image-quality metrics used to judge a fit (host side, numpy).

The reference imports ``skimage.metrics`` but only calls it from a commented block
(legacy_code/hash_experimentation.py:445-453: mean_squared_error, peak_signal_noise_ratio,
structural_similarity).  skimage is not available here; the definitions below follow its
published algorithms (data_range = 1 for [0,1]-normalised images, 7x7 uniform SSIM window,
K1 = 0.01, K2 = 0.03, sample covariance).  SSIM is applied slice-by-slice over the first two axes.
"""
from __future__ import annotations

import numpy as np


def mean_squared_error(a: np.ndarray, b: np.ndarray) -> float:
    return float(np.mean((np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)) ** 2))


def peak_signal_noise_ratio(truth: np.ndarray, test: np.ndarray, data_range: float = 1.0) -> float:
    err = mean_squared_error(truth, test)
    return float(10.0 * np.log10((data_range ** 2) / err))


def _ssim_plane(a: np.ndarray, b: np.ndarray, data_range: float, win: int) -> float:
    from scipy.ndimage import uniform_filter

    a = a.astype(np.float64)
    b = b.astype(np.float64)
    norm = (win * win) / (win * win - 1.0)
    ma, mb = uniform_filter(a, win), uniform_filter(b, win)
    va = norm * (uniform_filter(a * a, win) - ma * ma)
    vb = norm * (uniform_filter(b * b, win) - mb * mb)
    vab = norm * (uniform_filter(a * b, win) - ma * mb)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    s = ((2 * ma * mb + c1) * (2 * vab + c2)) / ((ma * ma + mb * mb + c1) * (va + vb + c2))
    p = (win - 1) // 2
    return float(s[p:-p, p:-p].mean())


def structural_similarity(truth: np.ndarray, test: np.ndarray, data_range: float = 1.0, win: int = 7) -> float:
    t = np.asarray(truth).reshape(truth.shape[0], truth.shape[1], -1)
    p = np.asarray(test).reshape(test.shape[0], test.shape[1], -1)
    return float(np.mean([_ssim_plane(t[..., k], p[..., k], data_range, win) for k in range(t.shape[-1])]))


def write_scores(path: str, truth: np.ndarray, test: np.ndarray, extra: dict = None) -> dict:
    """The scores.txt of the reference's commented block, made live."""
    scores = {"MSE": mean_squared_error(truth, test), "PSNR": peak_signal_noise_ratio(truth, test),
              "SSIM": structural_similarity(truth, test)}
    scores.update(extra or {})
    with open(path, "w") as f:
        for k, v in scores.items():
            f.write(f"{k} : {v}\n")
    return scores
