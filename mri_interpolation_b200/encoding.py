"""Multiresolution hash-grid encoding - same class surface as the reference's ``encoding.py``,
backed by the sm_100a kernels in csrc/hashgrid.cu through the C ABI.

Mirrors (reference file:line):
  PRIMES                      encoding.py:40
  fast_hash                   encoding.py:69-78   (torch utility kept for API parity; the kernels hash in-register)
  _HashGrid / _HashGridV2     encoding.py:81-128 / 194-270
  MultiResHashGrid / ...V2    encoding.py:131-191 / 273-336
  Frequency                   encoding.py:43-66   (no caller in the reference; kept importable)

Parameters are created by the same torch calls in the same order as the reference, so a given seed
yields a bit-identical ``state_dict`` (keys ``levels.{l}.embedding.weight``).  ``forward`` runs ONE
kernel over all levels; inputs must be CUDA fp32 (no CPU fallback).
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch
from torch import nn

from . import functional as Fn
from .pl_compat import pl

# --- constants ---
PRIMES = (1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737)
_KERNEL_DIMS = (2, 3, 4)


class Frequency(nn.Module):
    """NeRF positional encoding (encoding.py:43-66); plain torch, not on the hot path."""

    def __init__(self, dim: int, n_levels: int = 10):
        super().__init__()
        assert n_levels > 0
        self.n_levels = n_levels
        self.register_buffer("freqs", 2.0 ** torch.linspace(0.0, n_levels - 1, n_levels), persistent=False)
        self.input_dim = dim
        self.output_dim = dim * n_levels * 2

    def forward(self, x: torch.Tensor):
        x = x.unsqueeze(dim=-1) * self.freqs
        return torch.cat((torch.sin(x), torch.cos(x)), dim=-1).flatten(-2, -1)


@torch.no_grad()
def fast_hash(ind: torch.Tensor, primes: torch.Tensor, hashmap_size: int):
    """Spatial hash of integer corner indices (encoding.py:69-78): xor_d((ind_d * prime_d) mod 2^32) mod T."""
    d = ind.shape[-1]
    mixed = (ind * primes[:d]) & 0xFFFFFFFF
    acc = mixed[..., 0].clone()
    for i in range(1, d):
        acc ^= mixed[..., i]
    return acc % hashmap_size


def _corner_mask(dim: int) -> torch.Tensor:
    neigs = np.arange(1 << dim, dtype=np.int64).reshape((-1, 1))
    dims = np.arange(dim, dtype=np.int64).reshape((1, -1))
    return torch.tensor(neigs & (1 << dims) == 0, dtype=bool)


class _GridKernelMixin:
    """What HashGridFn needs from a grid object: dim, n_levels, n_features_per_level, per-level
    resolutions/rows and the cached table layouts."""

    def _init_kernel_state(self, dim, n_levels, n_features, resolutions, rows):
        if dim not in _KERNEL_DIMS:
            raise NotImplementedError(
                f"the B200 hash-grid kernels cover {list(_KERNEL_DIMS)}-D inputs (x,y[,z][,t]); got dim={dim}")
        object.__setattr__(self, "_resolutions", [tuple(float(r) for r in res) for res in resolutions])
        object.__setattr__(self, "_rows", [int(r) for r in rows])
        object.__setattr__(self, "_fwd_layout", Fn._TableLayout())
        object.__setattr__(self, "_bwd_layout", Fn._TableLayout())


class _SingleLevelView(_GridKernelMixin):
    def __init__(self, dim, n_features, resolution, rows):
        self.dim, self.n_levels, self.n_features_per_level = dim, 1, n_features
        self._init_kernel_state(dim, 1, n_features, [resolution], [rows])


class _HashGrid(nn.Module):
    """One resolution level (encoding.py:81-128)."""

    def __init__(self, dim: int, n_features: int, hashmap_size: int, resolution: float):
        super().__init__()
        self.dim = dim
        self.n_features = n_features
        self.hashmap_size = hashmap_size
        self.resolution = resolution
        assert self.dim <= len(PRIMES), f"HashGrid only supports < {len(PRIMES)}-D inputs"
        # look-up table: nn.Embedding's N(0,1) draw first, then U(-1e-4, 1e-4) (encoding.py:95-96)
        self.embedding = nn.Embedding(hashmap_size, n_features)
        nn.init.uniform_(self.embedding.weight, a=-0.0001, b=0.0001)
        self.register_buffer("primes", torch.tensor(PRIMES, dtype=torch.int64), persistent=False)
        self.register_buffer("bin_mask", _corner_mask(dim), persistent=False)

    def _axis_resolution(self):
        return (float(self.resolution),) * self.dim

    def forward(self, x: torch.Tensor):
        # x: (b..., dim) float32 in [0, 1]
        view = self.__dict__.get("_kernel_view")
        if view is None:
            view = _SingleLevelView(self.dim, self.n_features, self._axis_resolution(), self.hashmap_size)
            self.__dict__["_kernel_view"] = view
        return Fn.HashGridFn.apply(x, view, self.embedding.weight)


class _MultiResBase(_GridKernelMixin):
    def _finish(self, levels):
        self.levels = nn.ModuleList(levels)
        self.input_dim = self.dim
        self.output_dim = self.n_levels * self.n_features_per_level
        try:
            self._init_kernel_state(self.dim, self.n_levels, self.n_features_per_level,
                                    [lv._axis_resolution() for lv in levels], [lv.hashmap_size for lv in levels])
        except RuntimeError as e:
            # like the reference, a per-axis resolution that does not match `dim` only fails at forward time
            object.__setattr__(self, "_resolutions", None)
            object.__setattr__(self, "_deferred_error", e)

    def tables(self):
        return [lv.embedding.weight for lv in self.levels]

    def forward(self, x: torch.Tensor):
        """All levels in one launch; output (b..., n_levels * n_features_per_level) (encoding.py:190-191)."""
        if self._resolutions is None:
            raise RuntimeError(str(self._deferred_error))
        return Fn.HashGridFn.apply(x, self, *self.tables())

    @torch.no_grad()
    def corner_hashes(self, x: torch.Tensor):
        """(n, L, 2^D) hashes and weights straight from the CUDA kernel (parity probe)."""
        return Fn.hashgrid_corners(x, self)

    @torch.no_grad()
    def gathered_rows(self, x: torch.Tensor):
        """(encoding, (n, L, 2^D) table rows) from the production gather kernel itself (not the probe)."""
        return Fn.hashgrid_forward_rows(x, self)


class MultiResHashGrid(_MultiResBase, nn.Module):
    def __init__(
        self,
        dim: int,
        n_levels: int = 16,
        n_features_per_level: int = 2,
        log2_hashmap_size: int = 15,
        base_resolution: int = 16,
        finest_resolution: int = 512,
    ):
        """Instant-NGP style hash grid encoding (encoding.py:131-191).

        Output dimension is ``n_levels * n_features_per_level`` (``self.output_dim``).
        The growth factor divides by ``base_resolution - 1`` exactly like the reference (:168-171).
        """
        nn.Module.__init__(self)
        self.dim = dim
        self.n_levels = n_levels
        self.n_features_per_level = n_features_per_level
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.finest_resolution = finest_resolution

        b = math.exp((math.log(finest_resolution) - math.log(base_resolution)) / (base_resolution - 1))
        levels = []
        for level_idx in range(n_levels):
            resolution = math.floor(base_resolution * (b ** level_idx))
            hashmap_size = min(resolution ** dim, 2 ** log2_hashmap_size)
            levels.append(_HashGrid(dim=dim, n_features=n_features_per_level, hashmap_size=hashmap_size,
                                    resolution=resolution))
        self._finish(levels)


class _HashGridV2(pl.LightningModule):
    """One anisotropic level: per-axis resolution vector (encoding.py:194-270)."""

    def __init__(self, dim: int, n_features: int, hashmap_size: int, resolution: Sequence[float]):
        super().__init__()
        self.dim = dim
        self.n_features = n_features
        self.hashmap_size = hashmap_size
        self.resolution = torch.FloatTensor(resolution)
        assert self.dim <= len(PRIMES), f"HashGrid only supports < {len(PRIMES)}-D inputs"
        self.embedding = nn.Embedding(hashmap_size, n_features)
        nn.init.uniform_(self.embedding.weight, a=-0.0001, b=0.0001)
        self.register_buffer("primes", torch.tensor(PRIMES, dtype=torch.int64), persistent=False)
        self.register_buffer("bin_mask", _corner_mask(dim), persistent=False)

    def _axis_resolution(self):
        res = tuple(float(r) for r in self.resolution.tolist())
        if len(res) != self.dim:
            # the reference fails with a broadcast RuntimeError at forward time (encoding.py:245)
            raise RuntimeError(f"resolution {res} has {len(res)} axes but dim={self.dim}")
        return res

    def forward(self, x: torch.Tensor):
        view = self.__dict__.get("_kernel_view")
        if view is None:
            view = _SingleLevelView(self.dim, self.n_features, self._axis_resolution(), self.hashmap_size)
            self.__dict__["_kernel_view"] = view
        return Fn.HashGridFn.apply(x, view, self.embedding.weight)


class MultiResHashGridV2(_MultiResBase, pl.LightningModule):
    def __init__(
        self,
        dim: int,
        n_levels: int = 16,
        n_features_per_level: int = 2,
        log2_hashmap_size: int = 15,
        base_resolution: Sequence[int] = 16,
        finest_resolution: Sequence[int] = 512,
    ):
        """Anisotropic hash grid (encoding.py:273-336): base/finest are per-axis sequences, one growth
        factor per axis, table rows = min(max(res)**dim, 2**log2_hashmap_size)."""
        pl.LightningModule.__init__(self)
        self.dim = dim
        self.n_levels = n_levels
        self.n_features_per_level = n_features_per_level
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.finest_resolution = finest_resolution

        b_list = [math.exp((math.log(fr) - math.log(br)) / (br - 1)) for br, fr in zip(base_resolution, finest_resolution)]
        levels = []
        for level_idx in range(n_levels):
            resolution = [math.floor(br * (b ** level_idx)) for b, br in zip(b_list, base_resolution)]
            hashmap_size = min(max(resolution) ** dim, 2 ** log2_hashmap_size)
            levels.append(_HashGridV2(dim=dim, n_features=n_features_per_level, hashmap_size=hashmap_size,
                                      resolution=resolution))
        self._finish(levels)
