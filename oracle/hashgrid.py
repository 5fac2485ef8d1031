"""Oracle: multiresolution hash-grid encoding (CPU, torch fp32 + int64).

Restates reference ``encoding.py``:
  * PRIMES                         encoding.py:40
  * fast_hash                      encoding.py:69-78
  * _HashGrid / _HashGridV2        encoding.py:81-128 / 194-270
  * MultiResHashGrid / ...V2       encoding.py:131-191 / 273-336

The arithmetic is kept in the reference's order so that, on CPU, the results are
bit-identical to the reference (checked by oracle/make_golden.py):
  xs = x * res ; xi = trunc(xs) ; xf = xs - float(xi)
  corner n, axis d: bit d of n clear -> (xi_d, 1-xf_d) else (xi_d+1, xf_d)
  w_n = prod_d ws[n, d] ;  h_n = (xor_d ((ind[n,d]*PRIME[d]) & 0xFFFFFFFF)) % T
  out = sum_n table[h_n] * w_n
Test infrastructure only - never imported by the product package.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Sequence, Tuple, Union

import torch

# encoding.py:40
HASH_PRIMES = (1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737)


@dataclass(frozen=True)
class Level:
    """One resolution level: per-axis resolution and number of table rows."""

    resolution: Tuple[int, ...]
    rows: int


def geometry_isotropic(
    dim: int,
    n_levels: int = 16,
    log2_hashmap_size: int = 15,
    base_resolution: int = 16,
    finest_resolution: int = 512,
) -> List[Level]:
    """Level geometry of ``MultiResHashGrid`` (encoding.py:168-176).

    Growth factor uses the divisor ``base_resolution - 1`` (reference quirk),
    res_l = floor(base * b**l), rows_l = min(res_l**dim, 2**log2T).
    """
    growth = math.exp(
        (math.log(finest_resolution) - math.log(base_resolution)) / (base_resolution - 1)
    )
    out = []
    for lvl in range(n_levels):
        res = math.floor(base_resolution * (growth**lvl))
        out.append(Level((res,) * dim, min(res**dim, 2**log2_hashmap_size)))
    return out


def geometry_anisotropic(
    dim: int,
    n_levels: int,
    log2_hashmap_size: int,
    base_resolution: Sequence[int],
    finest_resolution: Sequence[int],
) -> List[Level]:
    """Level geometry of ``MultiResHashGridV2`` (encoding.py:310-321).

    One growth factor per axis; rows_l = min(max(res_l)**dim, 2**log2T).
    ``zip`` truncation of the reference is preserved (len(res) may be < dim,
    which the reference then fails on at forward time).
    """
    growth = [
        math.exp((math.log(fr) - math.log(br)) / (br - 1))
        for br, fr in zip(base_resolution, finest_resolution)
    ]
    out = []
    for lvl in range(n_levels):
        res = tuple(math.floor(br * (g**lvl)) for g, br in zip(growth, base_resolution))
        out.append(Level(res, min(max(res) ** dim, 2**log2_hashmap_size)))
    return out


def geometry(dim, n_levels, log2_hashmap_size, base_resolution, finest_resolution) -> List[Level]:
    """Dispatch the way HashMLP does (models.py:691-708): int -> V1 else V2."""
    if isinstance(base_resolution, int):
        return geometry_isotropic(dim, n_levels, log2_hashmap_size, base_resolution, finest_resolution)
    return geometry_anisotropic(dim, n_levels, log2_hashmap_size, base_resolution, finest_resolution)


def lower_corner_mask(dim: int) -> torch.Tensor:
    """(2**dim, dim) bool, True where corner n takes the LOWER cell index on axis d.

    encoding.py:101-106: mask[n, d] = ((n >> d) & 1) == 0.
    """
    n = torch.arange(1 << dim, dtype=torch.int64).reshape(-1, 1)
    d = torch.arange(dim, dtype=torch.int64).reshape(1, -1)
    return ((n >> d) & 1) == 0


def spatial_hash(corner_index: torch.Tensor, rows: int) -> torch.Tensor:
    """encoding.py:69-78 - int64 emulation of a uint32 multiply/xor hash, then % rows."""
    dim = corner_index.shape[-1]
    primes = torch.tensor(HASH_PRIMES[:dim], dtype=torch.int64, device=corner_index.device)
    mixed = (corner_index * primes) & 0xFFFFFFFF
    acc = mixed[..., 0].clone()
    for axis in range(1, dim):
        acc ^= mixed[..., axis]
    return acc % rows


def _scale(x: torch.Tensor, resolution: Sequence[int], anisotropic: bool) -> torch.Tensor:
    # V1 multiplies by a python int (encoding.py:111), V2 by an f32 vector (encoding.py:205,245).
    if anisotropic:
        return x * torch.tensor([float(r) for r in resolution], dtype=torch.float32, device=x.device)
    return x * resolution[0]


def corners(x: torch.Tensor, level: Level, anisotropic: bool = False):
    """Corner hashes (N.., C) int64 and D-linear weights (N.., C) f32 of one level.

    encoding.py:111-126.  No clamping: x == 1.0 gives cell index ``res`` and the
    upper corner ``res + 1``; both are hashed like any other.
    """
    dim = x.shape[-1]
    xs = _scale(x, level.resolution, anisotropic)
    cell = xs.long()
    frac = xs - cell.float()
    mask = lower_corner_mask(dim).to(x.device).reshape((1,) * (x.dim() - 1) + (1 << dim, dim))
    cell = cell.unsqueeze(-2)
    frac = frac.unsqueeze(-2)
    index = torch.where(mask, cell, cell + 1)
    per_axis = torch.where(mask, 1 - frac, frac)
    weight = per_axis.prod(dim=-1)
    return spatial_hash(index, level.rows), weight


def encode_level(x: torch.Tensor, table: torch.Tensor, level: Level, anisotropic: bool = False) -> torch.Tensor:
    """One level's (N.., F) features (encoding.py:108-128)."""
    h, w = corners(x, level, anisotropic)
    return torch.sum(torch.nn.functional.embedding(h, table) * w.unsqueeze(-1), dim=-2)


def encode(x: torch.Tensor, tables: Sequence[torch.Tensor], levels: Sequence[Level], anisotropic: bool = False) -> torch.Tensor:
    """All levels concatenated on the last axis (encoding.py:190-191)."""
    return torch.cat([encode_level(x, t, l, anisotropic) for t, l in zip(tables, levels)], dim=-1)


def table_gradients(x, grad_out, levels, n_features: int, anisotropic: bool = False):
    """Dense per-level table gradients: dTable[h_n] += w_n * dOut (autograd of encoding.py:127-128).

    Written as an explicit scatter in float64-free fp32 ``index_add_`` so it can be
    compared with autograd of :func:`encode` (tests do both).
    """
    grads = []
    flat_x = x.reshape(-1, x.shape[-1])
    flat_g = grad_out.reshape(flat_x.shape[0], -1)
    for li, level in enumerate(levels):
        h, w = corners(flat_x, level, anisotropic)
        g = flat_g[:, li * n_features : (li + 1) * n_features]
        contrib = (w.unsqueeze(-1) * g.unsqueeze(1)).reshape(-1, n_features)
        dt = torch.zeros(level.rows, n_features, dtype=torch.float32)
        dt.index_add_(0, h.reshape(-1), contrib)
        grads.append(dt)
    return grads


def init_tables(levels: Sequence[Level], n_features: int) -> List[torch.Tensor]:
    """Table init with the reference's RNG consumption (encoding.py:95-96).

    ``nn.Embedding`` draws N(0,1) first, then ``uniform_(-1e-4, 1e-4)`` overwrites.
    """
    out = []
    for level in levels:
        t = torch.empty(level.rows, n_features)
        torch.nn.init.normal_(t)
        torch.nn.init.uniform_(t, a=-0.0001, b=0.0001)
        out.append(t)
    return out
