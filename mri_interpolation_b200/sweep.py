"""Dense-grid query sweep (K7): reconstruct an up-sampled volume from a fitted model.

Replaces launcher.py:191-222 / datamodules.py:229-252 / notebook cells 26, 51, 60: the reference builds
the full (prod(shape), D) coordinate tensor on the host, wraps it in a DataLoader and concatenates
``trainer.predict`` outputs.  Here the coordinates are synthesised on the device from the flat voxel
index (bit-identical floats: the per-axis vectors come from ``torch.linspace`` itself), the query
volume is cut into contiguous C-order slabs, one per data-parallel rank, and no rank communicates.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
from torch import nn

from . import _lib
from . import functional as Fn
from .datamodules import mgrid_axes
from .models import HashMLP, _fusable_activation


def slab_range(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous C-order slab [first, first+count) of rank `rank` out of `world_size`."""
    first = (total * rank) // world_size
    last = (total * (rank + 1)) // world_size
    return first, last - first


def fold_batchnorm(lin: nn.Linear, bn: Optional[nn.BatchNorm1d]) -> Tuple[torch.Tensor, torch.Tensor]:
    """(weight, bias) of the Linear with an eval-mode BatchNorm1d folded in: BN(z) = (z - mean) / sqrt(var + eps) * gamma
    + beta is affine once the running statistics are frozen, so Linear -> BN is one Linear."""
    w, b = lin.weight, lin.bias
    if bn is None:
        return w, b
    scale = torch.rsqrt(bn.running_var + bn.eps)
    if bn.weight is not None:
        scale = scale * bn.weight
    shift = bn.bias if bn.bias is not None else torch.zeros_like(scale)
    return w * scale[:, None], (b - bn.running_mean) * scale + shift


def _fused_plan(model) -> Optional[dict]:
    """Describe the model for mri_hashmlp_sweep if it fits the fused kernel, else None.

    Decoder blocks may be Linear -> act [-> Dropout] (notebook variant) or the shipped Linear -> BatchNorm1d -> act ->
    Dropout (models.py:718-739); the latter only in eval mode with running statistics, where BN folds into the Linear."""
    if not isinstance(model, HashMLP) or len(model.decoder) != 2:
        return None
    enc = model.encoder
    if getattr(enc, "_resolutions", None) is None or enc.n_features_per_level not in (1, 2, 4):
        return None
    acts, lins = [], []
    for blk in model.decoder:
        mods = list(blk)
        if not isinstance(mods[0], nn.Linear) or mods[0].bias is None or len(mods) < 2:
            return None
        bn, rest = None, mods[1:]
        if isinstance(rest[0], nn.BatchNorm1d):
            bn, rest = rest[0], rest[1:]
            if model.training or bn.training or bn.running_mean is None or not rest:
                return None
        a = _fusable_activation(rest[0])
        if a is None or any(not isinstance(m, nn.Dropout) or (m.training and m.p > 0.0) for m in rest[1:]):
            return None
        acts.append(a)
        lins.append((mods[0], bn))
    (l0, _), (l1, _) = lins
    if l1.out_features != 1 or l0.out_features not in (16, 32, 64, 128):
        return None
    if acts[0] not in (_lib.ACT_GELU, _lib.ACT_RELU):
        return None  # the fused kernel is compiled for GELU / ReLU hidden activations
    return dict(acts=acts, l0=l0, l1=l1, blocks=lins)


@torch.no_grad()
def dense_sweep(model, shape: Sequence[int], batch_size: int = 1 << 20, norm_siren: bool = False, rank: int = 0,
                world_size: int = 1, fused: bool = True) -> torch.Tensor:
    """Query `model` on this rank's slab of the dense grid `shape`; returns (count, dim_out) on the device.

    Concatenating the slabs of ranks 0..W-1 and ``reshape(shape)`` gives the reference's
    ``torch.concat(trainer.predict(model, interp_loader)).reshape(shape)``.
    """
    shape = [int(s) for s in shape]
    total = int(np.prod(shape))
    first, count = slab_range(total, rank, world_size)
    device = next(model.parameters()).device
    if device.type != "cuda":
        raise _lib.MriB200Error("dense_sweep: the model must live on a CUDA device (no CPU fallback)")
    axes = mgrid_axes(shape, norm_siren)
    flat_axes = torch.cat(axes).to(device)
    cshape = (ctypes.c_int32 * len(shape))(*shape)
    was_training = model.training
    model.eval()
    try:
        plan = _fused_plan(model) if fused else None
        if plan is not None and len(shape) == model.encoder.dim:
            enc = model.encoder
            tables = enc.tables()
            enc._fwd_layout.refresh(tables, enc._resolutions, enc._rows)
            l0, l1 = plan["l0"], plan["l1"]
            (w0, b0), (w1, b1) = (fold_batchnorm(lin, bn) for lin, bn in plan["blocks"])
            packed = torch.cat([w0.reshape(-1), b0.reshape(-1), w1.reshape(-1), b1.reshape(-1)]).contiguous()
            dims = (ctypes.c_int32 * 3)(l0.in_features, l0.out_features, 1)
            out = torch.empty((count, 1), device=device, dtype=torch.float32)
            _lib.call("mri_hashmlp_sweep", flat_axes.data_ptr(), cshape, len(shape), first, count,
                      enc._fwd_layout.base, enc._fwd_layout.levels, enc.n_levels, enc.n_features_per_level,
                      packed.data_ptr(), dims, 2, plan["acts"][0], plan["acts"][1], out.data_ptr(), _lib.stream())
            return out
        outs = []
        for start in range(first, first + count, batch_size):
            n = min(batch_size, first + count - start)
            coords = torch.empty((n, len(shape)), device=device, dtype=torch.float32)
            _lib.call("mri_grid_coords", flat_axes.data_ptr(), cshape, len(shape), start, n, coords.data_ptr(),
                      _lib.stream())
            outs.append(model(coords))
        return torch.cat(outs) if outs else torch.empty((0, 1), device=device)
    finally:
        model.train(was_training)


def gather_slabs(local: torch.Tensor, shape: Sequence[int]) -> Optional[np.ndarray]:
    """Collect every rank's slab on rank 0 (host side, only to write the NIfTI) and reshape to `shape`."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local.reshape(tuple(shape) + tuple(local.shape[1:]) if local.shape[1] != 1 else tuple(shape)).cpu().numpy()
    world, rank = dist.get_world_size(), dist.get_rank()
    total = int(np.prod(shape))
    sizes = [slab_range(total, r, world)[1] for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros((pad, local.shape[1]), device=local.device, dtype=local.dtype)
    buf[: local.shape[0]] = local
    gathered = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0)
    if rank != 0:
        return None
    full = torch.cat([g[:s] for g, s in zip(gathered, sizes)])
    return full.reshape(tuple(shape) if full.shape[1] == 1 else tuple(shape) + (full.shape[1],)).cpu().numpy()
