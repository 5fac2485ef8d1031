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


class SlabSweeper:
    """Everything a sweep of one rank's slab needs, prepared ONCE: per-axis coordinate vectors on the device, the decoder
    packed for the fused kernel (eval-mode BatchNorm folded into its Linear), the output slab and - for ``run_to_host`` -
    a pinned host slab plus a copy stream.  ``run()`` then launches nothing but the sweep kernels, so a sweep that needs
    no communication also has no per-call host work to lose scaling on (round 1: linspace / allocation / BN folding
    inside the timed window cost 12 % at 8 GPUs).

    ``run_to_host()`` cuts the slab into whole axis-0 planes and copies chunk i to pinned host memory on a side stream
    while chunk i+1 is computed, so the end-to-end sweep costs little more than the kernels.
    """

    def __init__(self, model, shape: Sequence[int], batch_size: int = 1 << 20, norm_siren: bool = False, rank: int = 0,
                 world_size: int = 1, fused: bool = True):
        self.model = model
        self.shape = [int(s) for s in shape]
        self.total = int(np.prod(self.shape))
        self.first, self.count = slab_range(self.total, rank, world_size)
        self.batch_size = int(batch_size)
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise _lib.MriB200Error("dense_sweep: the model must live on a CUDA device (no CPU fallback)")
        self.flat_axes = torch.cat(mgrid_axes(self.shape, norm_siren)).to(self.device)
        self.cshape = (ctypes.c_int32 * len(self.shape))(*self.shape)
        self.plan = None
        was_training = model.training
        model.eval()
        try:
            plan = _fused_plan(model) if fused else None
            if plan is not None and len(self.shape) == model.encoder.dim:
                self.plan = plan
                self.refresh()
        finally:
            model.train(was_training)
        self.dim_out = 1 if self.plan is not None else None
        self.out = None
        self._host = None
        self._copy_stream = None

    def refresh(self) -> None:
        """Re-pack the decoder after the parameters changed (the hash tables are read in place)."""
        if self.plan is None:
            return
        plan = self.plan
        l0 = plan["l0"]
        with torch.no_grad():
            (w0, b0), (w1, b1) = (fold_batchnorm(lin, bn) for lin, bn in plan["blocks"])
            self.packed = torch.cat([w0.reshape(-1), b0.reshape(-1), w1.reshape(-1), b1.reshape(-1)]).contiguous()
        self.dims = (ctypes.c_int32 * 3)(l0.in_features, l0.out_features, 1)

    def _fused_range(self, first: int, count: int, out: torch.Tensor) -> None:
        enc = self.model.encoder
        tables = enc.tables()
        enc._fwd_layout.refresh(tables, enc._resolutions, enc._rows)
        _lib.call("mri_hashmlp_sweep", self.flat_axes.data_ptr(), self.cshape, len(self.shape), first, count,
                  enc._fwd_layout.base, enc._fwd_layout.levels, enc.n_levels, enc.n_features_per_level,
                  self.packed.data_ptr(), self.dims, 2, self.plan["acts"][0], self.plan["acts"][1], out.data_ptr(), _lib.stream())

    def _model_range(self, first: int, count: int) -> torch.Tensor:
        coords = torch.empty((count, len(self.shape)), device=self.device, dtype=torch.float32)
        _lib.call("mri_grid_coords", self.flat_axes.data_ptr(), self.cshape, len(self.shape), first, count, coords.data_ptr(),
                  _lib.stream())
        return self.model(coords)

    def _chunks(self, target: int):
        """[first, first+count) cut into pieces of about `target` voxels on whole axis-0 planes (the fused kernel walks
        whole planes axis-0-fastest; ragged ends of the slab stay separate, short pieces)."""
        plane = self.total // self.shape[0]
        lo, hi = self.first, self.first + self.count
        p_begin, p_end = (lo + plane - 1) // plane, hi // plane
        if p_end <= p_begin:
            return [(lo, hi - lo)] if hi > lo else []
        planes_per = max(16, (target + plane - 1) // plane)
        cuts = [lo] if lo < p_begin * plane else []
        p = p_begin
        while p < p_end:
            cuts.append(p * plane)
            p = min(p + planes_per, p_end)
        cuts.append(p_end * plane)
        if hi > p_end * plane:
            cuts.append(hi)
        return [(a, b - a) for a, b in zip(cuts, cuts[1:]) if b > a]

    @torch.no_grad()
    def run(self) -> torch.Tensor:
        """Sweep this rank's slab; returns (count, dim_out) on the device (the same buffer on every call)."""
        was_training = self.model.training
        self.model.eval()
        try:
            if self.plan is not None:
                if self.out is None:
                    self.out = torch.empty((self.count, 1), device=self.device, dtype=torch.float32)
                if self.count:
                    self._fused_range(self.first, self.count, self.out)
                return self.out
            outs = [self._model_range(start, min(self.batch_size, self.first + self.count - start))
                    for start in range(self.first, self.first + self.count, self.batch_size)]
            return torch.cat(outs) if outs else torch.empty((0, 1), device=self.device)
        finally:
            self.model.train(was_training)

    @torch.no_grad()
    def run_to_host(self, n_chunks: int = 8) -> torch.Tensor:
        """Sweep the slab into PINNED host memory: (count, dim_out) CPU tensor (reused across calls), valid on return."""
        was_training = self.model.training
        self.model.eval()
        try:
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(self.device)
            main, side = torch.cuda.current_stream(self.device), self._copy_stream
            if self.plan is not None:
                pieces = self._chunks(max(1, self.count // max(1, n_chunks)))
            else:
                pieces = [(s, min(self.batch_size, self.first + self.count - s))
                          for s in range(self.first, self.first + self.count, self.batch_size)]
            done = None
            for first, count in pieces:
                if self.plan is not None:
                    if self.out is None:
                        self.out = torch.empty((self.count, 1), device=self.device, dtype=torch.float32)
                    dev_piece = self.out[first - self.first:first - self.first + count]
                    self._fused_range(first, count, dev_piece)
                else:
                    dev_piece = self._model_range(first, count)
                if self._host is None:
                    self._host = torch.empty((self.count, dev_piece.shape[1]), dtype=torch.float32).pin_memory()
                ready = torch.cuda.Event()
                ready.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(ready)
                    self._host[first - self.first:first - self.first + count].copy_(dev_piece, non_blocking=True)
                    dev_piece.record_stream(side)
                    done = torch.cuda.Event()
                    done.record(side)
            if done is not None:
                done.synchronize()
            if self._host is None:
                self._host = torch.empty((0, 1), dtype=torch.float32)
            return self._host
        finally:
            self.model.train(was_training)


@torch.no_grad()
def dense_sweep(model, shape: Sequence[int], batch_size: int = 1 << 20, norm_siren: bool = False, rank: int = 0,
                world_size: int = 1, fused: bool = True, out_host: bool = False) -> torch.Tensor:
    """Query `model` on this rank's slab of the dense grid `shape`; returns (count, dim_out) on the device, or - with
    ``out_host=True`` - in pinned host memory, the device-to-host copy pipelined with the sweep (`SlabSweeper`).

    Concatenating the slabs of ranks 0..W-1 and ``reshape(shape)`` gives the reference's
    ``torch.concat(trainer.predict(model, interp_loader)).reshape(shape)``.
    """
    sweeper = SlabSweeper(model, shape, batch_size=batch_size, norm_siren=norm_siren, rank=rank, world_size=world_size,
                          fused=fused)
    return sweeper.run_to_host() if out_host else sweeper.run()


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
