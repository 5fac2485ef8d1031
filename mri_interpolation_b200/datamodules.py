"""Input pipeline - the reference's ``datamodules.py`` surface (MriImage / MriDataModule) with the
data resident on the device.

Mirrors: MriImage datamodules.py:123-172 (coords = linspace per axis, 'ij' meshgrid, C-order flatten;
intensities min-max normalised, optional [-1,1] ``norm_siren``), MriDataModule :175-252 incl.
``upsampling()``.  The reference feeds training through a per-sample ``__getitem__`` + collate
DataLoader with os.cpu_count() workers - its real end-to-end bottleneck.  Here the
loaders are ``DeviceBatchLoader``s: coords/pixels live in HBM once, a shuffled epoch is a device-side
permutation, a batch is one gather - same batches-per-epoch semantics (last batch may be short).
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import numpy as np
import torch
from torch.utils.data import Dataset

from . import nifti
from .pl_compat import pl


def mgrid_axes(shape: Sequence[int], norm_siren: bool = False):
    lo = -1 if norm_siren else 0
    return [torch.linspace(lo, 1, s) for s in shape]


def create_mgrid(shape: Sequence[int]) -> torch.Tensor:
    """utils.py:14-23 - (*shape, D) grid of linspace(0,1,s) coordinates."""
    return torch.stack(torch.meshgrid(*mgrid_axes(shape), indexing="ij"), dim=-1)


class ShuffledEpochs:
    """Index stream of shuffled epochs over ``n`` samples: ``epoch()`` returns this rank's sample indices of one epoch,
    batch after batch.  A batch is the set of samples at positions [i*B, (i+1)*B) of a fresh device-side permutation,
    exactly as with the reference's shuffled DataLoader.

    ``rank``/``world_size`` give each data-parallel rank a disjoint strided share of every epoch; the permutation is
    padded (wrapping around, like ``DistributedSampler``) to a multiple of ``world_size`` so that every rank sees the
    SAME number of batches of the SAME sizes - the optimiser step is collective, a rank with one batch more would hang.

    ``grid_shape`` (the C-order voxel grid the samples enumerate) switches on *locality-ordered batches*: inside a batch
    the samples are arranged with the axis-0 index running fastest (functional.locality_key).  The batch SET is
    unchanged and the loss is a mean over the batch, so training is the same up to fp32 summation order, but
    neighbouring rows of a batch now hit neighbouring hash-table rows: -32 % L2 sectors in the gather, and the scatter
    can merge duplicate updates (csrc/hash_device.cuh).  It costs one stable sort of small keys per epoch: the voxels
    are kept in locality order once, an epoch's permutation is turned into (batch id per voxel) and stably sorted by
    that id (measured on B200, 11.15 M voxels: 2.2 ms per epoch against 1.1 ms for torch.randperm alone).
    """

    def __init__(self, n: int, batch_size: int, device, seed: int = 1337, rank: int = 0, world_size: int = 1,
                 grid_shape: Optional[Sequence[int]] = None, locality: bool = True):
        self.n, self.batch_size = int(n), int(batch_size)
        self.device = torch.device(device)
        self.rank, self.world_size = rank, world_size
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(seed)
        self._loc_order = None
        if grid_shape is not None and locality:
            if int(np.prod(grid_shape)) != self.n:
                raise ValueError(f"grid_shape {tuple(grid_shape)} does not enumerate {self.n} samples")
            from .functional import locality_key
            flat = torch.arange(self.n, device=self.device)
            self._loc_order = torch.argsort(locality_key(flat, grid_shape, block=1))

    def local_count(self) -> int:
        return (self.n + self.world_size - 1) // self.world_size

    def batches_per_epoch(self, drop_last: bool = False) -> int:
        m = self.local_count()
        return m // self.batch_size if drop_last else (m + self.batch_size - 1) // self.batch_size

    def epoch(self) -> torch.Tensor:
        n, w, b = self.n, self.world_size, self.batch_size
        perm = torch.randperm(n, device=self.device, generator=self._gen)  # perm[p] = voxel at position p
        total = self.local_count() * w
        if total > n:
            perm = torch.cat([perm, perm[: total - n]])
        if self._loc_order is None:
            return perm[self.rank::w] if w > 1 else perm
        # position of every voxel (a wrapped-around voxel has two; both are kept)
        pos = torch.empty(n, dtype=torch.int64, device=self.device)
        pos[perm[:n]] = torch.arange(n, device=self.device)
        lo = self._loc_order
        vox, p = lo, pos[lo]
        if total > n:
            # the few wrapped-around voxels appear a second time, at positions n, n+1, ...: merge them in locality order
            extra = perm[n:]
            rank_of = torch.empty(n, dtype=torch.int64, device=self.device)
            rank_of[lo] = torch.arange(n, device=self.device)
            vox = torch.cat([lo, extra])
            p = torch.cat([p, torch.arange(n, total, device=self.device)])
            merged = torch.argsort(rank_of[vox], stable=True)
            vox, p = vox[merged], p[merged]
        if w > 1:
            mine = (p % w) == self.rank
            vox, p = vox[mine], p[mine] // w
        batch_id = p // b
        n_batches = (self.local_count() + b - 1) // b
        batch_id = batch_id.to(torch.uint8 if n_batches <= 255 else torch.int16 if n_batches <= 32767 else torch.int32)
        _, order = torch.sort(batch_id, stable=True)   # few key bits; voxels stay in locality order inside a batch
        return vox[order]


class DeviceBatchLoader:
    """Iterable of (coords, pixels) batches cut from tensors that already live on `device`.

    ``shuffle=True`` draws a fresh permutation per epoch (`ShuffledEpochs`: same batch sets as the reference's shuffled
    DataLoader, equal batch counts and sizes on every data-parallel rank); with ``grid_shape`` - the C-order voxel grid
    the rows of ``coords`` enumerate - the samples of a batch are arranged in locality order (axis-0 index fastest).
    """

    def __init__(self, coords: torch.Tensor, pixels: torch.Tensor, batch_size: int, shuffle: bool = False,
                 device: Optional[torch.device] = None, seed: int = 1337, rank: int = 0, world_size: int = 1,
                 drop_last: bool = False, grid_shape: Optional[Sequence[int]] = None, locality: bool = True,
                 grid_norm_siren: Optional[bool] = None):
        """``grid_norm_siren`` (not None) declares that ``coords`` IS the dense ``mgrid_axes(grid_shape, grid_norm_siren)``
        mesh in C order (what MriImage builds): shuffled CUDA batches are then produced by ``mri_gather_voxels`` - the
        coordinates rebuilt from the voxel index, bit-identical - in 0.015 ms instead of the 0.27 ms
        ``coords.index_select`` takes on the 11 M x 4 sample array (scripts/loader_cost.py)."""
        self.device = torch.device(device) if device is not None else coords.device
        self.coords = coords.to(self.device)
        self.pixels = pixels.to(self.device)
        self.batch_size = int(batch_size)
        self.shuffle = shuffle
        self.rank, self.world_size = rank, world_size
        self.drop_last = drop_last
        self.dataset = torch.utils.data.TensorDataset(self.coords, self.pixels)
        self.epoch = 0
        self.epochs = ShuffledEpochs(self.coords.shape[0], self.batch_size, self.device, seed, rank, world_size,
                                     grid_shape if shuffle else None, locality)
        self._voxels = None
        if (shuffle and grid_shape is not None and grid_norm_siren is not None and self.device.type == "cuda"
                and self.pixels.dtype == torch.float32 and self.pixels.numel() == self.coords.shape[0]
                and self.coords.shape[1] == len(grid_shape) and int(np.prod(grid_shape)) == self.coords.shape[0]):
            from .functional import VoxelSampler
            self._voxels = VoxelSampler(self.pixels.reshape(-1), grid_shape, norm_siren=bool(grid_norm_siren))

    def _draw_epoch(self) -> torch.Tensor:
        # drawn on the training stream: prefetching the next epoch on a side stream was measured SLOWER (Trainer.fit 371 vs
        # 454 M coords/s) - the sort / permutation kernels take SMs away from the one-wave persistent training kernels,
        # which then need a second wave
        return self.epochs.epoch()

    def __len__(self) -> int:
        return self.epochs.batches_per_epoch(self.drop_last)

    def epoch_indices(self) -> torch.Tensor:
        """This rank's sample indices of one shuffled epoch, batch after batch (advances the generator)."""
        return self._draw_epoch()

    def __iter__(self):
        if self.shuffle:
            order = self._draw_epoch()
            for i in range(len(self)):
                idx = order[i * self.batch_size:(i + 1) * self.batch_size]
                if self._voxels is not None:
                    yield self._voxels.batch(idx)
                else:
                    yield self.coords.index_select(0, idx), self.pixels.index_select(0, idx)
        else:
            if self.world_size > 1:
                raise RuntimeError("unshuffled loaders are not sharded; use sweep.dense_sweep for inference")
            for i in range(len(self)):
                sl = slice(i * self.batch_size, (i + 1) * self.batch_size)
                yield self.coords[sl], self.pixels[sl]
        self.epoch += 1


class PrefetchLoader:
    """Wraps an iterable of HOST batches (ideally pinned) and yields DEVICE batches: the H2D copy of batch
    i+1 runs on a side stream while batch i is being consumed (double buffering, no per-step allocation)."""

    def __init__(self, host_batches, device, depth: int = 2):
        self.host_batches = host_batches
        self.device = torch.device(device)
        self.depth = max(2, depth)
        self.copy_stream = torch.cuda.Stream(self.device)

    def __len__(self):
        return len(self.host_batches)

    def __iter__(self):
        main = torch.cuda.current_stream(self.device)
        slots = [None] * self.depth
        ready = [torch.cuda.Event() for _ in range(self.depth)]
        freed = [torch.cuda.Event() for _ in range(self.depth)]
        it = iter(self.host_batches)

        def issue(slot):
            try:
                batch = next(it)
            except StopIteration:
                return False
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(freed[slot])
                if slots[slot] is None:
                    slots[slot] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in batch)
                for dst, src in zip(slots[slot], batch):
                    if dst.shape != src.shape:
                        slots[slot] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in batch)
                        break
                for dst, src in zip(slots[slot], batch):
                    dst.copy_(src, non_blocking=True)
                ready[slot].record(self.copy_stream)
            return True

        for s in range(self.depth):
            freed[s].record(main)
        pending = []
        for s in range(self.depth - 1):
            if issue(s):
                pending.append(s)
        nxt = self.depth - 1
        while pending:
            cur = pending.pop(0)
            main.wait_event(ready[cur])
            if issue(nxt):
                pending.append(nxt)
                nxt = (nxt + 1) % self.depth
            yield slots[cur]
            freed[cur].record(main)


class MriImage(Dataset):
    """Coordinates in ``coords`` (M, D), intensities in ``pixels`` (M, 1) (datamodules.py:123-172)."""

    def __init__(self, config, image_path: str = None, norm_siren: bool = False, *args, **kwargs):
        super().__init__()
        path = image_path if image_path else config.image_path
        image = nifti.load(path).get_fdata(dtype=np.float32)
        axes = mgrid_axes(image.shape, norm_siren)
        mgrid = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1)
        pixels = torch.FloatTensor(image).flatten()
        if norm_siren:
            pixels = ((pixels - torch.min(pixels)) / (torch.max(pixels) - torch.min(pixels))) * 2 - 1
        else:
            pixels = (pixels - torch.min(pixels)) / (torch.max(pixels) - torch.min(pixels))
        coords = mgrid.reshape(len(pixels), config.dim_in)
        assert len(coords) == len(pixels)
        self.shape = tuple(image.shape)
        self.norm_siren = bool(norm_siren)
        self.coords = coords.contiguous()
        self.pixels = pixels.unsqueeze(-1)

    def __len__(self):
        return len(self.pixels)

    def __getitem__(self, idx):
        return self.coords[idx], self.pixels[idx]


class MriDataModule(pl.LightningDataModule):
    """ONE MRI image -> coords and pixels; train/val/test all see the same data (datamodules.py:175-252)."""

    def __init__(self, config=None, device=None, *args, **kwargs):
        super().__init__()
        self.train_ds = None
        self.val_ds = None
        self.test_ds = None
        self.config = config
        self.device = device

    def _dev(self):
        if self.device is not None:
            return torch.device(self.device)
        return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")

    def prepare_data(self) -> None:
        self.dataset = MriImage(config=self.config)
        self.train_ds = self.dataset
        self.test_ds = self.dataset
        self.val_ds = self.dataset

    def setup(self, stage=None):
        pass

    def _loader(self, ds, shuffle):
        rank, world = 0, 1
        if shuffle and torch.distributed.is_available() and torch.distributed.is_initialized():
            rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
        return DeviceBatchLoader(ds.coords, ds.pixels, self.config.batch_size, shuffle=shuffle, device=self._dev(),
                                 rank=rank, world_size=world, grid_shape=getattr(ds, "shape", None),
                                 grid_norm_siren=getattr(ds, "norm_siren", None))

    def train_dataloader(self):
        return self._loader(self.train_ds, True)

    def val_dataloader(self):
        return self._loader(self.val_ds, False)

    def test_dataloader(self):
        return self._loader(self.test_ds, False)

    def upsampling(self, shape, batch_size, norm_siren: bool = False):
        """Mock loader over a dense grid of ``shape`` (datamodules.py:229-252); for big grids prefer
        ``mri_interpolation_b200.sweep.dense_sweep`` which never materialises the coordinates."""
        axes = mgrid_axes(shape, norm_siren)
        mgrid = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1)
        fake_pix = torch.zeros(int(np.prod(shape)))
        coords = mgrid.reshape(len(fake_pix), len(shape))
        assert len(coords) == len(fake_pix)
        return DeviceBatchLoader(coords.contiguous(), fake_pix.unsqueeze(-1), batch_size, shuffle=False,
                                 device=self._dev())


class _OutOfScopeDataModule(pl.LightningDataModule):
    def __init__(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__} needs data that is not part of the hot-path scope")


class MNISTDataModule(_OutOfScopeDataModule): ...
class MriFramesDataModule(_OutOfScopeDataModule): ...
