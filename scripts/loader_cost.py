"""Device time of what DeviceBatchLoader does per batch on the sample volume (CUDA events): epoch draw, index_select of the
(11.15 M, 4) coordinate array and of the intensities, against mri_gather_voxels (coordinates synthesised from the index)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mri_interpolation_b200 import config as cfgmod, datamodules
from mri_interpolation_b200 import functional as Fn

dev = torch.device("cuda", 0)
cfg = cfgmod.HashConfig(); cfg.image_path, cfg.batch_size = bench.SAMPLE, 1 << 19
dm = datamodules.MriDataModule(config=cfg, device=dev); dm.prepare_data()
loader = dm.train_dataloader()
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
order = loader.epochs.epoch()
idx = order[: 1 << 19]
print("epoch draw ms", timed(lambda: loader.epochs.epoch(), 5), "per batch", timed(lambda: loader.epochs.epoch(), 5) / 21)
print("coords.index_select ms", timed(lambda: loader.coords.index_select(0, idx)))
print("pixels.index_select ms", timed(lambda: loader.pixels.index_select(0, idx)))
sampler = Fn.VoxelSampler(loader.pixels.reshape(-1), dm.dataset.shape)
print("mri_gather_voxels ms", timed(lambda: sampler.batch(idx)))
