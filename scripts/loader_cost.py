"""What one shuffled epoch of indices costs on the device, plain vs locality-ordered (DeviceBatchLoader.epoch_indices)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_interpolation_b200 import datamodules

shape = (352, 352, 6, 15)
n = int(np.prod(shape))
dev = torch.device("cuda", 0)
coords = torch.zeros(n, 1, device=dev)
for bs in (10_000, 1 << 19):
    for grid in (None, shape):
        ld = datamodules.DeviceBatchLoader(coords, coords, bs, shuffle=True, device=dev, grid_shape=grid)
        ld.epoch_indices(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            ld.epoch_indices()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 5 * 1e3
        print(f"batch {bs:7d} locality={grid is not None}: {ms:.2f} ms per epoch of {len(ld)} batches = {ms / len(ld) * 1e3:.1f} us per batch")
t0 = time.perf_counter()
for _ in range(5):
    torch.randperm(n, device=dev)
torch.cuda.synchronize()
print(f"torch.randperm({n}) alone: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms")
