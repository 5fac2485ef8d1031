"""Per-level timing of the stand-alone hash-grid scatter (and the whole gather) under different batch orders.
Shows which levels gain from a locality-ordered batch and where same-address reductions start to serialise."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mri_interpolation_b200 import _lib, encoding, functional as Fn  # noqa: E402

G4 = dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16, finest_resolution=2489)
shape = (352, 352, 6, 15)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
enc = encoding.MultiResHashGrid(4, **G4).to(dev)
n = 1 << 19
total = int(np.prod(shape))
pix = torch.rand(total, device=dev)
sampler = Fn.VoxelSampler(pix, shape)
index = torch.randint(0, total, (n,), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
go = torch.randn(n, 32, device=dev)
tables = [lv.embedding.weight for lv in enc.levels]
gbuf = torch.zeros(sum(t.numel() for t in tables) + 64, device=dev)
offs, off = [], 0
for t in tables:
    offs.append(off)
    off += (t.numel() + 3) // 4 * 4
levels = _lib.make_levels(enc._resolutions, enc._rows, offs)


def timed(fn, reps=5, inner=20):
    """`inner` back-to-back launches per timing (a single 20-40 us launch is below what the enqueue gaps and the ~2 us
    event resolution let one measure); tables and inputs are therefore L2-warm, as inside the training step."""
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / inner)
    return float(np.median(ts))


out = {}
for order in ("none", "x0fast", "blk4", "blk16"):
    idx = index if order == "none" else Fn.locality_sort(index, shape, block=int(order[3:]) if order.startswith("blk") else 1)
    x, _ = sampler.batch(idx)
    per = []
    for l in range(16):
        per.append(timed(lambda: _lib.call("mri_hashgrid_backward_levels", x.data_ptr(), n, 4, go.data_ptr(), gbuf.data_ptr(),
                                           levels, 16, 2, l, 1, _lib.stream())))
    with torch.no_grad():
        fwd = timed(lambda: enc(x))
    out[order] = {"bwd_per_level_us": [round(1e3 * v, 1) for v in per], "bwd_sum_ms": sum(per), "fwd_ms": fwd}
    print(order, out[order], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "level_probe.json"), "w"))
