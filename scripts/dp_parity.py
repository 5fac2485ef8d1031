#!/usr/bin/env python
"""Data-parallel parity gate, run under torchrun with W >= 2 ranks on W GPUs:
W-GPU training (each rank: its contiguous share of every global batch, NCCL all-reduce of the flat gradient
arena, fused Adam) must reproduce single-GPU training on the full global batches to summation-order noise.
Also checks that the dense-sweep slabs of the ranks tile the single-GPU sweep exactly."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mri_interpolation_b200 import distributed, models, sweep  # noqa: E402
from mri_interpolation_b200.optim import FusedAdam  # noqa: E402

rank, local_rank, world = distributed.init_from_env("nccl")
dev = torch.device("cuda", local_rank)
kw = dict(dim_in=4, n_levels=16, n_features_per_level=2, log2_hashmap_size=15, base_resolution=8, finest_resolution=256,
          dim_hidden=64, dim_out=1, n_layers=2, batch_norm=False, lr=5e-3)
steps, n_global = 8, 1 << 16
torch.manual_seed(1337)
dp_model = models.HashMLP(**kw).to(dev)
dp_opt = dp_model.configure_optimizers()
overlap = (not dp_opt.sharded) and os.environ.get("MRI_DP_OVERLAP", "1") == "1" and dp_opt.enable_overlap(dp_model.encoder, n_groups=3)
torch.manual_seed(1337)
ref_model = models.HashMLP(**kw).to(dev)
ref_opt = FusedAdam(ref_model.parameters(), lr=5e-3, data_parallel=False)
gen = torch.Generator(device=dev).manual_seed(7)  # same stream on every rank -> same global batches
first, count = distributed.split_batch(n_global, rank, world)
for step in range(steps):
    x = torch.rand(n_global, 4, device=dev, generator=gen)
    y = torch.rand(n_global, 1, device=dev, generator=gen)
    dp_model.training_step((x[first:first + count], y[first:first + count]), step).backward()
    dp_opt.step(); dp_opt.zero_grad()
    ref_model.training_step((x, y), step).backward()
    ref_opt.step(); ref_opt.zero_grad()
worst = 0.0
for (k, a), (_, b) in zip(dp_model.state_dict().items(), ref_model.state_dict().items()):
    if a.dtype.is_floating_point and a.numel() and float(b.norm()) > 0:
        worst = max(worst, float((a - b).norm() / b.norm()))
# replicas identical across ranks
flat = dp_opt.arena.data.clone()
torch.cuda.synchronize()
ref0 = flat.clone()
dist.broadcast(ref0, 0)
replica_diff = float((flat - ref0).abs().max())
# sweep slabs
shape = (24, 20, 6, 9)
local = sweep.dense_sweep(dp_model, shape, rank=rank, world_size=world)
full = sweep.gather_slabs(local, shape)
ok_sweep = True
if rank == 0:
    single = sweep.dense_sweep(dp_model, shape).reshape(shape).cpu().numpy()
    ok_sweep = bool((single == full).all())
res = {"world": world, "steps": steps, "global_batch": n_global, "max_rel_param_diff_vs_single_gpu": worst,
       "max_abs_replica_diff": replica_diff, "allreduces": dp_opt.allreduce_count, "overlap": bool(overlap), "sharded_p2p_adam": bool(dp_opt.sharded), "multimem": bool(getattr(dp_opt, "_grad_mc", 0)), "inkernel_sync": getattr(dp_opt, "_peer_flags", None) is not None, "sweep_slabs_tile_exactly": ok_sweep}
if rank == 0:
    print(json.dumps(res))
    out_dir = os.environ.get("MRI_DP_PARITY_OUT", os.path.join(ROOT, "gpurun_out"))
    os.makedirs(out_dir, exist_ok=True)
    json.dump(res, open(os.path.join(out_dir, f"dp_parity_w{world}.json"), "w"))
    assert worst < 1e-4 and replica_diff == 0.0 and ok_sweep and dp_opt.allreduce_count == steps, res
dist.destroy_process_group()
