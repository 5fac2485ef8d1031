"""CPU, world_size 2, gloo: the host-side data-parallel logic.

Each rank computes the oracle's gradients of the mean loss on ITS share of a global batch; the flat
gradient arena is summed with distributed.allreduce_sum_ and scaled by the returned 1/world - the
result must equal the single-process gradient of the global batch.  Also: dense-sweep slabs from the
two ranks tile the volume, and the shuffled loaders give disjoint shares."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, init_file, out_dir):
    sys.path.insert(0, ROOT)
    import torch.nn.functional as F
    from mri_interpolation_b200 import datamodules, distributed, sweep
    from oracle import hashgrid, networks

    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        kw = dict(dim_in=3, n_levels=3, n_features_per_level=2, log2_hashmap_size=9, base_resolution=4,
                  finest_resolution=16, dim_hidden=8, dim_out=1, n_layers=2)
        torch.manual_seed(1337)  # identical replicas
        params, levels = networks.hashmlp_init(**kw)
        params = {k: (v * (50.0 if "embedding" in k else 1.0)).requires_grad_() for k, v in params.items()
                  if not k.startswith("layers.")}
        gen = torch.Generator().manual_seed(9)
        n_global = 600
        x, y = torch.rand(n_global, 3, generator=gen), torch.rand(n_global, 1, generator=gen)
        first, count = distributed.split_batch(n_global, rank, world)
        loss = F.mse_loss(y[first:first + count], networks.hashmlp_forward(x[first:first + count], params, levels, 2, False))
        loss.backward()
        flat = torch.cat([p.grad.reshape(-1) for p in params.values()])  # the gradient arena
        scale = distributed.allreduce_sum_(flat)
        flat *= scale
        # single-process reference on the whole batch
        ref = {k: v.detach().clone().requires_grad_() for k, v in params.items()}
        F.mse_loss(y, networks.hashmlp_forward(x, ref, levels, 2, False)).backward()
        ref_flat = torch.cat([p.grad.reshape(-1) for p in ref.values()])
        ok_grad = bool(torch.allclose(flat, ref_flat, rtol=1e-5, atol=1e-8))
        # sweep slabs
        shape = (7, 5, 3)
        f, c = sweep.slab_range(int(np.prod(shape)), rank, world)
        spans = [None] * world
        dist.all_gather_object(spans, (f, c))
        # loader shares
        coords = torch.arange(101, dtype=torch.float32).reshape(-1, 1)
        ld = datamodules.DeviceBatchLoader(coords, coords.clone(), 16, shuffle=True, device="cpu", seed=3, rank=rank,
                                           world_size=world)
        mine = torch.cat([a for a, _ in ld]).flatten().tolist()
        shares = [None] * world
        dist.all_gather_object(shares, mine)
        # n % world != 0 with a full last batch on one rank only (n = 2*16+1) used to give rank 0 two batches
        # and rank 1 one - a hang in the collective optimiser step.  Every rank must see the same batch sizes.
        odd = torch.arange(33, dtype=torch.float32).reshape(-1, 1)
        sizes = [None] * world
        for grid in (None, (11, 3)):
            ld2 = datamodules.DeviceBatchLoader(odd, odd.clone(), 16, shuffle=True, device="cpu", seed=4, rank=rank, world_size=world,
                                                grid_shape=grid)
            dist.all_gather_object(sizes, [a.shape[0] for a, _ in ld2] + [len(ld2)])
            assert all(s == sizes[0] for s in sizes) and sizes[0] == [16, 1, 2], sizes
        if rank == 0:
            np.save(os.path.join(out_dir, "result.npy"),
                    np.asarray([ok_grad, scale == 0.5, spans == [sweep.slab_range(105, r, world) for r in range(world)],
                                sorted(set(sum(shares, []))) == list(range(101)) and all(len(sh) == 51 for sh in shares)], dtype=bool))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_and_sharding():
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "rendezvous")
        mp.spawn(_worker, args=(2, init_file, d), nprocs=2, join=True)
        res = np.load(os.path.join(d, "result.npy"))
        assert res.tolist() == [True, True, True, True]


def test_env_world_defaults(monkeypatch):
    from mri_interpolation_b200 import distributed
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    assert distributed.env_world() == (0, 0, 1)
    monkeypatch.setenv("RANK", "3"); monkeypatch.setenv("LOCAL_RANK", "3"); monkeypatch.setenv("WORLD_SIZE", "8")
    assert distributed.env_world() == (3, 3, 8)
    assert distributed.allreduce_sum_(torch.ones(4)) == 1.0  # not initialised -> no-op
