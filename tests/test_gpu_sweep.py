"""GPU parity: dense-grid sweep (coordinate synthesis, fused hash+decoder kernel, slab sharding) and a
short end-to-end fit on the sample ankle volume against the oracle (PSNR within 0.1 dB)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import SAMPLE, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_grid_coords_bit_exact_with_reference_recipe():
    from mri_interpolation_b200 import functional as Fn
    fx = load_golden("sweep_coords.npz")
    shape = tuple(int(s) for s in fx["shape"])
    axes = [torch.linspace(0, 1, s) for s in shape]
    total = int(np.prod(shape))
    full = Fn.grid_coords(axes, 0, total, DEV)
    assert np.array_equal(full.cpu().numpy(), fx["coords"])  # same floats as torch.linspace/meshgrid
    part = Fn.grid_coords(axes, 37, 200, DEV)
    assert np.array_equal(part.cpu().numpy(), fx["coords"][37:237])
    from oracle import sweep as osweep
    for shp, ns in (((352, 352, 29), False), ((9, 8, 6, 57), True), ((33, 17), False)):
        axes = [osweep.axis_values(s, ns) for s in shp]
        tot = int(np.prod(shp))
        first = max(0, tot - 5000)
        got = Fn.grid_coords(axes, first, tot - first, DEV)
        assert torch.equal(got.cpu(), osweep.grid_coords(shp, ns)[first:])


def _small_hashmlp(dim, hidden, F=2, seed=5):
    from mri_interpolation_b200 import models
    torch.manual_seed(seed)
    base, fin = (4, 32) if dim != 4 else (4, 20)
    net = models.HashMLP(dim_in=dim, n_levels=5, n_features_per_level=F, log2_hashmap_size=11, base_resolution=base,
                         finest_resolution=fin, dim_hidden=hidden, dim_out=1, n_layers=2, batch_norm=False)
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for lv in net.encoder.levels:
            lv.embedding.weight.copy_(torch.randn(lv.embedding.weight.shape, generator=gen) * 0.3)
    return net


@pytest.mark.parametrize("dim,shape,hidden,F", [(3, (20, 17, 9), 64, 2), (4, (9, 8, 3, 7), 32, 2), (2, (40, 31), 16, 4),
                                                (3, (11, 12, 13), 128, 1)])
def test_fused_sweep_matches_unfused_and_oracle(dim, shape, hidden, F):
    from mri_interpolation_b200 import sweep
    from oracle import networks, sweep as osweep
    net = _small_hashmlp(dim, hidden, F)
    params = {k: v.detach().clone() for k, v in net.state_dict().items()}
    levels = [type("L", (), {})() for _ in net.encoder.levels]
    from oracle import hashgrid
    levels = hashgrid.geometry(dim, 5, 11, net.base_resolution, net.finest_resolution)
    ref = osweep.dense_sweep(lambda c: networks.hashmlp_forward(c, params, levels, 2, False), shape, 1000)
    net = net.to(DEV)
    fused = sweep.dense_sweep(net, shape)
    unfused = sweep.dense_sweep(net, shape, batch_size=777, fused=False)
    assert fused.shape == (int(np.prod(shape)), 1)
    assert rel_err(fused.reshape(shape), ref) < 1e-5
    assert rel_err(unfused.reshape(shape), ref) < 1e-5
    # slabs of 3 "ranks" concatenate to the full sweep (no communication needed)
    parts = [sweep.dense_sweep(net, shape, rank=r, world_size=3) for r in range(3)]
    assert torch.equal(torch.cat(parts), fused)


def test_sweep_of_siren_and_trainer_predict_equivalence():
    from mri_interpolation_b200 import datamodules, models, sweep
    from mri_interpolation_b200.pl_compat import pl
    from oracle import networks, sweep as osweep
    torch.manual_seed(1337)
    net = models.SirenNet(dim_in=3, dim_hidden=32, n_layers=3)
    torch.manual_seed(1337)
    params, w0s = networks.siren_init(dim_in=3, dim_hidden=32, n_layers=3)
    shape = (12, 10, 7)
    ref = osweep.dense_sweep(lambda c: networks.siren_forward(c, params, w0s), shape, 500, norm_siren=True)
    net = net.to(DEV)
    out = sweep.dense_sweep(net, shape, batch_size=300, norm_siren=True)
    assert rel_err(out.reshape(shape), ref) < 1e-4
    # the reference's own route: upsampling() loader + trainer.predict + concat (launcher.py:204-217)
    dm = datamodules.MriDataModule(config=None)
    tr = pl.Trainer(accelerator="gpu", precision=32, logger=False)
    pred = torch.concat(tr.predict(net, dm.upsampling(shape, 256, norm_siren=True)))
    assert torch.equal(pred, out)


def test_short_fit_on_ankle_slice_psnr_parity(tmp_path):
    """Config-2 style fit (hash grid + 2x64 GELU decoder, Adam 5e-3) on the 2-D+t slice data[:, :, 3, :] of the
    sample volume: same init, same batch sequence, fixed step count, oracle on CPU vs kernels on GPU."""
    import torch.nn.functional as F
    from mri_interpolation_b200 import metrics, models, nifti, sweep
    from mri_interpolation_b200.pl_compat import pl
    from oracle import networks, sweep as osweep
    vol = nifti.load(SAMPLE).get_fdata(np.float32)[::4, ::4, 3, :]  # (88, 88, 15)
    shape = vol.shape
    coords = osweep.grid_coords(shape)
    pixels = osweep.normalise_intensities(torch.from_numpy(np.ascontiguousarray(vol)))
    kw = dict(dim_in=3, n_levels=8, n_features_per_level=2, log2_hashmap_size=14, base_resolution=(16, 16, 5),
              finest_resolution=(88, 88, 15), dim_hidden=64, dim_out=1, n_layers=2)
    steps, batch = 120, 4096
    gen = torch.Generator().manual_seed(1337)
    batches = [torch.randint(0, coords.shape[0], (batch,), generator=gen) for _ in range(steps)]

    torch.manual_seed(1337)
    params, levels = networks.hashmlp_init(**kw)
    ref = {k: v.clone().requires_grad_() for k, v in params.items() if not k.startswith("layers.")}
    ropt = torch.optim.Adam(list(ref.values()), lr=5e-3)
    for idx in batches:
        ropt.zero_grad()
        F.mse_loss(pixels[idx], networks.hashmlp_forward(coords[idx], ref, levels, 2, True)).backward()
        ropt.step()
    with torch.no_grad():
        ref_img = osweep.dense_sweep(lambda c: networks.hashmlp_forward(c, ref, levels, 2, True), shape, 1 << 15).numpy()

    torch.manual_seed(1337)
    net = models.HashMLP(**kw, batch_norm=False, lr=5e-3).to(DEV)
    opt = net.configure_optimizers()
    cd, pd = coords.to(DEV), pixels.to(DEV)
    for i, idx in enumerate(batches):
        idx = idx.to(DEV)
        opt.zero_grad()
        net.training_step((cd[idx], pd[idx]), i).backward()
        opt.step()
    img = sweep.dense_sweep(net, shape).reshape(shape).cpu().numpy()
    truth = pixels.reshape(shape).numpy()
    psnr_ref, psnr_gpu = metrics.peak_signal_noise_ratio(truth, ref_img), metrics.peak_signal_noise_ratio(truth, img)
    ssim_ref, ssim_gpu = metrics.structural_similarity(truth, ref_img), metrics.structural_similarity(truth, img)
    print(f"PSNR oracle {psnr_ref:.3f} dB, B200 {psnr_gpu:.3f} dB; SSIM oracle {ssim_ref:.4f}, B200 {ssim_gpu:.4f}")
    assert psnr_ref > 20.0  # the fit actually learned something
    assert abs(psnr_ref - psnr_gpu) < 0.1  # north_star: within 0.1 dB at a fixed step count
    assert abs(ssim_ref - ssim_gpu) < 0.01


def test_config2_full_volume_fit_psnr_parity():
    """north_star gate on BASELINE configs[1] ITSELF: hash grid G4 (hash_config.json geometry: L16 F2 T2^19 base 16 ->
    finest 2489, x,y,z,t) + 2x64 GELU decoder, Adam lr 5e-3 (config/base.py:83), the whole 352x352x6x15 sample volume,
    batch 10 000 (config/base.py:63), K = 1 116 steps = one shuffled epoch (launcher.py:156-165).  Same seeded init, same
    index stream; the oracle (the reference's PyTorch arithmetic, restated) runs eagerly on the same GPU, the product path
    runs the fused kernels on locality-ordered batches (same sets).  PSNR on the full volume must agree within 0.1 dB."""
    import torch.nn.functional as F
    from mri_interpolation_b200 import functional as Fn, metrics, models, nifti, sweep
    from oracle import networks, sweep as osweep
    vol = nifti.load(SAMPLE).get_fdata(np.float32)
    shape = vol.shape
    assert shape == (352, 352, 6, 15)
    pixels = osweep.normalise_intensities(torch.from_numpy(np.ascontiguousarray(vol))).to(DEV)
    coords = osweep.grid_coords(shape).to(DEV)
    kw = dict(dim_in=4, n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16,
              finest_resolution=2489, dim_hidden=64, dim_out=1, n_layers=2)
    batch = 10_000
    gen = torch.Generator().manual_seed(1337)
    perm = torch.randperm(coords.shape[0], generator=gen).to(DEV)
    batches = [perm[i:i + batch] for i in range(0, perm.shape[0], batch)]
    assert len(batches) == 1116

    torch.manual_seed(1337)
    params, levels = networks.hashmlp_init(**kw)
    ref = {k: v.to(DEV).requires_grad_() for k, v in params.items() if not k.startswith("layers.")}
    ropt = torch.optim.Adam(list(ref.values()), lr=5e-3)
    for idx in batches:
        ropt.zero_grad()
        F.mse_loss(pixels[idx], networks.hashmlp_forward(coords[idx], ref, levels, 2, False)).backward()
        ropt.step()
    with torch.no_grad():
        ref_img = torch.cat([networks.hashmlp_forward(coords[i:i + (1 << 18)], ref, levels, 2, False)
                             for i in range(0, coords.shape[0], 1 << 18)]).reshape(shape).cpu().numpy()
    del ref, ropt

    torch.manual_seed(1337)
    net = models.HashMLP(**kw, batch_norm=False, lr=5e-3).to(DEV)
    opt = net.configure_optimizers()
    assert type(net(coords[:64]).grad_fn).__name__.startswith("HashDecoderFn")  # the one-kernel-per-direction path
    for i, idx in enumerate(batches):
        idx = Fn.locality_sort(idx, shape, block=1)  # what DeviceBatchLoader(grid_shape=...) feeds: same set, axis-0 fastest
        opt.zero_grad()
        loss = net.training_step((coords[idx], pixels[idx]), i)
        loss.backward()
        opt.step()
    img = sweep.dense_sweep(net, shape).reshape(shape).cpu().numpy()
    truth = pixels.reshape(shape).cpu().numpy()
    psnr_ref, psnr_gpu = metrics.peak_signal_noise_ratio(truth, ref_img), metrics.peak_signal_noise_ratio(truth, img)
    ssim_ref, ssim_gpu = metrics.structural_similarity(truth, ref_img), metrics.structural_similarity(truth, img)
    print(f"config 2, 1116 steps: PSNR oracle {psnr_ref:.3f} dB, B200 {psnr_gpu:.3f} dB; SSIM oracle {ssim_ref:.4f}, B200 {ssim_gpu:.4f}")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "psnr_config2.json"), "w") as f:
            json.dump({"steps": len(batches), "batch": batch, "psnr_oracle_db": psnr_ref, "psnr_b200_db": psnr_gpu,
                       "ssim_oracle": ssim_ref, "ssim_b200": ssim_gpu}, f)
    assert psnr_ref > 20.0  # the fit learned something
    assert abs(psnr_ref - psnr_gpu) < 0.1  # north_star: within 0.1 dB at a fixed step count
    assert abs(ssim_ref - ssim_gpu) < 0.01


def test_voxel_sampler_matches_mriimage_semantics():
    from mri_interpolation_b200 import functional as Fn
    from oracle import sweep as osweep
    shape = (9, 8, 3, 5)
    gen = torch.Generator().manual_seed(4)
    vol = torch.rand(shape, generator=gen)
    pixels = osweep.normalise_intensities(vol)
    coords = osweep.grid_coords(shape)
    idx = torch.randint(0, coords.shape[0], (5000,), generator=gen)
    s = Fn.VoxelSampler(pixels.to(DEV), shape)
    x, y = s.batch(idx.to(DEV))
    assert torch.equal(x.cpu(), coords[idx]) and torch.equal(y.cpu(), pixels[idx])


def test_prefetch_loader_yields_host_batches_in_order():
    from mri_interpolation_b200.datamodules import PrefetchLoader
    gen = torch.Generator().manual_seed(8)
    host = [(torch.rand(1000 + 10 * i, 4, generator=gen).pin_memory(), torch.rand(1000 + 10 * i, 1, generator=gen).pin_memory())
            for i in range(7)]
    seen = []
    for x, y in PrefetchLoader(host, DEV):
        seen.append((x.clone(), y.clone()))  # slots are recycled: consume before asking for the next batch
    assert len(seen) == 7
    for (x, y), (hx, hy) in zip(seen, host):
        assert torch.equal(x.cpu(), hx) and torch.equal(y.cpu(), hy)


@pytest.mark.parametrize("dim,shape,n_levels,hidden,base,finest", [
    (4, (13, 11, 3, 7), 16, 64, 16, 200), (3, (21, 19, 10), 16, 64, 16, 200),
    (3, (37, 19, 10), 8, 64, (8, 6, 4), (96, 72, 12)),  # the notebook's shape: L = 8, anisotropic V2 levels
    (4, (32, 5, 3, 4), 16, 128, 16, 200), (3, (21, 19, 10), 4, 128, 16, 200), (4, (13, 11, 3, 7), 8, 128, 8, 300)])
def test_tensor_core_sweep_kernel_matches_model_forward_and_oracle(dim, shape, n_levels, hidden, base, finest):
    """F = 2 geometries (L in {4, 8, 16}, hidden in {64, 128}): mri_hashmlp_sweep takes the encoder+decoder tensor-core kernel
    with coordinates synthesised from the voxel index; slabs start at arbitrary voxels and need not be multiples of 16."""
    from mri_interpolation_b200 import models, sweep
    from oracle import hashgrid, networks, sweep as osweep
    torch.manual_seed(11)
    net = models.HashMLP(dim_in=dim, n_levels=n_levels, n_features_per_level=2, log2_hashmap_size=12, base_resolution=base,
                         finest_resolution=finest, dim_hidden=hidden, dim_out=1, n_layers=2, batch_norm=False)
    gen = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for lv in net.encoder.levels:
            lv.embedding.weight.copy_(torch.randn(lv.embedding.weight.shape, generator=gen) * 0.3)
    params = {k: v.detach().clone() for k, v in net.state_dict().items()}
    levels = hashgrid.geometry(dim, n_levels, 12, base, finest)
    aniso = not isinstance(base, int)
    ref = osweep.dense_sweep(lambda c: networks.hashmlp_forward(c, params, levels, 2, aniso), shape, 1000)
    net = net.to(DEV)
    assert sweep._fused_plan(net) is not None
    fused = sweep.dense_sweep(net, shape)
    unfused = sweep.dense_sweep(net, shape, batch_size=1000, fused=False)
    assert rel_err(fused.reshape(shape), ref) < 1e-5
    assert torch.equal(fused, unfused)  # same kernel arithmetic as the module's no-grad forward
    parts = [sweep.dense_sweep(net, shape, rank=r, world_size=5) for r in range(5)]
    assert torch.equal(torch.cat(parts), fused)


@pytest.mark.parametrize("world", [1, 3])
def test_slab_sweeper_pipelined_host_copy_equals_device_sweep(world):
    """SlabSweeper.run_to_host(): the slab arrives in pinned host memory chunk by chunk (whole axis-0 planes, copies on a
    side stream) and equals the device sweep bit for bit - fused hash model and SIREN, slabs that cut planes."""
    from mri_interpolation_b200 import models, sweep
    torch.manual_seed(5)
    hash_net = models.HashMLP(dim_in=4, n_levels=16, n_features_per_level=2, log2_hashmap_size=12, base_resolution=16,
                              finest_resolution=200, dim_hidden=64, dim_out=1, n_layers=2, batch_norm=False).to(DEV)
    siren = models.SirenNet(dim_in=3, dim_hidden=64, n_layers=3).to(DEV)
    for net, shape, ns in ((hash_net, (40, 11, 3, 7), False), (siren, (33, 19, 10), True)):
        whole = sweep.dense_sweep(net, shape, norm_siren=ns, batch_size=1500)
        parts = []
        for r in range(world):
            sw = sweep.SlabSweeper(net, shape, batch_size=1500, norm_siren=ns, rank=r, world_size=world)
            dev_out = sw.run().clone()
            host = sw.run_to_host(n_chunks=3)
            assert host.is_pinned() and torch.equal(host, dev_out.cpu())
            assert torch.equal(sw.run_to_host(n_chunks=5), dev_out.cpu())  # buffers are reused across calls
            covered = sw._chunks(max(1, sw.count // 3)) if sw.plan is not None else None
            if covered is not None:
                assert covered[0][0] == sw.first and sum(c for _, c in covered) == sw.count
            parts.append(host.clone())
        assert torch.equal(torch.cat(parts), whole.cpu())
    assert torch.equal(sweep.dense_sweep(hash_net, (40, 11, 3, 7), out_host=True), sweep.dense_sweep(hash_net, (40, 11, 3, 7)).cpu())


@pytest.mark.parametrize("n_levels", [16, 5])
def test_fused_sweep_folds_eval_batchnorm_of_the_shipped_decoder(n_levels):
    """The shipped HashMLP decoder (Linear -> BatchNorm1d -> GELU -> Dropout, models.py:718-739) in eval mode: BN folds
    into the Linear, so the fused sweep kernels (tensor-core for L=16, CUDA-core otherwise) serve it too."""
    from mri_interpolation_b200 import models, sweep
    torch.manual_seed(4)
    net = models.HashMLP(dim_in=3, n_levels=n_levels, n_features_per_level=2, log2_hashmap_size=12, base_resolution=8,
                         finest_resolution=128, dim_hidden=64, dim_out=1, n_layers=2, dropout=0.1)
    gen = torch.Generator().manual_seed(8)
    with torch.no_grad():
        for lv in net.encoder.levels:
            lv.embedding.weight.copy_(torch.randn(lv.embedding.weight.shape, generator=gen) * 0.3)
        for blk in net.decoder:
            bn = blk[1]
            assert isinstance(bn, torch.nn.BatchNorm1d)
            bn.running_mean.copy_(torch.randn(bn.num_features, generator=gen) * 0.2)
            bn.running_var.copy_(torch.rand(bn.num_features, generator=gen) + 0.3)
            bn.weight.copy_(torch.randn(bn.num_features, generator=gen))
            bn.bias.copy_(torch.randn(bn.num_features, generator=gen) * 0.1)
    net = net.to(DEV)
    net.train()
    assert sweep._fused_plan(net) is None  # batch statistics in training mode: not foldable
    shape = (19, 18, 7)
    fused = sweep.dense_sweep(net, shape)
    unfused = sweep.dense_sweep(net, shape, batch_size=999, fused=False)
    assert net.training  # dense_sweep restores the mode
    assert rel_err(fused, unfused) < 1e-5
    assert float(fused.abs().max()) > 1e-3


def test_gpu_metrics_match_the_numpy_oracle():
    """csrc/metrics.cu (MSE / PSNR / SSIM in float64 on the device) against oracle/sweep.py's numpy restatement of the
    skimage definitions; numpy inputs and CUDA tensors give the same numbers."""
    from mri_interpolation_b200 import metrics
    from oracle import sweep as osweep
    rng = np.random.default_rng(0)
    for shape in ((16, 16, 2, 3), (23, 9, 5), (7, 40)):
        a = rng.random(shape).astype(np.float32)
        b = np.clip(a + rng.normal(0, 0.05, a.shape).astype(np.float32), 0, 1)
        assert metrics.peak_signal_noise_ratio(a, b) == pytest.approx(osweep.psnr(a, b), abs=1e-9)
        assert metrics.structural_similarity(a, b) == pytest.approx(osweep.ssim_slices(a, b), abs=1e-9)
        assert metrics.mean_squared_error(a, b) == pytest.approx(float(np.mean((a.astype(np.float64) - b) ** 2)), rel=1e-12)
        ta, tb = torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV)
        assert metrics.structural_similarity(ta, tb) == metrics.structural_similarity(a, b)
    a = rng.random((16, 16, 2, 3)).astype(np.float32)
    assert metrics.peak_signal_noise_ratio(a, a + 0.01) == pytest.approx(40.0, abs=1e-3)
    assert metrics.structural_similarity(a, a) == pytest.approx(1.0)
    assert metrics.peak_signal_noise_ratio(a, a) == float("inf")


def test_gpu_metrics_on_the_sample_volume_and_linear_time_baseline(tmp_path):
    """interp.py:35-52 on the reference's own evaluation data (slice 3 of the ankle volume, as the script does): the
    GPU baseline is bit-identical to the numpy recipe, kept frames are reproduced exactly, and PSNR / SSIM of the
    re-interpolated volume agree with the oracle; write_scores emits the reference's scores.txt fields."""
    import interp
    from mri_interpolation_b200 import metrics, nifti
    from oracle import sweep as osweep
    data = nifti.load(SAMPLE).get_fdata(np.float32)
    data = np.ascontiguousarray((data / data.max())[:, :, 3, :])
    want = osweep.linear_time_baseline(data)
    got = interp.linear_time_baseline(data)
    assert got.is_cuda and got.shape == data.shape
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(got.cpu().numpy()[..., ::2], data[..., ::2])
    for t in (1, 2, 6):  # even / odd frame counts, single frame
        d = np.ascontiguousarray(data[:40, :30, :t])
        assert np.array_equal(metrics.linear_time_baseline(d).cpu().numpy(), osweep.linear_time_baseline(d))
    assert metrics.peak_signal_noise_ratio(data, got) == pytest.approx(osweep.psnr(data, want), abs=1e-9)
    assert metrics.structural_similarity(data, got) == pytest.approx(osweep.ssim_slices(data, want), abs=1e-9)
    scores = metrics.write_scores(str(tmp_path / "scores.txt"), data, got, {"Number of trainable parameters": 7})
    text = open(tmp_path / "scores.txt").read()
    assert "MSE : " in text and "PSNR : " in text and "SSIM : " in text and "Number of trainable parameters : 7" in text
    assert scores["PSNR"] == pytest.approx(osweep.psnr(data, want), abs=1e-9)


def test_mri_datamodule_batches_from_the_gather_kernel_equal_index_select(tmp_path):
    """MriDataModule's shuffled CUDA loader rebuilds the coordinates from the voxel index (mri_gather_voxels) instead of
    index_select on the materialised mesh: every batch is bit-identical to coords[idx], pixels[idx] of the same epoch."""
    from mri_interpolation_b200 import config as cfgmod, datamodules, nifti
    rng = np.random.default_rng(3)
    vol = rng.random((11, 7, 3, 5)).astype(np.float32)
    path = str(tmp_path / "vol.nii.gz")
    nifti.save(vol, path)
    cfg = cfgmod.HashConfig()
    cfg.image_path, cfg.batch_size, cfg.dim_in = path, 300, 4
    dm = datamodules.MriDataModule(config=cfg, device=DEV)
    dm.prepare_data()
    fast = dm.train_dataloader()
    assert fast._voxels is not None
    slow = datamodules.DeviceBatchLoader(dm.dataset.coords, dm.dataset.pixels, 300, shuffle=True, device=DEV,
                                         grid_shape=dm.dataset.shape)  # same seed, no grid declaration: index_select
    assert slow._voxels is None
    n = 0
    for (xa, ya), (xb, yb) in zip(fast, slow):
        assert torch.equal(xa, xb) and torch.equal(ya, yb)
        n += xa.shape[0]
    assert n == vol.size
