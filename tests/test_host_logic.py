"""CPU: host-side logic of the product package - construction parity with the reference (seeded
state_dicts bit-identical to the golden fixtures), level geometry, NIfTI reader, loaders, the
Lightning stand-in's orchestration, slab/batch partitioning."""
import os

import numpy as np
import pytest
import torch
from torch import nn

from conftest import SAMPLE, SIREN_CASES, load_golden
from mri_interpolation_b200 import functional as Fn
from mri_interpolation_b200 import config, datamodules, distributed, encoding, metrics, models, nifti, sweep
from mri_interpolation_b200.pl_compat import pl


@pytest.mark.parametrize("case", SIREN_CASES)
def test_sirennet_seeded_state_dict_equals_reference(case):
    fx = load_golden(f"siren_{case}.npz")
    kw = {k[3:]: fx[k].item() for k in fx if k.startswith("kw_")}
    torch.manual_seed(1337)
    net = models.SirenNet(**kw)
    sd = net.state_dict()
    ref_keys = {k[6:] for k in fx if k.startswith("param:")}
    assert set(sd.keys()) == ref_keys
    for k in ref_keys:
        assert np.array_equal(sd[k].numpy(), fx[f"param:{k}"]), k


def test_siren_param_counts_known_answers():
    assert sum(p.numel() for p in models.SirenNet(dim_in=2, dim_hidden=352, n_layers=4).parameters()) == 374177  # nb:825
    assert sum(p.numel() for p in models.SirenNet(dim_in=3, dim_hidden=1408, n_layers=4).parameters()) == 5958657  # nb:1347
    assert sum(p.numel() for p in models.SirenNet(dim_in=3, dim_hidden=1024, n_layers=8).parameters()) == 7352321
    assert sum(p.numel() for p in models.SirenNet(dim_in=4, dim_hidden=256, n_layers=5).parameters()) == 264705


def test_hashmlp_seeded_state_dict_equals_reference():
    fx = load_golden("hashmlp_small.npz")
    kw = {k[3:]: fx[k].item() for k in fx if k.startswith("kw_")}
    torch.manual_seed(1337)
    net = models.HashMLP(**kw)
    sd = net.state_dict()
    for k in fx:
        if k.startswith("param:") and "embedding" not in k:
            assert np.array_equal(sd[k[6:]].numpy(), fx[k]), k
    # same keys as the reference's HashMLP incl. BatchNorm buffers and the dead BaseMLP stack
    fb = load_golden("hashmlp_bn_small.npz")
    ref_keys = {k.split(":", 1)[1] for k in fb if k.startswith(("param:", "buffer:"))}
    assert set(sd.keys()) == ref_keys
    # embeddings: U(-1e-4, 1e-4) after the N(0,1) draw
    w = sd["encoder.levels.0.embedding.weight"]
    assert float(w.abs().max()) <= 1e-4


def test_notebook_variant_has_no_batchnorm():
    net = models.HashMLP(dim_in=3, n_levels=2, n_features_per_level=2, log2_hashmap_size=8, base_resolution=4,
                         finest_resolution=8, dim_hidden=16, n_layers=2, batch_norm=False)
    assert [type(m).__name__ for m in net.decoder[0]] == ["Linear", "GELU"]
    assert net.decoder[1][0].out_features == 1


def test_encoder_geometry_matches_reference():
    g = load_golden("geometry.npz")
    enc = encoding.MultiResHashGrid(dim=4, **config.g4_hash_kwargs())
    assert [lv.resolution for lv in enc.levels] == g["g4_res"].tolist()
    assert [lv.hashmap_size for lv in enc.levels] == g["g4_rows"].tolist()
    assert enc.output_dim == 32 and enc.input_dim == 4
    assert sum(p.numel() for p in enc.parameters()) == 15279648
    v2 = encoding.MultiResHashGridV2(dim=3, n_levels=8, n_features_per_level=2, log2_hashmap_size=23,
                                     base_resolution=(64, 64, 5), finest_resolution=(512, 512, 15))
    assert sum(p.numel() for p in v2.parameters()) == 6009032  # nb:2792
    assert np.array_equal(np.asarray([lv.resolution.tolist() for lv in v2.levels]), g["nbv2_res"].astype(float))
    assert list(v2.state_dict().keys())[0] == "levels.0.embedding.weight"


def test_shipped_hashconfig_tuple_mismatch_fails_at_forward_like_reference():
    # config/base.py:73-74 ships 3-tuples with dim_in=4: the reference constructs, then fails in forward
    enc = encoding.MultiResHashGridV2(dim=4, n_levels=2, n_features_per_level=1, log2_hashmap_size=10,
                                      base_resolution=(64, 64, 5), finest_resolution=(352, 352, 15))
    with pytest.raises(RuntimeError):
        enc(torch.rand(3, 4))


def test_fast_hash_utility_matches_oracle():
    from oracle import hashgrid
    gen = torch.Generator().manual_seed(3)
    ind = torch.randint(0, 3000, (50, 16, 4), generator=gen)
    a = encoding.fast_hash(ind.clone(), torch.tensor(encoding.PRIMES), 65536)
    assert torch.equal(a, hashgrid.spatial_hash(ind, 65536))


def test_out_of_scope_names_importable_but_not_constructible():
    # config/base.py:12-14 imports these six names; the first four are implemented (zoo.py, tests/test_zoo.py), the
    # PSF / random-Fourier-feature / Gabor experiments stay importable stubs
    for name in ("HashSirenNet", "ModulatedSirenNet", "MultiHashMLP", "MultiSiren", "SirenNet", "HashMLP"):
        assert isinstance(getattr(models, name), type)
    for name in ("RffNet", "GaborNet", "PsfSirenNet", "RealGaborLayer", "ComplexGaborLayer"):
        with pytest.raises(NotImplementedError):
            getattr(models, name)()


def test_nifti_reader_sample_volume_facts():
    img = nifti.load(SAMPLE)
    assert img.shape == (352, 352, 6, 15) and img.raw.dtype == np.int16
    assert img.slope == pytest.approx(27.2891, abs=1e-4) and img.inter == 0.0
    assert int(img.raw.max()) == 91 and np.unique(img.raw).size == 92
    assert (img.raw != 0).mean() == pytest.approx(0.479, abs=1e-3)
    data = img.get_fdata(np.float32)
    assert data.dtype == np.float32 and np.array_equal(data, img.raw.astype(np.float32) * np.float32(img.slope))


def test_nifti_roundtrip(tmp_path):
    a = np.random.default_rng(0).random((5, 4, 3, 2)).astype(np.float32)
    for name in ("a.nii", "a.nii.gz"):
        nifti.save(a, str(tmp_path / name))
        b = nifti.load(str(tmp_path / name))
        assert b.shape == a.shape and np.array_equal(b.get_fdata(np.float32), a)


def test_mriimage_matches_reference_recipe():
    cfg = config.HashConfig()
    ds = datamodules.MriImage(cfg)
    assert ds.coords.shape == (11151360, 4) and ds.pixels.shape == (11151360, 1)
    # C-order flatten: last axis (time) fastest; first step on the time axis = 1/14
    assert torch.equal(ds.coords[0], torch.zeros(4)) and float(ds.coords[1, 3]) == pytest.approx(1 / 14)
    assert float(ds.coords[15, 2]) == pytest.approx(1 / 5)
    assert float(ds.pixels.min()) == 0.0 and float(ds.pixels.max()) == 1.0
    raw = nifti.load(cfg.image_path).raw.reshape(-1)
    np.testing.assert_allclose(ds.pixels[:, 0].numpy(), raw / 91.0, rtol=0, atol=1e-7)
    x, y = ds[123456]
    assert x.shape == (4,) and y.shape == (1,)


def test_device_batch_loader_epoch_semantics():
    coords = torch.arange(103, dtype=torch.float32).reshape(-1, 1)
    pix = coords.clone()
    ld = datamodules.DeviceBatchLoader(coords, pix, 10, shuffle=True, device="cpu", seed=1)
    assert len(ld) == 11
    seen = torch.cat([x for x, _ in ld]).flatten()
    assert seen.shape[0] == 103 and torch.equal(seen.sort().values, coords.flatten())
    second = torch.cat([x for x, _ in ld]).flatten()
    assert not torch.equal(seen, second)  # new permutation every epoch
    plain = [x for x, _ in datamodules.DeviceBatchLoader(coords, pix, 50, device="cpu")]
    assert [p.shape[0] for p in plain] == [50, 50, 3] and torch.equal(torch.cat(plain), coords)
    # data-parallel shares are disjoint and cover the epoch
    shards = []
    for r in range(3):
        ld = datamodules.DeviceBatchLoader(coords, pix, 8, shuffle=True, device="cpu", seed=5, rank=r, world_size=3)
        shards.append(torch.cat([x for x, _ in ld]).flatten())
    allv = torch.cat(shards)
    # padded (wrap-around) to a multiple of the world size: every rank sees the same number and sizes of batches
    assert [s.shape[0] for s in shards] == [35, 35, 35] and torch.equal(allv.unique(), coords.flatten())


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("shape,bs", [((7, 5, 3), 16), ((6, 4, 2, 3), 25), ((9, 3), 27)])
def test_locality_ordered_batches_are_the_same_sets_in_axis0_fastest_order(world, shape, bs):
    """grid_shape switches on locality-ordered batches: per batch the SAME voxels as the plain shuffled loader (same
    seed), ordered with the axis-0 index fastest; ranks see equally many, equally sized batches."""
    n = int(np.prod(shape))
    coords = torch.arange(n, dtype=torch.float32).reshape(-1, 1)
    lens = set()
    for r in range(world):
        plain = datamodules.DeviceBatchLoader(coords, coords.clone(), bs, shuffle=True, device="cpu", seed=9, rank=r, world_size=world)
        local = datamodules.DeviceBatchLoader(coords, coords.clone(), bs, shuffle=True, device="cpu", seed=9, rank=r, world_size=world,
                                              grid_shape=shape)
        for epoch in range(2):
            pb, lb = [x.flatten().long() for x, _ in plain], [x.flatten().long() for x, _ in local]
            lens.add(tuple(b.shape[0] for b in lb))
            assert len(pb) == len(lb) == len(local)
            for a, b in zip(pb, lb):
                assert torch.equal(a.sort().values, b.sort().values)
                key = Fn.locality_key(b, shape, block=1)
                assert bool((key[1:] >= key[:-1]).all())
    assert len(lens) == 1  # identical batch sizes on every rank and in every epoch


def test_locality_key_walks_axis0_fastest():
    shape = (5, 4, 3)
    idx = torch.arange(int(np.prod(shape)))
    order = torch.argsort(Fn.locality_key(idx, shape, block=1))
    v0, v1, v2 = idx[order] // 12, idx[order] // 3 % 4, idx[order] % 3
    assert v0[:5].tolist() == [0, 1, 2, 3, 4] and v1[:5].tolist() == [0] * 5 and v2[:5].tolist() == [0] * 5
    assert v1[5:10].tolist() == [1] * 5 and v2[:20].tolist() == [0] * 20 and v2[20:40].tolist() == [1] * 20
    s = Fn.locality_sort(torch.tensor([[59, 0, 13, 12, 1]]), shape, block=1)
    assert s.tolist() == [[0, 12, 1, 13, 59]]


def test_slab_and_batch_partition():
    for total in (1, 7, 1000, 21559296):
        for w in (1, 2, 4, 8):
            spans = [sweep.slab_range(total, r, w) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
    assert distributed.split_batch(1 << 22, 3, 8) == (3 << 19, 1 << 19)


def test_upsampling_loader_matches_create_mgrid():
    dm = datamodules.MriDataModule(config=None, device="cpu")
    ld = dm.upsampling((4, 3, 5), batch_size=7)
    coords = torch.cat([x for x, _ in ld])
    assert torch.equal(coords, datamodules.create_mgrid((4, 3, 5)).reshape(-1, 3))


class _ToyModule(pl.LightningModule):
    def __init__(self):
        super().__init__()
        self.lin = nn.Linear(2, 1)
        self.steps = 0

    def forward(self, x):
        return self.lin(x)

    def training_step(self, batch, batch_idx):
        x, y = batch
        loss = nn.functional.mse_loss(self(x), y)
        self.log("train_loss", loss)
        self.steps += 1
        return loss

    def predict_step(self, batch, batch_idx):
        return self(batch[0])

    def configure_optimizers(self):
        return torch.optim.SGD(self.parameters(), lr=0.1)


def test_trainer_standin_fit_predict_logging(tmp_path):
    torch.manual_seed(0)
    x = torch.rand(64, 2)
    y = (x @ torch.tensor([[2.0], [-1.0]])) + 0.5
    loader = datamodules.DeviceBatchLoader(x, y, 16, shuffle=True, device="cpu")
    model = _ToyModule()
    tr = pl.Trainer(accelerator="cpu", max_epochs=30, precision=32, default_root_dir=str(tmp_path),
                    accumulate_grad_batches=None)
    tr.fit(model, loader)
    assert model.steps == 30 * 4 and tr.global_step == 120
    start = float(nn.functional.mse_loss(_ToyModule()(x), y))
    pred = torch.concat(tr.predict(model, datamodules.DeviceBatchLoader(x, y, 16, device="cpu")))
    assert pred.shape == (64, 1) and float(nn.functional.mse_loss(pred, y)) < 0.2 * start
    assert os.path.basename(model.logger.log_dir) == "version_0" and model.logger.version == 0
    assert os.path.isfile(os.path.join(model.logger.log_dir, "metrics.csv"))
    ck = [f for f in os.listdir(os.path.join(model.logger.log_dir, "checkpoints"))]
    assert ck and ck[0].endswith(".ckpt")
    again = _ToyModule.load_from_checkpoint(os.path.join(model.logger.log_dir, "checkpoints", ck[0]))
    assert torch.equal(again.lin.weight, model.lin.weight)
    # accumulate_grad_batches: fewer optimiser steps; dict schedule like launcher.py:159
    m2 = _ToyModule()
    tr2 = pl.Trainer(accelerator="cpu", max_epochs=2, precision=32, default_root_dir=str(tmp_path),
                     accumulate_grad_batches={0: 2})
    tr2.fit(m2, loader)
    assert tr2.global_step == 4 and m2.logger.version == 1
    with pytest.raises(NotImplementedError):
        pl.Trainer(precision=16)


def test_metrics_have_no_host_fallback():
    """metrics.* are CUDA kernels (csrc/metrics.cu); the numpy restatement is the oracle's (oracle/sweep.py)."""
    from mri_interpolation_b200 import MriB200Error
    if torch.cuda.is_available():
        pytest.skip("CUDA present: covered by tests/test_gpu_sweep.py::test_gpu_metrics_match_the_numpy_oracle")
    a = np.zeros((8, 8, 2), np.float32)
    for fn in (metrics.mean_squared_error, metrics.peak_signal_noise_ratio, metrics.structural_similarity):
        with pytest.raises(MriB200Error):
            fn(a, a)
    with pytest.raises(MriB200Error):
        metrics.linear_time_baseline(a)


def test_oracle_linear_time_baseline_is_the_reference_recipe():
    """interp.py:35-50 on a tiny volume, written out longhand: frames ::2 kept, sample f at continuous index f/2."""
    from oracle import sweep as osweep
    rng = np.random.default_rng(1)
    data = rng.random((3, 4, 7)).astype(np.float32)
    got = osweep.linear_time_baseline(data)
    kept = data[..., ::2]
    for f in range(7):
        pos = min(f / 2.0, kept.shape[-1] - 1)
        lo = int(np.floor(pos)); hi = min(lo + 1, kept.shape[-1] - 1); a = np.float32(pos - lo)
        np.testing.assert_array_equal(got[..., f], kept[..., lo] * (1 - a) + kept[..., hi] * a)
    np.testing.assert_array_equal(got[..., ::2], data[..., ::2])  # kept frames are reproduced exactly


def test_fold_batchnorm_equals_linear_then_eval_batchnorm():
    """sweep.fold_batchnorm: Linear -> BatchNorm1d(eval) collapses into one Linear (the shipped HashMLP decoder blocks,
    models.py:718-739, take the fused sweep kernel this way)."""
    import torch
    from torch import nn
    from mri_interpolation_b200.sweep import fold_batchnorm
    torch.manual_seed(3)
    for affine in (True, False):
        lin, bn = nn.Linear(32, 64), nn.BatchNorm1d(64, affine=affine)
        with torch.no_grad():
            bn.running_mean.copy_(torch.randn(64) * 0.3)
            bn.running_var.copy_(torch.rand(64) + 0.2)
            if affine:
                bn.weight.copy_(torch.randn(64))
                bn.bias.copy_(torch.randn(64))
        bn.eval()
        x = torch.randn(100, 32)
        w, b = fold_batchnorm(lin, bn)
        torch.testing.assert_close(torch.nn.functional.linear(x, w, b), bn(lin(x)), rtol=1e-5, atol=1e-5)
    w, b = fold_batchnorm(lin, None)
    assert w is lin.weight and b is lin.bias


def test_tensor_core_width_padding_rule():
    """siren_fused.padded_width: multiples of 128 run as they are, other widths >= 160 are zero-padded to the next
    multiple (the notebook's 352 -> 384), narrower ones stay on the fp32 path (0)."""
    from mri_interpolation_b200 import siren_fused
    assert [siren_fused.padded_width(h) for h in (128, 256, 1024)] == [128, 256, 1024]
    assert siren_fused.padded_width(352) == 384 and siren_fused.padded_width(200) == 256 and siren_fused.padded_width(160) == 256
    assert siren_fused.padded_width(64) == 0 and siren_fused.padded_width(159) == 0


def test_fused_step_declines_what_it_cannot_serve_on_cpu():
    """HashMLP.fused_training_step returns None (caller falls back to training_step + backward) for CPU batches; the
    decoder plan accepts exactly 2 Linear+activation blocks with one output."""
    torch.manual_seed(0)
    net = models.HashMLP(dim_in=3, n_levels=4, n_features_per_level=2, log2_hashmap_size=8, base_resolution=4,
                         finest_resolution=16, dim_hidden=64, dim_out=1, n_layers=2, batch_norm=False)
    assert net.fused_training_step((torch.rand(10, 3), torch.rand(10, 1)), 0) is None
    plan = net._fused_decoder_plan()
    assert plan is not None and plan[0].out_features == 64 and plan[1].out_features == 1
    bn = models.HashMLP(dim_in=3, n_levels=4, n_features_per_level=2, log2_hashmap_size=8, base_resolution=4,
                        finest_resolution=16, dim_hidden=64, dim_out=1, n_layers=2, batch_norm=True)
    assert bn._fused_decoder_plan() is None  # BatchNorm blocks: no one-kernel path
    three = models.HashMLP(dim_in=3, n_levels=4, n_features_per_level=2, log2_hashmap_size=8, base_resolution=4,
                           finest_resolution=16, dim_hidden=64, dim_out=1, n_layers=3, batch_norm=False)
    assert three._fused_decoder_plan() is None


def test_spectral_norm_variant_wraps_every_linear_like_the_legacy_recipe():
    """legacy_code/hash_experimentation.py:213-246: spectral_norm(Linear, n_power_iterations=4) in every block, Adam
    with weight_decay; state_dict carries the parametrization's original weight and the u / v vectors."""
    torch.manual_seed(0)
    net = models.HashMLP(dim_in=3, n_levels=4, n_features_per_level=2, log2_hashmap_size=8, base_resolution=4,
                         finest_resolution=16, dim_hidden=16, dim_out=1, n_layers=2, spectral_norm=True, weight_decay=1e-5)
    keys = list(net.decoder.state_dict())
    assert "0.0.parametrizations.weight.original" in keys and "0.0.parametrizations.weight.0._u" in keys
    lin = net.decoder[0][0]
    w = lin.weight
    assert abs(float(torch.linalg.matrix_norm(w.detach(), ord=2)) - 1.0) < 0.2  # W / sigma after 4 power iterations
    assert net.weight_decay == 1e-5 and isinstance(net.decoder[0][1], torch.nn.BatchNorm1d)


def test_multiply_based_modulo_formula_is_exact():
    """The arithmetic of csrc/common.cuh::exact_mod, replayed with numpy uint64: r = h - mulhi(h, floor(2^32 / rows)) * rows,
    one conditional subtraction - equal to h % rows for every 32-bit h and every non-power-of-two row count up to 2^31
    (the GPU tests assert the same on the kernels' own output)."""
    rng = np.random.default_rng(0)
    rows_list = [3, 5, 9, 81, 1000, 4489, 234256, 2247001, 3307949, (1 << 31) - 1, (1 << 31) - 19, (1 << 30) + 7]
    rows_list += [int(v) for v in rng.integers(3, 1 << 31, 200)]
    edge = np.array([0, 1, 2, (1 << 32) - 1, (1 << 32) - 2, 1 << 31, (1 << 31) - 1], dtype=np.uint64)
    for rows in rows_list:
        if rows & (rows - 1) == 0:
            continue
        magic = np.uint64((1 << 32) // rows)
        h = np.concatenate([edge, rng.integers(0, 1 << 32, 4000, dtype=np.uint64),
                            (np.arange(-3, 4, dtype=np.int64) + rows * rng.integers(1, max(2, (1 << 32) // rows), 50)[:, None]).reshape(-1).astype(np.uint64) & np.uint64(0xFFFFFFFF)])
        q = (h * magic) >> np.uint64(32)
        r = (h - q * np.uint64(rows)) & np.uint64(0xFFFFFFFF)
        assert (r < 2 * rows).all()
        r = np.where(r >= rows, r - np.uint64(rows), r)
        assert (r == h % np.uint64(rows)).all(), rows
