"""CPU: the oracle restatement against the committed golden vectors (generated from the reference's
own classes by oracle/make_golden.py).  Bit-exact for indices and forward outputs."""
import numpy as np
import pytest
import torch

from conftest import HASH_CASES, SIREN_CASES, load_golden, oracle_levels
from oracle import hashgrid, networks, sweep


@pytest.fixture(autouse=True)
def _single_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)  # golden vectors were made with 1 thread (reduction order)
    yield
    torch.set_num_threads(n)


def test_geometry_known_answers():
    g = load_golden("geometry.npz")
    lv = hashgrid.geometry_isotropic(4, 16, 19, 16, 2489)
    assert [l.resolution[0] for l in lv] == g["g4_res"].tolist()
    assert [l.rows for l in lv] == g["g4_rows"].tolist()
    assert g["g4_res"].tolist() == [16, 22, 31, 43, 61, 86, 120, 168, 236, 330, 462, 647, 907, 1269, 1777, 2489]
    assert sum(l.rows for l in lv) * 2 == 15279648  # 61.1 MB fp32
    lv2 = hashgrid.geometry_anisotropic(3, 8, 23, (64, 64, 5), (512, 512, 15))
    assert np.array_equal(np.asarray([l.resolution for l in lv2]), g["nbv2_res"])
    assert sum(l.rows for l in lv2) * 2 == 6009032  # nb:2792 "6.0 M"
    d3 = hashgrid.geometry_isotropic(3, 16, 19, 16, 512)
    assert [l.resolution[0] for l in d3] == [16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512]
    h, _ = hashgrid.corners(torch.ones(1, 4), lv[0])
    assert h[0].tolist() == g["g4_ones_hash_l0"].tolist()


@pytest.mark.parametrize("case", HASH_CASES)
def test_hashgrid_matches_reference_vectors(case):
    fx = load_golden(f"hashgrid_{case}.npz")
    levels = oracle_levels(fx)
    aniso = bool(fx["anisotropic"])
    x = torch.from_numpy(fx["x"])
    tables = [torch.from_numpy(fx[f"table{i}"]) for i in range(len(levels))]
    out = hashgrid.encode(x, tables, levels, aniso)
    assert torch.equal(out, torch.from_numpy(fx["out"]))  # bit-exact
    for li, lv in enumerate(levels):
        h, w = hashgrid.corners(x, lv, aniso)
        assert np.array_equal(h.numpy().astype(np.uint32), fx["hashes"][:, li])
        assert np.array_equal(w.numpy(), fx["weights"][:, li])
        assert int(h.max()) < lv.rows
    grads = hashgrid.table_gradients(x, torch.from_numpy(fx["grad_out"]), levels, int(fx["n_features"]), aniso)
    for li, g in enumerate(grads):
        np.testing.assert_allclose(g.numpy(), fx[f"grad{li}"], rtol=1e-5, atol=1e-6)


def test_hashgrid_backward_is_autograd_of_forward():
    fx = load_golden("hashgrid_v1_d3.npz")
    levels = oracle_levels(fx)
    x = torch.from_numpy(fx["x"])
    tables = [torch.from_numpy(fx[f"table{i}"]).clone().requires_grad_() for i in range(len(levels))]
    g = torch.from_numpy(fx["grad_out"])
    hashgrid.encode(x, tables, levels).backward(g)
    explicit = hashgrid.table_gradients(x, g, levels, int(fx["n_features"]))
    for t, e in zip(tables, explicit):
        torch.testing.assert_close(t.grad, e, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", SIREN_CASES)
def test_siren_matches_reference_vectors(case):
    fx = load_golden(f"siren_{case}.npz")
    kw = {k[3:]: fx[k].item() for k in fx if k.startswith("kw_")}
    torch.manual_seed(1337)
    params, w0s = networks.siren_init(**kw)
    for k, v in params.items():
        assert np.array_equal(v.numpy(), fx[f"param:{k}"]), k  # seed-1337 init is bit-identical
    x, y = torch.from_numpy(fx["x"]), torch.from_numpy(fx["y"])
    pred = networks.siren_forward(x, params, w0s)
    assert torch.equal(pred, torch.from_numpy(fx["pred"]))
    grads, _ = networks.siren_backward(x, params, w0s, networks.mse_grad(y, pred))
    for k in params:
        np.testing.assert_allclose(grads[k].numpy(), fx[f"grad:{k}"], rtol=2e-4, atol=1e-7)
    assert abs(float(networks.mse(y, pred)) - float(fx["loss"])) < 1e-7


def test_hashmlp_init_and_forward_match_reference():
    fx = load_golden("hashmlp_small.npz")
    kw = {k[3:]: fx[k].item() for k in fx if k.startswith("kw_")}
    torch.manual_seed(1337)
    params, levels = networks.hashmlp_init(**kw)
    # decoder / dead BaseMLP params keep their seed-1337 values in the fixture; tables were overwritten
    for k, v in params.items():
        if "embedding" not in k:
            assert np.array_equal(v.numpy(), fx[f"param:{k}"]), k
        else:
            params[k] = torch.from_numpy(fx[f"param:{k}"])
    pred = networks.hashmlp_forward(torch.from_numpy(fx["x"]), params, levels, kw["n_layers"], False)
    assert torch.equal(pred, torch.from_numpy(fx["pred"]))


@pytest.mark.parametrize("name", ["default", "tcnn_like", "l2"])
def test_adam_matches_torch_optim(name):
    fx = load_golden(f"adam_{name}.npz")
    p = torch.from_numpy(fx["p0"]).clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step, g in enumerate(fx["grads"], start=1):
        networks.adam_step(p, torch.from_numpy(g), m, v, step, float(fx["lr"]), float(fx["beta1"]), float(fx["beta2"]),
                           float(fx["eps"]), float(fx["weight_decay"]))
        np.testing.assert_allclose(p.numpy(), fx["params"][step - 1], rtol=1e-6, atol=1e-9)


def test_sweep_coords_match_reference_recipe():
    fx = load_golden("sweep_coords.npz")
    c = sweep.grid_coords(tuple(fx["shape"]))
    assert np.array_equal(c.numpy(), fx["coords"])
    assert np.array_equal(sweep.axis_values(29).numpy(), fx["lin29"])
    assert np.array_equal(sweep.axis_values(57, True).numpy(), fx["lin57m"])


def test_metrics_sanity():
    rng = np.random.default_rng(0)
    a = rng.random((32, 32, 3)).astype(np.float32)
    assert sweep.psnr(a, a + 0.1) == pytest.approx(20.0, abs=1e-4)
    assert sweep.ssim_slices(a, a) == pytest.approx(1.0, abs=1e-9)
    assert sweep.ssim_slices(a, rng.random((32, 32, 3)).astype(np.float32)) < 0.2
    frames = np.stack([np.full((2, 2), float(t)) for t in range(4)], -1)
    up = sweep.linear_time_interpolation(frames, 7)
    assert np.allclose(up[0, 0], np.linspace(0, 3, 7))


def test_ssim_closed_form_known_answers():
    """skimage is absent, so the SSIM restatement is pinned against cases with a closed form (skimage's definition,
    K1 = 0.01, K2 = 0.03, data_range 1): identical images -> 1; two constant images A, B -> (2AB + c1) / (A^2 + B^2 + c1)
    (zero variances and covariance); b = -a + 1 around mean 1/2 with small variance -> the luminance and contrast terms
    separately computable."""
    from oracle import sweep as osweep
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    rng = np.random.default_rng(0)
    a = rng.random((20, 18)).astype(np.float32)
    assert osweep.ssim2d(a, a) == pytest.approx(1.0, abs=1e-12)
    A, B = 0.3, 0.7
    got = osweep.ssim2d(np.full((15, 15), A, np.float32), np.full((15, 15), B, np.float32))
    assert got == pytest.approx((2 * A * B + c1) / (A * A + B * B + c1), rel=1e-6)
    # anti-correlated pair with equal means and variances: covariance = -variance inside every window
    x = (rng.random((16, 16)) - 0.5).astype(np.float64) * 0.2
    a2, b2 = (0.5 + x).astype(np.float32), (0.5 - x).astype(np.float32)
    from scipy.ndimage import uniform_filter
    win, n = 7, 49
    m = uniform_filter(a2.astype(np.float64), win)
    mb = uniform_filter(b2.astype(np.float64), win)
    va = (uniform_filter(a2.astype(np.float64) ** 2, win) - m * m) * n / (n - 1)
    vb = (uniform_filter(b2.astype(np.float64) ** 2, win) - mb * mb) * n / (n - 1)
    cov = (uniform_filter(a2.astype(np.float64) * b2.astype(np.float64), win) - m * mb) * n / (n - 1)
    assert np.allclose(cov[3:-3, 3:-3], -np.sqrt(va * vb)[3:-3, 3:-3], rtol=1e-4, atol=1e-9)
    want = (((2 * m * mb + c1) * (2 * cov + c2)) / ((m * m + mb * mb + c1) * (va + vb + c2)))[3:-3, 3:-3].mean()
    assert osweep.ssim2d(a2, b2) == pytest.approx(want, rel=1e-9)
    assert osweep.ssim2d(a2, b2) < 0.0  # strongly anti-correlated structure
