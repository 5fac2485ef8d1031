"""GPU parity: dense/SIREN/decoder kernels, MSE, Adam and whole-model training steps against the
oracle (torch fp32 on CPU) and the reference-generated golden vectors."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import SIREN_CASES, load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-3
DEV = "cuda"


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


ACTS = {"identity": lambda p, w0: p, "sine": lambda p, w0: torch.sin(w0 * p), "gelu": lambda p, w0: F.gelu(p),
        "relu": lambda p, w0: F.relu(p)}


@pytest.mark.parametrize("n,k,m", [(300, 3, 64), (1000, 64, 1), (257, 32, 64), (129, 4, 256), (513, 256, 256),
                                   (64, 48, 3), (1, 5, 7), (2050, 352, 352), (77, 64, 16), (90, 17, 33)])
@pytest.mark.parametrize("act", ["identity", "sine", "gelu", "relu"])
def test_dense_forward_backward_vs_torch(n, k, m, act):
    from mri_interpolation_b200 import functional as Fn
    gen = torch.Generator().manual_seed(n * 7 + k * 3 + m)
    x = (torch.rand(n, k, generator=gen) * 2 - 1).requires_grad_()
    w = ((torch.rand(m, k, generator=gen) * 2 - 1) / k ** 0.5).requires_grad_()
    b = ((torch.rand(m, generator=gen) * 2 - 1) * 0.1).requires_grad_()
    w0 = 30.0 if act == "sine" else 1.0
    y_ref = ACTS[act](F.linear(x, w, b), w0)
    gy = torch.randn(n, m, generator=gen)
    y_ref.backward(gy)
    xd, wd, bd = (t.detach().to(DEV).requires_grad_() for t in (x, w, b))
    y = Fn.dense(xd, wd, bd, act, w0)
    tol = 2e-4 if act == "sine" else 1e-5  # sin(30 * pre): argument rounding is amplified by w0
    assert rel_err(y, y_ref) < tol
    y.backward(gy.to(DEV))
    assert rel_err(xd.grad, x.grad) < max(tol, 1e-4)
    assert rel_err(wd.grad, w.grad) < max(tol, 1e-4)
    assert rel_err(bd.grad, b.grad) < max(tol, 1e-4)


def test_dense_no_bias_and_strided_input():
    from mri_interpolation_b200 import functional as Fn
    gen = torch.Generator().manual_seed(1)
    big = torch.rand(200, 40, generator=gen)
    w = torch.rand(24, 32, generator=gen) - 0.5
    x = big[:, 4:36]  # row stride 40
    ref = F.gelu(F.linear(x, w))
    out = Fn.dense(big.to(DEV)[:, 4:36], w.to(DEV), None, "gelu")
    assert rel_err(out, ref) < 1e-5


@pytest.mark.parametrize("case", SIREN_CASES)
def test_sirennet_matches_reference_vectors(case):
    from mri_interpolation_b200 import models
    fx = load_golden(f"siren_{case}.npz")
    kw = {k[3:]: fx[k].item() for k in fx if k.startswith("kw_")}
    torch.manual_seed(1337)
    net = models.SirenNet(**kw).to(DEV)
    x, y = torch.from_numpy(fx["x"]).to(DEV), torch.from_numpy(fx["y"]).to(DEV)
    pred = net(x)
    torch.testing.assert_close(pred.cpu(), torch.from_numpy(fx["pred"]), rtol=RTOL, atol=2e-5)
    loss = net.training_step((x, y), 0)
    assert abs(float(loss) - float(fx["loss"])) < 1e-5
    loss.backward()
    for name, p in net.named_parameters():
        if f"grad:{name}" in fx:
            assert rel_err(p.grad, torch.from_numpy(fx[f"grad:{name}"])) < RTOL, name


def test_hashmlp_notebook_variant_matches_reference_vectors():
    from mri_interpolation_b200 import models
    fx = load_golden("hashmlp_small.npz")
    kw = {k[3:]: fx[k].item() for k in fx if k.startswith("kw_")}
    net = models.HashMLP(**kw, batch_norm=False)
    net.load_state_dict({k[6:]: torch.from_numpy(v) for k, v in fx.items() if k.startswith("param:")}, strict=False)
    net = net.to(DEV)
    x, y = torch.from_numpy(fx["x"]).to(DEV), torch.from_numpy(fx["y"]).to(DEV)
    pred = net(x)
    torch.testing.assert_close(pred.cpu(), torch.from_numpy(fx["pred"]), rtol=RTOL, atol=1e-6)
    loss = net.training_step((x, y), 0)
    assert abs(float(loss) - float(fx["loss"])) < 1e-6
    loss.backward()
    for name, p in net.named_parameters():
        if f"grad:{name}" in fx and not name.startswith("layers."):
            assert rel_err(p.grad, torch.from_numpy(fx[f"grad:{name}"])) < RTOL, name
    # predict_step keeps the latents like the reference (models.py:746-751)
    out = net.predict_step((x, y), 0)
    assert torch.equal(out, pred.detach()) and len(net.get_latents()) == 1
    torch.testing.assert_close(net.get_latents()[0].cpu(), torch.from_numpy(fx["latents"]), rtol=RTOL, atol=1e-6)


def test_hashmlp_batchnorm_variant_matches_reference_vectors():
    from mri_interpolation_b200 import models
    fx = load_golden("hashmlp_bn_small.npz")
    kw = {k[3:]: fx[k].item() for k in fx if k.startswith("kw_")}
    net = models.HashMLP(**kw)  # shipped decoder: Linear -> BatchNorm1d -> GELU -> Dropout
    net.load_state_dict({k[6:]: torch.from_numpy(v) for k, v in fx.items() if k.startswith("param:")}, strict=False)
    net = net.to(DEV).train()
    x, y = torch.from_numpy(fx["x"]).to(DEV), torch.from_numpy(fx["y"]).to(DEV)
    import copy
    pred = copy.deepcopy(net)(x)  # a second forward on `net` itself would update the running statistics twice
    torch.testing.assert_close(pred.cpu(), torch.from_numpy(fx["pred"]), rtol=RTOL, atol=1e-5)
    net.training_step((x, y), 0).backward()
    scale = max(float(np.abs(fx[f"grad:{n}"]).max()) for n, _ in net.named_parameters() if f"grad:{n}" in fx)
    for name, p in net.named_parameters():
        if f"grad:{name}" in fx and not name.startswith("layers."):
            ref = torch.from_numpy(fx[f"grad:{name}"])
            if ".0.bias" in name:
                # a Linear bias feeding BatchNorm has a mathematically zero gradient: both sides are rounding noise
                assert float(p.grad.abs().max()) < 1e-4 * scale and float(ref.abs().max()) < 1e-4 * scale, name
            else:
                assert rel_err(p.grad, ref) < 5e-3, name
    for k, v in fx.items():
        if k.startswith("buffer:") and "running" in k:
            torch.testing.assert_close(dict(net.named_buffers())[k[7:]].cpu(), torch.from_numpy(v), rtol=1e-4, atol=1e-6)


def test_mse_loss_and_grad():
    from mri_interpolation_b200 import functional as Fn
    gen = torch.Generator().manual_seed(2)
    for shape in [(1, 1), (1000, 1), (333, 3), (70001, 1)]:
        y = torch.rand(shape, generator=gen)
        p = torch.rand(shape, generator=gen).requires_grad_()
        ref = F.mse_loss(y, p)
        ref.backward()
        pd = p.detach().to(DEV).requires_grad_()
        loss = Fn.mse_loss(y.to(DEV), pd)
        assert abs(float(loss) - float(ref)) < 1e-6 * max(1.0, float(ref))
        loss.backward()
        assert rel_err(pd.grad, p.grad) < 1e-6


@pytest.mark.parametrize("name", ["default", "tcnn_like", "l2"])
def test_fused_adam_matches_torch_trajectory(name):
    from mri_interpolation_b200.optim import FusedAdam
    fx = load_golden(f"adam_{name}.npz")
    p = torch.nn.Parameter(torch.from_numpy(fx["p0"]).to(DEV))
    opt = FusedAdam([p], lr=float(fx["lr"]), betas=(float(fx["beta1"]), float(fx["beta2"])), eps=float(fx["eps"]),
                    weight_decay=float(fx["weight_decay"]))
    for step, g in enumerate(fx["grads"]):
        p.grad.copy_(torch.from_numpy(g))
        opt.step()
        # SURVEY 7.5 gate: <= 1e-6 relative drift per step (torch's CPU kernels round a few ulps differently)
        np.testing.assert_allclose(p.detach().cpu().numpy(), fx["params"][step], rtol=2e-6 * (step + 1), atol=2e-9)
        assert float(p.grad.abs().max()) == 0.0  # gradient cleared in the same pass
        opt.zero_grad()


def test_fused_adam_odd_sizes_and_arena_views():
    from mri_interpolation_b200.optim import FusedAdam
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in [(3,), (5, 7), (1,), (1026,)]]
    ref = [torch.nn.Parameter(p.detach().cpu().clone()) for p in ps]
    opt = FusedAdam(ps, lr=1e-2)
    ropt = torch.optim.Adam(ref, lr=1e-2)
    assert all(p.data_ptr() >= opt.arena.data.data_ptr() for p in ps)
    for _ in range(3):
        for p, r in zip(ps, ref):
            g = torch.randn(r.shape)
            r.grad = g.clone()
            p.grad.copy_(g)
        opt.step()
        ropt.step()
    for p, r in zip(ps, ref):
        torch.testing.assert_close(p.detach().cpu(), r.detach(), rtol=2e-6, atol=1e-8)


def test_zero_grad_never_skips_a_dirty_arena():
    """zero_grad() may skip its memset only while the gradient arena is provably all zeros.
    (a) step(); backward(); zero_grad(); backward() must not double-accumulate;  (b) gradients that existed before the
    optimiser was built are copied into the arena and must be cleared by the first zero_grad();  (c) autograd's own
    accumulation (a parameter used by plain torch ops) also marks the arena dirty."""
    from mri_interpolation_b200 import models
    kw = dict(dim_in=3, n_levels=16, n_features_per_level=2, log2_hashmap_size=12, base_resolution=4,
              finest_resolution=64, dim_hidden=64, dim_out=1, n_layers=2)
    torch.manual_seed(3)
    x, y = torch.rand(777, 3, device=DEV), torch.rand(777, 1, device=DEV)

    def grads(net):
        return torch.cat([p.grad.reshape(-1) for n_, p in net.named_parameters() if not n_.startswith("layers.")]).clone()

    # (b) stale gradients before configure_optimizers()
    torch.manual_seed(1337)
    net = models.HashMLP(**kw, batch_norm=False, lr=5e-3).to(DEV)
    net.training_step((x, y), 0).backward()           # private .grad tensors, not in any arena yet
    opt = net.configure_optimizers()
    assert not opt._grads_clean
    opt.zero_grad()
    assert float(opt.arena.grad.abs().max()) == 0.0
    net.training_step((x, y), 0).backward()
    single = grads(net)
    # (a) step -> backward -> zero_grad -> backward
    opt.step()
    assert opt._grads_clean                              # the fused step cleared the arena
    net.training_step((x, y), 1).backward()
    assert not opt._grads_clean                          # ... and this backward dirtied it again
    opt.zero_grad()
    assert float(opt.arena.grad.abs().max()) == 0.0
    net.training_step((x, y), 2).backward()
    once = grads(net)
    opt.zero_grad()
    opt.zero_grad()                                      # free: nothing accumulated in between
    net.training_step((x, y), 2).backward()
    torch.testing.assert_close(grads(net), once, rtol=1e-5, atol=1e-9)
    assert float(single.abs().max()) > 0
    # (c) torch-side accumulation
    ps = [torch.nn.Parameter(torch.randn(50, device=DEV))]
    from mri_interpolation_b200.optim import FusedAdam
    o2 = FusedAdam(ps, lr=1e-2)
    assert o2._grads_clean
    (ps[0] * 2).sum().backward()
    assert not o2._grads_clean
    o2.zero_grad()
    assert float(ps[0].grad.abs().max()) == 0.0


def test_training_steps_track_the_oracle():
    """5 full steps (hash grid + decoder + MSE + Adam) from identical init and batches: parameters stay
    within summation-order noise of the oracle's torch.optim.Adam run."""
    from mri_interpolation_b200 import models
    from oracle import hashgrid, networks
    kw = dict(dim_in=3, n_levels=6, n_features_per_level=2, log2_hashmap_size=12, base_resolution=4,
              finest_resolution=64, dim_hidden=32, dim_out=1, n_layers=2)
    torch.manual_seed(1337)
    net = models.HashMLP(**kw, batch_norm=False, lr=5e-3)
    torch.manual_seed(1337)
    params, levels = networks.hashmlp_init(**kw)
    ref = {k: v.clone().requires_grad_() for k, v in params.items() if not k.startswith("layers.")}
    ropt = torch.optim.Adam(list(ref.values()), lr=5e-3)
    net = net.to(DEV)
    opt = net.configure_optimizers()
    gen = torch.Generator().manual_seed(3)
    for step in range(5):
        x, y = torch.rand(2000, 3, generator=gen), torch.rand(2000, 1, generator=gen)
        ropt.zero_grad()
        lr_ = F.mse_loss(y, networks.hashmlp_forward(x, ref, levels, 2, False))
        lr_.backward()
        ropt.step()
        opt.zero_grad()
        loss = net.training_step((x.to(DEV), y.to(DEV)), step)
        loss.backward()
        opt.step()
        assert abs(float(loss) - float(lr_)) < 1e-5
    sd = net.state_dict()
    for k, v in ref.items():
        assert rel_err(sd[k], v.detach()) < 1e-4, k


@pytest.mark.parametrize("k0,h,act", [(32, 64, "gelu"), (16, 64, "gelu"), (64, 32, "relu"), (16, 32, "relu"), (32, 32, "gelu")])
@pytest.mark.parametrize("n", [1, 127, 1000, 40000])
def test_fused_decoder2_vs_torch(k0, h, act, n):
    """Fused decoder (hidden layer on tensor cores, bf16x3 split precision) vs torch fp32 on CPU."""
    from mri_interpolation_b200 import functional as Fn
    gen = torch.Generator().manual_seed(k0 + h + n)
    torch.manual_seed(k0 * 1000 + h * 10 + n)
    f = {"gelu": F.gelu, "relu": F.relu}[act]
    enc = (torch.randn(n, k0, generator=gen)).requires_grad_()
    l1, l2 = torch.nn.Linear(k0, h), torch.nn.Linear(h, 1)
    pre1 = l1(enc)
    pre2 = l2(f(pre1))
    y_ref = f(pre2)
    gy = torch.randn(n, 1, generator=gen)
    if act == "relu":
        # rows sitting on a ReLU kink (|pre| ~ rounding error) have an ill-defined derivative: leave them out
        near_kink = (pre1.detach().abs().min(dim=1, keepdim=True).values < 1e-4) | (pre2.detach().abs() < 1e-4)
        gy = gy.masked_fill(near_kink, 0.0)
    y_ref.backward(gy)
    code = Fn.activation_code(act)
    encd = enc.detach().to(DEV).requires_grad_()
    ps = [p.detach().to(DEV).requires_grad_() for p in (l1.weight, l1.bias, l2.weight, l2.bias)]
    assert Fn.decoder2_supported(k0, h, code)
    y = Fn.Decoder2Fn.apply(encd, *ps, code, code)
    torch.testing.assert_close(y.cpu(), y_ref.detach(), rtol=1e-3, atol=5e-6)  # north_star bound: 1e-3 relative
    if n >= 100:
        assert rel_err(y, y_ref) < 5e-5
    y.backward(gy.to(DEV))
    tol = 1e-4 if n >= 100 else 1e-3
    assert rel_err(encd.grad, enc.grad) < tol
    for p, r in zip(ps, (l1.weight, l1.bias, l2.weight, l2.bias)):
        assert rel_err(p.grad, r.grad) < tol


def test_sharded_adam_kernel_emulated_two_ranks_on_one_gpu():
    """mri_adam_step_sharded with both 'ranks' living on this GPU (two gradient arenas, two parameter replicas,
    launched one after the other - no kernel waits on another): every replica must end up with the Adam update of the
    SUMMED gradient scaled by 1/world, exactly like the all-reduce path."""
    import ctypes
    from mri_interpolation_b200 import _lib
    n, world = 8192, 2
    gen = torch.Generator(device=DEV).manual_seed(3)
    p0 = torch.randn(n, device=DEV, generator=gen) * 0.1
    params = [p0.clone() for _ in range(world)]
    grads = [torch.randn(n, device=DEV, generator=gen) for _ in range(world)]
    ref_p, ref_g = p0.clone(), (grads[0] + grads[1])
    ref_m, ref_v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    _lib.call("mri_adam_step", ref_p.data_ptr(), ref_g.clone().data_ptr(), ref_m.data_ptr(), ref_v.data_ptr(), n, 1, 5e-3, 0.9, 0.999,
              1e-8, 0.0, 1.0 / world, 0, _lib.stream())
    shard = n // world
    pg = (ctypes.c_uint64 * world)(*[g.data_ptr() for g in grads])
    pp = (ctypes.c_uint64 * world)(*[p.data_ptr() for p in params])
    for r in range(world):
        m, v = torch.zeros(shard, device=DEV), torch.zeros(shard, device=DEV)
        _lib.call("mri_adam_step_sharded", pg, pp, 0, 0, world, r, m.data_ptr(), v.data_ptr(), r * shard, shard, 1, 5e-3, 0.9, 0.999,
                  1e-8, 0.0, 1.0 / world, 0, _lib.stream())
    for p in params:
        torch.testing.assert_close(p, ref_p, rtol=1e-6, atol=1e-9)
    assert torch.equal(params[0], params[1])


# geometries of the fused encoder+decoder kernels: all-T levels; coarse levels with non-power-of-two row counts
# (res^D < T: the `%` wrap); anisotropic V2 levels (max(res)^D rows)
FUSED_GEOMETRIES = [
    dict(dim_in=4, log2_hashmap_size=12, base_resolution=16, finest_resolution=200),
    dict(dim_in=4, log2_hashmap_size=16, base_resolution=8, finest_resolution=300),
    dict(dim_in=3, log2_hashmap_size=14, base_resolution=8, finest_resolution=300),
    dict(dim_in=3, log2_hashmap_size=13, base_resolution=(8, 6, 4), finest_resolution=(200, 150, 40)),
]


# (n_levels, hidden width) the one-kernel forward / backward cover besides the headline (16, 64): the notebook's L = 8
# anisotropic model (nb cell 37), 128-wide decoders (hash_config.json's n_neurons), L = 4 (half a k-tile, zero padded)
FUSED_SHAPES = [(16, 64)] + [(8, 64), (4, 64), (16, 128), (8, 128), (4, 128)]
FUSED_CASES = [(g, 16, 64) for g in range(4)] + [(g, L, H) for (L, H) in FUSED_SHAPES[1:] for g in (1, 3)]


@pytest.mark.parametrize("gi,n_levels,hidden", FUSED_CASES)
@pytest.mark.parametrize("act", ["gelu", "relu"])
def test_fused_hashdecoder_backward_matches_unfused_and_oracle(gi, n_levels, hidden, act):
    """F = 2 models (L in {4, 8, 16}, hidden in {64, 128}): encoder+decoder forward and backward in one kernel each vs the
    separate-kernel path vs the oracle."""
    import copy
    from mri_interpolation_b200 import models
    from oracle import networks
    if act == "relu" and (n_levels, hidden) == (16, 64) and gi not in (1, 2):
        pytest.skip("ReLU on the headline geometry: two geometries are enough")
    geo = FUSED_GEOMETRIES[gi]
    dim = geo["dim_in"]
    kw = dict(n_levels=n_levels, n_features_per_level=2, dim_hidden=hidden, dim_out=1, n_layers=2, **geo)
    torch.manual_seed(1337)
    net = models.HashMLP(**kw, batch_norm=False, activation=torch.nn.ReLU if act == "relu" else torch.nn.GELU)
    oact = F.relu if act == "relu" else F.gelu
    gen = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for lv in net.encoder.levels:
            lv.embedding.weight.copy_(torch.randn(lv.embedding.weight.shape, generator=gen) * 0.2)
    params = {k: v.detach().clone().requires_grad_() for k, v in net.state_dict().items() if not k.startswith("layers.")}
    levels = networks.hashgrid.geometry(dim, n_levels, geo["log2_hashmap_size"], geo["base_resolution"], geo["finest_resolution"])
    aniso = not isinstance(geo["base_resolution"], int)
    if geo["log2_hashmap_size"] != 12:
        assert any(lv.rows & (lv.rows - 1) for lv in levels)  # the non-power-of-two wrap is exercised
    for n in (1, 33, 5000):
        x, y = torch.rand(n, dim, generator=gen), torch.rand(n, 1, generator=gen)
        if act == "relu":
            # a pre-activation within rounding distance of the ReLU kink opens its gate in one summation order and not
            # in the other (fp32 vs split bf16): such a sample moves 16 x 2^D table rows by its whole contribution, in
            # ANY two implementations.  The comparison runs on samples whose gates are unambiguous.
            with torch.no_grad():
                tabs = [params[f"encoder.levels.{i}.embedding.weight"] for i in range(n_levels)]
                pre1 = F.linear(networks.hashgrid.encode(x, tabs, levels, aniso), params["decoder.0.0.weight"], params["decoder.0.0.bias"])
                pre2 = F.linear(F.relu(pre1), params["decoder.1.0.weight"], params["decoder.1.0.bias"])  # the output has a ReLU too
            keep = (pre1.abs().min(dim=1).values > 1e-4) & (pre2.abs().min(dim=1).values > 1e-4)
            x, y = x[keep], y[keep]
            n = x.shape[0]
            if n == 0:
                continue
        for p in params.values():
            p.grad = None
        F.mse_loss(y, networks.hashmlp_forward(x, params, levels, 2, aniso, oact)).backward()
        fused, plain = copy.deepcopy(net).to(DEV), copy.deepcopy(net).to(DEV)
        plain.fuse_backward = False
        assert type(fused(x.to(DEV)).grad_fn).__name__.startswith("HashDecoderFn")
        assert not type(plain(x.to(DEV)).grad_fn).__name__.startswith("HashDecoderFn")
        lf = fused.training_step((x.to(DEV), y.to(DEV)), 0)
        lf.backward()
        lp = plain.training_step((x.to(DEV), y.to(DEV)), 0)
        lp.backward()
        assert abs(float(lf) - float(lp)) < 1e-6
        for (name, pf), (_, pp) in zip(fused.named_parameters(), plain.named_parameters()):
            if name.startswith("layers."):
                continue
            ref = params[name].grad
            scale = float(ref.abs().max()) + 1e-12
            tol = 2e-3 * scale
            assert float((pf.grad.cpu() - ref).abs().max()) < tol, (n, name)
            assert float((pf.grad - pp.grad).abs().max()) < tol, (n, name)
            if n >= 33:
                assert rel_err(pf.grad, ref) < 1e-3, (n, name)


@pytest.mark.parametrize("geo", FUSED_GEOMETRIES)
def test_fused_kernels_address_exactly_the_reference_rows(geo):
    """Index exactness of the two kernels the bench times: (forward) the encoding written by hashdecoder_mma_fwd_kernel is
    bit-identical to the stand-alone gather, whose addressed rows are asserted against the oracle's hashes; (backward)
    the support of the table gradient left by hashdecoder_mma_bwd_kernel is exactly the oracle's set of hashed rows
    with a non-zero corner weight."""
    from mri_interpolation_b200 import models
    from oracle import networks
    dim = geo["dim_in"]
    kw = dict(n_levels=16, n_features_per_level=2, dim_hidden=64, dim_out=1, n_layers=2, **geo)
    torch.manual_seed(1337)
    net = models.HashMLP(**kw, batch_norm=False)
    gen = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for lv in net.encoder.levels:
            lv.embedding.weight.copy_(torch.randn(lv.embedding.weight.shape, generator=gen) * 0.2)
    levels = networks.hashgrid.geometry(dim, 16, geo["log2_hashmap_size"], geo["base_resolution"], geo["finest_resolution"])
    aniso = not isinstance(geo["base_resolution"], int)
    net = net.to(DEV)
    n = 777
    x, y = torch.rand(n, dim, generator=gen), torch.rand(n, 1, generator=gen)
    x[:4] = 0.0
    x[4:8] = 1.0
    pred = net(x.to(DEV))
    assert type(pred.grad_fn).__name__.startswith("HashDecoderFn")
    enc_rows, rows = net.encoder.gathered_rows(x.to(DEV))
    enc_fused = pred.grad_fn.saved_tensors[1]  # the (n, 32) encoding the fused forward kernel wrote for its backward
    assert torch.equal(enc_fused, enc_rows)
    F.mse_loss(y.to(DEV), pred).backward()
    for li, lv in enumerate(levels):
        h, w = networks.hashgrid.corners(x, lv, aniso)
        assert torch.equal(rows[:, li].cpu(), h), li
        want = torch.unique(h[w != 0])
        got = torch.nonzero(net.encoder.levels[li].embedding.weight.grad.abs().sum(1)).flatten().cpu()
        assert torch.equal(got, want), f"level {li}: fused scatter touched other rows"


@pytest.mark.parametrize("geo", [FUSED_GEOMETRIES[1], FUSED_GEOMETRIES[2]])
@pytest.mark.parametrize("order", ["line", "line_descending", "dense_duplicates"])
def test_fused_backward_merges_axis0_runs_of_locality_ordered_batches(geo, order):
    """Batches ordered along axis 0 (functional.locality_sort) put samples of one axis-0 line in consecutive rows; the
    fused backward sums their coinciding corner updates in registers before the reduction (hash_device.cuh
    merge_line_runs).  Grid voxels of a small volume, dense enough that runs of equal / adjacent cells, line changes
    inside a tile, repeated voxels and a ragged tail all occur; gradients must equal the oracle's and be independent of
    the order of the batch."""
    import copy
    from mri_interpolation_b200 import functional as Fn, models
    from oracle import networks
    dim = geo["dim_in"]
    shape = (37, 5, 3, 2)[:dim]
    kw = dict(n_levels=16, n_features_per_level=2, dim_hidden=64, dim_out=1, n_layers=2, **geo)
    torch.manual_seed(1337)
    net = models.HashMLP(**kw, batch_norm=False)
    gen = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for lv in net.encoder.levels:
            lv.embedding.weight.copy_(torch.randn(lv.embedding.weight.shape, generator=gen) * 0.2)
    params = {k: v.detach().clone().requires_grad_() for k, v in net.state_dict().items() if not k.startswith("layers.")}
    levels = networks.hashgrid.geometry(dim, 16, geo["log2_hashmap_size"], geo["base_resolution"], geo["finest_resolution"])
    total = int(np.prod(shape))
    if order == "dense_duplicates":
        index = torch.randint(0, total, (3 * total + 5,), generator=gen)  # every voxel ~3 times: equal cells at every level
    else:
        index = torch.randperm(total, generator=gen)[: total - 7]
    index = Fn.locality_sort(index, shape, block=1)
    if order == "line_descending":
        index = index.flip(0)
    axes = [torch.linspace(0, 1, s) for s in shape]
    rem, cols = index.clone(), []
    for d in range(dim - 1, -1, -1):
        cols.append(axes[d][rem % shape[d]])
        rem = rem // shape[d]
    x = torch.stack(cols[::-1], dim=-1).contiguous()
    y = torch.rand(x.shape[0], 1, generator=gen)
    F.mse_loss(y, networks.hashmlp_forward(x, params, levels, 2, False)).backward()
    fused = copy.deepcopy(net).to(DEV)
    assert type(fused(x.to(DEV)).grad_fn).__name__.startswith("HashDecoderFn")
    fused.training_step((x.to(DEV), y.to(DEV)), 0).backward()
    shuffled = copy.deepcopy(net).to(DEV)
    perm = torch.randperm(x.shape[0], generator=gen)
    shuffled.training_step((x[perm].to(DEV), y[perm].to(DEV)), 0).backward()
    for (name, pf), (_, ps) in zip(fused.named_parameters(), shuffled.named_parameters()):
        if name.startswith("layers."):
            continue
        ref = params[name].grad
        assert rel_err(pf.grad, ref) < 1e-3, name
        assert rel_err(pf.grad, ps.grad) < 1e-4, name


@pytest.mark.parametrize("geo,act", [(FUSED_GEOMETRIES[0], "gelu"), (FUSED_GEOMETRIES[1], "relu"), (FUSED_GEOMETRIES[2], "relu"),
                                     (FUSED_GEOMETRIES[3], "gelu")])
def test_fused_hashdecoder_forward_matches_two_kernel_path_and_oracle(geo, act):
    """Encoder+decoder forward in ONE kernel (mri_hashdecoder_forward): the encoding it writes is bit-identical to
    mri_hashgrid_forward, the output equals the gather + decoder kernels and the oracle."""
    from mri_interpolation_b200 import _lib, functional as Fn, models
    from oracle import hashgrid, networks
    dim = geo["dim_in"]
    aniso = not isinstance(geo["base_resolution"], int)
    kw = dict(n_levels=16, n_features_per_level=2, dim_hidden=64, dim_out=1, n_layers=2, **geo)
    torch.manual_seed(7)
    net = models.HashMLP(**kw, batch_norm=False, activation=torch.nn.ReLU if act == "relu" else torch.nn.GELU)
    gen = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for lv in net.encoder.levels:
            lv.embedding.weight.copy_(torch.randn(lv.embedding.weight.shape, generator=gen) * 0.2)
    params = {k: v.detach().clone() for k, v in net.state_dict().items()}
    levels = hashgrid.geometry(dim, 16, geo["log2_hashmap_size"], geo["base_resolution"], geo["finest_resolution"])
    net = net.to(DEV)
    enc_mod = net.encoder
    l1, l2, a1, a2 = net._fused_decoder_plan()
    for n in (1, 15, 16, 4099):
        x = torch.rand(n, dim, generator=gen)
        x[0, 0] = 1.0  # the reference does not clamp at the upper edge
        xd = x.to(DEV)
        with torch.no_grad():
            enc_ref = enc_mod(xd)
            y_two = Fn.Decoder2Fn.apply(enc_ref, l1.weight, l1.bias, l2.weight, l2.bias, a1, a2)
        tables = enc_mod.tables()
        enc_mod._fwd_layout.refresh(tables, enc_mod._resolutions, enc_mod._rows)
        enc = torch.full((n, 32), float("nan"), device=DEV)
        y = torch.empty(n, 1, device=DEV)
        pre2 = torch.empty(n, device=DEV)
        _lib.call("mri_hashdecoder_forward", xd.data_ptr(), n, dim, enc_mod._fwd_layout.base, enc_mod._fwd_layout.levels, 16, 2,
                  32, 64, l1.weight.data_ptr(), l1.bias.data_ptr(), l2.weight.data_ptr(), l2.bias.data_ptr(), a1, a2,
                  enc.data_ptr(), y.data_ptr(), pre2.data_ptr(), _lib.stream())
        assert torch.equal(enc, enc_ref), n
        torch.testing.assert_close(y, y_two, rtol=1e-6, atol=1e-7)
        ref = networks.hashmlp_forward(x, params, levels, 2, aniso, F.relu if act == "relu" else F.gelu)
        assert float((y.cpu() - ref).abs().max()) < 1e-3 * (float(ref.abs().max()) + 1e-6), n
        # module path: training forward keeps enc for the backward, no-grad forward writes only y
        y_mod = net(xd)
        with torch.no_grad():
            y_eval = net(xd)
        assert torch.equal(y_mod.detach(), y) and torch.equal(y_eval, y)


def test_eager_pytorch_on_the_same_gpu_is_the_comparator_not_the_product():
    """the reference's PyTorch arithmetic (oracle port) run eagerly on the B200 next to the CUDA path, same
    G4 + 2x64 model and batch; reported (gpurun_out/eager_comparator.json when that directory exists), and the
    kernel path has to be well ahead of it."""
    import json
    import os
    from mri_interpolation_b200 import models
    from oracle import networks
    g4 = dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16, finest_resolution=2489)
    n = 1 << 17
    torch.manual_seed(1337)
    net = models.HashMLP(dim_in=4, dim_hidden=64, dim_out=1, n_layers=2, batch_norm=False, lr=5e-3, **g4).to(DEV)
    opt = net.configure_optimizers()
    params, levels = networks.hashmlp_init(dim_in=4, dim_hidden=64, dim_out=1, n_layers=2, **g4)
    params = {k: v.to(DEV).requires_grad_() for k, v in params.items() if not k.startswith("layers.")}
    eager_opt = torch.optim.Adam(list(params.values()), lr=5e-3)
    gen = torch.Generator(device=DEV).manual_seed(1)
    x, y = torch.rand(n, 4, device=DEV, generator=gen), torch.rand(n, 1, device=DEV, generator=gen)

    def ours():
        loss = net.training_step((x, y), 0)
        loss.backward()
        opt.step()
        opt.zero_grad()

    def eager():
        eager_opt.zero_grad()
        F.mse_loss(y, networks.hashmlp_forward(x, params, levels, 2, False)).backward()
        eager_opt.step()

    def timed(fn, reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    t_ours, t_eager = timed(ours, 20), timed(eager, 5)
    if os.path.isdir("gpurun_out"):
        with open("gpurun_out/eager_comparator.json", "w") as f:
            json.dump({"coords_per_step": n, "ms_per_step_cuda_path": t_ours, "ms_per_step_eager_pytorch_same_gpu": t_eager,
                       "coords_per_s_cuda_path": n / t_ours * 1e3, "coords_per_s_eager_pytorch_same_gpu": n / t_eager * 1e3}, f)
    assert t_ours * 3 < t_eager, (t_ours, t_eager)


@pytest.mark.parametrize("batch_norm", [False, True])
def test_hashmlp_spectral_norm_legacy_recipe_tracks_torch(batch_norm):
    """the legacy decoder (legacy_code/hash_experimentation.py:213-246) - spectral_norm(Linear, 4 power
    iterations) [-> BatchNorm1d] -> GELU, Adam with weight_decay 1e-5.  Three training steps on the B200 path against the
    same modules in plain PyTorch on the CPU (oracle hash encoding + torch.nn decoder + torch.optim.Adam), identical
    initial state including the power-iteration vectors."""
    import copy
    from mri_interpolation_b200 import models
    from oracle import hashgrid
    kw = dict(dim_in=3, n_levels=6, n_features_per_level=2, log2_hashmap_size=11, base_resolution=4, finest_resolution=64,
              dim_hidden=32, dim_out=1, n_layers=2)
    torch.manual_seed(5)
    net = models.HashMLP(**kw, batch_norm=batch_norm, spectral_norm=True, weight_decay=1e-5, lr=2e-3)
    with torch.no_grad():
        for lv in net.encoder.levels:
            lv.embedding.weight.mul_(2000.0)  # +-0.2: the decoder sees a signal
    levels = hashgrid.geometry(3, 6, 11, 4, 64)
    ref_tables = [lv.embedding.weight.detach().clone().requires_grad_() for lv in net.encoder.levels]
    ref_dec = copy.deepcopy(net.decoder)  # spectral-norm parametrised Linear (+ BatchNorm1d) blocks, same u / v buffers
    ref_opt = torch.optim.Adam(ref_tables + list(ref_dec.parameters()), lr=2e-3, weight_decay=1e-5)
    net = net.to(DEV)
    opt = net.configure_optimizers()
    assert opt.defaults["weight_decay"] == 1e-5
    gen = torch.Generator().manual_seed(2)
    for step in range(3):
        x, y = torch.rand(512, 3, generator=gen), torch.rand(512, 1, generator=gen)
        z = hashgrid.encode(x, ref_tables, levels)
        for block in ref_dec:
            z = block(z)
        lref = F.mse_loss(y, z)
        ref_opt.zero_grad()
        lref.backward()
        ref_opt.step()
        loss = net.training_step((x.to(DEV), y.to(DEV)), step)
        loss.backward()
        opt.step()
        opt.zero_grad()
        assert abs(float(loss) - float(lref)) < 1e-4 * max(1.0, abs(float(lref))), (step, float(loss), float(lref))
    for t, lv in zip(ref_tables, net.encoder.levels):
        assert rel_err(lv.embedding.weight, t) < 1e-3
    ref_sd, sd = ref_dec.state_dict(), net.decoder.state_dict()
    for k, v in ref_sd.items():
        if batch_norm and k.endswith(".0.bias"):
            continue  # a Linear bias in front of BatchNorm has zero true gradient: Adam turns its rounding noise into +-lr steps
        if v.dtype.is_floating_point and v.numel() > 1:
            # ... and those bias steps feed the running means, hence the wider bound with BatchNorm
            assert rel_err(sd[k], v) < (1e-2 if batch_norm else 2e-3), k


@pytest.mark.parametrize("gi", [0, 1, 2, 3])
@pytest.mark.parametrize("act", ["gelu", "relu"])
def test_one_kernel_training_step_matches_training_step_plus_backward_and_oracle(gi, act):
    """HashMLP.fused_training_step (mri_hashmlp_mse_step: gather + decoder + MSE + decoder backward + scatter in ONE
    kernel) against training_step + loss.backward() (two kernels + MSE kernel) and the oracle: same loss, same gradients;
    gradients accumulate over two calls like autograd's."""
    import copy
    from mri_interpolation_b200 import models
    from mri_interpolation_b200.pl_compat import training_step_and_backward
    from oracle import networks
    geo = FUSED_GEOMETRIES[gi]
    dim = geo["dim_in"]
    kw = dict(n_levels=16, n_features_per_level=2, dim_hidden=64, dim_out=1, n_layers=2, **geo)
    torch.manual_seed(1337)
    net = models.HashMLP(**kw, batch_norm=False, activation=torch.nn.ReLU if act == "relu" else torch.nn.GELU)
    oact = F.relu if act == "relu" else F.gelu
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for lv in net.encoder.levels:
            lv.embedding.weight.copy_(torch.randn(lv.embedding.weight.shape, generator=gen) * 0.2)
    params = {k: v.detach().clone().requires_grad_() for k, v in net.state_dict().items() if not k.startswith("layers.")}
    aniso = not isinstance(geo["base_resolution"], int)
    levels = networks.hashgrid.geometry(dim, 16, geo["log2_hashmap_size"], geo["base_resolution"], geo["finest_resolution"])
    for n in (1, 17, 4099):
        x, y = torch.rand(n, dim, generator=gen), torch.rand(n, 1, generator=gen)
        if act == "relu":  # unambiguous gates only (see the backward test above)
            with torch.no_grad():
                tabs = [params[f"encoder.levels.{i}.embedding.weight"] for i in range(16)]
                pre1 = F.linear(networks.hashgrid.encode(x, tabs, levels, aniso), params["decoder.0.0.weight"], params["decoder.0.0.bias"])
                pre2 = F.linear(F.relu(pre1), params["decoder.1.0.weight"], params["decoder.1.0.bias"])
            keep = (pre1.abs().min(dim=1).values > 1e-4) & (pre2.abs().min(dim=1).values > 1e-4)
            x, y = x[keep], y[keep]
            n = x.shape[0]
            if n == 0:
                continue
        for p in params.values():
            p.grad = None
        lref = F.mse_loss(y, networks.hashmlp_forward(x, params, levels, 2, aniso, oact))
        lref.backward()
        one, two, three = copy.deepcopy(net).to(DEV), copy.deepcopy(net).to(DEV), copy.deepcopy(net).to(DEV)
        one.fuse_step, one.direct_step = True, False      # one kernel
        two.fuse_step, two.direct_step = False, False     # autograd: forward kernel, MSE kernel, backward kernel
        three.fuse_step, three.direct_step = False, True  # the same three kernels called directly (the default)
        batch = (x.to(DEV), y.to(DEV))
        assert one.fused_training_step(batch, 0) is None  # separately allocated gradients: no common level layout
        for m in (one, two, three):
            m.configure_optimizers()                      # ... the flat arenas give the tables and their gradients one layout
        l1 = one.fused_training_step(batch, 0)
        assert l1 is not None and l1.grad_fn is None and two.fused_training_step(batch, 0) is None
        l2 = training_step_and_backward(two, batch, 0)
        l3 = three.fused_training_step(batch, 0)
        assert l3 is not None and l3.grad_fn is None and abs(float(l3) - float(l2)) <= 1e-6 * abs(float(l2))  # atomic partial sums
        for (name, p3), (_, p2) in zip(three.named_parameters(), two.named_parameters()):
            if not name.startswith("layers."):
                # same kernels on the same inputs: equal up to the order in which the reductions land
                assert rel_err(p3.grad, p2.grad) < 1e-6, name
        assert abs(float(l1) - float(lref)) < 1e-5 * max(1.0, float(lref)) and abs(float(l1) - float(l2)) < 1e-6
        for (name, p1), (_, p2) in zip(one.named_parameters(), two.named_parameters()):
            if name.startswith("layers."):
                continue
            ref = params[name].grad
            scale = float(ref.abs().max()) + 1e-12
            assert float((p1.grad.cpu() - ref).abs().max()) < 2e-3 * scale, (n, name)
            assert float((p1.grad - p2.grad).abs().max()) < 2e-3 * scale, (n, name)
            if n >= 17:
                assert rel_err(p1.grad, ref) < 1e-3, (n, name)
        # a second call accumulates (autograd semantics): gradients double
        g_before = {k: p.grad.clone() for k, p in one.named_parameters() if p.grad is not None and not k.startswith("layers.")}
        one.fused_training_step(batch, 1)
        for k, p in one.named_parameters():
            if k in g_before:
                assert rel_err(p.grad, 2 * g_before[k]) < 1e-5, k


def test_trainer_takes_the_one_kernel_step_and_matches_the_autograd_loop(tmp_path):
    """Trainer.fit with the one-kernel step (default) and with MRI_FUSED_STEP switched off train to the same parameters."""
    import copy
    from mri_interpolation_b200 import _lib, models
    from mri_interpolation_b200.datamodules import DeviceBatchLoader
    from mri_interpolation_b200.pl_compat import pl
    kw = dict(dim_in=4, n_levels=16, n_features_per_level=2, log2_hashmap_size=12, base_resolution=16, finest_resolution=200,
              dim_hidden=64, dim_out=1, n_layers=2)
    torch.manual_seed(1337)
    net = models.HashMLP(**kw, batch_norm=False, lr=5e-3)
    gen = torch.Generator().manual_seed(3)
    coords, pix = torch.rand(5000, 4, generator=gen), torch.rand(5000, 1, generator=gen)
    results = []
    for fuse in (True, False):
        m = copy.deepcopy(net)
        m.fuse_step = fuse  # opt-in (default off: measured slower than the two-kernel step); fuse False -> the direct step
        loader = DeviceBatchLoader(coords, pix, 1024, shuffle=True, device=DEV, seed=11)
        tr = pl.Trainer(accelerator="gpu", max_epochs=2, precision=32, default_root_dir=str(tmp_path), enable_checkpointing=False,
                        cuda_graph=False)
        before = _lib.launch_count
        tr.fit(m, loader)
        results.append((m, _lib.launch_count - before, float(tr.callback_metrics["train_loss"])))
    (a, la, lossa), (b, lb, lossb) = results
    assert la < lb  # fewer launches: one kernel instead of forward + MSE + backward
    assert abs(lossa - lossb) < 1e-5
    for (k, p), (_, q) in zip(a.state_dict().items(), b.state_dict().items()):
        if p.dtype.is_floating_point and p.numel() and float(q.norm()) > 0:
            assert rel_err(p, q) < 1e-4, k


@pytest.mark.parametrize("n_levels,hidden", [(16, 64), (8, 64), (16, 128)])
def test_direct_step_equals_the_autograd_loop_over_several_optimiser_steps(n_levels, hidden):
    """pl_compat.training_step_and_backward with the direct step (default) and with it switched off (autograd: the same
    kernels behind two Functions) train to the same parameters; x.requires_grad or a custom criterion fall back."""
    import copy
    from mri_interpolation_b200 import models
    from mri_interpolation_b200.pl_compat import training_step_and_backward
    kw = dict(dim_in=4, n_levels=n_levels, n_features_per_level=2, log2_hashmap_size=12, base_resolution=16, finest_resolution=200,
              dim_hidden=hidden, dim_out=1, n_layers=2)
    torch.manual_seed(1337)
    net = models.HashMLP(**kw, batch_norm=False, lr=5e-3)
    gen = torch.Generator().manual_seed(9)
    direct, auto = copy.deepcopy(net).to(DEV), copy.deepcopy(net).to(DEV)
    auto.direct_step = False
    od, oa = direct.configure_optimizers(), auto.configure_optimizers()
    for step in range(4):
        x, y = torch.rand(3000, 4, generator=gen).to(DEV), torch.rand(3000, 1, generator=gen).to(DEV)
        ld = training_step_and_backward(direct, (x, y), step)
        la = training_step_and_backward(auto, (x, y), step)
        assert abs(float(ld) - float(la)) <= 1e-5 * abs(float(la))
        od.step(); od.zero_grad(); oa.step(); oa.zero_grad()
    for (k, p), (_, q) in zip(direct.state_dict().items(), auto.state_dict().items()):
        if p.dtype.is_floating_point and p.numel() and float(q.norm()) > 0:
            assert rel_err(p, q) < 1e-5, k
    x = torch.rand(100, 4, generator=gen).to(DEV).requires_grad_()
    assert direct.fused_training_step((x, torch.rand(100, 1, device=DEV)), 0) is None
    direct.criterion = torch.nn.functional.l1_loss
    assert direct.fused_training_step((x.detach(), torch.rand(100, 1, device=DEV)), 0) is None


def test_locality_ordered_and_shuffled_batches_train_to_the_same_parameters():
    """The batch ORDER is the loader's choice, the batch SET is the reference's: five Adam steps on locality-ordered
    batches (axis-0 index fastest, what ShuffledEpochs produces) and on the same batches in random order end in the same
    parameters up to fp32 summation order (G4-shaped model on a small volume, headline kernels with the register merge)."""
    import copy
    from mri_interpolation_b200 import functional as Fn, models
    from mri_interpolation_b200.pl_compat import training_step_and_backward
    kw = dict(dim_in=4, n_levels=16, n_features_per_level=2, log2_hashmap_size=14, base_resolution=16, finest_resolution=512,
              dim_hidden=64, dim_out=1, n_layers=2)
    torch.manual_seed(1337)
    net = models.HashMLP(**kw, batch_norm=False, lr=5e-3)
    gen = torch.Generator().manual_seed(21)
    with torch.no_grad():
        for lv in net.encoder.levels:
            lv.embedding.weight.copy_(torch.randn(lv.embedding.weight.shape, generator=gen) * 0.1)
    shape = (48, 40, 4, 5)
    total = int(np.prod(shape))
    pix = torch.rand(total, generator=gen).to(DEV)
    sampler = Fn.VoxelSampler(pix, shape)
    a, b = copy.deepcopy(net).to(DEV), copy.deepcopy(net).to(DEV)
    oa, ob = a.configure_optimizers(), b.configure_optimizers()
    for step in range(5):
        idx = torch.randperm(total, generator=gen)[:6000]
        sorted_idx = Fn.locality_sort(idx, shape, block=1).to(DEV)
        shuffled_idx = idx[torch.randperm(6000, generator=gen)].to(DEV)
        assert torch.equal(torch.sort(sorted_idx).values, torch.sort(shuffled_idx).values)
        la = training_step_and_backward(a, sampler.batch(sorted_idx), step)
        lb = training_step_and_backward(b, sampler.batch(shuffled_idx), step)
        assert abs(float(la) - float(lb)) <= 1e-5 * abs(float(lb))
        oa.step(); oa.zero_grad(); ob.step(); ob.zero_grad()
    for (k, p), (_, q) in zip(a.state_dict().items(), b.state_dict().items()):
        if p.dtype.is_floating_point and p.numel() and float(q.norm()) > 0:
            assert rel_err(p, q) < 1e-5, k
