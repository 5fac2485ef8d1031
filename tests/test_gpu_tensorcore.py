"""GPU parity: tcgen05 SIREN layers (bf16x3 split-precision and plain bf16) against fp64 torch matmul."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_split_planes_reconstruct_fp32():
    from mri_interpolation_b200 import tc
    x = torch.randn(1000, 64, device=DEV) * 3
    hi, lo = tc.split(x)
    rec = hi.float() + lo.float()
    assert float(((rec - x).abs() / x.abs().clamp_min(1e-20)).max()) < 2 ** -15
    assert torch.equal(hi, x.to(torch.bfloat16))


@pytest.mark.parametrize("n,k,m", [(128, 64, 64), (300, 128, 128), (1000, 256, 256), (4096, 1024, 1024), (777, 192, 64),
                                   (130, 64, 512)])
@pytest.mark.parametrize("passes", [3, 1])
def test_tc_layer_identity_matches_fp64(n, k, m, passes):
    from mri_interpolation_b200 import tc
    from mri_interpolation_b200._lib import ACT_IDENTITY
    gen = torch.Generator(device=DEV).manual_seed(n + k + m)
    a = torch.rand(n, k, device=DEV, generator=gen) * 2 - 1
    w = (torch.rand(m, k, device=DEV, generator=gen) * 2 - 1) / k ** 0.5
    b = torch.rand(m, device=DEV, generator=gen) - 0.5
    ref = a.double() @ w.double().t() + b.double()
    a_hi, a_lo = tc.split(a)
    w_hi, w_lo = tc.split(w)
    oh, ol, of, _ = tc.layer(a_hi, a_lo, w_hi, w_lo, b, ACT_IDENTITY, 1.0, passes=passes, want_f32=True)
    err = rel_err(of, ref)
    assert err < (2e-5 if passes == 3 else 6e-3), err
    if passes == 3:
        assert rel_err(oh.float() + ol.float(), ref) < 3e-5
    else:
        assert rel_err(oh.float(), ref) < 1e-2


def test_tc_layer_sine_epilogue_and_aux():
    from mri_interpolation_b200 import tc
    from mri_interpolation_b200._lib import ACT_SINE
    n, k, m, w0 = 2048, 256, 256, 30.0
    gen = torch.Generator(device=DEV).manual_seed(5)
    a = torch.rand(n, k, device=DEV, generator=gen) * 2 - 1
    bound = (6.0 / k) ** 0.5 / w0
    w = (torch.rand(m, k, device=DEV, generator=gen) * 2 - 1) * bound
    b = (torch.rand(m, device=DEV, generator=gen) * 2 - 1) * bound
    pre = a.double() @ w.double().t() + b.double()
    a_hi, a_lo = tc.split(a)
    w_hi, w_lo = tc.split(w)
    oh, ol, of, aux = tc.layer(a_hi, a_lo, w_hi, w_lo, b, ACT_SINE, w0, passes=3, want_f32=True, want_aux=True)
    assert float((of.double() - torch.sin(w0 * pre)).abs().max()) < 1e-4
    assert float((aux.double() - w0 * torch.cos(w0 * pre)).abs().max()) < 1e-4 * w0
    mul = torch.rand(n, m, device=DEV, generator=gen)
    _, _, of2, _ = tc.layer(a_hi, a_lo, w_hi, w_lo, b, ACT_SINE, w0, passes=3, mul=mul, want_planes=False, want_f32=True)
    assert float((of2 - of * mul).abs().max()) < 1e-6


@pytest.mark.parametrize("n,k,m", [(256, 64, 128), (1000, 128, 128), (5000, 256, 256), (20000, 1024, 1024), (333, 192, 384)])
@pytest.mark.parametrize("passes", [3, 1])
def test_tc_wgrad_matches_fp64(n, k, m, passes):
    from mri_interpolation_b200 import tc
    gen = torch.Generator(device=DEV).manual_seed(n + 3 * k + m)
    g = torch.randn(n, m, device=DEV, generator=gen)
    x = torch.rand(n, k, device=DEV, generator=gen) * 2 - 1
    g_hi, g_lo = tc.split(g)
    x_hi, x_lo = tc.split(x)
    gw = torch.zeros(m, k, device=DEV)
    gb = torch.zeros(m, device=DEV)
    tc.wgrad(g_hi, g_lo, x_hi, x_lo, gw, gb, passes=passes)
    ref_w = g.double().t() @ x.double()
    ref_b = (g_hi.double() + (g_lo.double() if passes == 3 else 0)).sum(0)
    assert rel_err(gw, ref_w) < (2e-5 if passes == 3 else 8e-3)
    assert rel_err(gb, ref_b) < 1e-5
    tc.wgrad(g_hi, g_lo, x_hi, x_lo, gw, None, passes=passes)  # accumulates
    assert rel_err(gw, 2 * ref_w) < (2e-5 if passes == 3 else 8e-3)


def test_tc_dgrad_via_transposed_weights_and_mul_split():
    from mri_interpolation_b200 import tc
    from mri_interpolation_b200._lib import ACT_IDENTITY
    n, k, m = 3000, 256, 256
    gen = torch.Generator(device=DEV).manual_seed(9)
    dout = torch.randn(n, m, device=DEV, generator=gen)
    actp = torch.randn(n, m, device=DEV, generator=gen)
    w = torch.randn(m, k, device=DEV, generator=gen) / k ** 0.5
    prev = torch.randn(n, k, device=DEV, generator=gen)
    g_hi, g_lo = tc.mul_split(dout, actp)
    assert rel_err(g_hi.float() + g_lo.float(), dout * actp) < 1e-4
    wt_hi, wt_lo = tc.split(w.t().contiguous())
    oh, ol, of, _ = tc.layer(g_hi, g_lo, wt_hi, wt_lo, None, ACT_IDENTITY, 1.0, passes=3, mul=prev, want_f32=True)
    ref = ((dout * actp).double() @ w.double()) * prev.double()
    assert rel_err(of, ref) < 3e-5 and rel_err(oh.float() + ol.float(), ref) < 5e-5


@pytest.mark.parametrize("kw,n", [(dict(dim_in=3, dim_hidden=256, n_layers=3), 3000),
                                  (dict(dim_in=4, dim_hidden=256, n_layers=5), 5000),
                                  (dict(dim_in=3, dim_hidden=1024, n_layers=8), 4096),
                                  (dict(dim_in=3, dim_hidden=352, n_layers=4), 3000),   # the notebook's SIREN (nb:837): 352 -> 384 tiles
                                  (dict(dim_in=4, dim_hidden=200, n_layers=3), 1500)])  # 200 -> 256
def test_sirennet_tensor_core_path_matches_oracle(kw, n):
    """Whole network fwd + bwd in the split-precision mode vs the fp32 oracle: <= 1e-3 relative (north_star); widths
    that are not a multiple of 128 run zero-padded on the same tiles."""
    import torch.nn.functional as F
    from mri_interpolation_b200 import models
    from oracle import networks
    torch.manual_seed(1337)
    net = models.SirenNet(**kw)
    torch.manual_seed(1337)
    params, w0s = networks.siren_init(**kw)
    gen = torch.Generator().manual_seed(2)
    x = torch.rand(n, kw["dim_in"], generator=gen) * 2 - 1
    y = torch.rand(n, 1, generator=gen)
    ref = {k: v.clone().requires_grad_() for k, v in params.items()}
    pred_ref = networks.siren_forward(x, ref, w0s)
    F.mse_loss(y, pred_ref).backward()
    net = net.to(DEV)
    assert net._tensor_core_mode() == "bf16x3"
    pred = net(x.to(DEV))
    assert rel_err(pred, pred_ref.detach()) < 1e-3
    loss = net.training_step((x.to(DEV), y.to(DEV)), 0)
    loss.backward()
    for name, p in net.named_parameters():
        assert rel_err(p.grad, ref[name].grad) < 1e-3, name
    # the CUDA-core fp32 path agrees too, and plain bf16 stays within its stated tolerance
    net.precision = "fp32"
    assert rel_err(net(x.to(DEV)), pred_ref.detach()) < 1e-3
    net.precision = "bf16"
    with torch.no_grad():
        assert rel_err(net(x.to(DEV)), pred_ref.detach()) < 5e-2  # bf16 mode: stated tolerance 5e-2


@pytest.mark.parametrize("n,k,m", [(3000, 256, 256), (1000, 128, 512), (5000, 1024, 1024), (700, 512, 64)])
@pytest.mark.parametrize("passes", [3, 1])
def test_tc_dgrad_on_untransposed_weights(n, k, m, passes):
    from mri_interpolation_b200 import tc
    gen = torch.Generator(device=DEV).manual_seed(n + k + m)
    g = torch.randn(n, m, device=DEV, generator=gen)
    w = torch.randn(m, k, device=DEV, generator=gen) / m ** 0.5
    mul = torch.randn(n, k, device=DEV, generator=gen)
    g_hi, g_lo = tc.split(g)
    w_hi, w_lo = tc.split(w)
    cs = torch.zeros(k, device=DEV)
    oh, ol, of = tc.dgrad(g_hi, g_lo if passes == 3 else None, w_hi, w_lo if passes == 3 else None, passes=passes, mul=mul,
                          want_f32=True, colsum=cs)
    ref = (g.double() @ w.double()) * mul.double()
    assert rel_err(cs, of.double().sum(0)) < 1e-4
    assert rel_err(of, ref) < (3e-5 if passes == 3 else 8e-3)
    if passes == 3:
        assert rel_err(oh.float() + ol.float(), ref) < 5e-5
