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
