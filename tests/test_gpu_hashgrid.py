"""GPU parity: hash-grid kernels (through the C ABI via the nn.Module surface) against the oracle and
the reference-generated golden vectors.  Indices bit-exact; values within 1e-3 relative (fp32)."""
import numpy as np
import pytest
import torch

from conftest import HASH_CASES, build_encoder, load_golden, oracle_levels

pytestmark = pytest.mark.gpu

RTOL = 1e-3  # north_star tolerance for fp32 outputs/gradients
DEV = "cuda"


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("case", HASH_CASES)
def test_corner_hashes_bit_exact_and_weights(case):
    fx = load_golden(f"hashgrid_{case}.npz")
    enc = build_encoder(fx, DEV)
    h, w = enc.corner_hashes(torch.from_numpy(fx["x"]).to(DEV))
    assert np.array_equal(h.cpu().numpy().astype(np.uint32), fx["hashes"])  # bit-exact indices, reference corner order
    np.testing.assert_allclose(w.cpu().numpy(), fx["weights"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("case", HASH_CASES)
def test_production_gather_addresses_exactly_the_reference_rows(case):
    """The rows the SHIPPED gather kernel reads (hashgrid_fwd_kernel / encode_half_level with its row sink switched on,
    not the corner probe) equal the reference's hashes bit for bit, and the instrumented launch returns the same
    encoding as the plain one."""
    fx = load_golden(f"hashgrid_{case}.npz")
    enc = build_encoder(fx, DEV)
    x = torch.from_numpy(fx["x"]).to(DEV)
    out, rows = enc.gathered_rows(x)
    assert np.array_equal(rows.cpu().numpy().astype(np.uint32), fx["hashes"])
    with torch.no_grad():
        assert torch.equal(out, enc(x))


@pytest.mark.parametrize("case", HASH_CASES)
def test_production_scatter_touches_exactly_the_reference_rows(case):
    """Support of the table gradient written by the shipped scatter kernel == the set of reference rows with a
    non-zero corner weight (one-hot probing of the production reduction path: a wrong index lands on another row)."""
    fx = load_golden(f"hashgrid_{case}.npz")
    enc = build_encoder(fx, DEV)
    x = torch.from_numpy(fx["x"]).to(DEV)
    out = enc(x)
    out.backward(torch.ones_like(out))
    hashes, weights = fx["hashes"].astype(np.int64), fx["weights"]
    for li, lv in enumerate(enc.levels):
        got = torch.nonzero(lv.embedding.weight.grad.abs().sum(1)).flatten().cpu().numpy()
        want = np.unique(hashes[:, li][weights[:, li] != 0])
        assert np.array_equal(got, want), f"level {li}"


@pytest.mark.parametrize("case", HASH_CASES)
def test_forward_backward_match_reference_vectors(case):
    fx = load_golden(f"hashgrid_{case}.npz")
    enc = build_encoder(fx, DEV)
    x = torch.from_numpy(fx["x"]).to(DEV)
    out = enc(x)
    ref = torch.from_numpy(fx["out"])
    assert out.shape == ref.shape
    torch.testing.assert_close(out.cpu(), ref, rtol=RTOL, atol=1e-6)
    assert rel_err(out, ref) < 1e-5
    out.backward(torch.from_numpy(fx["grad_out"]).to(DEV))
    for li, lv in enumerate(enc.levels):
        g = lv.embedding.weight.grad
        gref = torch.from_numpy(fx[f"grad{li}"])
        torch.testing.assert_close(g.cpu(), gref, rtol=RTOL, atol=1e-5)
        assert rel_err(g, gref) < 1e-5


def test_single_level_module_forward():
    fx = load_golden("hashgrid_v1_d3.npz")
    enc = build_encoder(fx, DEV)
    x = torch.from_numpy(fx["x"]).to(DEV)
    F = int(fx["n_features"])
    for li in (0, 2, 5):
        out = enc.levels[li](x)
        torch.testing.assert_close(out.cpu(), torch.from_numpy(fx["out"][:, li * F:(li + 1) * F]), rtol=RTOL, atol=1e-6)


def test_backward_accumulates_into_existing_grad_and_returns_fresh_otherwise():
    fx = load_golden("hashgrid_v1_d4.npz")
    enc = build_encoder(fx, DEV)
    x = torch.from_numpy(fx["x"]).to(DEV)
    g = torch.from_numpy(fx["grad_out"]).to(DEV)
    enc(x).backward(g)            # fresh grads
    first = [lv.embedding.weight.grad.clone() for lv in enc.levels]
    enc(x).backward(g)            # direct accumulation into the existing buffers
    for lv, f in zip(enc.levels, first):
        torch.testing.assert_close(lv.embedding.weight.grad, 2 * f, rtol=1e-5, atol=1e-6)


def test_leading_batch_dims_and_ragged_sizes():
    fx = load_golden("hashgrid_v2_d3.npz")
    enc = build_encoder(fx, DEV)
    x = torch.from_numpy(fx["x"]).to(DEV)
    ref = enc(x)
    x3 = x[:264].reshape(4, 66, 3)
    assert torch.equal(enc(x3), ref[:264].reshape(4, 66, -1))
    for n in (1, 31, 257):
        assert torch.equal(enc(x[:n].contiguous()), ref[:n])
    empty = enc(torch.empty(0, 3, device=DEV))
    assert empty.shape == (0, enc.output_dim)
    # a non-contiguous / offset input view still works
    assert torch.equal(enc(x[3:100]), ref[3:100])


@pytest.mark.parametrize("dim,kw", [
    (4, dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16, finest_resolution=2489)),
    (3, dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16, finest_resolution=512)),
])
def test_full_geometry_against_oracle_on_seeded_inputs(dim, kw):
    """Headline geometry (G4 / D3): oracle on CPU at a size it finishes in seconds."""
    from mri_interpolation_b200 import encoding
    from oracle import hashgrid
    torch.manual_seed(1337)
    enc = encoding.MultiResHashGrid(dim, **kw)
    gen = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for lv in enc.levels:
            lv.embedding.weight.copy_(torch.randn(lv.embedding.weight.shape, generator=gen) * 0.1)
    n = 4096
    x = torch.rand(n, dim, generator=gen)
    x[:8] = 1.0
    x[8:16] = 0.0
    levels = hashgrid.geometry_isotropic(dim, kw["n_levels"], kw["log2_hashmap_size"], kw["base_resolution"],
                                         kw["finest_resolution"])
    tables = [lv.embedding.weight.detach().clone() for lv in enc.levels]
    ref = hashgrid.encode(x, tables, levels)
    enc = enc.to(DEV)
    out = enc(x.to(DEV))
    assert rel_err(out, ref) < 1e-5
    torch.testing.assert_close(out.cpu(), ref, rtol=RTOL, atol=1e-6)
    h, _ = enc.corner_hashes(x.to(DEV))
    for li, lv in enumerate(levels):
        ho, _ = hashgrid.corners(x, lv)
        assert torch.equal(h[:, li].cpu(), ho), f"level {li} hashes differ"
    out_rows, rows = enc.gathered_rows(x.to(DEV))  # the production gather path itself
    assert torch.equal(out_rows, out.detach())
    for li, lv in enumerate(levels):
        ho, _ = hashgrid.corners(x, lv)
        assert torch.equal(rows[:, li].cpu(), ho), f"level {li}: rows gathered by the production kernel differ"
    go = torch.randn(ref.shape, generator=gen)
    out.backward(go.to(DEV))
    gref = hashgrid.table_gradients(x, go, levels, 2)
    for lv, g in zip(enc.levels, gref):
        assert rel_err(lv.embedding.weight.grad, g) < 1e-5


def test_size_independent_properties_at_full_batch():
    """2^20 coordinates through G4: partition of unity, linearity in the tables, gradient column sums."""
    from mri_interpolation_b200 import encoding
    enc = encoding.MultiResHashGrid(4, n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16,
                                    finest_resolution=2489).to(DEV)
    n = 1 << 20
    gen = torch.Generator(device=DEV).manual_seed(11)
    x = torch.rand(n, 4, device=DEV, generator=gen)
    with torch.no_grad():
        for lv in enc.levels:
            lv.embedding.weight.fill_(1.0)
        ones = enc(x)
        assert float((ones - 1).abs().max()) < 1e-5  # weights of the 16 corners sum to 1 on every level
        for lv in enc.levels:
            lv.embedding.weight.copy_(torch.randn(lv.embedding.weight.shape, device=DEV, generator=gen))
        a = enc(x)
        for lv in enc.levels:
            lv.embedding.weight.mul_(-2.5)
        b = enc(x)
        assert rel_err(b, -2.5 * a) < 1e-6
    go = torch.randn(n, 32, device=DEV, generator=gen)
    enc(x).backward(go)
    col = go.double().sum(0).reshape(16, 2)
    for li, lv in enumerate(enc.levels):
        got = lv.embedding.weight.grad.double().sum(0)
        assert float((got - col[li]).abs().max()) < 2e-2 * float(col[li].abs().max() + 10.0)


def test_cpu_input_is_rejected():
    from mri_interpolation_b200 import MriB200Error
    fx = load_golden("hashgrid_v1_d3.npz")
    enc = build_encoder(fx, DEV)
    with pytest.raises(MriB200Error):
        enc(torch.from_numpy(fx["x"]))


@pytest.mark.parametrize("dim,base,finest,log2", [(2, 3, 1500, 22), (3, 5, 150, 24), (4, 3, 40, 24)])
def test_multiply_based_modulo_is_exact_on_non_power_of_two_tables(dim, base, finest, log2):
    """The coarse levels here have res^D < T rows, not powers of two (9 ... 3 307 949): the kernels reduce the 32-bit hash with exact_mod
    (mulhi by floor(2^32 / rows) + one conditional subtraction) instead of a division - rows addressed by the shipped gather
    kernel and the probe must equal the oracle's `%` bit for bit, for hashes over the whole 32-bit range."""
    from mri_interpolation_b200 import encoding
    from oracle import hashgrid
    enc = encoding.MultiResHashGrid(dim, n_levels=6, n_features_per_level=2, log2_hashmap_size=log2, base_resolution=base,
                                    finest_resolution=finest).to(DEV)
    levels = hashgrid.geometry_isotropic(dim, 6, log2, base, finest)
    assert sum(1 for lv in levels if lv.rows & (lv.rows - 1)) >= 3
    gen = torch.Generator().manual_seed(dim)
    x = torch.rand(20000, dim, generator=gen)
    x[:64] = 1.0
    x[64:128] = 0.0
    _, rows = enc.gathered_rows(x.to(DEV))
    h, _ = enc.corner_hashes(x.to(DEV))
    for li, lv in enumerate(levels):
        ho, _ = hashgrid.corners(x, lv)
        assert torch.equal(rows[:, li].cpu(), ho), f"level {li} ({lv.rows} rows)"
        assert torch.equal(h[:, li].cpu(), ho), f"level {li} ({lv.rows} rows): probe"
