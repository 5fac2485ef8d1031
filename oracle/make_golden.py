"""Pin the oracle against the reference itself and write tests/golden/*.npz.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
For every case it (1) runs the reference's own classes (encoding.py / models.py imported
unmodified through oracle/_ref_import.py), (2) runs the oracle restatement on the same
seeded inputs, (3) asserts the two are BIT-IDENTICAL on CPU, (4) stores inputs + reference
outputs as a fixture.  The fixtures are what travels to the GPU box.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import _ref_import, hashgrid, networks, sweep  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# (name, dim, n_levels, F, log2T, base, finest)  - ints -> V1, tuples -> V2
HASH_CASES = [
    ("v1_d4", 4, 4, 2, 12, 4, 24),
    ("v1_d3", 3, 6, 2, 10, 4, 40),
    ("v1_d2_f4", 2, 5, 4, 9, 8, 64),
    ("v1_d3_f1", 3, 3, 1, 14, 16, 64),
    ("v2_d3", 3, 4, 2, 13, (16, 16, 5), (88, 88, 15)),
    ("v2_d4_f1", 4, 3, 1, 12, (8, 8, 3, 5), (40, 44, 6, 15)),
]


def edge_coords(dim: int, levels, n_random: int, gen: torch.Generator) -> torch.Tensor:
    """Uniform randoms plus the edge set {0, 1, k/res, largest float below 1, linspace grid points}."""
    rows = [torch.rand(n_random, dim, generator=gen)]
    rows.append(torch.zeros(1, dim))
    rows.append(torch.ones(1, dim))
    rows.append(torch.full((1, dim), float(np.nextafter(np.float32(1), np.float32(0)))))
    for level in levels:
        res = torch.tensor([float(r) for r in level.resolution]).expand(dim) if len(level.resolution) == dim else None
        if res is None:
            continue
        k = torch.randint(0, 1 << 30, (6, dim), generator=gen) % (res.long() + 1)
        rows.append((k.float() / res).clamp(0, 1))
    lin = torch.linspace(0, 1, 15)
    idx = torch.randint(0, 15, (32, dim), generator=gen)
    rows.append(lin[idx])
    # mixed rows: some axes exactly 1, others random
    mix = torch.rand(8, dim, generator=gen)
    mix[:, 0] = 1.0
    rows.append(mix)
    return torch.cat(rows).contiguous()


def build_ref_encoder(ref_encoding, dim, n_levels, F, log2T, base, finest):
    cls = ref_encoding.MultiResHashGrid if isinstance(base, int) else ref_encoding.MultiResHashGridV2
    return cls(dim=dim, n_levels=n_levels, n_features_per_level=F, log2_hashmap_size=log2T,
               base_resolution=base, finest_resolution=finest)


def golden_hash(ref_encoding):
    for name, dim, n_levels, F, log2T, base, finest in HASH_CASES:
        aniso = not isinstance(base, int)
        torch.manual_seed(1337)
        ref = build_ref_encoder(ref_encoding, dim, n_levels, F, log2T, base, finest)
        torch.manual_seed(1337)
        levels = hashgrid.geometry(dim, n_levels, log2T, base, finest)
        tables = hashgrid.init_tables(levels, F)
        # geometry + init parity
        for li, lv in enumerate(levels):
            r = ref.levels[li]
            assert r.hashmap_size == lv.rows, (name, li)
            rr = r.resolution.tolist() if aniso else [r.resolution] * dim
            assert [int(v) for v in rr] == list(lv.resolution), (name, li)
            assert torch.equal(r.embedding.weight.detach(), tables[li]), (name, li, "init")
        # make the tables less tiny so outputs are well scaled for tolerance tests
        gen = torch.Generator().manual_seed(4242)
        tables = [torch.randn(t.shape, generator=gen) * 0.1 for t in tables]
        with torch.no_grad():
            for li, t in enumerate(tables):
                ref.levels[li].embedding.weight.copy_(t)
        x = edge_coords(dim, levels, 200, gen)
        # reference forward/backward
        out_ref = ref(x)
        gout = torch.randn(out_ref.shape, generator=gen)
        out_ref.backward(gout)
        gref = [ref.levels[li].embedding.weight.grad.clone() for li in range(n_levels)]
        # oracle forward
        out_or = hashgrid.encode(x, tables, levels, aniso)
        assert torch.equal(out_or, out_ref.detach()), (name, "forward not bit-identical")
        # hashes / weights per level
        hs, ws = [], []
        for li, lv in enumerate(levels):
            h, w = hashgrid.corners(x, lv, aniso)
            # reference hashes through its own fast_hash on its own corner set
            r = ref.levels[li]
            xs = x * (r.resolution if aniso else r.resolution)
            xi = xs.long().unsqueeze(-2)
            inds = torch.where(r.bin_mask.unsqueeze(0), xi, xi + 1)
            h_ref = ref_encoding.fast_hash(inds, r.primes, r.hashmap_size)
            assert torch.equal(h, h_ref), (name, li, "hash")
            hs.append(h)
            ws.append(w)
        g_or = hashgrid.table_gradients(x, gout, levels, F, aniso)
        for a, b in zip(g_or, gref):
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
        fx = dict(
            dim=dim, n_levels=n_levels, n_features=F, log2_hashmap_size=log2T,
            base=np.asarray(base), finest=np.asarray(finest), anisotropic=aniso,
            resolutions=np.asarray([lv.resolution for lv in levels], dtype=np.int64),
            rows=np.asarray([lv.rows for lv in levels], dtype=np.int64),
            x=x.numpy(), out=out_ref.detach().numpy(), grad_out=gout.numpy(),
            hashes=torch.stack(hs, 1).numpy().astype(np.uint32),
            weights=torch.stack(ws, 1).numpy(),
        )
        for li in range(n_levels):
            fx[f"table{li}"] = tables[li].numpy()
            fx[f"grad{li}"] = gref[li].numpy()
        np.savez_compressed(os.path.join(GOLD, f"hashgrid_{name}.npz"), **fx)
        print(f"hashgrid {name}: N={x.shape[0]} levels={[(lv.resolution, lv.rows) for lv in levels]}")


def golden_geometry(ref_encoding):
    """Headline geometries: G4 (hash_config.json through the python API), the notebook V2 config
    (6 009 032 params, nb:2792) and HashConfig's shipped tuples (config/base.py:70-74)."""
    out = {}
    g4 = dict(dim=4, n_levels=16, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16,
              finest_resolution=2489)
    torch.manual_seed(0)
    ref = ref_encoding.MultiResHashGrid(**g4)
    lv = hashgrid.geometry_isotropic(4, 16, 19, 16, 2489)
    assert [l.rows for l in lv] == [r.hashmap_size for r in ref.levels]
    assert [l.resolution[0] for l in lv] == [r.resolution for r in ref.levels]
    assert sum(l.rows for l in lv) * 2 == 15279648
    out["g4_res"] = np.asarray([l.resolution[0] for l in lv])
    out["g4_rows"] = np.asarray([l.rows for l in lv])
    # all-ones corner set at level 0 (SURVEY 8c weak pin), from the reference
    r0 = ref.levels[0]
    x = torch.ones(1, 4)
    xi = (x * r0.resolution).long().unsqueeze(-2)
    inds = torch.where(r0.bin_mask.unsqueeze(0), xi, xi + 1)
    h = ref_encoding.fast_hash(inds, r0.primes, r0.hashmap_size)[0]
    assert h.tolist() == [52480, 52481, 17105, 17104, 25781, 25780, 60260, 60261, 4117, 4116, 40900, 40901,
                          47520, 47521, 13937, 13936]
    ho, _ = hashgrid.corners(x, lv[0])
    assert torch.equal(ho[0], h)
    out["g4_ones_hash_l0"] = h.numpy()
    nbv2 = dict(dim=3, n_levels=8, n_features_per_level=2, log2_hashmap_size=23, base_resolution=(64, 64, 5),
                finest_resolution=(512, 512, 15))
    ref2 = ref_encoding.MultiResHashGridV2(**nbv2)
    lv2 = hashgrid.geometry_anisotropic(3, 8, 23, (64, 64, 5), (512, 512, 15))
    assert [l.rows for l in lv2] == [r.hashmap_size for r in ref2.levels]
    assert sum(p.numel() for p in ref2.parameters()) == 6009032 == sum(l.rows for l in lv2) * 2
    out["nbv2_res"] = np.asarray([l.resolution for l in lv2])
    out["nbv2_rows"] = np.asarray([l.rows for l in lv2])
    d3 = hashgrid.geometry_isotropic(3, 16, 19, 16, 512)
    ref3 = ref_encoding.MultiResHashGrid(dim=3, n_levels=16, n_features_per_level=2, log2_hashmap_size=19,
                                         base_resolution=16, finest_resolution=512)
    assert [l.resolution[0] for l in d3] == [r.resolution for r in ref3.levels]
    out["d3_res"] = np.asarray([l.resolution[0] for l in d3])
    out["d3_rows"] = np.asarray([l.rows for l in d3])
    np.savez_compressed(os.path.join(GOLD, "geometry.npz"), **out)
    print("geometry: G4 res", out["g4_res"].tolist())


def golden_siren(ref_models):
    cases = [
        ("small", dict(dim_in=3, dim_hidden=32, dim_out=1, n_layers=3, w0=30.0, w0_initial=30.0), 97, False),
        ("d4_h64", dict(dim_in=4, dim_hidden=64, dim_out=1, n_layers=4, w0=30.0, w0_initial=30.0), 130, False),
        ("d2_out3", dict(dim_in=2, dim_hidden=48, dim_out=3, n_layers=2, w0=20.0, w0_initial=25.0), 64, True),
    ]
    for name, kw, n, siren_norm in cases:
        torch.manual_seed(1337)
        ref = ref_models.SirenNet(**kw)
        torch.manual_seed(1337)
        params, w0s = networks.siren_init(**kw)
        sd = ref.state_dict()
        for k, v in params.items():
            assert torch.equal(sd[k], v), (name, k, "init")
        assert set(sd.keys()) == set(params.keys()), (sorted(sd.keys()), sorted(params.keys()))
        gen = torch.Generator().manual_seed(99)
        x = torch.rand(n, kw["dim_in"], generator=gen)
        if siren_norm:
            x = x * 2 - 1
        y = torch.rand(n, kw["dim_out"], generator=gen)
        pred = ref(x)
        loss = torch.nn.functional.mse_loss(y, pred)
        loss.backward()
        pred_or = networks.siren_forward(x, params, w0s)
        assert torch.equal(pred_or, pred.detach()), (name, "forward")
        grads, _ = networks.siren_backward(x, params, w0s, networks.mse_grad(y, pred_or))
        fx = dict(x=x.numpy(), y=y.numpy(), pred=pred.detach().numpy(), loss=float(loss),
                  w0s=np.asarray(w0s), **{f"kw_{k}": v for k, v in kw.items()})
        for k, p in ref.named_parameters():
            torch.testing.assert_close(grads[k], p.grad, rtol=2e-4, atol=1e-7)
            fx[f"param:{k}"] = p.detach().numpy()
            fx[f"grad:{k}"] = p.grad.numpy()
        np.savez_compressed(os.path.join(GOLD, f"siren_{name}.npz"), **fx)
        print(f"siren {name}: loss={float(loss):.6f}")
    # parameter-count known answers (nb:825, nb:1347)
    assert sum(p.numel() for p in ref_models.SirenNet(dim_in=2, dim_hidden=352, n_layers=4).parameters()) == 374177
    assert sum(p.numel() for p in ref_models.SirenNet(dim_in=3, dim_hidden=1408, n_layers=4).parameters()) == 5958657


def golden_hashmlp(ref_models):
    """HashMLP construction (RNG order + state_dict keys) and the intended forward on the
    BN-free notebook variant (decoder blocks Linear->GELU)."""
    kw = dict(dim_in=3, n_levels=4, n_features_per_level=2, log2_hashmap_size=11, base_resolution=4,
              finest_resolution=32, dim_hidden=16, dim_out=1, n_layers=2)
    torch.manual_seed(1337)
    ref = ref_models.HashMLP(**kw)
    torch.manual_seed(1337)
    params, levels = networks.hashmlp_init(**kw)
    sd = ref.state_dict()
    for k, v in params.items():
        assert torch.equal(sd[k], v), (k, "hashmlp init")
    extra = set(sd.keys()) - set(params.keys())
    assert all(".1." in k for k in extra), extra  # only BatchNorm1d entries are not modelled by the oracle
    gen = torch.Generator().manual_seed(5)
    x = torch.rand(150, 3, generator=gen)
    y = torch.rand(150, 1, generator=gen)
    # scale tables up so the decoder sees non-trivial inputs
    for li in range(4):
        params[f"encoder.levels.{li}.embedding.weight"] = torch.randn(levels[li].rows, 2, generator=gen) * 0.3
    with torch.no_grad():
        for li in range(4):
            ref.encoder.levels[li].embedding.weight.copy_(params[f"encoder.levels.{li}.embedding.weight"])
    # intended forward with the reference's own modules, BN/Dropout skipped (nb cell 37)
    z = ref.encoder(x)
    h = z
    for blk in ref.decoder:
        h = torch.nn.functional.gelu(blk[0](h))
    loss = torch.nn.functional.mse_loss(y, h)
    loss.backward()
    pred_or = networks.hashmlp_forward(x, params, levels, 2, False)
    assert torch.equal(pred_or, h.detach())
    fx = dict(x=x.numpy(), y=y.numpy(), pred=h.detach().numpy(), loss=float(loss), latents=z.detach().numpy(),
              **{f"kw_{k}": v for k, v in kw.items()})
    for k, p in ref.named_parameters():
        fx[f"param:{k}"] = p.detach().numpy()
        if p.grad is not None:
            fx[f"grad:{k}"] = p.grad.numpy()
    np.savez_compressed(os.path.join(GOLD, "hashmlp_small.npz"), **fx)
    # BatchNorm variant exactly as models.py:718-739 builds it (train mode), per-block loop forward
    torch.manual_seed(7)
    ref_bn = ref_models.HashMLP(**kw)
    with torch.no_grad():
        for li in range(4):
            ref_bn.encoder.levels[li].embedding.weight.copy_(params[f"encoder.levels.{li}.embedding.weight"])
    ref_bn.train()
    h = ref_bn.encoder(x)
    for blk in ref_bn.decoder:
        h = blk(h)
    loss_bn = torch.nn.functional.mse_loss(y, h)
    loss_bn.backward()
    fxb = dict(x=x.numpy(), y=y.numpy(), pred=h.detach().numpy(), loss=float(loss_bn),
               **{f"kw_{k}": v for k, v in kw.items()})
    for k, p in ref_bn.named_parameters():
        fxb[f"param:{k}"] = p.detach().numpy()
        if p.grad is not None:
            fxb[f"grad:{k}"] = p.grad.numpy()
    for k, b in ref_bn.named_buffers():
        if "running" in k or "num_batches" in k:
            fxb[f"buffer:{k}"] = b.detach().numpy()
    np.savez_compressed(os.path.join(GOLD, "hashmlp_bn_small.npz"), **fxb)
    print(f"hashmlp: loss={float(loss):.6f} bn_loss={float(loss_bn):.6f}")


def golden_zoo(ref_models):
    """Model-zoo variants on the hot-path operators: fixtures straight from the reference's classes
    (ModulatedSirenNet models.py:263-322, MultiSiren :888-956); the tcnn-backed classes cannot run in the reference."""
    kw = dict(dim_in=3, dim_hidden=32, dim_out=1, n_layers=3, w0=30.0, w0_initial=30.0)
    torch.manual_seed(1337)
    ref = ref_models.ModulatedSirenNet(**kw)
    gen = torch.Generator().manual_seed(5)
    x = torch.rand(83, 3, generator=gen) * 2 - 1
    y = torch.rand(83, 1, generator=gen)
    pred = ref(x.clone())
    loss = torch.nn.functional.mse_loss(y, pred)
    loss.backward()
    fx = dict(x=x.numpy(), y=y.numpy(), pred=pred.detach().numpy(), loss=float(loss), **{f"kw_{k}": v for k, v in kw.items()})
    for k, p in ref.named_parameters():
        fx[f"param:{k}"] = p.detach().numpy()
        if p.grad is not None:
            fx[f"grad:{k}"] = p.grad.numpy()
    np.savez_compressed(os.path.join(GOLD, "modulated_siren_small.npz"), **fx)
    print(f"modulated siren: loss={float(loss):.6f}")
    torch.manual_seed(1337)
    ms = ref_models.MultiSiren(dim_in=3, dim_hidden=16, dim_out=1, n_layers=2, n_frames=3, lr=1e-4)
    xf = torch.rand(1, 40, 3, generator=gen) * 2 - 1
    yf = torch.rand(1, 40, 1, generator=gen)
    loss = ms.training_step((xf, yf, 2), 0)
    loss.backward()
    fx = dict(x=xf.numpy(), y=yf.numpy(), loss=float(loss), pred=ms(xf[0], 2).detach().numpy())
    for k, p in ms.named_parameters():
        fx[f"param:{k}"] = p.detach().numpy()
        if p.grad is not None:
            fx[f"grad:{k}"] = p.grad.numpy()
    np.savez_compressed(os.path.join(GOLD, "multi_siren_small.npz"), **fx)
    print(f"multi siren: loss={float(loss):.6f}")


def golden_adam():
    gen = torch.Generator().manual_seed(11)
    for name, kw in (("default", dict(lr=5e-3)), ("tcnn_like", dict(lr=1e-2, betas=(0.9, 0.99), eps=1e-15)),
                     ("l2", dict(lr=1e-4, weight_decay=1e-5))):
        p0 = torch.randn(1000, generator=gen) * 0.1
        p = torch.nn.Parameter(p0.clone())
        opt = torch.optim.Adam([p], **kw)
        po, m, v = p0.clone(), torch.zeros(1000), torch.zeros(1000)
        gs, traj = [], []
        for step in range(1, 6):
            g = torch.randn(1000, generator=gen) * (1e-3 if step % 2 else 1.0)
            g[::7] = 0.0  # untouched rows still decay (dense Adam)
            p.grad = g.clone()
            opt.step()
            b1, b2 = kw.get("betas", (0.9, 0.999))
            networks.adam_step(po, g, m, v, step, kw["lr"], b1, b2, kw.get("eps", 1e-8), kw.get("weight_decay", 0.0))
            torch.testing.assert_close(po, p.detach(), rtol=1e-6, atol=1e-9)
            gs.append(g.numpy())
            traj.append(p.detach().clone().numpy())
        np.savez_compressed(os.path.join(GOLD, f"adam_{name}.npz"), p0=p0.numpy(), grads=np.stack(gs),
                            params=np.stack(traj), lr=kw["lr"], beta1=kw.get("betas", (0.9, 0.999))[0],
                            beta2=kw.get("betas", (0.9, 0.999))[1], eps=kw.get("eps", 1e-8),
                            weight_decay=kw.get("weight_decay", 0.0))
    print("adam ok")


def golden_sweep():
    shape = (7, 5, 3, 4)
    c = sweep.grid_coords(shape)
    axes = [torch.linspace(0, 1, s) for s in shape]
    ref = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(-1, 4)  # utils.py:14-23 recipe
    assert torch.equal(c, ref)
    np.savez_compressed(os.path.join(GOLD, "sweep_coords.npz"), shape=np.asarray(shape), coords=c.numpy(),
                        lin29=torch.linspace(0, 1, 29).numpy(), lin352=torch.linspace(0, 1, 352).numpy(),
                        lin57m=torch.linspace(-1, 1, 57).numpy())
    print("sweep ok")


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(1)  # reduction order independent of the thread count
    ref_encoding, ref_models = _ref_import.load()
    golden_geometry(ref_encoding)
    golden_hash(ref_encoding)
    golden_siren(ref_models)
    golden_hashmlp(ref_models)
    golden_zoo(ref_models)
    golden_adam()
    golden_sweep()
    print("oracle pinned against the reference; fixtures in", GOLD)


if __name__ == "__main__":
    main()
