"""Model-zoo variants on the hot-path operators (mri_interpolation_b200/zoo.py): modulated SIRENs and the
tiny-cuda-nn shaped front-ends.  Fixtures for the pure-torch classes come from the reference itself
(oracle/make_golden.py::golden_zoo); the tcnn-backed classes cannot run in the reference (tinycudann import disabled,
models.py:10), so they are checked against the oracle's hash grid + a plain torch MLP on the same weights."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden

DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _kw(fx):
    return {k[3:]: (v.item() if hasattr(v, "item") else v) for k, v in fx.items() if k.startswith("kw_")}


def test_seeded_zoo_state_dicts_equal_the_reference():
    from mri_interpolation_b200 import models
    fx = load_golden("modulated_siren_small.npz")
    torch.manual_seed(1337)
    net = models.ModulatedSirenNet(**_kw(fx))
    sd = net.state_dict()
    names = [k[6:] for k in fx if k.startswith("param:")]
    assert set(names) == set(sd.keys())
    for k in names:
        assert np.array_equal(sd[k].numpy(), fx["param:" + k]), k
    fx = load_golden("multi_siren_small.npz")
    torch.manual_seed(1337)
    ms = models.MultiSiren(dim_in=3, dim_hidden=16, dim_out=1, n_layers=2, n_frames=3, lr=1e-4)
    sd = ms.state_dict()
    for k in [k[6:] for k in fx if k.startswith("param:")]:
        assert np.array_equal(sd[k].numpy(), fx["param:" + k]), k


def test_tcnn_style_geometry_and_names():
    from mri_interpolation_b200 import models
    cfg = json.load(open(os.path.join(ROOT, "config", "hash_config.json")))
    enc = models.TcnnStyleEncoding(4, cfg["encoding"])
    # res_l = floor(16 * 1.4^l): the G4 geometry of SURVEY 8a; 15 279 648 table parameters
    assert [lv.resolution for lv in enc.levels] == [16, 22, 31, 43, 61, 86, 120, 168, 236, 330, 462, 647, 907, 1269, 1777, 2489]
    assert sum(p.numel() for p in enc.parameters()) == 15279648 and enc.n_output_dims == 32
    net = models.TcnnStyleNetwork(32, 1, cfg["network"])
    assert [tuple(l.weight.shape) for l in net.layers] == [(128, 32), (128, 128), (1, 128)] and all(l.bias is None for l in net.layers)
    small = dict(encoding=dict(cfg["encoding"], log2_hashmap_size=10, n_levels=4), network=cfg["network"])
    for cls, args in ((models.HashSirenNet, (small,)), (models.MultiHashMLP, (4, 1, 2, small, 1e-3))):
        assert isinstance(cls(*args), torch.nn.Module)
    t = models.TcnnHashMLP(3, 4, 2, 10, 8, 1.5, dim_hidden=32)
    assert len(t.decoder.layers) == 9  # n_hidden_layers = BaseMLP's default n_layers (8), as in the reference
    with pytest.raises(NotImplementedError):
        models.RffNet()


@pytest.mark.gpu
def test_modulated_siren_matches_reference_vectors():
    from mri_interpolation_b200 import models
    fx = load_golden("modulated_siren_small.npz")
    torch.manual_seed(1337)
    net = models.ModulatedSirenNet(**_kw(fx)).to(DEV)
    x, y = torch.from_numpy(fx["x"]).to(DEV), torch.from_numpy(fx["y"]).to(DEV)
    pred = net(x)
    assert rel_err(pred, torch.from_numpy(fx["pred"])) < 1e-3
    loss = net.training_step((x, y), 0)
    assert abs(float(loss) - float(fx["loss"])) < 1e-4
    loss.backward()
    checked = 0
    for name, p in net.named_parameters():
        if "grad:" + name in fx:
            assert rel_err(p.grad, torch.from_numpy(fx["grad:" + name])) < 1e-3, name
            checked += 1
    assert checked == 14  # modulator (3 x 2) + siren (4 x 2)


@pytest.mark.gpu
def test_multi_siren_frame_training_step_matches_reference_vectors():
    from mri_interpolation_b200 import models
    fx = load_golden("multi_siren_small.npz")
    torch.manual_seed(1337)
    ms = models.MultiSiren(dim_in=3, dim_hidden=16, dim_out=1, n_layers=2, n_frames=3, lr=1e-4).to(DEV)
    x, y = torch.from_numpy(fx["x"]).to(DEV), torch.from_numpy(fx["y"]).to(DEV)
    assert rel_err(ms(x[0], 2), torch.from_numpy(fx["pred"])) < 1e-3
    loss = ms.training_step((x, y, 2), 0)
    assert abs(float(loss) - float(fx["loss"])) < 1e-4
    loss.backward()
    for name, p in ms.named_parameters():
        if "grad:" + name in fx and p.grad is not None:
            assert rel_err(p.grad, torch.from_numpy(fx["grad:" + name])) < 1e-3, name
    assert ms.encoders[0].layers[0].weight.grad is None  # only the selected frame's encoder takes part


@pytest.mark.gpu
def test_tcnn_style_models_match_oracle_grid_plus_torch_mlp():
    """TcnnHashMLP / MultiHashMLP / HashSirenNet on the kernels vs the oracle's hash grid (reference Python semantics,
    tcnn geometry rule) + plain torch layers on the same weights: outputs and every gradient within 1e-3."""
    from mri_interpolation_b200 import models
    from oracle import hashgrid
    torch.manual_seed(3)
    net = models.TcnnHashMLP(3, 5, 2, 11, 6, 1.6, dim_hidden=32)
    gen = torch.Generator().manual_seed(4)
    with torch.no_grad():
        for lv in net.encoder.levels:
            lv.embedding.weight.copy_(torch.randn(lv.embedding.weight.shape, generator=gen) * 0.3)
    levels = [hashgrid.Level((int(lv.resolution),) * 3, lv.hashmap_size) for lv in net.encoder.levels]
    ref = {k: v.detach().clone().requires_grad_() for k, v in net.state_dict().items() if not k.startswith("layers.")}
    x, y = torch.rand(500, 3, generator=gen), torch.rand(500, 1, generator=gen)

    def ref_forward(xx):
        tables = [ref[f"encoder.levels.{l}.embedding.weight"] for l in range(5)]
        z = hashgrid.encode(xx, tables, levels)
        n_lin = len(net.decoder.layers)
        for i in range(n_lin):
            z = F.linear(z, ref[f"decoder.layers.{i}.weight"])
            if i < n_lin - 1:
                z = F.relu(z)
        return z

    F.mse_loss(y, ref_forward(x)).backward()
    net = net.to(DEV)
    out = net(x.to(DEV))
    assert rel_err(out, ref_forward(x).detach()) < 1e-3
    net.training_step((x.to(DEV), y.to(DEV)), 0).backward()
    for name, p in net.named_parameters():
        if name in ref:
            assert rel_err(p.grad, ref[name].grad) < 1e-3, name
    # per-frame hash model: one training step touches the chosen frame's tables and the shared decoder only
    cfg = dict(encoding=dict(otype="HashGrid", n_levels=4, n_features_per_level=2, log2_hashmap_size=10, base_resolution=8,
                             per_level_scale=1.5, interpolation="Linear"),
               network=dict(otype="FullyFusedMLP", activation="ReLU", output_activation="None", n_neurons=32, n_hidden_layers=2))
    mh = models.MultiHashMLP(3, 1, 2, cfg, 1e-3).to(DEV)
    opt = mh.configure_optimizers()
    before = [lv.embedding.weight.detach().clone() for lv in mh.encoders[0].levels]
    xb, yb = torch.rand(1, 300, 3, device=DEV), torch.rand(1, 300, 1, device=DEV)
    mh.training_step((xb, yb, 1), 0).backward()
    assert all(float(lv.embedding.weight.grad.abs().max()) == 0.0 for lv in mh.encoders[0].levels)
    assert any(float(lv.embedding.weight.grad.abs().max()) > 0.0 for lv in mh.encoders[1].levels)
    opt.step()
    assert mh.predict_step((xb, yb, 1), 0).shape == (300, 1) and len(mh.get_latents()) == 1
    hs = models.HashSirenNet(cfg, dim_in=3, dim_hidden=32, n_layers=2).to(DEV)
    hs.training_step((xb[0], yb[0]), 0).backward()
    assert all(p.grad is not None for n_, p in hs.named_parameters() if n_.startswith(("encoding", "modulator", "siren")))
