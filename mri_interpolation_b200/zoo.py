"""The rest of the reference's model zoo that sits on the hot-path operators, re-targeted at the B200
kernels: modulated SIRENs and the tiny-cuda-nn shaped front-ends.

Mirrors (reference file:line):
  Modulator           models.py:236-260   ReLU MLP whose layer i sees [hidden_{i-1}, z]; returns every hidden
  ModulatedSirenNet   models.py:263-322   SIREN whose layer outputs are multiplied element-wise by the modulator's
  HashSirenNet        models.py:325-394   same, the modulator fed by a hash-grid encoding of the coordinates
  TcnnHashMLP         models.py:587-655   hash grid + fully fused ReLU MLP, configured the tiny-cuda-nn way
  MultiSiren          models.py:888-956   one SIREN encoder per frame + a shared SIREN decoder
  MultiHashMLP        models.py:959-1027  one hash grid per frame + a shared MLP decoder

The reference builds the last four on `tinycudann` objects (`tcnn.Encoding`, `tcnn.Network`), whose import is commented
out (models.py:10): as shipped they raise NameError.  Here `TcnnStyleEncoding` / `TcnnStyleNetwork` accept the same
config dictionaries (config/hash_config.json) and run on this package's kernels:
  * the grid uses the reference's PYTHON grid semantics (encoding.py: every level hashed, no +0.5 offset, true modulo)
    with tcnn's geometry rule res_l = floor(base * per_level_scale^l) - tiny-cuda-nn's own dense-coarse-level indexing
    is a different function and is NOT reproduced: there is no reference output to pin it against (unpinned);
  * the network is a bias-free MLP like tcnn's FullyFusedMLP, every layer one fused dense kernel.
Constructors create their parameters with the same torch calls in the same order as the reference wherever the
reference's code is pure torch (Modulator, ModulatedSirenNet, MultiSiren), so seeded state_dicts are identical.
"""
from __future__ import annotations

import math
from typing import Sequence

import torch
from torch import nn

from . import functional as Fn
from ._lib import ACT_GELU, ACT_IDENTITY, ACT_RELU, ACT_SINE
from .encoding import _HashGrid, _MultiResBase
from .models import BaseMLP, SirenNet, cast_tuple
from .optim import FusedAdam
from .pl_compat import pl


class Modulator(nn.Module):
    """Modulator of 'Modulated periodic activations for generalizable local functional representations' (models.py:236-260)."""

    def __init__(self, dim_in, dim_hidden, n_layers):
        super().__init__()
        self.layers = nn.ModuleList([])
        for ind in range(n_layers):
            is_first = ind == 0
            dim = dim_in if is_first else (dim_hidden + dim_in)
            self.layers.append(nn.Sequential(nn.Linear(dim, dim_hidden), nn.ReLU()))

    def forward(self, z):
        x = z
        hiddens = []
        for layer in self.layers:
            lin = layer[0]
            x = Fn.dense(x, lin.weight, lin.bias, ACT_RELU)  # Linear + ReLU in one kernel
            hiddens.append(x)
            x = torch.cat((x, z), dim=1)
        return tuple(hiddens)


def _modulated_forward(siren: SirenNet, mods, x, n_layers: int):
    mods = cast_tuple(mods, n_layers)
    for layer, mod in zip(siren.layers, mods):
        x = layer(x) * mod
    return siren.last_layer(x)


class ModulatedSirenNet(SirenNet):
    """SIREN with every sine layer multiplied element-wise by the matching modulator layer (models.py:263-322)."""

    def __init__(self, dim_in: int = 3, dim_hidden: int = 64, dim_out: int = 1, n_layers: int = 4, w0: float = 30.0,
                 w0_initial: float = 30.0, sigma: float = 6.0, use_bias: bool = True, final_activation: nn = None,
                 lr: float = 1e-4):
        super().__init__()  # the reference builds a default SirenNet first (and draws its parameters from the RNG)
        self.dim_in, self.dim_hidden, self.dim_out, self.n_layers = dim_in, dim_hidden, dim_out, n_layers
        self.w0, self.w0_initial, self.sigma, self.use_bias = w0, w0_initial, sigma, use_bias
        self.final_activation = final_activation
        self.lr = lr
        self.losses = []
        self.modulator = Modulator(dim_in=dim_in, dim_hidden=dim_hidden, n_layers=n_layers)
        self.siren = SirenNet(dim_in=dim_in, dim_hidden=dim_hidden, dim_out=dim_out, n_layers=n_layers, w0=w0,
                              w0_initial=w0_initial, sigma=sigma, use_bias=use_bias, final_activation=final_activation, lr=lr)

    def forward(self, x):
        return _modulated_forward(self.siren, self.modulator(x), x, self.n_layers)


# --------------------------------------------------------------------------- tiny-cuda-nn shaped building blocks
class TcnnStyleEncoding(_MultiResBase, nn.Module):
    """`tcnn.Encoding(n_input_dims, encoding_config)` look-alike for `"otype": "HashGrid"` configs.

    res_l = floor(base_resolution * per_level_scale^l), rows_l = min(res_l^D, 2^log2_hashmap_size), F features per
    level; tables initialised like the reference's Python grid (N(0,1) draw, then U(-1e-4, 1e-4)).  `n_output_dims`
    = n_levels * n_features_per_level.  Runs on the hash-grid kernels with the reference's Python grid semantics."""

    def __init__(self, n_input_dims: int, encoding_config: dict, dtype=torch.float32):
        nn.Module.__init__(self)
        otype = str(encoding_config.get("otype", "HashGrid"))
        if otype.lower() not in ("hashgrid", "grid"):
            raise NotImplementedError(f"TcnnStyleEncoding covers HashGrid encodings, not {otype!r}")
        if dtype != torch.float32:
            raise NotImplementedError("the B200 hash-grid kernels are fp32")
        interp = str(encoding_config.get("interpolation", "Linear")).lower()
        if interp != "linear":
            raise NotImplementedError(f"interpolation {interp!r}: the kernels interpolate linearly (encoding.py:108-128)")
        self.dim = int(n_input_dims)
        self.n_levels = int(encoding_config["n_levels"])
        self.n_features_per_level = int(encoding_config["n_features_per_level"])
        self.log2_hashmap_size = int(encoding_config["log2_hashmap_size"])
        self.base_resolution = int(encoding_config["base_resolution"])
        self.per_level_scale = float(encoding_config.get("per_level_scale", 2.0))
        levels = []
        for l in range(self.n_levels):
            res = math.floor(self.base_resolution * (self.per_level_scale ** l))
            rows = min(res ** self.dim, 2 ** self.log2_hashmap_size)
            levels.append(_HashGrid(dim=self.dim, n_features=self.n_features_per_level, hashmap_size=rows, resolution=res))
        self._finish(levels)
        self.n_input_dims = self.dim
        self.n_output_dims = self.output_dim


_TCNN_ACT = {"relu": ACT_RELU, "none": ACT_IDENTITY, "identity": ACT_IDENTITY, "sine": ACT_SINE, "gelu": ACT_GELU}


class TcnnStyleNetwork(nn.Module):
    """`tcnn.Network(n_input_dims, n_output_dims, network_config)` look-alike (FullyFusedMLP / CutlassMLP): bias-free
    Linear layers, `n_hidden_layers` hidden layers of `n_neurons`, `activation` / `output_activation`."""

    def __init__(self, n_input_dims: int, n_output_dims: int, network_config: dict):
        super().__init__()
        self.n_input_dims, self.n_output_dims = int(n_input_dims), int(n_output_dims)
        width = int(network_config.get("n_neurons", 64))
        n_hidden = int(network_config.get("n_hidden_layers", 2))
        try:
            self.act = _TCNN_ACT[str(network_config.get("activation", "ReLU")).lower()]
            self.out_act = _TCNN_ACT[str(network_config.get("output_activation", "None")).lower()]
        except KeyError as e:
            raise NotImplementedError(f"activation {e} is not fused by the dense kernel") from None
        dims = [self.n_input_dims] + [width] * n_hidden + [self.n_output_dims]
        self.layers = nn.ModuleList([nn.Linear(a, b, bias=False) for a, b in zip(dims, dims[1:])])

    def forward(self, x):
        last = len(self.layers) - 1
        for i, lin in enumerate(self.layers):
            x = Fn.dense(x, lin.weight, None, self.out_act if i == last else self.act, 1.0)
        return x


class HashSirenNet(SirenNet):
    """Modulated SIREN whose modulator reads a hash-grid encoding of the coordinates (models.py:325-394).
    ``config`` is the tiny-cuda-nn style dictionary of config/hash_config.json (its "encoding" entry)."""

    def __init__(self, config, dim_in: int = 3, dim_hidden: int = 64, dim_out: int = 1, n_layers: int = 4, w0: float = 30.0,
                 w0_initial: float = 30.0, sigma: float = 6.0, use_bias: bool = True, final_activation: nn = None,
                 lr: float = 1e-4):
        super().__init__()
        self.config = config  # the reference reads self.config without assigning it (models.py:371): intended
        self.dim_in, self.dim_hidden, self.dim_out, self.n_layers = dim_in, dim_hidden, dim_out, n_layers
        self.w0, self.w0_initial, self.sigma, self.use_bias = w0, w0_initial, sigma, use_bias
        self.final_activation = final_activation
        self.lr = lr
        self.losses = []
        self.encoding = TcnnStyleEncoding(n_input_dims=dim_in, encoding_config=config["encoding"], dtype=torch.float32)
        self.modulator = Modulator(dim_in=config["encoding"]["n_levels"] * config["encoding"]["n_features_per_level"],
                                   dim_hidden=dim_hidden, n_layers=n_layers)
        self.siren = SirenNet(dim_in=dim_in, dim_hidden=dim_hidden, dim_out=dim_out, n_layers=n_layers, w0=w0,
                              w0_initial=w0_initial, use_bias=use_bias, final_activation=final_activation, lr=lr)

    def forward(self, x):
        return _modulated_forward(self.siren, self.modulator(self.encoding(x)), x, self.n_layers)


class TcnnHashMLP(BaseMLP):
    """Hash grid + fully fused ReLU MLP configured the tiny-cuda-nn way (models.py:587-655).  Like the reference the
    decoder has ``self.n_layers`` hidden layers, i.e. BaseMLP's default (8)."""

    def __init__(self, dim_in: int, n_levels: int, n_features_per_level: int, log2_hashmap_size: int, base_resolution: int,
                 per_level_scale: float, interplation_method: str = 'linear', dim_hidden: int = 64, dim_out: int = 1,
                 lr: float = 1e-4):
        super().__init__()
        self.dim_in, self.n_levels, self.n_features_per_level = dim_in, n_levels, n_features_per_level
        self.log2_hashmap_size, self.base_resolution, self.per_level_scale = log2_hashmap_size, base_resolution, per_level_scale
        self.interpolation_method = interplation_method
        self.dim_hidden, self.dim_out, self.lr = dim_hidden, dim_out, lr
        self.latents = []
        self.encoder = TcnnStyleEncoding(n_input_dims=dim_in, encoding_config={
            "otype": "HashGrid", "n_levels": n_levels, "n_features_per_level": n_features_per_level,
            "log2_hashmap_size": log2_hashmap_size, "base_resolution": base_resolution, "per_level_scale": per_level_scale,
            "interpolation": interplation_method}, dtype=torch.float32)
        self.decoder = TcnnStyleNetwork(n_input_dims=self.encoder.n_output_dims, n_output_dims=dim_out, network_config={
            "otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": dim_hidden,
            "n_hidden_layers": self.n_layers})

    def forward(self, x):
        return self.decoder(self.encoder(x))

    def predict_step(self, batch, batch_idx):
        x, y = batch
        z = self.encoder(x)
        self.latents.append(z)
        return self.decoder(z)

    def get_latents(self):
        return self.latents


class _PerFrameModel(pl.LightningModule):
    """Shared Lightning glue of the legacy per-frame models (models.py:888-1027): batch = (x, y, frame_idx) with a leading
    batch axis of 1 (one batch = one whole frame), MSE(y_pred, y), Adam with weight_decay 1e-5."""

    def forward(self, x, frame_idx):
        return self.decoder(self.encoders[int(frame_idx)](x))

    def configure_optimizers(self):
        self.optimizer = FusedAdam(self.parameters(), lr=self.lr, weight_decay=1e-5)
        return self.optimizer

    def _frame(self, batch):
        x, y, frame_idx = batch
        return x.squeeze(0), y.squeeze(0), int(frame_idx)

    def training_step(self, batch, batch_idx):
        x, y, frame_idx = self._frame(batch)
        y_pred = self.decoder(self.encoders[frame_idx](x))
        loss = Fn.mse_loss(y_pred, y)
        self.losses.append(loss.detach())
        self.log("train_loss", loss)
        return loss

    def predict_step(self, batch, batch_idx):
        x, y, frame_idx = self._frame(batch)
        z = self.encoders[frame_idx](x)
        if hasattr(self, "latents"):
            self.latents.append(z)
        return self.decoder(z)


class MultiSiren(_PerFrameModel):
    """One SIREN encoder per frame + a shared SIREN decoder (models.py:888-956)."""

    def __init__(self, dim_in, dim_hidden, dim_out, n_layers, n_frames, lr, *args, **kwargs):
        super().__init__()
        self.dim_in, self.dim_hidden, self.dim_out = dim_in, dim_hidden, dim_out
        self.n_layers, self.n_frames, self.lr = n_layers, n_frames, lr
        self.losses = []
        self.encoders = nn.ModuleList()
        for _ in range(n_frames):
            self.encoders.append(SirenNet(dim_in=dim_in, dim_hidden=dim_hidden, dim_out=dim_hidden, n_layers=n_layers))
        self.decoder = SirenNet(dim_in=dim_hidden, dim_hidden=dim_hidden, dim_out=dim_out, n_layers=n_layers)
        self.automatic_optimization = True


class MultiHashMLP(_PerFrameModel):
    """One hash grid per frame + a shared MLP decoder, tiny-cuda-nn style config (models.py:959-1027)."""

    def __init__(self, dim_in, dim_out, n_frames, config, lr, *args, **kwargs):
        super().__init__()
        self.config = config
        self.dim_in, self.dim_out, self.n_frames, self.lr = dim_in, dim_out, n_frames, lr
        self.losses = []
        self.latents = []
        self.encoders = nn.ModuleList()
        for _ in range(n_frames):
            self.encoders.append(TcnnStyleEncoding(n_input_dims=dim_in, encoding_config=config["encoding"]))
        self.decoder = TcnnStyleNetwork(
            n_input_dims=config["encoding"]["n_levels"] * config["encoding"]["n_features_per_level"],
            n_output_dims=dim_out, network_config=config["network"])
        self.automatic_optimization = True

    def get_latents(self):
        return self.latents
