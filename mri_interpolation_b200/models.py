"""Coordinate-network models - the reference's ``models.py`` class surface on the B200 kernels.

Mirrors (reference file:line):
  BaseMLP      models.py:20-95    Lightning glue: training_step (MSE), configure_optimizers (Adam), predict_step
  Sine         models.py:108-114
  SirenLayer   models.py:117-156  init U(+-1/dim_in) (first) or U(+-sqrt(sigma/dim_in)/w0), weight then bias
  SirenNet     models.py:160-233  n_layers sine layers + identity-activated last layer (no w0 on the last)
  HashMLP      models.py:658-754  hash-grid encoder + [Linear -> (BatchNorm1d) -> GELU -> Dropout] blocks;
                                  forward = encoder then block-by-block loop (nb cell 37,
                                  legacy_code/hash_experimentation.py:237-241; the shipped forward calls a
                                  ModuleList and cannot run)

Constructors consume the torch RNG in the reference's order, so ``torch.manual_seed(1337)`` gives the
same initial state_dict (same keys, shapes and values) as the reference.  All compute runs through
the C ABI (csrc/*.cu); inputs/parameters must be CUDA fp32 - there is no CPU fallback.
"""
from __future__ import annotations

import math
import os
from typing import Any, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import encoding
from . import functional as Fn
from ._lib import ACT_GELU, ACT_IDENTITY, ACT_RELU, ACT_SINE
from .optim import FusedAdam
from .pl_compat import pl


def _fusable_activation(mod) -> Any:
    """Activation code if `mod` is an activation the dense kernel fuses, else None."""
    if isinstance(mod, nn.Identity):
        return ACT_IDENTITY
    if isinstance(mod, nn.ReLU):
        return ACT_RELU
    if isinstance(mod, nn.GELU) and getattr(mod, "approximate", "none") == "none":
        return ACT_GELU
    return None


def sync_batchnorm_(module: nn.Module) -> int:
    """Under data-parallel training (torch.distributed initialised, world_size > 1) replace every BatchNorm1d by a
    SyncBatchNorm that shares its parameters and buffers (same state_dict keys).  The shipped HashMLP decoder has
    BatchNorm (models.py:718-739): per-rank batch statistics would make W-GPU training differ from single-GPU training
    on the global batch, and the running statistics - buffers, which the gradient arena does not synchronise - would
    drift apart between the replicas (a sharded sweep would then stitch slabs of slightly different decoders).
    Returns the number of layers converted."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return 0
    count = 0
    for parent in module.modules():
        for name, child in list(parent.named_children()):
            if isinstance(child, nn.modules.batchnorm._BatchNorm) and not isinstance(child, nn.SyncBatchNorm):
                parent._modules[name] = nn.SyncBatchNorm.convert_sync_batchnorm(child)
                count += 1
    return count


class BaseMLP(pl.LightningModule):
    """Fully connected network, base class of the other models (models.py:20-95)."""

    def __init__(
        self,
        dim_in: int = 2,
        dim_out: int = 1,
        dim_hidden: int = 128,
        n_layers: int = 8,
        activation: torch.nn = nn.ReLU,
        criterion: F = Fn.mse_loss,
        lr: float = 1e-4,
        *args,
        **kwargs,
    ) -> None:
        super().__init__()
        self.dim_in = dim_in
        self.dim_hidden = dim_hidden
        self.dim_out = dim_out
        self.n_layers = n_layers
        self.activation = activation
        self.criterion = criterion
        self.lr = lr

        layers = []
        for i in range(n_layers):
            layers.append(
                nn.Linear(
                    in_features=dim_in if i == 0 else dim_hidden,
                    out_features=dim_out if i == (n_layers - 1) else dim_hidden,
                )
            )
            layers.append(activation())
        self.layers = torch.nn.Sequential(*layers)

    def forward(self, x) -> Any:
        # the reference's BaseMLP.forward recurses forever (models.py:58-59); intended: the layer stack
        mods = list(self.layers)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Linear):
                act = _fusable_activation(mods[i + 1]) if i + 1 < len(mods) else None
                if act is not None:
                    x = Fn.dense(x, m.weight, m.bias, act)
                    i += 2
                    continue
                x = Fn.dense(x, m.weight, m.bias, ACT_IDENTITY)
            else:
                x = m(x)
            i += 1
        return x

    def training_step(self, batch, batch_idx) -> torch.FloatTensor:
        x, y = batch
        y_pred = self.forward(x)
        loss = self.criterion(y, y_pred)
        self.log("train_loss", loss)
        return loss

    def configure_optimizers(self):
        # torch.optim.Adam(self.parameters(), lr) semantics (models.py:68-70), one fused kernel
        sync_batchnorm_(self)
        self.optimizer = FusedAdam(self.parameters(), lr=self.lr)
        return self.optimizer

    def predict_step(self, batch, batch_idx):
        x, y = batch
        return self(x)

    def lr_schedulers(self):
        self.scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer=self.optimizer, T_max=10)
        return self.scheduler

    def set_parameters(self, theta):
        """Manually set parameters from a matching list (models.py:87-95; used for meta learning)."""
        p_dict = self.state_dict()
        for p, thet in zip(p_dict, theta):
            p_dict[p] = thet.data
        self.load_state_dict(p_dict)
        self.eval()
        self.train()


# utils for siren
def exists(val):
    return val is not None


def cast_tuple(val, repeat=1):
    return val if isinstance(val, tuple) else ((val,) * repeat)


class Sine(nn.Module):
    def __init__(self, w0=30.0):
        super().__init__()
        self.w0 = w0

    def forward(self, x):
        return torch.sin(self.w0 * x)


class SirenLayer(nn.Module):
    """One SIREN layer (models.py:117-156): activation(x W^T + b), sine by default, fused in one kernel."""

    def __init__(
        self,
        dim_in: int,
        dim_out: int = 1,
        w0: float = 30.0,
        sigma: float = 6.0,
        is_first: bool = False,
        use_bias: bool = True,
        activation: nn = None,
    ):
        super().__init__()
        self.dim_in = dim_in
        self.is_first = is_first

        weight = torch.zeros(dim_out, dim_in)
        bias = torch.zeros(dim_out) if use_bias else None
        self.init_(weight, bias, sigma=sigma, w0=w0)

        self.weight = nn.Parameter(weight)
        self.bias = nn.Parameter(bias) if use_bias else None
        self.activation = Sine(w0) if activation is None else activation

    def init_(self, weight, bias, sigma, w0):
        dim = self.dim_in
        w_std = (1 / dim) if self.is_first else (math.sqrt(sigma / dim) / w0)
        weight.uniform_(-w_std, w_std)
        if exists(bias):
            bias.uniform_(-w_std, w_std)

    def forward(self, x):
        act = self.activation
        if isinstance(act, Sine):
            return Fn.dense(x, self.weight, self.bias, ACT_SINE, act.w0)
        code = _fusable_activation(act)
        if code is not None:
            return Fn.dense(x, self.weight, self.bias, code)
        return act(Fn.dense(x, self.weight, self.bias, ACT_IDENTITY))


class SirenNet(BaseMLP):
    """SIREN (models.py:160-233): implicit representation with periodic activations.

    dim_in/dim_hidden/dim_out/n_layers, w0 (hidden layers), w0_initial (first layer), sigma (init
    spread), use_bias, final_activation (None = identity), lr.  ``layers`` holds the sine layers,
    ``last_layer`` the output layer, ``losses`` a free list the callers append to.
    """

    def __init__(
        self,
        dim_in: int = 3,
        dim_hidden: int = 64,
        dim_out: int = 1,
        n_layers: int = 4,
        w0: float = 30.0,
        w0_initial: float = 30.0,
        sigma: float = 6.0,
        use_bias: bool = True,
        final_activation: nn = None,
        lr: float = 1e-4,
        *args,
        **kwargs,
    ):
        # BaseMLP.__init__() runs with its defaults first (models.py:192) and draws its 8 Linear
        # layers from the RNG; they are then replaced by the SIREN layers, as in the reference.
        super().__init__()
        self.n_layers = n_layers
        self.dim_hidden = dim_hidden
        self.sigma = sigma
        self.losses = []
        self.lr = lr

        self.layers = nn.ModuleList([])
        for ind in range(n_layers):
            is_first = ind == 0
            layer_w0 = w0_initial if is_first else w0
            layer_dim_in = dim_in if is_first else dim_hidden
            self.layers.append(
                SirenLayer(dim_in=layer_dim_in, dim_out=dim_hidden, w0=layer_w0, sigma=self.sigma,
                           use_bias=use_bias, is_first=is_first)
            )

        final_activation = nn.Identity() if not exists(final_activation) else final_activation
        self.last_layer = SirenLayer(dim_in=dim_hidden, dim_out=dim_out, w0=w0, sigma=self.sigma,
                                     use_bias=use_bias, activation=final_activation)
        # "auto": hidden->hidden layers that fit the tcgen05 tiles (width % 128 == 0) run on the tensor cores in
        # the split-precision fp32-parity mode "bf16x3"; "fp32" forces the CUDA-core path; "bf16" = 1 MMA/slice.
        self.precision = kwargs.get("precision", "auto")

    def _tensor_core_mode(self):
        mode = self.precision
        if mode == "fp32":
            return None
        from . import siren_fused
        ok = self.__dict__.get("_tc_eligible")
        if ok is None:
            ok = siren_fused.eligible(self)
            self.__dict__["_tc_eligible"] = ok
        if not ok:
            if mode in ("bf16x3", "bf16"):
                raise RuntimeError("this SirenNet does not fit the tensor-core tiles (equal hidden widths >= 160, or a multiple of 128, are needed)")
            return None
        return "bf16x3" if mode == "auto" else mode

    def forward(self, x):
        # d/dx exists on the fp32 CUDA-core path only: a caller that differentiates with respect to the coordinates gets
        # the same behaviour whatever the hidden width is
        mode = None if (torch.is_grad_enabled() and x.requires_grad) else self._tensor_core_mode()
        if mode is not None:
            from . import siren_fused
            return siren_fused.forward(self, x, mode)
        for layer in self.layers:
            x = layer(x)
        return self.last_layer(x)


class HashMLP(BaseMLP):
    """Hash-grid encoder + small MLP decoder (models.py:658-754).

    ``base_resolution`` int -> ``MultiResHashGrid``; sequence -> ``MultiResHashGridV2`` (:691-708).
    ``batch_norm=True`` reproduces the shipped decoder blocks Linear -> BatchNorm1d -> activation ->
    Dropout (:718-739); ``batch_norm=False`` is the notebook's Linear -> activation variant (nb cell 37).
    ``spectral_norm=True`` is the legacy recipe (legacy_code/hash_experimentation.py:213-234, the block the shipped
    file keeps commented at :720-730): every Linear wrapped in ``torch.nn.utils.parametrizations.spectral_norm(...,
    n_power_iterations=4, eps=1e-12, dim=None)``; the power iteration and W / sigma stay torch's (two matrix-vector
    products on a 64 x 32 weight), the Linear itself runs on this package's dense kernel with the normalised weight;
    ``weight_decay`` (legacy: 1e-5, L2-coupled Adam, :244-246) goes to the fused Adam kernel.
    ``n_layers`` reaches BaseMLP through **kwargs exactly like the reference (default 8).
    """

    def __init__(self,
                 dim_in: int,
                 n_levels: int,
                 n_features_per_level: int,
                 log2_hashmap_size: int,
                 base_resolution: Tuple[int, ...],
                 finest_resolution: Tuple[int, ...],
                 interplation_method: str = 'linear',
                 dim_hidden: int = 64,
                 dim_out: int = 1,
                 activation: nn = nn.GELU,
                 dropout: float = 0.0,
                 lr: float = 1e-4,
                 *args,
                 batch_norm: bool = True,
                 spectral_norm: bool = False,
                 weight_decay: float = 0.0,
                 **kwargs):
        super().__init__(*args, **kwargs)
        self.spectral_norm = spectral_norm
        self.weight_decay = weight_decay
        self.dim_in = dim_in
        self.n_levels = n_levels
        self.n_features_per_level = n_features_per_level
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.finest_resolution = finest_resolution
        self.interpolation_method = interplation_method
        self.dim_hidden = dim_hidden
        self.dim_out = dim_out
        self.dropout = dropout
        self.lr = lr
        self.batch_norm = batch_norm
        self.latents = []  # encoder outputs kept by predict_step for visualisation (:689,749)
        self.keep_latents = True
        self.fuse_backward = os.environ.get("MRI_FUSED_BACKWARD", "1") == "1"
        # forward + loss + backward as ONE kernel (fused_training_step): opt-in.  Measured on B200 (G4, 2^19 coords): 1.335 ms
        # against 0.23 + 0.43 ms for the forward and backward kernels - at the backward's 12 warps per SM the gather's L2
        # latency is exposed instead of overlapping the scatter (profiles/README.md), so two kernels stay the default.
        self.fuse_step = os.environ.get("MRI_FUSED_STEP", "0") == "1"
        # training_step + backward as three direct kernel calls (no autograd engine): default on
        self.direct_step = os.environ.get("MRI_DIRECT_STEP", "1") == "1"

        if isinstance(self.base_resolution, int):
            self.encoder = encoding.MultiResHashGrid(
                dim=self.dim_in, n_levels=self.n_levels, n_features_per_level=self.n_features_per_level,
                log2_hashmap_size=self.log2_hashmap_size, base_resolution=self.base_resolution,
                finest_resolution=self.finest_resolution)
        else:
            self.encoder = encoding.MultiResHashGridV2(
                dim=self.dim_in, n_levels=self.n_levels, n_features_per_level=self.n_features_per_level,
                log2_hashmap_size=self.log2_hashmap_size, base_resolution=self.base_resolution,
                finest_resolution=self.finest_resolution)

        self.encoding_dim_out = self.n_levels * self.n_features_per_level

        self.decoder = torch.nn.ModuleList()
        for i in range(self.n_layers):
            in_features = self.encoding_dim_out if i == 0 else self.dim_hidden
            out_features = self.dim_out if i == (self.n_layers - 1) else self.dim_hidden
            lin = torch.nn.Linear(in_features=in_features, out_features=out_features)
            if spectral_norm:
                lin = torch.nn.utils.parametrizations.spectral_norm(lin, n_power_iterations=4, eps=1e-12, dim=None)
            mods = [lin]
            if batch_norm:
                mods.append(torch.nn.BatchNorm1d(num_features=out_features))
            mods.append(activation())
            if batch_norm:
                mods.append(torch.nn.Dropout(p=dropout, inplace=False))
            self.decoder.append(torch.nn.Sequential(*mods))

    def configure_optimizers(self):
        # models.py:68-70; the legacy recipe's Adam(weight_decay=1e-5) (legacy_code/hash_experimentation.py:244-246) when asked
        sync_batchnorm_(self)
        self.optimizer = FusedAdam(self.parameters(), lr=self.lr, weight_decay=self.weight_decay)
        return self.optimizer

    @staticmethod
    def _run_block(block, x):
        mods = list(block)
        lin = mods[0]
        rest = mods[1:]
        if len(rest) >= 1:
            act = _fusable_activation(rest[0])
            tail_inert = all(isinstance(m, nn.Dropout) and (m.p == 0.0 or not m.training) for m in rest[1:])
            if act is not None and tail_inert:
                return Fn.dense(x, lin.weight, lin.bias, act)
        x = Fn.dense(x, lin.weight, lin.bias, ACT_IDENTITY)
        for m in rest:  # BatchNorm1d & co: batch statistics need a second pass
            x = m(x)
        return x

    def _fused_decoder_plan(self):
        """(lin1, lin2, act1, act2) when the decoder is two Linear+activation blocks with one output: the shape the
        stand-alone fused decoder (csrc/decoder.cu, K0 in {16,32,64}, H in {32,64}) and / or the encoder+decoder kernels
        (csrc/hashdecoder_*.cu, F = 2, L in {4,8,16}, H in {64,128}) cover; each caller checks its own kernel's geometry."""
        plan = self.__dict__.get("_decoder_plan", 0)
        if plan != 0:
            return plan
        plan = None
        if len(self.decoder) == 2:
            blocks = [list(b) for b in self.decoder]
            ok = all(len(b) >= 2 and isinstance(b[0], nn.Linear) and b[0].bias is not None and
                     _fusable_activation(b[1]) is not None and
                     all(isinstance(m, nn.Dropout) and m.p == 0.0 for m in b[2:]) for b in blocks)
            if ok:
                l1, l2 = blocks[0][0], blocks[1][0]
                a1, a2 = _fusable_activation(blocks[0][1]), _fusable_activation(blocks[1][1])
                if l2.out_features == 1 and l2.in_features == l1.out_features and a1 in (ACT_GELU, ACT_RELU):
                    plan = (l1, l2, a1, a2)
        self.__dict__["_decoder_plan"] = plan
        return plan

    def decode(self, z):
        plan = self._fused_decoder_plan() if z.is_cuda else None
        if plan is not None and Fn.decoder2_supported(plan[0].in_features, plan[0].out_features, plan[2]):
            l1, l2, a1, a2 = plan
            return Fn.Decoder2Fn.apply(z, l1.weight, l1.bias, l2.weight, l2.bias, a1, a2)
        for block in self.decoder:
            z = self._run_block(block, z)
        return z

    def forward(self, x):
        plan = self._fused_decoder_plan() if x.is_cuda else None
        enc = self.encoder
        if (plan is not None and self.fuse_backward and getattr(enc, "_resolutions", None) is not None
                and Fn.hashdecoder_supported(enc.dim, enc.n_levels, enc.n_features_per_level, plan[0].out_features, plan[2])):
            # encoder + decoder as one autograd node: the backward is ONE kernel (decoder backward + table scatter)
            l1, l2, a1, a2 = plan
            return Fn.HashDecoderFn.apply(x, enc, l1.weight, l1.bias, l2.weight, l2.bias, a1, a2, *enc.tables())
        return self.decode(self.encoder(x))

    def fused_training_step(self, batch, batch_idx):
        """``training_step`` + ``loss.backward()`` without the autograd engine: returns the loss (a detached 0-dim tensor,
        logged as ``train_loss``) with the gradients already accumulated in the parameters' .grad buffers, or None when
        this model / batch has no such path (the caller then runs training_step + backward).  Two variants:
        * direct (default, ``MRI_DIRECT_STEP=0`` switches it off): the SAME three kernels autograd would launch - fused
          forward, MSE, fused backward - called back to back; bit-identical results, a quarter of the Python time;
        * one kernel for all of it (csrc/hashdecoder_step.cu; opt-in with ``MRI_FUSED_STEP=1`` / ``fuse_step = True``,
          headline geometry, measured slower on the GPU).
        The Trainer stand-in asks for this path by itself (pl_compat.training_step_and_backward); models.py:61-66, 741-744."""
        x, y = batch
        plan = self._fused_decoder_plan() if (x.is_cuda and self.fuse_backward and (self.fuse_step or self.direct_step)) else None
        enc = self.encoder
        if (plan is None or not torch.is_grad_enabled() or getattr(enc, "_resolutions", None) is None
                or getattr(enc, "_grad_group_hook", None) is not None or self.criterion is not Fn.mse_loss
                or y.numel() != x.reshape(-1, x.shape[-1]).shape[0] or x.requires_grad
                or not Fn.hashdecoder_supported(enc.dim, enc.n_levels, enc.n_features_per_level, plan[0].out_features, plan[2])):
            return None
        l1, l2, a1, a2 = plan
        tabs = enc.tables()
        params = list(tabs) + [l1.weight, l1.bias, l2.weight, l2.bias]
        if not all(p.is_leaf and p.requires_grad for p in params):
            return None
        one_kernel = (self.fuse_step
                      and Fn.hashmlp_mse_step_supported(enc.dim, enc.n_levels, enc.n_features_per_level, plan[0].out_features, plan[2])
                      # the one-kernel step takes ONE level layout for the tables and their gradients (true for the flat arenas)
                      and not any(t.grad is None or t.grad.data_ptr() - tabs[0].grad.data_ptr() != t.data_ptr() - tabs[0].data_ptr()
                                  for t in tabs))
        if one_kernel:
            loss, _ = Fn.hashmlp_mse_step(x, y, enc, l1.weight, l1.bias, l2.weight, l2.bias, a1, a2)
        elif self.direct_step:
            loss = Fn.hashmlp_mse_direct_step(x, y, enc, l1.weight, l1.bias, l2.weight, l2.bias, a1, a2)
        else:
            return None
        self.log("train_loss", loss)
        return loss

    def predict_step(self, batch, batch_idx):
        x, y = batch
        z = self.encoder(x)
        if self.keep_latents:
            self.latents.append(z)
        return self.decode(z)

    def get_latents(self):
        return self.latents


class _OutOfScope(pl.LightningModule):
    """Reference model-zoo variants that are not on the north-star hot path (point-spread-function, random
    Fourier feature and Gabor experiments; rff / PSF helpers are absent): the names stay importable, construction raises."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            f"{type(self).__name__} is outside the B200 hot-path scope (SIREN, hash-grid + MLP, dense sweep, "
            f"Adam); see DESIGN.md 'Out of scope'")


class PsfSirenNet(_OutOfScope): ...
class RffNet(_OutOfScope): ...
class RealGaborLayer(_OutOfScope): ...
class ComplexGaborLayer(_OutOfScope): ...
class GaborNet(_OutOfScope): ...


# the model-zoo variants that sit on the hot-path operators (modulated SIRENs, tiny-cuda-nn shaped front-ends) live in
# zoo.py; imported last because zoo.py builds on the classes above
from .zoo import (HashSirenNet, ModulatedSirenNet, Modulator, MultiHashMLP, MultiSiren, TcnnHashMLP,  # noqa: E402,F401
                  TcnnStyleEncoding, TcnnStyleNetwork)
