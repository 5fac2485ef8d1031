"""Oracle: SIREN sine MLP, HashMLP decoder, MSE loss, Adam (CPU, torch fp32).

Restates reference ``models.py``:
  * Sine / SirenLayer (init_, forward)   models.py:108-156
  * SirenNet (layers + last_layer)       models.py:160-233
  * BaseMLP default ReLU stack (RNG use) models.py:46-56
  * BaseMLP.training_step (MSE)          models.py:61-66
  * configure_optimizers (Adam defaults) models.py:68-70
  * HashMLP decoder                      models.py:712-739, intended forward =
    notebook cell 37 / legacy_code/hash_experimentation.py:237-241
Test infrastructure only - never imported by the product package.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import hashgrid


# --------------------------------------------------------------------------- init
def consume_basemlp_rng(dim_in=2, dim_out=1, dim_hidden=128, n_layers=8) -> Dict[str, torch.Tensor]:
    """Replay the RNG draws of BaseMLP.__init__'s default ``nn.Linear`` stack (models.py:46-56).

    SirenNet calls ``super().__init__()`` with no arguments (models.py:192) so the
    defaults above apply; HashMLP forwards ``n_layers`` & co through **kwargs.
    Returns the (dead) parameters with the reference's state_dict keys.
    """
    dead = {}
    for i in range(n_layers):
        lin = torch.nn.Linear(dim_in if i == 0 else dim_hidden, dim_out if i == n_layers - 1 else dim_hidden)
        dead[f"layers.{2 * i}.weight"] = lin.weight.detach()
        dead[f"layers.{2 * i}.bias"] = lin.bias.detach()
    return dead


def siren_layer_init(dim_in: int, dim_out: int, w0: float, sigma: float, is_first: bool, use_bias: bool = True):
    """models.py:136-151: weight then bias drawn from U(-b, b); b = 1/dim (first) else sqrt(sigma/dim)/w0."""
    bound = (1 / dim_in) if is_first else (math.sqrt(sigma / dim_in) / w0)
    weight = torch.zeros(dim_out, dim_in)
    bias = torch.zeros(dim_out) if use_bias else None
    weight.uniform_(-bound, bound)
    if bias is not None:
        bias.uniform_(-bound, bound)
    return weight, bias


def siren_init(dim_in=3, dim_hidden=64, dim_out=1, n_layers=4, w0=30.0, w0_initial=30.0, sigma=6.0, use_bias=True):
    """SirenNet parameter construction in the reference's RNG order (models.py:192-228).

    Returns (params, w0s): params is an ordered dict with the reference's state_dict keys
    (layers.{i}.weight/bias, last_layer.weight/bias); w0s lists the per-hidden-layer w0.
    """
    consume_basemlp_rng()  # models.py:192 - discarded, overwritten by self.layers = ModuleList
    params: Dict[str, torch.Tensor] = {}
    w0s = []
    for i in range(n_layers):
        first = i == 0
        layer_w0 = w0_initial if first else w0
        w, b = siren_layer_init(dim_in if first else dim_hidden, dim_hidden, layer_w0, sigma, first, use_bias)
        params[f"layers.{i}.weight"] = w
        if b is not None:
            params[f"layers.{i}.bias"] = b
        w0s.append(layer_w0)
    w, b = siren_layer_init(dim_hidden, dim_out, w0, sigma, False, use_bias)
    params["last_layer.weight"] = w
    if b is not None:
        params["last_layer.bias"] = b
    return params, w0s


# ------------------------------------------------------------------------ forward
def siren_forward(x: torch.Tensor, params: Dict[str, torch.Tensor], w0s: Sequence[float]) -> torch.Tensor:
    """models.py:230-233 with SirenLayer.forward (153-156) and Sine (113-114).

    Hidden layers: sin(w0 * (x W^T + b)); last layer: plain affine (Identity activation, no w0).
    """
    h = x
    for i, w0 in enumerate(w0s):
        h = torch.sin(w0 * F.linear(h, params[f"layers.{i}.weight"], params.get(f"layers.{i}.bias")))
    return F.linear(h, params["last_layer.weight"], params.get("last_layer.bias"))


def siren_backward(x, params, w0s, grad_y):
    """Closed-form gradients (autograd of models.py:153-156): returns dict of dW/db and dX.

    dPre = dOut * w0 cos(w0 pre); dX = dPre W; dW = dPre^T X; db = sum_N dPre.
    """
    acts = [x]
    pres = []
    h = x
    for i, w0 in enumerate(w0s):
        pre = F.linear(h, params[f"layers.{i}.weight"], params.get(f"layers.{i}.bias"))
        pres.append(pre)
        h = torch.sin(w0 * pre)
        acts.append(h)
    grads = {}
    grads["last_layer.weight"] = grad_y.t() @ h
    if "last_layer.bias" in params:
        grads["last_layer.bias"] = grad_y.sum(0)
    d = grad_y @ params["last_layer.weight"]
    for i in reversed(range(len(w0s))):
        dpre = d * (w0s[i] * torch.cos(w0s[i] * pres[i]))
        grads[f"layers.{i}.weight"] = dpre.t() @ acts[i]
        if f"layers.{i}.bias" in params:
            grads[f"layers.{i}.bias"] = dpre.sum(0)
        d = dpre @ params[f"layers.{i}.weight"]
    return grads, d


def decoder_init(enc_dim: int, dim_hidden: int, dim_out: int, n_layers: int):
    """Decoder ``nn.Linear`` draws in the reference order (models.py:712-739). Keys decoder.{i}.0.*"""
    params = {}
    for i in range(n_layers):
        lin = torch.nn.Linear(enc_dim if i == 0 else dim_hidden, dim_out if i == n_layers - 1 else dim_hidden)
        params[f"decoder.{i}.0.weight"] = lin.weight.detach()
        params[f"decoder.{i}.0.bias"] = lin.bias.detach()
    return params


def hashmlp_init(dim_in, n_levels, n_features_per_level, log2_hashmap_size, base_resolution, finest_resolution,
                 dim_hidden=64, dim_out=1, n_layers=8):
    """HashMLP parameter construction in the reference's RNG order (models.py:677-739).

    BaseMLP.__init__ runs first with (n_layers, ...) from **kwargs and its defaults
    dim_in=2, dim_hidden=128, dim_out=1 (models.py:25-56), then the per-level
    embeddings, then the decoder Linears.
    """
    params = dict(consume_basemlp_rng(n_layers=n_layers))
    levels = hashgrid.geometry(dim_in, n_levels, log2_hashmap_size, base_resolution, finest_resolution)
    for li, t in enumerate(hashgrid.init_tables(levels, n_features_per_level)):
        params[f"encoder.levels.{li}.embedding.weight"] = t
    params.update(decoder_init(n_levels * n_features_per_level, dim_hidden, dim_out, n_layers))
    return params, levels


def decoder_forward(z: torch.Tensor, params, n_layers: int, activation=F.gelu) -> torch.Tensor:
    """Intended HashMLP forward, BN/Dropout-free notebook variant (nb cell 37):
    every block is Linear -> GELU, including the output block."""
    h = z
    for i in range(n_layers):
        h = activation(F.linear(h, params[f"decoder.{i}.0.weight"], params[f"decoder.{i}.0.bias"]))
    return h


def hashmlp_forward(x, params, levels, n_layers: int, anisotropic: bool, activation=F.gelu):
    tables = [params[f"encoder.levels.{i}.embedding.weight"] for i in range(len(levels))]
    return decoder_forward(hashgrid.encode(x, tables, levels, anisotropic), params, n_layers, activation)


# --------------------------------------------------------------------- loss / Adam
def mse(y: torch.Tensor, y_pred: torch.Tensor) -> torch.Tensor:
    """models.py:64: criterion(y, y_pred) with F.mse_loss (mean over all elements)."""
    return F.mse_loss(y, y_pred)


def mse_grad(y: torch.Tensor, y_pred: torch.Tensor) -> torch.Tensor:
    """d mse / d y_pred = 2 (y_pred - y) / numel."""
    return 2.0 * (y_pred - y) / y_pred.numel()


def adam_step(p, g, m, v, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """One dense torch.optim.Adam update (models.py:68-70 defaults), in place.

    Follows torch's single-tensor path: L2 (coupled) weight decay, bias-corrected,
    eps added OUTSIDE the sqrt:  denom = sqrt(v)/sqrt(1-b2^t) + eps ; p -= lr/(1-b1^t) * m/denom.
    ``step`` is 1-based.
    """
    if weight_decay != 0.0:
        g = g.add(p, alpha=weight_decay)
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1**step
    bc2 = 1 - beta2**step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))
    return p, m, v
