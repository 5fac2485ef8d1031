"""torch.autograd glue over the C ABI (include/mri_b200.h).  CUDA fp32 only, no fallback.

Gradient convention ("direct accumulation"): when a parameter already owns a ``.grad`` buffer
(always the case once the model's parameters live in a FlatArena, see optim.py) the backward
kernels accumulate straight into it - they are ``+=`` kernels (red.global.add) anyway - and the
autograd Function returns ``None`` for that input, so no dense (T_l, F) gradient is ever
materialised, copied or re-added.  Otherwise a zeroed buffer is allocated, filled and returned
like any autograd gradient.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_IDENTITY, ACT_RELU, ACT_SINE, MriB200Error

_ACT_BY_NAME = {"identity": ACT_IDENTITY, "sine": ACT_SINE, "gelu": ACT_GELU, "relu": ACT_RELU}


def activation_code(act) -> int:
    if isinstance(act, int):
        return act
    return _ACT_BY_NAME[act]


_grad_write_epoch = 0  # bumped whenever a backward kernel is handed a .grad buffer to accumulate into


def grad_write_epoch() -> int:
    """Monotonic counter of direct gradient accumulations; optim.FusedAdam compares it with the value it saw when it
    last cleared the gradient arena to know whether the arena is still all zeros."""
    return _grad_write_epoch


def _direct_grad(p: torch.Tensor) -> Optional[torch.Tensor]:
    global _grad_write_epoch
    if not p.is_leaf:  # e.g. a spectral-norm parametrised weight: autograd carries the gradient on to its leaves
        return None
    g = getattr(p, "grad", None)
    if g is not None and g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() and g.shape == p.shape:
        _grad_write_epoch += 1
        return g
    return None


# ----------------------------------------------------------------------------- hash grid
class _TableLayout:
    """Base pointer + per-level float offsets of a list of (rows_l, F) tables living anywhere
    in one device's memory (normally consecutive views of the FlatArena)."""

    __slots__ = ("key", "base", "levels")

    def __init__(self):
        self.key = None
        self.base = 0
        self.levels = None

    def refresh(self, tensors: Sequence[torch.Tensor], resolutions, rows):
        ptrs = tuple(t.data_ptr() for t in tensors)
        if ptrs == self.key:
            return
        base = min(ptrs)
        offs = []
        for p in ptrs:
            if (p - base) % 4:
                raise MriB200Error("hash tables must be 4-byte aligned relative to each other")
            offs.append((p - base) // 4)
        self.levels = _lib.make_levels(resolutions, rows, offs)
        self.base = base
        self.key = ptrs


def _no_coordinate_gradient(ctx, what: str) -> None:
    """The reference feeds coordinates from a DataLoader and never differentiates with respect to them;
    the kernels produce table gradients only.  Asking for d/dx must fail loudly, not return a silent None."""
    if ctx.needs_input_grad[0]:
        raise MriB200Error(f"{what}: the gradient with respect to the input coordinates is not implemented "
                           f"(detach the coordinates; only the hash tables and the decoder receive gradients)")


class HashGridFn(torch.autograd.Function):
    """encoding.py:108-128,190-191 forward; autograd of :127-128 backward (tables only - the
    reference never needs d/dx because coordinates come from a DataLoader)."""

    @staticmethod
    def forward(ctx, x, grid, *tables):
        _no_coordinate_gradient(ctx, "hash-grid encoding")
        n_levels, nf, dim = grid.n_levels, grid.n_features_per_level, grid.dim
        x2 = _lib.require_cuda_f32(x, "hashgrid input").reshape(-1, dim).contiguous()
        for t in tables:
            _lib.require_cuda_f32(t, "hash table")
            if not t.is_contiguous():
                raise MriB200Error("hash tables must be contiguous")
        grid._fwd_layout.refresh(tables, grid._resolutions, grid._rows)
        out = torch.empty((x2.shape[0], n_levels * nf), device=x.device, dtype=torch.float32)
        _lib.call("mri_hashgrid_forward", x2.data_ptr(), x2.shape[0], dim, grid._fwd_layout.base,
                  grid._fwd_layout.levels, n_levels, nf, out.data_ptr(), _lib.stream())
        ctx.grid = grid
        ctx.tables = tables
        ctx.save_for_backward(x2)
        ctx.lead_shape = x.shape[:-1]
        return out.reshape(*x.shape[:-1], n_levels * nf)

    @staticmethod
    def backward(ctx, grad_out):
        grid = ctx.grid
        (x2,) = ctx.saved_tensors
        n_levels, nf, dim = grid.n_levels, grid.n_features_per_level, grid.dim
        go = grad_out.reshape(-1, n_levels * nf).contiguous()
        tables = ctx.tables
        need = [t.requires_grad for t in tables]
        if not any(need):
            return (None, None) + (None,) * len(tables)
        direct = [_direct_grad(t) for t in tables]
        if all(d is not None for d in direct):
            targets, ret = direct, (None,) * len(tables)
        else:
            sizes = [(t.numel() + 3) // 4 * 4 for t in tables]
            flat = torch.zeros(sum(sizes), device=go.device, dtype=torch.float32)
            targets, off = [], 0
            for t, s in zip(tables, sizes):
                targets.append(flat[off:off + t.numel()].view_as(t))
                off += s
            ret = tuple(targets)
        grid._bwd_layout.refresh(targets, grid._resolutions, grid._rows)
        hook = getattr(grid, "_grad_group_hook", None)
        groups = getattr(grid, "_grad_groups", None)
        if hook is not None and groups and all(d is not None for d in direct):
            # level groups: the optimiser all-reduces a finished group while the next group's scatter runs
            hook(-1)
            for gi, (lo, hi) in enumerate(groups):
                _lib.call("mri_hashgrid_backward_levels", x2.data_ptr(), x2.shape[0], dim, go.data_ptr(),
                          grid._bwd_layout.base, grid._bwd_layout.levels, n_levels, nf, lo, hi - lo, _lib.stream())
                hook(gi)
        else:
            _lib.call("mri_hashgrid_backward", x2.data_ptr(), x2.shape[0], dim, go.data_ptr(), grid._bwd_layout.base,
                      grid._bwd_layout.levels, n_levels, nf, _lib.stream())
        return (None, None) + ret


def hashgrid_corners(x: torch.Tensor, grid) -> Tuple[torch.Tensor, torch.Tensor]:
    """Parity probe: (n, L, 2^D) int64 hashes and f32 weights from the CUDA kernel."""
    x2 = _lib.require_cuda_f32(x, "x").reshape(-1, grid.dim).contiguous()
    n, c = x2.shape[0], 1 << grid.dim
    hashes = torch.empty((n, grid.n_levels, c), device=x.device, dtype=torch.int32)
    weights = torch.empty((n, grid.n_levels, c), device=x.device, dtype=torch.float32)
    levels = _lib.make_levels(grid._resolutions, grid._rows, [0] * grid.n_levels)
    _lib.call("mri_hashgrid_corners", x2.data_ptr(), n, grid.dim, levels, grid.n_levels, hashes.data_ptr(),
              weights.data_ptr(), _lib.stream())
    return hashes.to(torch.int64) & 0xFFFFFFFF, weights


def hashgrid_forward_rows(x: torch.Tensor, grid) -> Tuple[torch.Tensor, torch.Tensor]:
    """The production gather kernel (mri_hashgrid_forward's code path) with the table row of every corner recorded:
    returns (encoding (n, L*F), rows (n, L, 2^D) int64 in the reference's corner order)."""
    x2 = _lib.require_cuda_f32(x, "x").reshape(-1, grid.dim).contiguous()
    n, c = x2.shape[0], 1 << grid.dim
    tables = grid.tables()
    grid._fwd_layout.refresh(tables, grid._resolutions, grid._rows)
    out = torch.empty((n, grid.n_levels * grid.n_features_per_level), device=x.device, dtype=torch.float32)
    rows = torch.full((n, grid.n_levels, c), -1, device=x.device, dtype=torch.int32)
    _lib.call("mri_hashgrid_forward_rows", x2.data_ptr(), n, grid.dim, grid._fwd_layout.base, grid._fwd_layout.levels,
              grid.n_levels, grid.n_features_per_level, out.data_ptr(), rows.data_ptr(), _lib.stream())
    return out, rows.to(torch.int64) & 0xFFFFFFFF


# -------------------------------------------------------------------------------- dense
class DenseFn(torch.autograd.Function):
    """y = act(x W^T + b) - SirenLayer.forward (models.py:153-156) / decoder Linear+act."""

    @staticmethod
    def forward(ctx, x, weight, bias, act: int, w0: float):
        _lib.require_cuda_f32(x, "dense input")
        _lib.require_cuda_f32(weight, "dense weight")
        m, k = weight.shape
        x2 = x.reshape(-1, k)
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        if not weight.is_contiguous():
            raise MriB200Error("dense weight must be contiguous")
        n = x2.shape[0]
        y = torch.empty((n, m), device=x.device, dtype=torch.float32)
        needs_grad = x.requires_grad or weight.requires_grad or (bias is not None and bias.requires_grad)
        keep_pre = needs_grad and act != ACT_IDENTITY
        pre = torch.empty_like(y) if keep_pre else None
        _lib.call("mri_dense_forward", x2.data_ptr(), x2.stride(0), weight.data_ptr(), _lib.ptr(bias), n, k, m, act,
                  float(w0), y.data_ptr(), _lib.ptr(pre), _lib.stream())
        ctx.act, ctx.w0 = act, float(w0)
        ctx.weight, ctx.bias = weight, bias
        ctx.save_for_backward(x2, pre)
        ctx.x_needs_grad = x.requires_grad
        ctx.x_shape = x.shape
        return y.reshape(*x.shape[:-1], m)

    @staticmethod
    def backward(ctx, grad_y):
        x2, pre = ctx.saved_tensors
        weight, bias = ctx.weight, ctx.bias
        m, k = weight.shape
        n = x2.shape[0]
        gy = grad_y.reshape(n, m).contiguous()
        dpre = torch.empty_like(gy)
        gx = torch.empty((n, k), device=gy.device, dtype=torch.float32) if ctx.x_needs_grad else None
        gw_direct = _direct_grad(weight)
        gw = gw_direct if gw_direct is not None else torch.zeros_like(weight)
        gb = gb_direct = None
        if bias is not None:
            gb_direct = _direct_grad(bias)
            gb = gb_direct if gb_direct is not None else torch.zeros_like(bias)
        _lib.call("mri_dense_backward", x2.data_ptr(), x2.stride(0), weight.data_ptr(), _lib.ptr(pre), gy.data_ptr(),
                  n, k, m, ctx.act, ctx.w0, dpre.data_ptr(), _lib.ptr(gx), gw.data_ptr(), _lib.ptr(gb), _lib.stream(),
                  kernels=3 if gx is not None else 2)
        return (gx.reshape(ctx.x_shape) if gx is not None else None,
                None if gw_direct is not None else gw,
                None if (bias is None or gb_direct is not None) else gb,
                None, None)


def dense(x, weight, bias=None, act="identity", w0: float = 1.0):
    return DenseFn.apply(x, weight, bias, activation_code(act), float(w0))


class Decoder2Fn(torch.autograd.Function):
    """Fused 2-layer decoder enc -> H -> 1 (csrc/decoder.cu): HashMLP's decoder blocks without BatchNorm."""

    @staticmethod
    def forward(ctx, enc, w1, b1, w2, b2, act1: int, act2: int):
        _lib.require_cuda_f32(enc, "decoder input")
        h, k0 = w1.shape
        e2 = enc.reshape(-1, k0).contiguous()
        n = e2.shape[0]
        y = torch.empty((n, 1), device=enc.device, dtype=torch.float32)
        needs_grad = enc.requires_grad or w1.requires_grad
        pre2 = torch.empty((n,), device=enc.device, dtype=torch.float32) if needs_grad else None
        _lib.call("mri_decoder2_forward", e2.data_ptr(), n, k0, h, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                  b2.data_ptr(), act1, act2, y.data_ptr(), _lib.ptr(pre2), _lib.stream())
        ctx.acts = (act1, act2)
        ctx.params = (w1, b1, w2, b2)
        ctx.save_for_backward(e2, pre2)
        ctx.enc_shape = enc.shape
        return y.reshape(*enc.shape[:-1], 1)

    @staticmethod
    def backward(ctx, grad_y):
        e2, pre2 = ctx.saved_tensors
        w1, b1, w2, b2 = ctx.params
        h, k0 = w1.shape
        n = e2.shape[0]
        gy = grad_y.reshape(n).contiguous()
        genc = torch.empty_like(e2)
        grads, ret = [], []
        for p in (w1, b1, w2, b2):
            d = _direct_grad(p)
            grads.append(d if d is not None else torch.zeros_like(p))
            ret.append(None if d is not None else grads[-1])
        _lib.call("mri_decoder2_backward", e2.data_ptr(), n, k0, h, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                  pre2.data_ptr(), gy.data_ptr(), ctx.acts[0], ctx.acts[1], genc.data_ptr(), grads[0].data_ptr(),
                  grads[1].data_ptr(), grads[2].data_ptr(), grads[3].data_ptr(), _lib.stream())
        return (genc.reshape(ctx.enc_shape), ret[0], ret[1], ret[2], ret[3], None, None)


# MRI_FUSED_FORWARD=0 keeps the two-kernel forward (gather, then decoder) for profiling / A-B comparisons
FUSED_FORWARD = os.environ.get("MRI_FUSED_FORWARD", "1") != "0"


class HashDecoderFn(torch.autograd.Function):
    """Hash-grid encoder + 2-layer decoder as ONE autograd node: forward = ONE kernel (gather feeding the tensor-core
    decoder in registers), backward = ONE kernel (decoder backward with the table scatter fused in: dEnc never touches memory)."""

    @staticmethod
    def forward(ctx, x, grid, w1, b1, w2, b2, act1: int, act2: int, *tables):
        _no_coordinate_gradient(ctx, "fused hash-grid + decoder")
        n_levels, nf, dim = grid.n_levels, grid.n_features_per_level, grid.dim
        x2 = _lib.require_cuda_f32(x, "hashgrid input").reshape(-1, dim).contiguous()
        n = x2.shape[0]
        for t in tables:
            _lib.require_cuda_f32(t, "hash table")
        grid._fwd_layout.refresh(tables, grid._resolutions, grid._rows)
        h, k0 = w1.shape
        y = torch.empty((n, 1), device=x.device, dtype=torch.float32)
        train = any(ctx.needs_input_grad)  # all False under torch.no_grad(): nothing is kept for a backward
        pre2 = torch.empty((n,), device=x.device, dtype=torch.float32) if train else None
        if FUSED_FORWARD:
            # ONE kernel: the gather feeds the decoder's tensor-core fragments in registers; the encoding is only
            # written when the backward will need it
            enc = torch.empty((n, n_levels * nf), device=x.device, dtype=torch.float32) if train else None
            _lib.call("mri_hashdecoder_forward", x2.data_ptr(), n, dim, grid._fwd_layout.base, grid._fwd_layout.levels,
                      n_levels, nf, k0, h, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), act1, act2,
                      _lib.ptr(enc), y.data_ptr(), _lib.ptr(pre2), _lib.stream())
        else:
            enc = torch.empty((n, n_levels * nf), device=x.device, dtype=torch.float32)
            _lib.call("mri_hashgrid_forward", x2.data_ptr(), n, dim, grid._fwd_layout.base, grid._fwd_layout.levels, n_levels,
                      nf, enc.data_ptr(), _lib.stream())
            _lib.call("mri_decoder2_forward", enc.data_ptr(), n, k0, h, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                      b2.data_ptr(), act1, act2, y.data_ptr(), _lib.ptr(pre2), _lib.stream())
        ctx.grid, ctx.acts = grid, (act1, act2)
        ctx.params, ctx.tables = (w1, b1, w2, b2), tables
        ctx.save_for_backward(x2, enc, pre2)
        return y.reshape(*x.shape[:-1], 1)

    @staticmethod
    def backward(ctx, grad_y):
        grid = ctx.grid
        x2, enc, pre2 = ctx.saved_tensors
        w1, b1, w2, b2 = ctx.params
        tables = ctx.tables
        n_levels, nf, dim = grid.n_levels, grid.n_features_per_level, grid.dim
        h, k0 = w1.shape
        n = x2.shape[0]
        gy = grad_y.reshape(n).contiguous()
        pgrads, pret = [], []
        for p in (w1, b1, w2, b2):
            d = _direct_grad(p)
            pgrads.append(d if d is not None else torch.zeros_like(p))
            pret.append(None if d is not None else pgrads[-1])
        direct = [_direct_grad(t) for t in tables]
        if all(d is not None for d in direct):
            targets, tret = direct, (None,) * len(tables)
        else:
            sizes = [(t.numel() + 3) // 4 * 4 for t in tables]
            flat = torch.zeros(sum(sizes), device=gy.device, dtype=torch.float32)
            targets, off = [], 0
            for t, s in zip(tables, sizes):
                targets.append(flat[off:off + t.numel()].view_as(t))
                off += s
            tret = tuple(targets)
        grid._bwd_layout.refresh(targets, grid._resolutions, grid._rows)
        _lib.call("mri_hashdecoder_backward", x2.data_ptr(), n, dim, enc.data_ptr(), k0, h, w1.data_ptr(), b1.data_ptr(),
                  w2.data_ptr(), pre2.data_ptr(), gy.data_ptr(), ctx.acts[0], ctx.acts[1], grid._bwd_layout.base,
                  grid._bwd_layout.levels, n_levels, nf, pgrads[0].data_ptr(), pgrads[1].data_ptr(), pgrads[2].data_ptr(),
                  pgrads[3].data_ptr(), _lib.stream())
        return (None, None, pret[0], pret[1], pret[2], pret[3], None, None) + tret


def hashmlp_mse_step_supported(dim: int, n_levels: int, n_features: int, h: int, act1: int) -> bool:
    return bool(_lib.lib().mri_hashmlp_mse_step_supported(int(dim), int(n_levels), int(n_features), int(h), int(act1)))


def _grad_buffer(p: torch.Tensor) -> torch.Tensor:
    """The buffer a backward kernel accumulates into: the parameter's .grad (normally a view of the optimiser's flat
    gradient arena), created as zeros when the parameter has none yet."""
    g = _direct_grad(p)
    if g is None:
        if not p.is_leaf:
            raise MriB200Error("fused training step: parameters must be leaf tensors")
        p.grad = torch.zeros_like(p)
        g = _direct_grad(p)
    return g


@torch.no_grad()
def hashmlp_mse_direct_step(x: torch.Tensor, target: torch.Tensor, grid, w1, b1, w2, b2, act1: int, act2: int):
    """training_step + loss.backward() of HashMLP under the MSE loss WITHOUT the autograd engine: the three kernels the
    autograd path launches (mri_hashdecoder_forward, mri_mse_loss_grad, mri_hashdecoder_backward) called back to back,
    gradients accumulated into the parameters' .grad buffers.  Same kernels, same arithmetic, results identical up to the
    order in which the atomic reductions land - but ~0.15 ms of Python per step instead of ~0.6 ms (two autograd Functions, the engine's worker-thread hand-off and the
    loss-gradient multiply), which is what makes the launcher's loop GPU-bound at 2^19 coordinates per step."""
    n_levels, nf, dim = grid.n_levels, grid.n_features_per_level, grid.dim
    x2 = _lib.require_cuda_f32(x, "hashgrid input").reshape(-1, dim).contiguous()
    n = x2.shape[0]
    t2 = _lib.require_cuda_f32(target, "target").reshape(-1).contiguous()
    if t2.shape[0] != n:
        raise MriB200Error(f"fused training step: {n} coordinates but {t2.shape[0]} targets (one output per coordinate)")
    tables = grid.tables()
    grid._fwd_layout.refresh(tables, grid._resolutions, grid._rows)
    gtables = [_grad_buffer(t) for t in tables]
    grid._bwd_layout.refresh(gtables, grid._resolutions, grid._rows)
    gw1, gb1, gw2, gb2 = (_grad_buffer(p) for p in (w1, b1, w2, b2))
    h, k0 = w1.shape
    dev = x.device
    loss = torch.zeros((), device=dev, dtype=torch.float32)
    if n == 0:
        return loss
    tail = torch.empty((3, n), device=dev, dtype=torch.float32)  # y | pre2 | dL/dy
    y, pre2, gy = tail[0], tail[1], tail[2]
    enc = torch.empty((n, n_levels * nf), device=dev, dtype=torch.float32)
    s = _lib.stream()
    _lib.call("mri_hashdecoder_forward", x2.data_ptr(), n, dim, grid._fwd_layout.base, grid._fwd_layout.levels, n_levels, nf, k0, h,
              w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), act1, act2, enc.data_ptr(), y.data_ptr(), pre2.data_ptr(), s)
    _lib.call("mri_mse_loss_grad", y.data_ptr(), t2.data_ptr(), n, 1.0 / n, gy.data_ptr(), loss.data_ptr(), s)
    _lib.call("mri_hashdecoder_backward", x2.data_ptr(), n, dim, enc.data_ptr(), k0, h, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
              pre2.data_ptr(), gy.data_ptr(), act1, act2, grid._bwd_layout.base, grid._bwd_layout.levels, n_levels, nf,
              gw1.data_ptr(), gb1.data_ptr(), gw2.data_ptr(), gb2.data_ptr(), s)
    return loss


@torch.no_grad()
def hashmlp_mse_step(x: torch.Tensor, target: torch.Tensor, grid, w1, b1, w2, b2, act1: int, act2: int,
                     want_pred: bool = False):
    """loss = F.mse_loss(target, decoder(encoder(x))) AND its backward in ONE kernel (mri_hashmlp_mse_step): the gradients
    of the tables and the decoder accumulate into the parameters' .grad buffers, nothing else is written.  Returns
    (loss 0-dim tensor, predictions (n, 1) or None).  The mean's gradient 2 (y - t) / n needs no reduction over the batch
    first, so forward, loss and backward of a tile run back to back on one set of registers."""
    n_levels, nf, dim = grid.n_levels, grid.n_features_per_level, grid.dim
    x2 = _lib.require_cuda_f32(x, "hashgrid input").reshape(-1, dim).contiguous()
    t2 = _lib.require_cuda_f32(target, "target").reshape(-1).contiguous()
    n = x2.shape[0]
    if t2.shape[0] != n:
        raise MriB200Error(f"fused training step: {n} coordinates but {t2.shape[0]} targets (one output per coordinate)")
    tables = grid.tables()
    grid._fwd_layout.refresh(tables, grid._resolutions, grid._rows)
    gtables = [_grad_buffer(t) for t in tables]
    grid._bwd_layout.refresh(gtables, grid._resolutions, grid._rows)
    gw1, gb1, gw2, gb2 = (_grad_buffer(p) for p in (w1, b1, w2, b2))
    h, k0 = w1.shape
    loss = torch.zeros((), device=x.device, dtype=torch.float32)
    pred = torch.empty((n, 1), device=x.device, dtype=torch.float32) if want_pred else None
    if n > 0:
        _lib.call("mri_hashmlp_mse_step", x2.data_ptr(), t2.data_ptr(), n, dim, grid._fwd_layout.base, grid._fwd_layout.levels,
                  n_levels, nf, k0, h, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), act1, act2, 1.0 / n,
                  grid._bwd_layout.base, grid._bwd_layout.levels, gw1.data_ptr(), gb1.data_ptr(), gw2.data_ptr(), gb2.data_ptr(),
                  loss.data_ptr(), _lib.ptr(pred), _lib.stream())
    return loss, pred


def hashdecoder_supported(dim: int, n_levels: int, n_features: int, h: int, act1: int) -> bool:
    return bool(_lib.lib().mri_hashdecoder_supported(int(dim), int(n_levels), int(n_features), int(h), int(act1)))


def decoder2_supported(k0: int, h: int, act1: int) -> bool:
    return bool(_lib.lib().mri_decoder2_supported(int(k0), int(h), int(act1)))


# ---------------------------------------------------------------------------------- loss
class MseFn(torch.autograd.Function):
    """F.mse_loss(y, y_pred) (models.py:64) with the gradient produced in the same pass.
    ``inv_count`` defaults to 1/numel; data-parallel training passes 1/(global numel)."""

    @staticmethod
    def forward(ctx, pred, target, inv_count: Optional[float]):
        _lib.require_cuda_f32(pred, "prediction")
        _lib.require_cuda_f32(target, "target")
        if pred.shape != target.shape:
            raise MriB200Error(f"mse_loss: shapes differ {tuple(pred.shape)} vs {tuple(target.shape)}")
        p, t = pred.contiguous(), target.contiguous()
        inv = float(inv_count) if inv_count is not None else 1.0 / max(p.numel(), 1)
        loss = torch.zeros((), device=p.device, dtype=torch.float32)
        grad = torch.empty_like(p) if pred.requires_grad else None
        _lib.call("mri_mse_loss_grad", p.data_ptr(), t.data_ptr(), p.numel(), inv, _lib.ptr(grad), loss.data_ptr(),
                  _lib.stream())
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        if grad is None:
            return None, None, None
        return grad * g, None, None


def mse_loss(a: torch.Tensor, b: torch.Tensor, inv_count: Optional[float] = None) -> torch.Tensor:
    """Drop-in for ``F.mse_loss(y, y_pred)`` as BaseMLP.training_step calls it (target first)."""
    if b.requires_grad or not a.requires_grad:
        return MseFn.apply(b, a, inv_count)
    return MseFn.apply(a, b, inv_count)


# --------------------------------------------------------------------------------- sweep
def grid_coords(axes: Sequence[torch.Tensor], first: int, count: int, device) -> torch.Tensor:
    """Coordinates of voxels [first, first+count) of the C-order grid spanned by ``axes``."""
    shape = [int(a.numel()) for a in axes]
    flat_axes = torch.cat([a.reshape(-1).to(torch.float32) for a in axes]).to(device)
    out = torch.empty((count, len(shape)), device=device, dtype=torch.float32)
    import ctypes
    cshape = (ctypes.c_int32 * len(shape))(*shape)
    _lib.call("mri_grid_coords", flat_axes.data_ptr(), cshape, len(shape), int(first), int(count), out.data_ptr(),
              _lib.stream())
    return out


def locality_key(index: torch.Tensor, shape: Sequence[int], block: int = 4) -> torch.Tensor:
    """Sort key that puts voxels sharing hash-table sectors next to each other in a batch.

    Axis 0 is the only axis whose hash prime is 1 (encoding.py:40), so at every level the rows of cells that are
    neighbours along axis 0 differ in their low bits only - they share 32-byte sectors / 128-byte lines - while a step
    along any other axis lands on an unrelated row.  The key therefore walks axis 0 fastest inside strips of ``block``
    axis-1 voxels (coarse levels also share cells across a few axis-1 neighbours), the remaining axes slowest.  The
    loss is a mean over the batch: the order of a batch changes nothing but the memory access pattern."""
    shape = [int(s) for s in shape]
    rem = index
    v = []
    for s in reversed(shape):
        v.append(rem % s)
        rem = rem // s
    v = v[::-1]  # per-axis voxel indices, axis 0 first
    if len(shape) == 1:
        return v[0]
    key = torch.zeros_like(index)
    for d in range(len(shape) - 1, 1, -1):
        key = key * shape[d] + v[d]
    if block > 1:
        key = ((key * ((shape[1] + block - 1) // block) + v[1] // block) * shape[0] + v[0]) * block + v[1] % block
    else:
        key = (key * shape[1] + v[1]) * shape[0] + v[0]
    return key


def locality_sort(index: torch.Tensor, shape: Sequence[int], block: int = 4) -> torch.Tensor:
    """``index`` (..., n) voxel indices -> same sets, each row ordered by `locality_key`."""
    order = torch.argsort(locality_key(index, shape, block), dim=-1)
    return torch.gather(index, -1, order)


class VoxelSampler:
    """Device-resident volume -> training batches from voxel indices (coordinates synthesised in-kernel)."""

    def __init__(self, pixels: torch.Tensor, shape: Sequence[int], norm_siren: bool = False):
        from .datamodules import mgrid_axes
        import ctypes
        self.pixels = _lib.require_cuda_f32(pixels, "pixels").reshape(-1).contiguous()
        self.shape = [int(s) for s in shape]
        self.axes = torch.cat(mgrid_axes(self.shape, norm_siren)).to(pixels.device)
        self._cshape = (ctypes.c_int32 * len(self.shape))(*self.shape)
        self.total = int(self.pixels.numel())

    def batch(self, index: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if index.dtype != torch.int64 or not index.is_cuda:
            raise MriB200Error("VoxelSampler.batch: index must be a CUDA int64 tensor")
        n = index.numel()
        x = torch.empty((n, len(self.shape)), device=index.device, dtype=torch.float32)
        y = torch.empty((n, 1), device=index.device, dtype=torch.float32)
        _lib.call("mri_gather_voxels", self.axes.data_ptr(), self._cshape, len(self.shape), index.data_ptr(), n,
                  self.pixels.data_ptr(), x.data_ptr(), y.data_ptr(), _lib.stream())
        return x, y
