"""Whole-network SIREN forward/backward on the tcgen05 tensor cores (K3/K4 for wide layers).

Layer plan for SirenNet(dim_in -> H x n_layers -> dim_out) (models.py:160-233):
  layer 0        dim_in -> H   K = 3/4: CUDA-core fused dense+sine (csrc/dense.cu), fp32
  layers 1..L-1  H -> H        tcgen05 GEMM tiles, operands as bf16 (hi, lo) planes, epilogue = +bias,
                               sin(w0 .) -> next layer's planes, w0 cos(w0 .) kept for the backward pass
  last layer     H -> dim_out  CUDA-core dense (a few columns), fp32
Backward mirrors it: dPre planes flow through dgrad GEMMs (W^T planes, epilogue multiplies by the stored
activation derivative), weight gradients are split-K tcgen05 GEMMs over the batch (MN-major operands).

``precision``: "bf16x3" = split-precision 3-MMA mode (fp32 parity, <= 1e-3), "bf16" = single MMA.

Hidden widths that are not a multiple of the 128-wide tiles (the notebook's 4 x 352 SIREN, nb:837) run on the same
kernels zero-padded to the next multiple of 128: padded units have zero weights and bias, so they output sin(0) = 0,
contribute nothing downstream and receive exactly zero gradient - the result is that of the unpadded network, at
(H_pad / H)^2 of its flops (1.19x for 352 -> 384) instead of the fp32 CUDA-core GEMM's ~15x lower rate.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib, tc
from . import functional as Fn
from ._lib import ACT_IDENTITY, ACT_SINE, MriB200Error

PASSES = {"bf16x3": 3, "bf16": 1}
TILE = 128
MIN_PADDED_WIDTH = 160  # below this the padding overhead (H_pad / H)^2 exceeds 2.5x: the fp32 path keeps narrow nets


def padded_width(h: int) -> int:
    """Width the tensor-core tiles run at for a hidden width h (h itself when it already fits)."""
    if h % TILE == 0:
        return h
    return (h + TILE - 1) // TILE * TILE if h >= MIN_PADDED_WIDTH else 0


def _pad2(t: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    out = torch.zeros((rows, cols), device=t.device, dtype=t.dtype)
    out[: t.shape[0], : t.shape[1]] = t
    return out


def _pad1(t: torch.Tensor, n: int) -> torch.Tensor:
    out = torch.zeros((n,), device=t.device, dtype=t.dtype)
    out[: t.shape[0]] = t
    return out


def eligible(net) -> bool:
    """True when every hidden->hidden layer fits the tensor-core tiles and the layers are plain sine layers."""
    from .models import Sine
    layers = list(net.layers)
    if len(layers) < 2:
        return False
    for i, l in enumerate(layers):
        if not isinstance(l.activation, Sine) or l.bias is None:
            return False
        if i >= 1:
            m, k = l.weight.shape
            if m != k or m != layers[0].weight.shape[0]:
                return False
            hp = padded_width(m)
            if hp == 0 or not (tc.supported(hp, hp) and tc.wgrad_supported(hp, hp)):
                return False
    last = net.last_layer
    if layers[0].weight.shape[1] > 4 or last.weight.shape[0] > 4:
        return False  # first/last-layer stream kernels cover dim_in <= 4, dim_out <= 4
    return isinstance(last.activation, torch.nn.Identity) and last.bias is not None


class SirenTcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w0s, passes, *params):
        n_hidden = len(w0s)
        ws, bs = params[0::2], params[1::2]
        _lib.require_cuda_f32(x, "SIREN input")
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        n = x2.shape[0]
        dev = x.device
        train = any(p.requires_grad for p in params)
        h_dim, d_in = ws[0].shape
        m_out = ws[-1].shape[0]
        if d_in > 4 or m_out > 4:
            raise MriB200Error("tensor-core SIREN path handles dim_in <= 4 and dim_out <= 4; use precision='fp32'")
        three = passes == 3
        h_real = h_dim
        h_dim = padded_width(h_real)
        if h_dim == 0:
            raise MriB200Error(f"hidden width {h_real} does not fit the tensor-core tiles; use precision='fp32'")
        if h_dim != h_real:  # zero-padded copies of this step's parameters (a few hundred KB; the planes are re-split anyway)
            with torch.no_grad():
                ws = [_pad2(ws[0], h_dim, d_in)] + [_pad2(w, h_dim, h_dim) for w in ws[1:-1]] + [_pad2(ws[-1], m_out, h_dim)]
                bs = [_pad1(b, h_dim) for b in bs[:-1]] + [bs[-1]]
        # first layer (K = dim_in): planes written directly, no fp32 activations, no split pass
        a_hi = torch.empty((n, h_dim), device=dev, dtype=torch.bfloat16)
        a_lo = torch.empty_like(a_hi) if three else None
        aux0 = torch.empty((n, h_dim), device=dev, dtype=torch.float32) if train else None
        _lib.call("mri_siren_first_forward", x2.data_ptr(), x2.stride(0), ws[0].data_ptr(), bs[0].data_ptr(), n, d_in, h_dim,
                  float(w0s[0]), a_hi.data_ptr(), _lib.ptr(a_lo), _lib.ptr(aux0), _lib.stream())
        acts = [(a_hi, a_lo)]  # acts[i] = output planes of layer i = input of layer i+1
        auxs: List[Optional[torch.Tensor]] = [aux0]
        wplanes = [None]
        for i in range(1, n_hidden):
            w_hi, w_lo = tc.split(ws[i], need_lo=three)  # weights change every optimiser step: re-split
            wplanes.append((w_hi, w_lo))
            oh, ol, _, aux = tc.layer(acts[-1][0], acts[-1][1], w_hi, w_lo, bs[i], ACT_SINE, w0s[i], passes=passes,
                                      want_planes=True, want_f32=False, want_aux=train)
            acts.append((oh, ol))
            auxs.append(aux)
            if not train:
                acts = acts[-1:]
        y = torch.empty((n, m_out), device=dev, dtype=torch.float32)
        _lib.call("mri_siren_last_forward", acts[-1][0].data_ptr(), _lib.ptr(acts[-1][1]), ws[-1].data_ptr(), bs[-1].data_ptr(), n,
                  h_dim, m_out, y.data_ptr(), _lib.stream())
        if train:
            ctx.saved = (x2, acts, auxs, wplanes)
            ctx.params = params
            ctx.padded = (ws, bs) if h_dim != h_real else None
            ctx.w0s, ctx.passes = w0s, passes
        ctx.train = train
        ctx.n_params = len(params)
        return y.reshape(*lead, m_out)

    @staticmethod
    def backward(ctx, grad_y):
        if not ctx.train:
            return (None,) * (3 + ctx.n_params)
        x2, acts, auxs, wplanes = ctx.saved
        params, w0s, passes = ctx.params, ctx.w0s, ctx.passes
        ws, bs = params[0::2], params[1::2]
        n_hidden = len(w0s)
        n = x2.shape[0]
        dev = x2.device
        padded = ctx.padded
        if padded is not None:  # the kernels see the zero-padded network; gradients are cut back to the real shapes below
            ws, bs = padded
        h_dim, d_in = ws[0].shape
        m_out = ws[-1].shape[0]
        three = passes == 3
        grads, direct = [], []
        for p in params:
            d = Fn._direct_grad(p)
            direct.append(d is not None)
            grads.append(d if d is not None else torch.zeros_like(p))
        real_grads = grads
        if padded is not None:
            grads = [torch.zeros_like(t) for pair in zip(ws, bs) for t in pair]
        gw, gb = grads[0::2], grads[1::2]
        gy = grad_y.reshape(n, m_out).contiguous()
        last = n_hidden - 1
        # output layer + head of the hidden backward in one pass over (n, H):
        #   dPre_last = (gy W_out) * aux_last -> planes; db_last_hidden; dW_out = gy^T h; db_out
        g_hi = torch.empty((n, h_dim), device=dev, dtype=torch.bfloat16)
        g_lo = torch.empty_like(g_hi) if three else None
        _lib.call("mri_siren_last_backward", gy.data_ptr(), ws[-1].data_ptr(), auxs[last].data_ptr(), acts[last][0].data_ptr(),
                  _lib.ptr(acts[last][1]), n, h_dim, m_out, g_hi.data_ptr(), _lib.ptr(g_lo), gb[last].data_ptr(),
                  gw[-1].data_ptr(), gb[-1].data_ptr(), _lib.stream(), kernels=3)
        dpre0 = None
        for i in range(last, 0, -1):
            x_hi, x_lo = acts[i - 1]
            tc.wgrad(g_hi, g_lo, x_hi, x_lo, gw[i], None, passes=passes)  # bias gradients come from the epilogues
            w_hi, w_lo = wplanes[i]  # the forward's planes, read as an MN-major operand: no transpose
            if i > 1:
                g_hi, g_lo, _ = tc.dgrad(g_hi, g_lo, w_hi, w_lo, passes=passes, mul=auxs[i - 1], want_planes=True,
                                         colsum=gb[i - 1])
            else:
                _, _, dpre0 = tc.dgrad(g_hi, g_lo, w_hi, w_lo, passes=passes, mul=auxs[0], want_planes=False, want_f32=True)
        if n_hidden == 1:
            raise MriB200Error("tensor-core SIREN path needs at least two sine layers")
        _lib.call("mri_siren_first_backward", dpre0.data_ptr(), x2.data_ptr(), x2.stride(0), n, d_in, h_dim, gw[0].data_ptr(),
                  gb[0].data_ptr(), _lib.stream())  # dW0 and db0 in one pass over dPre0
        if padded is not None:
            for real, pad in zip(real_grads, grads):  # padded rows / columns hold exact zeros
                real.add_(pad[: real.shape[0], : real.shape[1]] if real.dim() == 2 else pad[: real.shape[0]])
            grads = real_grads
        out = [None if d else g for d, g in zip(direct, grads)]
        return (None, None, None) + tuple(out)


def forward(net, x: torch.Tensor, precision: str) -> torch.Tensor:
    if precision not in PASSES:
        raise MriB200Error(f"unknown tensor-core precision {precision!r} (use 'bf16x3' or 'bf16')")
    params = []
    w0s = []
    for l in net.layers:
        params += [l.weight, l.bias]
        w0s.append(float(l.activation.w0))
    params += [net.last_layer.weight, net.last_layer.bias]
    return SirenTcFn.apply(x, tuple(w0s), PASSES[precision], *params)
