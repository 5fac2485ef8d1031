"""Whole-network SIREN forward/backward on the tcgen05 tensor cores (K3/K4 for wide layers).

Layer plan for SirenNet(dim_in -> H x n_layers -> dim_out) (models.py:160-233):
  layer 0        dim_in -> H   K = 3/4: CUDA-core fused dense+sine (csrc/dense.cu), fp32
  layers 1..L-1  H -> H        tcgen05 GEMM tiles, operands as bf16 (hi, lo) planes, epilogue = +bias,
                               sin(w0 .) -> next layer's planes, w0 cos(w0 .) kept for the backward pass
  last layer     H -> dim_out  CUDA-core dense (a few columns), fp32
Backward mirrors it: dPre planes flow through dgrad GEMMs (W^T planes, epilogue multiplies by the stored
activation derivative), weight gradients are split-K tcgen05 GEMMs over the batch (MN-major operands).

``precision``: "bf16x3" = split-precision 3-MMA mode (fp32 parity, <= 1e-3), "bf16" = single MMA.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib, tc
from . import functional as Fn
from ._lib import ACT_IDENTITY, ACT_SINE, MriB200Error

PASSES = {"bf16x3": 3, "bf16": 1}


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
            if not (tc.supported(k, m) and tc.supported(m, k) and tc.wgrad_supported(k, m)):
                return False
    last = net.last_layer
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
        h_dim = ws[0].shape[0]
        # layer 0 on the CUDA cores
        h0 = torch.empty((n, h_dim), device=dev, dtype=torch.float32)
        pre0 = torch.empty_like(h0) if train else None
        _lib.call("mri_dense_forward", x2.data_ptr(), x2.stride(0), ws[0].data_ptr(), bs[0].data_ptr(), n, x2.shape[1],
                  h_dim, ACT_SINE, float(w0s[0]), h0.data_ptr(), _lib.ptr(pre0), _lib.stream())
        a_hi, a_lo = tc.split(h0, need_lo=(passes == 3))
        del h0
        acts = [(a_hi, a_lo)]  # acts[i] = input planes of layer i+1
        auxs: List[Optional[torch.Tensor]] = [None]
        wplanes = [None]
        last_f32 = None
        for i in range(1, n_hidden):
            w_hi, w_lo = tc.split(ws[i], need_lo=(passes == 3))  # weights change every optimiser step: re-split
            wplanes.append((w_hi, w_lo))
            is_last_hidden = i == n_hidden - 1
            oh, ol, of, aux = tc.layer(acts[-1][0], acts[-1][1], w_hi, w_lo, bs[i], ACT_SINE, w0s[i], passes=passes,
                                       want_planes=(not is_last_hidden) or train, want_f32=is_last_hidden, want_aux=train)
            acts.append((oh, ol))
            auxs.append(aux)
            if is_last_hidden:
                last_f32 = of
            if not train:
                acts = acts[-1:]
        m_out = ws[-1].shape[0]
        y = torch.empty((n, m_out), device=dev, dtype=torch.float32)
        _lib.call("mri_dense_forward", last_f32.data_ptr(), last_f32.stride(0), ws[-1].data_ptr(), bs[-1].data_ptr(), n,
                  h_dim, m_out, ACT_IDENTITY, 1.0, y.data_ptr(), None, _lib.stream())
        if train:
            ctx.saved = (x2, pre0, acts, auxs, last_f32, wplanes)
            ctx.params = params
            ctx.w0s, ctx.passes = w0s, passes
        ctx.train = train
        return y.reshape(*lead, m_out)

    @staticmethod
    def backward(ctx, grad_y):
        if not ctx.train:
            return (None,) * (3 + len(ctx.params))
        x2, pre0, acts, auxs, last_f32, wplanes = ctx.saved
        params, w0s, passes = ctx.params, ctx.w0s, ctx.passes
        ws, bs = params[0::2], params[1::2]
        n_hidden = len(w0s)
        n = x2.shape[0]
        dev = x2.device
        h_dim = ws[0].shape[0]
        m_out = ws[-1].shape[0]
        grads = []
        direct = []
        for p in params:
            d = Fn._direct_grad(p)
            direct.append(d is not None)
            grads.append(d if d is not None else torch.zeros_like(p))
        gw, gb = grads[0::2], grads[1::2]
        gy = grad_y.reshape(n, m_out).contiguous()
        # last layer (identity): dH = gy W_last ; dW_last += gy^T h ; db_last += colsum(gy)
        dpre_scratch = torch.empty_like(gy)
        dh = torch.empty((n, h_dim), device=dev, dtype=torch.float32)
        _lib.call("mri_dense_backward", last_f32.data_ptr(), last_f32.stride(0), ws[-1].data_ptr(), None, gy.data_ptr(), n,
                  h_dim, m_out, ACT_IDENTITY, 1.0, dpre_scratch.data_ptr(), dh.data_ptr(), gw[-1].data_ptr(),
                  gb[-1].data_ptr(), _lib.stream(), kernels=3)
        g_hi, g_lo = tc.mul_split(dh, auxs[n_hidden - 1], need_lo=(passes == 3))
        del dh
        dh0 = None
        for i in range(n_hidden - 1, 0, -1):
            x_hi, x_lo = acts[i - 1]
            tc.wgrad(g_hi, g_lo, x_hi, x_lo, gw[i], gb[i], passes=passes)
            w_hi, w_lo = wplanes[i]  # the forward's planes, read as an MN-major operand: no transpose
            if i > 1:
                g_hi, g_lo, _ = tc.dgrad(g_hi, g_lo, w_hi, w_lo, passes=passes, mul=auxs[i - 1], want_planes=True)
            else:
                _, _, dh0 = tc.dgrad(g_hi, g_lo, w_hi, w_lo, passes=passes, want_planes=False, want_f32=True)
        # layer 0 (CUDA cores): dpre0 = dh0 * w0 cos(w0 pre0); dW0, db0
        scratch = torch.empty_like(dh0)
        _lib.call("mri_dense_backward", x2.data_ptr(), x2.stride(0), ws[0].data_ptr(), pre0.data_ptr(), dh0.data_ptr(), n,
                  x2.shape[1], h_dim, ACT_SINE, float(w0s[0]), scratch.data_ptr(), None, gw[0].data_ptr(), gb[0].data_ptr(),
                  _lib.stream(), kernels=2)
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
