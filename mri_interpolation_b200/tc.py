"""Thin Python handles over the tcgen05 SIREN-layer entry points (csrc/siren_tc.cu).

Operands live as two bf16 planes (hi, lo) with x = hi + lo; ``passes=3`` multiplies them as
A_hi W_hi + A_lo W_hi + A_hi W_lo in fp32 TMEM accumulators (fp32-parity mode), ``passes=1`` is plain bf16.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ACT_IDENTITY, ACT_SINE, MriB200Error


def supported(k: int, m: int) -> bool:
    return bool(_lib.lib().mri_siren_tc_supported(int(k), int(m)))


def split(x: torch.Tensor, need_lo: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """fp32 tensor -> (hi, lo) bf16 planes of the same shape."""
    _lib.require_cuda_f32(x, "split input")
    x = x.contiguous()
    hi = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    lo = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if need_lo else None
    _lib.call("mri_siren_tc_split", x.data_ptr(), x.numel(), hi.data_ptr(), _lib.ptr(lo), _lib.stream())
    return hi, lo


def layer(a_hi, a_lo, w_hi, w_lo, bias, act: int, w0: float, passes: int = 3, mul=None, want_planes: bool = True,
          want_f32: bool = False, want_aux: bool = False):
    """acc = A W^T (+bias) on the tensor cores with the fused epilogue; returns (out_hi, out_lo, out_f32, aux)."""
    n, k = a_hi.shape
    m = w_hi.shape[0]
    if w_hi.shape[1] != k:
        raise MriB200Error(f"tc.layer: A is (n,{k}) but W is {tuple(w_hi.shape)}")
    for t in (a_hi, a_lo, w_hi, w_lo):
        if t is not None and (t.dtype != torch.bfloat16 or not t.is_cuda or not t.is_contiguous()):
            raise MriB200Error("tc.layer: operands must be contiguous CUDA bf16 planes")
    dev = a_hi.device
    out_hi = torch.empty((n, m), device=dev, dtype=torch.bfloat16) if want_planes else None
    out_lo = torch.empty((n, m), device=dev, dtype=torch.bfloat16) if (want_planes and passes == 3) else None
    out_f32 = torch.empty((n, m), device=dev, dtype=torch.float32) if want_f32 else None
    aux = torch.empty((n, m), device=dev, dtype=torch.float32) if want_aux else None
    _lib.call("mri_siren_tc_layer", a_hi.data_ptr(), _lib.ptr(a_lo), w_hi.data_ptr(), _lib.ptr(w_lo), _lib.ptr(bias),
              n, k, m, int(act), float(w0), int(passes), _lib.ptr(mul), _lib.ptr(out_hi), _lib.ptr(out_lo),
              _lib.ptr(out_f32), _lib.ptr(aux), _lib.stream())
    return out_hi, out_lo, out_f32, aux


def mul_split(a: torch.Tensor, b: Optional[torch.Tensor], need_lo: bool = True):
    """(hi, lo) planes of a * b (b may be None)."""
    _lib.require_cuda_f32(a, "mul_split input")
    a = a.contiguous()
    if b is not None:
        b = _lib.require_cuda_f32(b, "mul_split factor").contiguous()
    hi = torch.empty(a.shape, device=a.device, dtype=torch.bfloat16)
    lo = torch.empty(a.shape, device=a.device, dtype=torch.bfloat16) if need_lo else None
    _lib.call("mri_siren_tc_mul_split", a.data_ptr(), _lib.ptr(b), a.numel(), hi.data_ptr(), _lib.ptr(lo), _lib.stream())
    return hi, lo


def wgrad_supported(k: int, m: int) -> bool:
    return m % 128 == 0 and k % 64 == 0 and k >= 64


def wgrad(g_hi, g_lo, x_hi, x_lo, grad_w: torch.Tensor, grad_b: Optional[torch.Tensor], passes: int = 3) -> None:
    """grad_w (m,k) += G^T X ; grad_b (m) += colsum(G) on the tensor cores (split-K over the batch)."""
    n, m = g_hi.shape
    k = x_hi.shape[1]
    if x_hi.shape[0] != n or tuple(grad_w.shape) != (m, k):
        raise MriB200Error("tc.wgrad: shape mismatch")
    _lib.call("mri_siren_tc_wgrad", g_hi.data_ptr(), _lib.ptr(g_lo), x_hi.data_ptr(), _lib.ptr(x_lo), n, k, m, int(passes),
              grad_w.data_ptr(), _lib.ptr(grad_b), _lib.stream(), kernels=2 if grad_b is not None else 1)


def dgrad(g_hi, g_lo, w_hi, w_lo, passes: int = 3, mul=None, want_planes: bool = True, want_f32: bool = False,
          colsum: Optional[torch.Tensor] = None):
    """dX (n, k) = G (n, m) . W (m, k) [* mul] with W's planes as the forward uses them (no transpose);
    ``colsum`` (k,) is incremented by the column sums of the result (bias gradient of the previous layer)."""
    n, m = g_hi.shape
    if w_hi.shape[0] != m:
        raise MriB200Error(f"tc.dgrad: G is (n,{m}) but W is {tuple(w_hi.shape)}")
    k = w_hi.shape[1]
    dev = g_hi.device
    out_hi = torch.empty((n, k), device=dev, dtype=torch.bfloat16) if want_planes else None
    out_lo = torch.empty((n, k), device=dev, dtype=torch.bfloat16) if (want_planes and passes == 3) else None
    out_f32 = torch.empty((n, k), device=dev, dtype=torch.float32) if want_f32 else None
    _lib.call("mri_siren_tc_dgrad", g_hi.data_ptr(), _lib.ptr(g_lo), w_hi.data_ptr(), _lib.ptr(w_lo), n, k, m, int(passes),
              _lib.ptr(mul), _lib.ptr(out_hi), _lib.ptr(out_lo), _lib.ptr(out_f32), _lib.ptr(colsum), _lib.stream())
    return out_hi, out_lo, out_f32
