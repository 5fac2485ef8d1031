"""ctypes binding of libmri_b200.so (the C ABI declared in include/mri_b200.h).

There is NO CPU or PyTorch fallback: if the library cannot be loaded, or a tensor is not a CUDA
fp32 tensor, the ops raise.  The library is built in-tree by mri_interpolation_b200/build.py.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import torch

from . import build as _build

MAX_DIM = 4
MAX_LEVELS = 32

ACT_IDENTITY, ACT_SINE, ACT_GELU, ACT_RELU = 0, 1, 2, 3


class MriB200Error(RuntimeError):
    pass


class Level(ctypes.Structure):
    """mirror of mri_level_t"""

    _fields_ = [
        ("resolution", ctypes.c_float * MAX_DIM),
        ("rows", ctypes.c_uint32),
        ("reserved", ctypes.c_uint32),
        ("offset", ctypes.c_uint64),
    ]


_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_I = ctypes.c_int
_F = ctypes.c_float
_D = ctypes.c_double

# name -> argtypes; every function returns int (status) unless listed in _SPECIAL
SIGNATURES = {
    "mri_hashgrid_forward": [_P, _I64, _I, _P, ctypes.POINTER(Level), _I, _I, _P, _P],
    "mri_hashgrid_forward_rows": [_P, _I64, _I, _P, ctypes.POINTER(Level), _I, _I, _P, _P, _P],
    "mri_hashgrid_backward": [_P, _I64, _I, _P, _P, ctypes.POINTER(Level), _I, _I, _P],
    "mri_hashgrid_backward_levels": [_P, _I64, _I, _P, _P, ctypes.POINTER(Level), _I, _I, _I, _I, _P],
    "mri_hashgrid_corners": [_P, _I64, _I, ctypes.POINTER(Level), _I, _P, _P, _P],
    "mri_dense_forward": [_P, _I64, _P, _P, _I64, _I, _I, _I, _F, _P, _P, _P],
    "mri_dense_backward": [_P, _I64, _P, _P, _P, _I64, _I, _I, _I, _F, _P, _P, _P, _P, _P],
    "mri_mse_loss_grad": [_P, _P, _I64, _F, _P, _P, _P],
    "mri_adam_step": [_P, _P, _P, _P, _I64, _I64, _D, _D, _D, _D, _D, _D, _I, _P],
    "mri_adam_step_captured": [_P, _P, _P, _P, _I64, _P, _P, _D, _D, _D, _D, _D, _D, _I, _P],
    "mri_adam_step_sharded": [ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint64, ctypes.c_uint64,
                              _I, _I, _P, _P, _I64, _I64, _I64,
                              _D, _D, _D, _D, _D, _D, _I, _P],
    "mri_adam_step_sharded_sync": [ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint64, ctypes.c_uint64,
                                   ctypes.POINTER(ctypes.c_uint64), _I, _I, _P, _P, _I64, _I64, _I64,
                                   _D, _D, _D, _D, _D, _D, _I, _P],
    "mri_grid_coords": [_P, ctypes.POINTER(ctypes.c_int32), _I, _I64, _I64, _P, _P],
    "mri_gather_voxels": [_P, ctypes.POINTER(ctypes.c_int32), _I, _P, _I64, _P, _P, _P, _P],
    "mri_hashmlp_sweep": [_P, ctypes.POINTER(ctypes.c_int32), _I, _I64, _I64, _P, ctypes.POINTER(Level), _I, _I, _P,
                          ctypes.POINTER(ctypes.c_int32), _I, _I, _I, _P, _P],
    "mri_decoder2_supported": [_I, _I, _I],
    "mri_decoder2_forward": [_P, _I64, _I, _I, _P, _P, _P, _P, _I, _I, _P, _P, _P],
    "mri_decoder2_backward": [_P, _I64, _I, _I, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P],
    "mri_hashdecoder_supported": [_I, _I, _I, _I, _I],
    "mri_hashdecoder_backward": [_P, _I64, _I, _P, _I, _I, _P, _P, _P, _P, _P, _I, _I, _P, ctypes.POINTER(Level), _I, _I, _P, _P, _P,
                                 _P, _P],
    "mri_hashdecoder_forward": [_P, _I64, _I, _P, ctypes.POINTER(Level), _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P],
    "mri_hashmlp_mse_step_supported": [_I, _I, _I, _I, _I],
    "mri_hashmlp_mse_step": [_P, _P, _I64, _I, _P, ctypes.POINTER(Level), _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _F, _P,
                             ctypes.POINTER(Level), _P, _P, _P, _P, _P, _P, _P],
    "mri_siren_tc_supported": [_I, _I],
    "mri_siren_tc_split": [_P, _I64, _P, _P, _P],
    "mri_siren_tc_layer": [_P, _P, _P, _P, _P, _I64, _I, _I, _I, _F, _I, _P, _P, _P, _P, _P, _P],
    "mri_siren_tc_dgrad": [_P, _P, _P, _P, _I64, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    "mri_siren_tc_wgrad": [_P, _P, _P, _P, _I64, _I, _I, _I, _P, _P, _P],
    "mri_siren_tc_mul_split": [_P, _P, _I64, _P, _P, _P],
    "mri_siren_first_forward": [_P, _I64, _P, _P, _I64, _I, _I, _F, _P, _P, _P, _P],
    "mri_siren_first_backward": [_P, _P, _I64, _I64, _I, _I, _P, _P, _P],
    "mri_siren_last_forward": [_P, _P, _P, _P, _I64, _I, _I, _P, _P],
    "mri_siren_last_backward": [_P, _P, _P, _P, _P, _I64, _I, _I, _P, _P, _P, _P, _P, _P],
    "mri_sq_err_sum": [_P, _P, _I64, _P, _P],
    "mri_ssim_sum": [_P, _P, _I, _I, _I64, _I, _D, _P, _P],
    "mri_linear_time_interp": [_P, _I64, _I, _P, _P],
    "mri_probe_red_rate": [_P, _I64, _I, _I, ctypes.POINTER(ctypes.c_int64), _P],
}
_SPECIAL = {
    "mri_version": ([], _I),
    "mri_sm_count": ([], _I),
    "mri_last_error": ([], ctypes.c_char_p),
}

_lib: Optional[ctypes.CDLL] = None
launch_count = 0  # number of C-ABI compute calls issued (bench.py reports kernels launched)


def library_path() -> str:
    return _build.LIB_PATH


def lib() -> ctypes.CDLL:
    """Load (building first if the .so is absent and nvcc is available) and bind the library."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.isfile(path):
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            raise MriB200Error(
                f"libmri_b200.so is missing and could not be built ({e}); run "
                f"`python -m mri_interpolation_b200.build` - there is no CPU fallback") from e
    try:
        handle = ctypes.CDLL(path)
    except OSError as e:
        raise MriB200Error(f"cannot load {path}: {e}; there is no CPU fallback") from e
    for name, argtypes in SIGNATURES.items():
        fn = getattr(handle, name, None)
        if fn is None:
            raise MriB200Error(f"{path} does not export {name}; rebuild with python -m mri_interpolation_b200.build --force")
        fn.argtypes = argtypes
        fn.restype = _I
    for name, (argtypes, restype) in _SPECIAL.items():
        fn = getattr(handle, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = handle
    return handle


def exported_symbols():
    return list(SIGNATURES) + list(_SPECIAL)


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().mri_last_error()
        raise MriB200Error(f"{what} failed (status {status}): {msg.decode() if msg else ''}")


def call(name: str, *args, kernels: int = 1) -> None:
    """Invoke a C-ABI entry point; `kernels` = CUDA kernels that call launches (for bench.py's count)."""
    global launch_count
    launch_count += kernels
    check(getattr(lib(), name)(*args), name)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream() -> int:
    """cudaStream_t of torch's current stream on the current device.  The raw accessor skips the Stream object that
    torch.cuda.current_stream() builds (25 us per call in the Trainer loop's profile, four calls per training step)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise MriB200Error(f"{name}: expected a tensor")
    if not t.is_cuda:
        raise MriB200Error(
            f"{name}: tensor is on {t.device}; the B200 backend runs on CUDA only (no CPU fallback) - move the "
            f"module and its inputs to a CUDA device")
    if t.dtype != torch.float32:
        raise MriB200Error(f"{name}: expected float32, got {t.dtype}")
    return t


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def make_levels(resolutions: Sequence[Sequence[float]], rows: Sequence[int], offsets: Sequence[int]):
    n = len(rows)
    if n > MAX_LEVELS:
        raise MriB200Error(f"at most {MAX_LEVELS} levels are supported, got {n}")
    arr = (Level * n)()
    for i in range(n):
        res = list(resolutions[i])
        if len(res) > MAX_DIM:
            raise MriB200Error(f"hash grid supports at most {MAX_DIM} input dims")
        for d in range(MAX_DIM):
            arr[i].resolution[d] = float(res[d]) if d < len(res) else 0.0
        arr[i].rows = int(rows[i])
        arr[i].reserved = 0
        arr[i].offset = int(offsets[i])
    return arr
