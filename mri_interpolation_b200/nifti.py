"""Minimal NIfTI-1 reader/writer (nibabel is not available in the target image).

Covers what the reference does with nibabel: ``nib.load(path).shape`` (config/base.py:22),
``image.get_fdata(dtype=np.float32)`` (datamodules.py:137: raw * scl_slope + scl_inter, Fortran
order on disk) and ``nib.save(nib.Nifti1Image(im, affine=np.eye(4)), path)`` (launcher.py:189,219).
Single-file ``.nii`` / ``.nii.gz`` only.
"""
from __future__ import annotations

import gzip
import struct
from typing import Tuple

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
           768: np.uint32}
_CODES = {np.dtype(v).name: k for k, v in _DTYPES.items()}


def _open(path: str, mode: str):
    return gzip.open(path, mode) if path.endswith(".gz") else open(path, mode)


class NiftiImage:
    def __init__(self, raw: np.ndarray, slope: float, inter: float, pixdim: Tuple[float, ...], header: bytes = b""):
        self.raw = raw
        self.slope = slope
        self.inter = inter
        self.pixdim = pixdim
        self.header_bytes = header

    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(self.raw.shape)

    def get_fdata(self, dtype=np.float32) -> np.ndarray:
        data = self.raw.astype(dtype)
        # nibabel applies scaling only when the header asks for it (slope 0 or NaN = none)
        if self.slope not in (0.0, 1.0) and not np.isnan(self.slope) or (self.inter != 0.0 and not np.isnan(self.inter)):
            slope = dtype(self.slope if self.slope != 0.0 and not np.isnan(self.slope) else 1.0)
            inter = dtype(0.0 if np.isnan(self.inter) else self.inter)
            data = data * slope + inter
        return data


def load(path: str) -> NiftiImage:
    with _open(path, "rb") as f:
        blob = f.read()
    if len(blob) < 348:
        raise ValueError(f"{path}: too short for a NIfTI-1 header")
    endian = "<"
    if struct.unpack("<i", blob[:4])[0] != 348:
        endian = ">"
        if struct.unpack(">i", blob[:4])[0] != 348:
            raise ValueError(f"{path}: not a NIfTI-1 file (sizeof_hdr != 348)")
    if blob[344:347] not in (b"n+1", b"ni1"):
        raise ValueError(f"{path}: bad NIfTI-1 magic {blob[344:348]!r}")
    dim = struct.unpack(endian + "8h", blob[40:56])
    datatype, bitpix = struct.unpack(endian + "2h", blob[70:74])
    pixdim = struct.unpack(endian + "8f", blob[76:108])
    vox_offset, slope, inter = struct.unpack(endian + "3f", blob[108:120])
    if datatype not in _DTYPES:
        raise ValueError(f"{path}: unsupported NIfTI datatype code {datatype}")
    ndim = dim[0]
    shape = tuple(int(d) for d in dim[1:1 + ndim])
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(endian)
    count = int(np.prod(shape))
    start = int(vox_offset)
    raw = np.frombuffer(blob, dtype=dt, count=count, offset=start).reshape(shape, order="F")
    return NiftiImage(np.ascontiguousarray(raw).astype(dt.newbyteorder("=")), float(slope), float(inter),
                      tuple(pixdim[1:1 + ndim]), blob[:348])


def save(array: np.ndarray, path: str) -> None:
    """Write ``array`` as a single-file NIfTI-1 with an identity affine (launcher.py:189)."""
    arr = np.asarray(array)
    if arr.dtype.name not in _CODES:
        arr = arr.astype(np.float32)
    if arr.ndim > 7:
        raise ValueError("NIfTI-1 holds at most 7 dimensions")
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    dim = [arr.ndim] + list(arr.shape) + [1] * (7 - arr.ndim)
    struct.pack_into("<8h", hdr, 40, *dim)
    struct.pack_into("<2h", hdr, 70, _CODES[arr.dtype.name], arr.dtype.itemsize * 8)
    struct.pack_into("<8f", hdr, 76, 1.0, *([1.0] * 7))
    struct.pack_into("<3f", hdr, 108, 352.0, 1.0, 0.0)  # vox_offset, scl_slope, scl_inter
    struct.pack_into("<2h", hdr, 252, 0, 2)  # qform_code 0, sform_code 2 (aligned)
    struct.pack_into("<4f", hdr, 280, 1.0, 0.0, 0.0, 0.0)  # srow_x
    struct.pack_into("<4f", hdr, 296, 0.0, 1.0, 0.0, 0.0)  # srow_y
    struct.pack_into("<4f", hdr, 312, 0.0, 0.0, 1.0, 0.0)  # srow_z
    hdr[344:348] = b"n+1\x00"
    with _open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(b"\x00" * 4)  # extension flag -> data starts at byte 352
        f.write(np.asfortranarray(arr).tobytes(order="F"))
