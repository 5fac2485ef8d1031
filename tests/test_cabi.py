"""CPU: the C-ABI library builds, loads and exports every symbol include/mri_b200.h declares.
No compute call is made here (no GPU); argument validation that happens before any CUDA call is checked."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from mri_interpolation_b200 import _lib, build


def header_functions():
    text = open(os.path.join(ROOT, "include", "mri_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mri_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    path = build.build()
    assert os.path.isfile(path)
    lib = _lib.lib()
    assert lib.mri_version() == 100


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.lib()
    names = header_functions()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/mri_b200.h but not exported"
    assert set(names) == set(_lib.exported_symbols()), "ctypes bindings out of sync with the header"


def test_level_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.Level) == 32
    assert _lib.Level.rows.offset == 16 and _lib.Level.offset.offset == 24


def test_invalid_arguments_are_reported_not_crashed():
    lib = _lib.lib()
    lv = _lib.make_levels([(4.0, 4.0, 4.0)], [64], [0])
    assert lib.mri_hashgrid_forward(None, 10, 3, None, lv, 1, 2, None, None) == -1
    assert b"null" in lib.mri_last_error()
    assert lib.mri_hashgrid_forward(256, 10, 7, 256, lv, 1, 2, 256, None) == -2  # dim 7 unsupported
    assert lib.mri_hashgrid_forward(256, 10, 3, 256, lv, 1, 3, 256, None) == -2  # F=3 unsupported
    assert lib.mri_adam_step(None, None, None, None, 10, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0, 0, None) == -1
    assert lib.mri_adam_step(256, 256, 256, 256, 10, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0, 0, None) == -1  # step < 1
    assert lib.mri_dense_forward(None, 4, None, None, 8, 4, 4, 0, 1.0, None, None, None) == -1
    with pytest.raises(_lib.MriB200Error):
        _lib.check(-1, "probe")


def test_cpu_tensors_fail_loudly():
    import torch
    from mri_interpolation_b200 import encoding, models
    enc = encoding.MultiResHashGrid(3, n_levels=2, log2_hashmap_size=8, base_resolution=4, finest_resolution=8)
    with pytest.raises(_lib.MriB200Error, match="no CPU fallback"):
        enc(torch.rand(5, 3))
    net = models.SirenNet(dim_in=2, dim_hidden=8, n_layers=2)
    with pytest.raises(_lib.MriB200Error, match="no CPU fallback"):
        net(torch.rand(5, 2))
    with pytest.raises(_lib.MriB200Error):
        net.configure_optimizers()
