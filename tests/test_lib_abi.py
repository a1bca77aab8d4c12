"""The C-ABI library loads and exports every symbol include/py4cast_b200.h declares
(no compute calls: there is no GPU in the `not gpu` run)."""
import ctypes
import os
import re

import pytest

from py4cast_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "py4cast_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(p4c_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built():
    assert os.path.isfile(_lib.lib_path()), "run __graft_entry__.build() first"


def test_every_declared_symbol_is_exported_and_bound():
    declared = _declared_symbols()
    assert len(declared) >= 20
    lib = ctypes.CDLL(_lib.lib_path())
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes prototype in _lib.SYMBOLS"
    assert set(_lib.SYMBOLS) == set(declared)


def test_struct_layouts_match_the_header():
    # sizes computed from the C declarations (LP64): see include/py4cast_b200.h
    assert ctypes.sizeof(_lib.Src) == 64
    assert ctypes.sizeof(_lib.Mlp) == 8 * 4 + 3 * 64 + 6 * 8 + 8 + 8 + 8
    assert ctypes.sizeof(_lib.AdamWState) == 24
    assert ctypes.sizeof(_lib.Pack) == 7 * 8 + 6 * 4
    assert ctypes.sizeof(_lib.WGrad) == 7 * 8 + 4 * 4
    assert ctypes.sizeof(_lib.SegPart) == 3 * 8 + 2 * 4 + 8 + 2 * 4 + 8 + 8 + 2 * 4
    # and against the compiled library itself
    lib = _lib.load()
    for which, st in enumerate((_lib.Src, _lib.Mlp, _lib.MlpBwd, _lib.ArX, _lib.AdamWState, _lib.Pack, _lib.WGrad,
                                _lib.SegPart)):
        assert lib.p4c_sizeof(which) == ctypes.sizeof(st), st.__name__
    assert lib.p4c_sizeof(99) == -1


def test_host_only_entry_points():
    lib = _lib.load()
    assert lib.p4c_version() >= 100
    lib.p4c_reset_launch_count()
    assert lib.p4c_launch_count() == 0
    # argument validation happens before any CUDA call
    m = _lib.Mlp()
    rc = lib.p4c_mlp_forward(ctypes.byref(m), None, None, None, None, None, None)
    assert rc == -1 and b"mlp" in lib.p4c_last_error()
    assert lib.p4c_loss_partial_floats(8, 327680, 21, 0) > 0


def test_no_cpu_fallback():
    import torch

    from py4cast_b200 import ops

    with pytest.raises(_lib.P4CError, match="CUDA"):
        ops.csr_build(torch.zeros(4, dtype=torch.int64), 2)
