"""CPU test: the C-ABI shared library loads and exports every symbol include/idoc_b200.h declares
(no compute calls without a GPU), and argument validation fails loudly."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "idoc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(idoc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "idoc_b200", "libidoc_b200.so"))
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), s


def test_user_family_registry():
    """csrc/families/*.cuh are compiled in and found by name (no device needed for the lookup)."""
    import idoc_b200 as ib
    fams = ib._lib.user_families()
    assert fams["unicycle"] == (16, 3, 2, 9)
    assert ib.ilqr.FAMILIES["unicycle"] == 16 and ib._lib.lib.idoc_ilqr_theta_dim(16, 3, 2) == 9
    assert ib._lib.lib.idoc_ilqr_family_info(len(fams), None, None, None, None, None) != 0


def test_no_cpu_fallback_and_loud_errors():
    import idoc_b200 as ib
    L = ib._lib
    assert L.lib.idoc_version().startswith(b"idoc_b200")
    assert L.lib.idoc_lqr_supported(L.F64, 8, 4) == 1 and L.lib.idoc_lqr_supported(L.F64, 7, 5) == 2 and L.lib.idoc_lqr_supported(L.F64, 140, 5) == 0
    assert L.lib.idoc_lqr_ws_bytes(L.F32, 65536, 100, 8, 4) == 65536 * 100 * 92 * 4 + 65536 * 4
    if L.device_count() == 0:
        from oracle import lqr_np as O
        p = O.random_params(np.random.default_rng(0), 4, 2, 3)
        with pytest.raises(L.IdocError):
            ib.lqr.build(3).direct(ib.lqr.Params(p.x0, ib.lqr.LQR(*p.lqr)))
    # argument validation happens before any CUDA call
    rc = L.lib.idoc_lqr_fwd(None, 5, 1, 1, 8, 4, *([None] * 19), 0)
    assert rc == -1 and b"dtype" in L.lib.idoc_last_error()
    rc = L.lib.idoc_lqr_fwd(None, L.F32, 1, 1, 8, 4, *([None] * 19), 0)
    assert rc == -1 and b"null" in L.lib.idoc_last_error()
    rc = L.lib.idoc_lqr_fwd(None, L.F32, 1, 1, 199, 9, *([None] * 19), 0)
    assert rc == -2
