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


def test_run_time_family_plugin_builds_and_exports_the_ilqr_entry_points():
    """idoc_b200.ilqr.define_family on a machine without a GPU: nvcc cross-compiles the plugin, it loads next to the product
    library and exports every idoc_ilqr_* entry point with exactly one family in its table (no compute calls here)."""
    import ctypes as C
    import idoc_b200 as ib
    src = """
struct AbiProbeF {
  static constexpr int N = 2, M = 1, TD = 2;
  template <class S> __device__ static void dyn(const S* th, const S* x, const S* u, S* f) { f[0] = x[0] + th[0] * x[1]; f[1] = x[1] + th[0] * u[0]; }
  template <class S> __device__ static S cost(const S* th, const S* x, const S* u) { return S(0.5) * (x[0] * x[0] + th[1] * u[0] * u[0]); }
  template <class S> __device__ static S costf(const S* th, const S* x) { return S(0.5) * (x[0] * x[0] + x[1] * x[1]); }
};
IDOC_FAMILY(abi_probe, AbiProbeF)
"""
    so = ib.ilqr.define_family("abi_probe", src)
    h = C.CDLL(so)
    for name in ib._lib._ILQR_PROTOS:
        assert hasattr(h, name), name
    fams = {}
    hp = ib.ilqr.PLUGINS["abi_probe"]
    name, pid, n, m, td = C.c_char_p(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
    assert hp.idoc_ilqr_family_count() == 1
    assert hp.idoc_ilqr_family_info(0, C.byref(name), C.byref(pid), C.byref(n), C.byref(m), C.byref(td)) == 0
    assert (name.value.decode(), pid.value, n.value, m.value, td.value) == ("abi_probe", 16, 2, 1, 2)
    assert hp.idoc_ilqr_theta_dim(16, 2, 1) == 2
