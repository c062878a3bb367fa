"""ctypes binding of the C ABI in include/idoc_b200.h (libidoc_b200.so, built in-tree by
`python __graft_entry__.py` / `make -C idoc_b200/csrc`).

There is no CPU fallback: importing this module without the built library, or calling a solver
without a CUDA device, raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libidoc_b200.so")

F32, F64 = 0, 1
_DT = {np.dtype(np.float32): F32, np.dtype(np.float64): F64}

ERRORS = {-1: "invalid argument", -2: "unsupported (n, m, dtype)", -3: "CUDA error", -4: "workspace too small",
          -5: "NCCL error"}


class IdocError(RuntimeError):
    pass


class IdocNumericalWarning(RuntimeWarning):
    """A solve finished with IDOC_STATUS_NONFINITE (outputs are NaN, as the reference would propagate them) or
    IDOC_STATUS_CHOL_CLAMP (a Cholesky pivot had to be clamped: the result is not trustworthy)."""


STATUS_SHIFT, STATUS_NONFINITE, STATUS_CHOL_CLAMP = 1, 2, 4


def report_status(status, what):
    """Warn (IdocNumericalWarning) when any problem of a finished solve carries a failure bit.  `status`: int32 array."""
    import warnings
    st = np.asarray(status)
    bad = st & (STATUS_NONFINITE | STATUS_CHOL_CLAMP)
    if bad.any():
        nf, cl = int(((st & STATUS_NONFINITE) != 0).sum()), int(((st & STATUS_CHOL_CLAMP) != 0).sum())
        warnings.warn(f"{what}: {nf} of {st.size} problem(s) non-finite (NaN/Inf inputs or a failed factorisation; outputs are "
                      f"NaN), {cl} with a clamped Cholesky pivot (indefinite Guu beyond the eigen-shift)", IdocNumericalWarning,
                      stacklevel=3)
    return st


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'"
        " at the repository root). idoc_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_vp, _i, _i64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t


_ILQR_PROTOS = {}  # entry points that a run-time family plugin exports too (load_ilqr_plugin)


def _proto(name, restype, argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = argtypes
    if name.startswith("idoc_ilqr_"):
        _ILQR_PROTOS[name] = (restype, argtypes)
    return f


_proto("idoc_version", C.c_char_p, [])
_proto("idoc_last_error", C.c_char_p, [])
_proto("idoc_lqr_supported", _i, [_i, _i, _i])
_proto("idoc_lqr_ws_bytes", _sz, [_i, _i64, _i, _i, _i])
_proto("idoc_lqr_vjp_ws_bytes", _sz, [_i, _i64, _i, _i, _i])
_proto("idoc_lqr_vjp_reduce_ws_bytes", _sz, [_i, _i64, _i, _i, _i])
_proto("idoc_lqr_fwd", _i, [_vp, _i, _i64, _i, _i, _i] + [_vp] * 11 + [_vp] * 3 + [_vp, _vp, _vp, _vp, _vp, _sz])
_proto("idoc_lqr_backward", _i, [_vp, _i, _i64, _i, _i, _i] + [_vp] * 10 + [_vp, _vp, _vp, _vp, _vp, _sz])
_proto("idoc_lqr_vjp", _i, [_vp, _i, _i64, _i, _i, _i] + [_vp] * 4 + [_vp] * 4 + [_vp] * 3 + [_vp] * 11 + [_vp, _sz, _i])
_proto("idoc_lqr_kkt", _i, [_vp, _i, _i64, _i, _i, _i] + [_vp] * 11 + [_vp] * 3 + [_vp] * 3)
_proto("idoc_tilqr_fwd", _i, [_vp, _i, _i64, _i, _i, _i] + [_vp] * 6 + [_vp] * 3 + [_vp, _vp, _sz])
_proto("idoc_tilqr_vjp", _i, [_vp, _i, _i64, _i, _i, _i] + [_vp] * 3 + [_vp] * 4 + [_vp] * 3 + [_vp] * 6 + [_vp, _sz])
_proto("idoc_blqr_ws_bytes", _sz, [_i, _i64, _i, _i, _i, _i])
_proto("idoc_blqr_vjp_ws_bytes", _sz, [_i, _i64, _i, _i, _i, _i])
_proto("idoc_blqr_fwd", _i, [_vp, _i, _i64, _i, _i, _i, _i] + [_vp] * 11 + [_vp] * 3 + [_vp] * 4 + [_vp, _sz])
_proto("idoc_blqr_vjp", _i, [_vp, _i, _i64, _i, _i, _i, _i] + [_vp] * 7 + [_vp] * 3 + [_vp] * 3 + [_vp] * 11 + [_vp, _sz])
_proto("idoc_lqr_host_last_error", C.c_char_p, [])
_proto("idoc_lqr_host_create", _i, [C.POINTER(_vp), _i, _i64, _i, _i, _i, _i, _i])
_proto("idoc_lqr_host_solve_vjp", _i, [_vp, _i64] + [_vp] * 11 + [_vp] * 3 + [_vp] * 11 + [_i, C.POINTER(C.c_float)])
_proto("idoc_lqr_host_finish", _i, [_vp])
_proto("idoc_lqr_host_elapsed_ms", _i, [_vp, C.POINTER(C.c_float)])
_proto("idoc_lqr_host_streams", _i, [_vp, C.POINTER(_vp), C.POINTER(_vp)])
_proto("idoc_lqr_host_bytes", _i, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)])
_proto("idoc_lqr_host_destroy", _i, [_vp])
_proto("idoc_ilqr_last_error", C.c_char_p, [])
_proto("idoc_ilqr_theta_dim", _i, [_i, _i, _i])
_proto("idoc_ilqr_family_count", _i, [])
_proto("idoc_ilqr_family_info", _i, [_i, C.POINTER(C.c_char_p)] + [C.POINTER(_i)] * 4)
_proto("idoc_ilqr_ws_bytes", _sz, [_i, _i64, _i, _i, _i, _i])
_proto("idoc_ilqr_vjp_ws_bytes", _sz, [_i, _i64, _i, _i, _i, _i])
_proto("idoc_ilqr_solve", _i, [_vp, _i, _i, _i64, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, C.c_double, _i, C.c_double,
                               C.c_double, C.c_double, _i, _vp, _vp, _vp, _sz, _i])
_proto("idoc_ilqr_linearize", _i, [_vp, _i, _i, _i64, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _i] + [_vp] * 10)
_proto("idoc_ilqr_vjp", _i, [_vp, _i, _i, _i64, _i, _i, _i, _i, _vp, _i] + [_vp] * 9 + [_vp, _sz])
_proto("idoc_ilqr_solve_box", _i, [_vp, _i, _i, _i64, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, C.c_double, _i,
                                   C.c_double, C.c_double, C.c_double, _i, _vp, _vp, _vp, _sz])
_proto("idoc_ilqr_vjp_box", _i, [_vp, _i, _i, _i64, _i, _i, _i, _vp, _i] + [_vp] * 11 + [_vp, _sz])
_proto("idoc_bilqr_box_supported", _i, [_i, _i, _i, _i])
_proto("idoc_bilqr_solve_box", _i, [_vp, _i, _i, _i64, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, C.c_double, _i,
                                    C.c_double, C.c_double, C.c_double, _i, _vp, _vp, _vp, _sz])
_proto("idoc_bilqr_vjp_box", _i, [_vp, _i, _i, _i64, _i, _i, _i, _i, _vp, _i] + [_vp] * 11 + [_vp, _sz])
_proto("idoc_lqr_generate", _i, [_vp, _i, _i64, _i, _i, _i, C.c_uint64] + [_vp] * 11)
_proto("idoc_device_count", _i, [])
_proto("idoc_set_device", _i, [_i])
_proto("idoc_malloc", _i, [C.POINTER(_vp), _sz])
_proto("idoc_free", _i, [_vp])
_proto("idoc_host_alloc", _i, [C.POINTER(_vp), _sz])
_proto("idoc_host_free", _i, [_vp])
_proto("idoc_memset", _i, [_vp, _i, _sz, _vp])
_proto("idoc_memcpy_h2d", _i, [_vp, _vp, _sz, _vp])
_proto("idoc_memcpy_d2h", _i, [_vp, _vp, _sz, _vp])
_proto("idoc_memcpy_d2d", _i, [_vp, _vp, _sz, _vp])
_proto("idoc_stream_create", _i, [C.POINTER(_vp)])
_proto("idoc_stream_destroy", _i, [_vp])
_proto("idoc_stream_sync", _i, [_vp])
_proto("idoc_event_create", _i, [C.POINTER(_vp)])
_proto("idoc_event_destroy", _i, [_vp])
_proto("idoc_event_record", _i, [_vp, _vp])
_proto("idoc_stream_wait_event", _i, [_vp, _vp])
_proto("idoc_event_sync", _i, [_vp])
_proto("idoc_event_elapsed_ms", _i, [_vp, _vp, C.POINTER(C.c_float)])
_proto("idoc_mem_info", _i, [C.POINTER(_sz), C.POINTER(_sz)])
_proto("idoc_device_props", _i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.c_char_p, _i])
_proto("idoc_launch_count", _i64, [])
_proto("idoc_fma_peak", _i, [_vp, _i, C.POINTER(C.c_double)])
_proto("idoc_nccl_unique_id", _i, [_vp])
_proto("idoc_nccl_init", _i, [_vp, _i, _i, C.POINTER(_vp)])
_proto("idoc_nccl_allreduce_sum", _i, [_vp, _vp, _i, _vp, _sz])
_proto("idoc_nccl_destroy", _i, [_vp])
_proto("idoc_profile_begin", _i, [])
_proto("idoc_profile_end", _i, [_i, C.POINTER(_i), C.POINTER(C.c_float), C.POINTER(_i)])

KERNEL_NAMES = ["lqr_riccati", "lqr_rollout", "lqr_costate_fixup", "lqr_adj_bwd", "lqr_adj_fwd", "lqr_kkt", "lqr_generate",
                "tilqr_fwd", "tilqr_vjp", "ilqr", "misc", "ilqr_fused", "grad_reduce"]


def profile_begin():
    check(lib.idoc_profile_begin(), "profile_begin")


def profile_end(max_records=65536):
    """-> {kernel name: [ms, ...]} for the launches since profile_begin()."""
    ids = (_i * max_records)()
    ms = (C.c_float * max_records)()
    cnt = _i()
    check(lib.idoc_profile_end(max_records, ids, ms, C.byref(cnt)), "profile_end")
    out = {}
    for j in range(cnt.value):
        out.setdefault(KERNEL_NAMES[ids[j]], []).append(ms[j])
    return out


def check(rc, what=""):
    if rc != 0:
        # two translation units keep their own last-error text: ask the one whose entry point failed first
        first, second = (lib.idoc_ilqr_last_error, lib.idoc_last_error) if "ilqr" in what else (lib.idoc_last_error, lib.idoc_ilqr_last_error)
        if "host" in what:
            first = lib.idoc_lqr_host_last_error
        msg = first().decode() or second().decode()
        raise IdocError(f"{what}: {ERRORS.get(rc, rc)}: {msg}")


def dtype_code(dt):
    try:
        return _DT[np.dtype(dt)]
    except KeyError:
        raise IdocError(f"unsupported dtype {dt}: idoc_b200 computes in float32 or float64") from None


def fma_peak_tflops(dtype):
    """Measured FP32 / FP64 FMA throughput of the current device (TFLOP/s)."""
    v = C.c_double()
    check(lib.idoc_fma_peak(None, dtype_code(dtype), C.byref(v)), "fma_peak")
    return v.value


def mem_info():
    """(free, total) bytes of the current device."""
    f, t = _sz(), _sz()
    check(lib.idoc_mem_info(C.byref(f), C.byref(t)), "mem_info")
    return f.value, t.value


def device_count():
    return lib.idoc_device_count()


def require_device():
    if device_count() < 1:
        raise IdocError("no CUDA device: idoc_b200 has no CPU fallback")


class Stream:
    def __init__(self):
        p = _vp()
        check(lib.idoc_stream_create(C.byref(p)), "stream_create")
        self.ptr = p.value

    def sync(self):
        check(lib.idoc_stream_sync(self.ptr), "stream_sync")

    def wait(self, event):
        check(lib.idoc_stream_wait_event(self.ptr, event.ptr), "stream_wait_event")

    def __del__(self):
        try:
            lib.idoc_stream_destroy(self.ptr)
        except Exception:
            pass


class Event:
    def __init__(self):
        p = _vp()
        check(lib.idoc_event_create(C.byref(p)), "event_create")
        self.ptr = p.value

    def record(self, stream=None):
        check(lib.idoc_event_record(self.ptr, stream.ptr if stream is not None else None), "event_record")

    def sync(self):
        check(lib.idoc_event_sync(self.ptr), "event_sync")

    def elapsed_ms(self, later):
        ms = C.c_float()
        check(lib.idoc_event_elapsed_ms(self.ptr, later.ptr, C.byref(ms)), "event_elapsed")
        return ms.value

    def __del__(self):
        try:
            lib.idoc_event_destroy(self.ptr)
        except Exception:
            pass


def sync():
    check(lib.idoc_stream_sync(None), "device_sync")


class DeviceArray:
    """A dense row-major array in device memory (owned)."""

    def __init__(self, shape, dtype, zero=False, _view_of=None, _offset=0):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        self._base = _view_of  # a view keeps its owner alive and never frees
        if _view_of is not None:
            assert _offset % 16 == 0 and _offset + self.nbytes <= _view_of.nbytes
            self.ptr = _view_of.ptr + _offset
            return
        p = _vp()
        check(lib.idoc_malloc(C.byref(p), max(self.nbytes, 16)), "malloc")
        self.ptr = p.value
        if zero:
            self.zero_()

    def view(self, shape, byte_offset=0):
        """A DeviceArray over a 16-byte-aligned sub-range of this one (same dtype, not owning)."""
        return DeviceArray(shape, self.dtype, _view_of=self, _offset=int(byte_offset))

    def zero_(self, stream=None):
        check(lib.idoc_memset(self.ptr, 0, self.nbytes, stream.ptr if stream is not None else None), "memset")
        return self

    @classmethod
    def from_numpy(cls, a, dtype=None, stream=None):
        a = np.ascontiguousarray(a, dtype=dtype if dtype is not None else a.dtype)
        d = cls(a.shape, a.dtype)
        d.copy_from(a, stream)
        return d

    def copy_from(self, a, stream=None):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        assert a.nbytes == self.nbytes, (a.shape, self.shape)
        check(lib.idoc_memcpy_h2d(self.ptr, a.ctypes.data, self.nbytes, stream.ptr if stream is not None else None), "h2d")

    def numpy(self, out=None, stream=None):
        if out is None:
            out = np.empty(self.shape, dtype=self.dtype)
        check(lib.idoc_memcpy_d2h(out.ctypes.data, self.ptr, self.nbytes, stream.ptr if stream is not None else None), "d2h")
        return out

    def free(self):
        if self.ptr is not None:
            if self._base is None:
                lib.idoc_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray:
    """Page-locked host memory exposed as a NumPy array (`.a`)."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = _vp()
        check(lib.idoc_host_alloc(C.byref(p), max(self.nbytes, 16)), "host_alloc")
        self.ptr = p.value
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.a = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self.ptr is not None:
            self.a = None
            lib.idoc_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def ptr(x):
    """Device pointer of a DeviceArray (or None)."""
    return None if x is None else x.ptr


def user_families():
    """{name: (problem id, n, m, theta_dim)} of the user-defined iLQR families compiled into the library
    (csrc/families/*.cuh)."""
    out = {}
    for i in range(lib.idoc_ilqr_family_count()):
        name, pid, n, m, td = C.c_char_p(), _i(), _i(), _i(), _i()
        check(lib.idoc_ilqr_family_info(i, C.byref(name), C.byref(pid), C.byref(n), C.byref(m), C.byref(td)), "family_info")
        out[name.value.decode()] = (pid.value, n.value, m.value, td.value)
    return out


def load_ilqr_plugin(path):
    """A run-time family plugin (idoc_b200.ilqr.define_family): a shared library exporting the idoc_ilqr_* entry points for its
    own problem family, calling the LQR entry points of the product library.  Loaded RTLD_LOCAL."""
    h = C.CDLL(path)
    for name, (restype, argtypes) in _ILQR_PROTOS.items():
        f = getattr(h, name)
        f.restype = restype
        f.argtypes = argtypes
    return h


def check_with(handle, rc, what=""):
    """check() for an entry point of `handle` (the product library or a family plugin: each has its own last-error text)"""
    if rc != 0:
        msg = handle.idoc_ilqr_last_error().decode() or lib.idoc_last_error().decode()
        raise IdocError(f"{what}: {ERRORS.get(rc, rc)}: {msg}")
