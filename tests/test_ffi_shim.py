"""CPU test of the XLA-FFI shim source (ffi/idoc_b200_ffi.cc).  jaxlib's headers are absent from this image, so the shim is
compiled (syntax + types only) against a minimal stand-in of the header surface it uses (ffi/mock/), whose DEFINE macro
static_asserts that every handler's parameter list matches its Ctx / Arg / Ret / Attr binding; and every C-ABI function
the shim forwards to is checked against include/idoc_b200.h by the compiler.  Also: every handler symbol the JAX front end
(idoc_b200/jax_api.py) registers is defined by the shim."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_compiles_against_the_header_surface_and_the_c_abi():
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-Wno-unused-function",
           "-I", os.path.join(ROOT, "ffi", "mock"), "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
           os.path.join(ROOT, "ffi", "idoc_b200_ffi.cc")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


def test_front_end_registers_only_handlers_the_shim_defines():
    cc = open(os.path.join(ROOT, "ffi", "idoc_b200_ffi.cc")).read()
    defined = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((idoc_\w+_ffi)", cc))
    py = open(os.path.join(ROOT, "idoc_b200", "jax_api.py")).read()
    names = re.search(r"for _name in \(([^)]*)\)", py).group(1)
    wanted = {"idoc_%s_ffi" % n for n in re.findall(r'"(\w+)"', names)}
    assert wanted and wanted <= defined, wanted - defined
    called = set(re.findall(r'ffi_call\("(idoc_\w+)"', py))
    assert called <= {w[:-4] for w in wanted}, called
