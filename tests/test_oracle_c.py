"""CPU test: the C restatement (oracle/lqr_c.c, the bench's CPU baseline) against the NumPy oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import lqr_np as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def clib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    from oracle import lqr_c
    return lqr_c


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-10), (np.float32, 2e-4)])
@pytest.mark.parametrize("shape", [(8, 4, 20, 5), (3, 2, 5, 3), (4, 2, 50, 2)])
def test_c_oracle_matches_numpy_oracle(clib, dtype, tol, shape):
    n, m, T, B = shape
    rng = np.random.default_rng(11)
    p = O.random_params(rng, n, m, T, batch=(B,), dtype=dtype)
    so, go = O.direct(p, return_gains=True)
    s, K, k = clib.fwd(p)
    rel = lambda a, b: np.max(np.abs(a - b)) / np.max(np.abs(b))
    assert rel(s.X, so.X) < tol and rel(s.U, so.U) < tol and rel(s.Nu, so.Nu) < tol and rel(K, go.K) < tol
    w = O.State(*[rng.standard_normal(a.shape).astype(dtype) for a in so])
    g = clib.vjp(p, s, w)
    g_o = O.vjp_exact(p, so, w)
    assert rel(g.x0, g_o.x0) < 10 * tol
    for a, b in zip(g.lqr, g_o.lqr):
        assert rel(a, b) < 10 * tol


def test_c_oracle_golden_shift(clib):
    z = np.load(os.path.join(ROOT, "tests", "golden", "lqr_shift_n4m2T10.npz"))
    p = O.Params(z["x0"][None], O.LQR(**{k: z[k][None] for k in O.LQR._fields}))
    s, K, k = clib.fwd(p)
    rel = lambda a, b: np.max(np.abs(a - b)) / np.max(np.abs(b))
    assert rel(s.X[0], z["X"]) < 1e-10 and rel(s.Nu[0], z["Nu"]) < 1e-10 and rel(K[0], z["K"]) < 1e-10
