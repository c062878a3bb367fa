"""CPU tests: the NumPy oracle (oracle/lqr_np.py) against golden vectors produced by running the
unmodified reference sources (tests/golden/make_golden.py), plus independent checks of the exact VJP
(dense KKT solve, finite differences) — the reference's own test strategy, tests/test_lqr.py:35-90."""
import glob
import os

import numpy as np
import pytest

from oracle import lqr_np as O

TOL = 1e-10  # north_star fp64 tolerance


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def load_case(path):
    z = np.load(path)
    lqr = O.LQR(**{k: z[k] for k in O.LQR._fields})
    return z, O.Params(z["x0"], lqr)


LQR_CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "lqr_*.npz")))


@pytest.mark.parametrize("path", LQR_CASES, ids=[os.path.basename(p)[:-4] for p in LQR_CASES])
def test_forward_matches_reference(path):
    z, p = load_case(path)
    T = p.lqr.A.shape[0]
    s, gains = O.direct(p, T, return_gains=True)
    _, (dC, dc) = O.backward(p.lqr, T, return_expected_change=True)
    assert rel(gains.K, z["K"]) < TOL and rel(gains.k, z["k"]) < TOL
    assert rel(s.X, z["X"]) < TOL and rel(s.U, z["U"]) < TOL and rel(s.Nu, z["Nu"]) < TOL
    assert abs(dC - z["dC"]) <= TOL * abs(z["dC"]) + 1e-300 and abs(dc - z["dc"]) <= TOL * abs(z["dc"])
    res = O.kkt(O.State(z["sp_X"], z["sp_U"], z["sp_Nu"]), p)
    assert rel(res.X, z["kkt_X"]) < TOL and rel(res.U, z["kkt_U"]) < TOL and rel(res.Nu, z["kkt_Nu"]) < TOL


@pytest.mark.parametrize("path", [p for p in LQR_CASES if "shift" not in p],
                         ids=[os.path.basename(p)[:-4] for p in LQR_CASES if "shift" not in p])
def test_vjp_matches_reference_kkt_dense_solve(path):
    z, p = load_case(path)
    s = O.State(z["X"], z["U"], z["Nu"])
    for pre, w in (("g_", O.State(z["w_X"], z["w_U"], z["w_Nu"])),
                   ("g0_", O.State(z["X"], z["U"], np.zeros_like(z["Nu"])))):
        g = O.vjp_exact(p, s, w)
        assert rel(g.x0, z[pre + "x0"]) < 1e-9
        for name in O.LQR._fields:
            want = z[pre + name]
            got = getattr(g.lqr, name)
            scale = max(np.max(np.abs(want)), 1e-300)
            assert np.max(np.abs(got - want)) / scale < 1e-9, (name, np.max(np.abs(got - want)) / scale)


def test_shift_branch_is_exercised():
    z, p = load_case([c for c in LQR_CASES if "shift_n4m2" in c][0])
    _, aux = O.backward(p.lqr, p.lqr.A.shape[0], return_aux=True)
    assert (aux["shift"] > 0).sum() == 3


def test_batched_equals_loop():
    rng = np.random.default_rng(3)
    p = O.random_params(rng, 4, 2, 9, batch=(5,))
    s = O.direct(p)
    w = O.State(*[rng.standard_normal(a.shape) for a in s])
    g = O.vjp_exact(p, s, w)
    for b in range(5):
        pb = O.Params(p.x0[b], O.LQR(*[a[b] for a in p.lqr]))
        sb = O.direct(pb)
        gb = O.vjp_exact(pb, sb, O.State(*[a[b] for a in w]))
        for a, c in zip(s, sb):
            assert rel(a[b], c) < 1e-13
        for a, c in zip(g.lqr, gb.lqr):
            assert rel(a[b], c) < 1e-12
        assert rel(g.x0[b], gb.x0) < 1e-12


def test_kkt_residual_and_reference_tolerances():
    # tests/test_lqr.py:46-49: mean |kkt| < 1e-5 on the reference's own parameterisation
    rng = np.random.default_rng(42)
    for (n, m, T) in ((3, 2, 5), (4, 2, 50)):
        p = O.reference_test_params(rng, n, m, T)
        s = O.direct(p)
        vals = O.check_kkt(O.kkt(s, p))
        assert max(vals) < 1e-12


def test_vjp_finite_differences():
    rng = np.random.default_rng(7)
    n, m, T = 3, 2, 6
    p = O.random_params(rng, n, m, T)
    s = O.direct(p)
    w = O.State(*[rng.standard_normal(a.shape) for a in s])
    g = O.vjp_exact(p, s, w)

    def loss(pp):
        pp = O.Params(pp.x0, pp.lqr.symm())
        ss = O.direct(pp)
        return sum(np.sum(a * b) for a, b in zip(ss, w))

    eps = 1e-6
    # directional derivative along a random direction in every leaf
    dirs = O.Params(rng.standard_normal(p.x0.shape), O.LQR(*[rng.standard_normal(a.shape) for a in p.lqr]))
    plus = O.Params(p.x0 + eps * dirs.x0, O.LQR(*[a + eps * d for a, d in zip(p.lqr, dirs.lqr)]))
    minus = O.Params(p.x0 - eps * dirs.x0, O.LQR(*[a - eps * d for a, d in zip(p.lqr, dirs.lqr)]))
    fd = (loss(plus) - loss(minus)) / (2 * eps)
    an = np.sum(g.x0 * dirs.x0) + sum(np.sum(a * d) for a, d in zip(g.lqr, dirs.lqr))
    assert abs(fd - an) / abs(an) < 1e-6
    # Q[0], q[0] never enter the problem: exactly zero gradient
    assert np.all(g.lqr.Q[0] == 0) and np.all(g.lqr.q[0] == 0)
