"""GPU parity tests for the shared-control batch iLQR (idoc_b200.bilqr) against golden vectors produced by running
the reference's bilqr.py (its own test size batch_size=1, n=m=10, T=20, and a 3-system case with the exact VJP)."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "bilqr_*.npz")))
rel = lambda a, b: float(np.max(np.abs(np.asarray(a, np.float64) - b)) / max(np.max(np.abs(b)), 1e-300))
FIELDS = ("Q", "q", "Qf", "R", "r", "A", "B")


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_bilqr_matches_reference_golden(path):
    import idoc_b200 as ib
    ib._lib.require_device()
    z = np.load(path)
    T, bs, n = z["X"].shape
    m = z["U"].shape[1]
    cs = ib.bilqr.Problem(family="relu_ax", horizon=T, state_dim=n, control_dim=m)
    solver = ib.bilqr.build(cs, maxiter=int(z["maxiter"]), thres=1e-8, line_search=ib.make_line_search())
    params = ib.bilqr.Params(z["x0"], tuple(z["th_" + k] for k in FIELDS))
    s, pull = solver.vjp(ib.typs.State(z["X0"], z["U0"], np.zeros_like(z["X0"])), params)
    assert s.X.shape == z["X"].shape
    assert rel(s.X, z["X"]) < 1e-10 and rel(s.U, z["U"]) < 1e-10 and rel(s.Nu, z["Nu"]) < 1e-10
    ib.utils.check_kkt(solver.kkt, s, params, thres=1e-9)                     # tests/test_bilqr.py:121
    r = solver.kkt(ib.typs.State(z["sp_X"], z["sp_U"], z["sp_Nu"]), params)
    assert rel(r.X, z["kkt_X"]) < 1e-10 and rel(r.U, z["kkt_U"]) < 1e-10 and rel(r.Nu, z["kkt_Nu"]) < 1e-10
    if "g_x0" in z.files:
        g = pull(ib.typs.State(z["w_X"], z["w_U"], z["w_Nu"]))
        assert rel(g.x0, z["g_x0"]) < 1e-10
        for k, got in zip(FIELDS, g.theta):
            assert rel(got, z["g_" + k]) < 1e-10, k
