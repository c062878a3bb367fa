"""CPU tests: oracle/ilqr_np.py against golden vectors produced by running the reference's ilqr.py /
line_search.py (tests/golden/ilqr_*.npz)."""
import glob
import os

import numpy as np
import pytest

from oracle import ilqr_np as IO
from oracle import lqr_np as O

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ilqr_*.npz")))
rel = lambda a, b: np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_ilqr_oracle_matches_reference(path):
    z = np.load(path)
    th = IO.ReluRnn.Theta(**{k: z["th_" + k] for k in IO.ReluRnn.FIELDS})
    x0 = z["x0"]
    X0, c0 = IO.simulate(IO.ReluRnn, th, x0, z["U0"])
    assert rel(X0, z["X0"]) < 1e-13 and abs(c0 - z["c0"]) < 1e-12 * abs(z["c0"])
    s, it = IO.ilqr(IO.ReluRnn, th, x0, z["X0"], z["U0"], maxiter=int(z["maxiter"]), thres=1e-8, line_search=bool(z["use_ls"]))
    tol = 1e-9
    assert rel(s.X, z["X"]) < tol and rel(s.U, z["U"]) < tol and rel(s.Nu, z["Nu"]) < tol
    r = IO.kkt(IO.ReluRnn, th, x0, O.State(z["sp_X"], z["sp_U"], z["sp_Nu"]))
    assert rel(r.X, z["kkt_X"]) < 1e-10 and rel(r.U, z["kkt_U"]) < 1e-10 and rel(r.Nu, z["kkt_Nu"]) < 1e-10
    if "g_x0" in z.files:
        gx0, gth = IO.vjp_exact_relu(th, x0, O.State(z["X"], z["U"], z["Nu"]), O.State(z["w_X"], z["w_U"], z["w_Nu"]))
        assert rel(gx0, z["g_x0"]) < 1e-8
        for k in IO.ReluRnn.FIELDS:
            assert rel(getattr(gth, k), z["g_" + k]) < 1e-8, k
