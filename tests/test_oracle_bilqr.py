"""CPU tests: oracle/bilqr_np.py against golden vectors produced by running the reference's bilqr.py."""
import glob
import os

import numpy as np
import pytest

from oracle import bilqr_np as BO
from oracle import lqr_np as O

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "bilqr_*.npz")))
rel = lambda a, b: np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_bilqr_oracle_matches_reference(path):
    z = np.load(path)
    th = BO.ReluAx.Theta(**{k: z["th_" + k] for k in BO.ReluAx.FIELDS})
    s, it = BO.bilqr(BO.ReluAx, th, z["x0"], z["X0"], z["U0"], maxiter=int(z["maxiter"]), thres=1e-8, line_search=True)
    assert rel(s.X, z["X"]) < 1e-9 and rel(s.U, z["U"]) < 1e-9 and rel(s.Nu, z["Nu"]) < 1e-9
    r = BO.kkt(BO.ReluAx, th, z["x0"], O.State(z["sp_X"], z["sp_U"], z["sp_Nu"]))
    assert rel(r.X, z["kkt_X"]) < 1e-10 and rel(r.U, z["kkt_U"]) < 1e-10 and rel(r.Nu, z["kkt_Nu"]) < 1e-10
    if "g_x0" in z.files:
        gx0, gth = BO.vjp_exact_relu_ax(th, z["x0"], O.State(z["X"], z["U"], z["Nu"]), O.State(z["w_X"], z["w_U"], z["w_Nu"]))
        assert rel(gx0, z["g_x0"]) < 1e-8
        for k in BO.ReluAx.FIELDS:
            assert rel(getattr(gth, k), z["g_" + k]) < 1e-8, k
