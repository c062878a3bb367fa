"""CPU tests: oracle/tilqr_np.py against golden vectors from the reference's tilqr.py (tests/golden)."""
import glob
import os

import numpy as np
import pytest

from oracle import tilqr_np as TO

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "tilqr_*.npz")))
rel = lambda a, b: np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_tilqr_oracle_matches_reference(path):
    z = np.load(path)
    T = z["X"].shape[0]
    th = TO.Params(z["x0"], TO.TILQR(Q=z["Q"], Qf=z["Qf"], R=z["R"], A=z["A"], B=z["B"]))
    s = TO.direct(th, T)
    assert rel(s.X, z["X"]) < 1e-10 and rel(s.U, z["U"]) < 1e-10 and rel(s.Nu, z["Nu"]) < 1e-10
    r = TO.kkt(TO.State(z["sp_X"], z["sp_U"], z["sp_Nu"]), th)
    assert rel(r.X, z["kkt_X"]) < 1e-10 and rel(r.U, z["kkt_U"]) < 1e-10 and rel(r.Nu, z["kkt_Nu"]) < 1e-10
    g = TO.vjp_exact(th, TO.State(z["X"], z["U"], z["Nu"]), TO.State(z["w_X"], z["w_U"], z["w_Nu"]), T)
    assert rel(g.x0, z["g_x0"]) < 1e-8
    for name in TO.TILQR._fields:
        assert rel(getattr(g.lqr, name), z["g_" + name]) < 1e-8, name
