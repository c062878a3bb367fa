"""GPU parity tests for the shared-control batch LQR (idoc_b200.blqr) against golden vectors produced by running
the reference's blqr.py (incl. its own test size batch_size=30, n=3, m=2, T=5: a lifted state of 90)."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "blqr_*.npz")))
rel = lambda a, b: float(np.max(np.abs(np.asarray(a, np.float64) - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_blqr_matches_reference_golden(path):
    import idoc_b200 as ib
    ib._lib.require_device()
    z = np.load(path)
    p = ib.blqr.Params(z["x0"], ib.blqr.BLQR(**{k: z[k] for k in ib.blqr.BLQR._fields}))
    T = z["A"].shape[0]
    solver = ib.blqr.build(T)
    s, pull = solver.vjp(p)
    assert s.X.shape == z["X"].shape and s.U.shape == z["U"].shape
    assert rel(s.X, z["X"]) < 1e-10 and rel(s.U, z["U"]) < 1e-10 and rel(s.Nu, z["Nu"]) < 1e-10
    ib.utils.check_kkt(solver.kkt, s, p, thres=1e-10)                  # tests/test_blqr.py:52-55
    gains = ib.blqr.backward(p.blqr, T)
    assert rel(gains.K, z["K"]) < 1e-10 and rel(gains.k, z["k"]) < 1e-10
    g = pull(ib.typs.State(z["w_X"], z["w_U"], z["w_Nu"]))
    assert rel(g.x0, z["g_x0"]) < 1e-10
    for name in ib.blqr.BLQR._fields:
        assert getattr(g.blqr, name).shape == z["g_" + name].shape
        assert rel(getattr(g.blqr, name), z["g_" + name]) < 1e-10, name
