"""CPU tests: oracle/blqr_np.py (block-diagonal lift) against golden vectors from the reference's blqr.py."""
import glob
import os

import numpy as np
import pytest

from oracle import blqr_np as BO
from oracle import lqr_np as O

CASES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "blqr_*.npz")))
rel = lambda a, b: np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_blqr_oracle_matches_reference(path):
    z = np.load(path)
    p = BO.Params(z["x0"], BO.BLQR(**{k: z[k] for k in BO.BLQR._fields}))
    s, gains = BO.direct(p, return_gains=True)
    assert rel(s.X, z["X"]) < 1e-10 and rel(s.U, z["U"]) < 1e-10 and rel(s.Nu, z["Nu"]) < 1e-10
    assert rel(gains.K, z["K"]) < 1e-10 and rel(gains.k, z["k"]) < 1e-10
    g = BO.vjp_exact(p, O.State(z["X"], z["U"], z["Nu"]), O.State(z["w_X"], z["w_U"], z["w_Nu"]))
    assert rel(g.x0, z["g_x0"]) < 1e-8
    for name in BO.BLQR._fields:
        assert rel(getattr(g.blqr, name), z["g_" + name]) < 1e-8, name
