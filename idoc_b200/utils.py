"""Helpers mirroring idoc/utils.py (NumPy; host side)."""
import numpy as np

from . import typs


def init_stable(rng: np.random.Generator, state_dim: int):
    """0.5 * orthonormal factor of a Gaussian matrix: spectral radius 0.5 (idoc/utils.py:9-13)."""
    Q, _ = np.linalg.qr(rng.standard_normal((state_dim, state_dim)))
    return 0.5 * Q


def check_kkt(kkt, s: typs.State, theta, thres=1e-5):
    """idoc/utils.py:16-27."""
    r = kkt(s, theta)
    vals = [float(np.mean(np.abs(z))) for z in (r.X, r.U, r.Nu)]
    assert all(v < thres for v in vals), vals
    return vals


def relative_difference(z1, z2):
    """idoc/utils.py:30-34."""
    return float(np.mean(np.abs(z1 - z2) / (np.abs(z1) + np.abs(z2) + 1e-9)))


def print_and_check(err, thres=1e-3):
    """idoc/utils.py:37-39."""
    assert err < thres, err
