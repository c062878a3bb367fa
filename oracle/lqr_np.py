"""CPU ORACLE (test infrastructure, not product code) — NumPy fp64 restatement of the
reference's time-varying LQR, `/root/reference/idoc/lqr.py`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import this module.  The product path (`idoc_b200/`) never does.

Pinning: the forward solve (`backward`/`forward`/`adjoint`/`kkt`/`direct`) is checked
against golden vectors produced by executing the UNMODIFIED reference sources
(`/root/reference/idoc/lqr.py`) on a NumPy stand-in for jax (`tests/golden/make_golden.py`;
jax itself is not installable in this image).  The implicit VJP is checked against the
dense solve of the linear system jaxopt's `custom_root` poses, assembled from the
reference's own `kkt` function (same script).  The reference's literal `jax.grad`
output cannot be produced here (no jax): that last link is "parity unpinned".

Every array may carry leading batch axes: a single problem is `Q[T,n,n]`, a batch is
`Q[B,T,n,n]` (what `jax.vmap` of the reference would see).  Field names and order follow
the reference pytrees (lqr.py:16-59, typs.py:6-10).
"""
from typing import NamedTuple
import numpy as np


class State(NamedTuple):  # typs.py:6-10
    X: np.ndarray
    U: np.ndarray
    Nu: np.ndarray


class Gains(NamedTuple):  # lqr.py:16-20
    K: np.ndarray
    k: np.ndarray


class LQR(NamedTuple):  # lqr.py:23-35
    Q: np.ndarray
    q: np.ndarray
    Qf: np.ndarray
    qf: np.ndarray
    M: np.ndarray
    R: np.ndarray
    r: np.ndarray
    A: np.ndarray
    B: np.ndarray
    d: np.ndarray

    def symm(self):  # lqr.py:37-52
        return self._replace(Q=_sym(self.Q), Qf=_sym(self.Qf), R=_sym(self.R))


class Params(NamedTuple):  # lqr.py:55-59
    x0: np.ndarray
    lqr: LQR


EPS = 1e-8  # lqr.py:73


def _sym(Z):
    return 0.5 * (Z + np.swapaxes(Z, -1, -2))


def _mv(Mx, v):
    return np.einsum("...ij,...j->...i", Mx, v)


def _mtv(Mx, v):
    return np.einsum("...ji,...j->...i", Mx, v)


def _T(Z):
    return np.swapaxes(Z, -1, -2)


_COMPUTE = [np.float64]


class compute_dtype:
    """Context manager: run the restatement in another floating-point type (np.float32 gives the rounding-error budget of
    the reference ALGORITHM in single precision, i.e. what the reference itself would lose on the same inputs without
    jax_enable_x64; used only to state the fp32 error budget in the parity tests)."""

    def __init__(self, dt):
        self.dt = dt

    def __enter__(self):
        _COMPUTE.append(self.dt)

    def __exit__(self, *a):
        _COMPUTE.pop()


def _f64(t):
    """cast to the compute type (float64 unless inside `compute_dtype`)"""
    if isinstance(t, tuple):
        return type(t)(*[_f64(c) for c in t])
    return np.asarray(t, dtype=_COMPUTE[-1])


def backward(lqr: LQR, horizon: int, *, return_expected_change=False, return_aux=False):
    """Riccati recursion, lqr.py:62-107.  Step body is lqr.py:77-95 line for line."""
    lqr = _f64(lqr)
    A, B, d = lqr.A, lqr.B, lqr.d
    Q, q, Qf, qf = lqr.Q, lqr.q, lqr.Qf, lqr.qf
    R, r, M = lqr.R, lqr.r, lqr.M
    m = R.shape[-1]
    V, v = Qf, qf
    dC = np.zeros(Qf.shape[:-2], Qf.dtype)
    dc = np.zeros(Qf.shape[:-2], Qf.dtype)
    Ks, ks, shifts, Vs, vs = [], [], [], [], []
    for t in range(horizon - 1, -1, -1):  # scan over flip(arange(T)), lqr.py:97-99
        At, Bt = A[..., t, :, :], B[..., t, :, :]
        ATt, BTt = _T(At), _T(Bt)
        Vs.append(V)
        vs.append(v)
        Gxx = _sym(Q[..., t, :, :] + ATt @ V @ At)  # lqr.py:80
        Guu = _sym(R[..., t, :, :] + BTt @ V @ Bt)  # lqr.py:81
        Gxu = M[..., t, :, :] + ATt @ V @ Bt  # lqr.py:82
        Vd = _mv(V, d[..., t, :])  # lqr.py:83
        gx = q[..., t, :] + _mv(ATt, v) + _mv(ATt, Vd)  # lqr.py:84
        gu = r[..., t, :] + _mv(BTt, v) + _mv(BTt, Vd)  # lqr.py:85
        min_eval = np.linalg.eigh(Guu)[0][..., 0]  # lqr.py:86-87
        shift = np.maximum(0.0, EPS - min_eval).astype(Guu.dtype)
        Gtuu = Guu + shift[..., None, None] * np.eye(m, dtype=Guu.dtype)  # lqr.py:88
        K = -np.linalg.solve(Gtuu, _T(Gxu))  # lqr.py:89
        k = -np.linalg.solve(Gtuu, gu[..., None])[..., 0]  # lqr.py:90
        KT = _T(K)
        V = _sym(Gxx + Gxu @ K + KT @ _T(Gxu) + KT @ Guu @ K)  # lqr.py:91 (unshifted Guu)
        v = gx + _mv(Gxu, k) + _mv(KT, gu) + _mv(KT, _mv(Guu, k))  # lqr.py:92
        dC = dC + 0.5 * np.einsum("...i,...i->...", _mv(Guu, k), k)  # lqr.py:93
        dc = dc + np.einsum("...i,...i->...", gu, k)  # lqr.py:94
        Ks.append(K)
        ks.append(k)
        shifts.append(shift)
    gains = Gains(K=np.stack(Ks[::-1], axis=-3), k=np.stack(ks[::-1], axis=-2))  # lqr.py:100
    out = (gains,)
    if return_expected_change:
        out = out + ((dC, dc),)  # expected_change(alpha) = alpha^2 dC + alpha dc, lqr.py:104-105
    if return_aux:
        out = out + (dict(shift=np.stack(shifts[::-1], axis=-1),
                          Vnext=np.stack(Vs[::-1], axis=-3), vnext=np.stack(vs[::-1], axis=-2)),)
    return out[0] if len(out) == 1 else out


def forward(lqr: LQR, x0, gains: Gains):
    """Closed-loop rollout, lqr.py:153-164.  X[t] = x_{t+1} (x0 excluded)."""
    lqr, x0, gains = _f64(lqr), _f64(x0), _f64(gains)
    T = lqr.A.shape[-3]
    x = x0
    Xs, Us = [], []
    for t in range(T):
        u = _mv(gains.K[..., t, :, :], x) + gains.k[..., t, :]  # lqr.py:159
        x = _mv(lqr.A[..., t, :, :], x) + _mv(lqr.B[..., t, :, :], u) + lqr.d[..., t, :]  # lqr.py:160
        Xs.append(x)
        Us.append(u)
    return np.stack(Xs, axis=-2), np.stack(Us, axis=-2)


def adjoint(X, U, lqr: LQR, horizon: int):
    """Co-states, lqr.py:110-125."""
    lqr, X, U = _f64(lqr), _f64(X), _f64(U)
    nu = _mv(lqr.Qf, X[..., -1, :]) + lqr.qf  # lqr.py:121
    Nus = [nu]
    for t in range(horizon - 1, 0, -1):  # t = horizon - _t, _t = 1..T-1, lqr.py:116-119
        nu = (_mtv(lqr.A[..., t, :, :], nu) + _mv(lqr.Q[..., t, :, :], X[..., t - 1, :])
              + lqr.q[..., t, :] + _mv(lqr.M[..., t, :, :], U[..., t, :]))
        Nus.append(nu)
    return np.stack(Nus[::-1], axis=-2)


def kkt(s: State, params: Params) -> State:
    """KKT residual, lqr.py:128-150 (applies .symm() to the problem, lqr.py:130)."""
    s, params = _f64(s), _f64(params)
    x0, lqr = params.x0, params.lqr.symm()
    X, U, Nu = s.X, s.U, s.Nu
    A, B, d, Q, q, Qf, qf, R, r, M = lqr.A, lqr.B, lqr.d, lqr.Q, lqr.q, lqr.Qf, lqr.qf, lqr.R, lqr.r, lqr.M
    dLdX = np.concatenate(
        (q[..., 1:, :] + _mv(M[..., 1:, :, :], U[..., 1:, :]) + _mv(Q[..., 1:, :, :], X[..., :-1, :])
         + _mtv(A[..., 1:, :, :], Nu[..., 1:, :]) - Nu[..., :-1, :],
         (qf + _mv(Qf, X[..., -1, :]) - Nu[..., -1, :])[..., None, :]), axis=-2)
    sX = np.concatenate((x0[..., None, :], X[..., :-1, :]), axis=-2)
    dLdU = _mtv(B, Nu) + _mv(R, U) + r + _mtv(M, sX)
    dLdNu = d + _mv(A, sX) + _mv(B, U) - X
    return State(X=dLdX, U=dLdU, Nu=dLdNu)


def direct(params: Params, horizon=None, *, return_gains=False):
    """lqr.build(T).direct, lqr.py:170-175: backward -> forward -> adjoint."""
    params = _f64(params)
    T = params.lqr.A.shape[-3] if horizon is None else horizon
    gains = backward(params.lqr, T)
    X, U = forward(params.lqr, params.x0, gains)
    Nu = adjoint(X, U, params.lqr, T)
    s = State(X, U, Nu)
    return (s, gains) if return_gains else s


def vjp_exact(params: Params, s: State, w: State) -> Params:
    """Exact implicit-function VJP of `direct` (what jaxopt `custom_root(kkt)` computes,
    lqr.py:177, with the linear solve done exactly instead of by CG).

    J = d kkt / d s is symmetric (Hessian of the Lagrangian), so  J u = -w  is the KKT system
    of another LQR with the same quadratic blocks and linear terms taken from w
    (SURVEY.md section 3.4):  x0'=0, q'[0]=0, q'[t]=Xb[t-1], qf'=Xb[T-1], r'=Ub, d'=Nub.
    Parameter cotangents are  -(d kkt/d theta)^T u  written out as outer products.
    `params` must be symmetric in Q, R, Qf (kkt symmetrises; direct does not, lqr.py:130,170).
    """
    params, s, w = _f64(params), _f64(s), _f64(w)
    lqr = params.lqr.symm()
    T = lqr.A.shape[-3]
    zeros_q0 = np.zeros_like(w.X[..., :1, :])
    # J u = -w  with the adjoint problem's linear terms = +w gives u = -(solution); we fold the
    # sign by solving with linear terms +w and using u = solution (then theta_bar = +(dF/dtheta)^T-terms
    # below carry the matching sign; verified against the dense solve in tests/test_oracle.py).
    adj = LQR(Q=lqr.Q, q=np.concatenate((zeros_q0, w.X[..., :-1, :]), axis=-2), Qf=lqr.Qf, qf=w.X[..., -1, :],
              M=lqr.M, R=lqr.R, r=w.U, A=lqr.A, B=lqr.B, d=w.Nu)
    u = direct(Params(x0=np.zeros_like(params.x0), lqr=adj), T)
    uX, uU, uNu = u.X, u.U, u.Nu
    X, U, Nu = s.X, s.U, s.Nu
    sX = np.concatenate((params.x0[..., None, :], X[..., :-1, :]), axis=-2)
    suX = np.concatenate((np.zeros_like(params.x0)[..., None, :], uX[..., :-1, :]), axis=-2)
    op = lambda a, b: a[..., :, None] * b[..., None, :]
    Ab = op(Nu, suX) + op(uNu, sX)
    Bb = op(Nu, uU) + op(uNu, U)
    db = uNu
    Qb = _sym(op(suX, sX))
    qb = suX
    Rb = _sym(op(uU, U))
    rb = uU
    Mb = op(suX, U) + op(sX, uU)
    Qfb = _sym(op(uX[..., -1, :], X[..., -1, :]))
    qfb = uX[..., -1, :]
    x0b = _mv(lqr.M[..., 0, :, :], uU[..., 0, :]) + _mtv(lqr.A[..., 0, :, :], uNu[..., 0, :])
    return Params(x0=x0b, lqr=LQR(Q=Qb, q=qb, Qf=Qfb, qf=qfb, M=Mb, R=Rb, r=rb, A=Ab, B=Bb, d=db))


# ---------------------------------------------------------------- helpers restating utils.py
def init_stable(rng, n):
    """utils.py:9-13: 0.5 * Q-factor of a Gaussian matrix (spectral radius 0.5)."""
    Rm = rng.standard_normal((n, n))
    Qm, _ = np.linalg.qr(Rm)
    return 0.5 * Qm


def relative_difference(z1, z2):
    """utils.py:30-34."""
    return np.mean(np.abs(z1 - z2) / (np.abs(z1) + np.abs(z2) + 1e-9))


def check_kkt(res: State, thres=1e-5):
    """utils.py:16-27."""
    vals = [np.mean(np.abs(res.X)), np.mean(np.abs(res.U)), np.mean(np.abs(res.Nu))]
    assert all(v < thres for v in vals), vals
    return vals


def reference_test_params(rng, n, m, T, batch=()):
    """Parameterisation of tests/test_lqr.py:9-32 (NumPy rng instead of jax.random)."""
    def tile(z):
        return np.broadcast_to(z, tuple(batch) + (T,) + z.shape).copy()
    A0 = np.stack([init_stable(rng, n) for _ in range(int(np.prod(batch)) or 1)]).reshape(tuple(batch) + (n, n))
    B0 = rng.standard_normal(tuple(batch) + (n, m))
    A = np.repeat(A0[..., None, :, :], T, axis=-3)
    B = np.repeat(B0[..., None, :, :], T, axis=-3)
    lqr = LQR(Q=tile(np.eye(n)), q=tile(0.2 * np.ones(n)),
              Qf=np.broadcast_to(np.eye(n), tuple(batch) + (n, n)).copy(),
              qf=np.broadcast_to(0.2 * np.ones(n), tuple(batch) + (n,)).copy(),
              M=tile(1e-4 * np.ones((n, m))), R=tile(1e-4 * np.eye(m)), r=tile(1e-4 * np.ones(m)),
              A=A, B=B, d=tile(np.ones(n)))
    return Params(x0=rng.standard_normal(tuple(batch) + (n,)), lqr=lqr)


def random_params(rng, n, m, T, batch=(), dtype=np.float64):
    """SURVEY.md section 8(d) synthetic inputs: fresh draw per timestep (time-varying),
    A_t = init_stable, B,d,q,r,qf,x0 ~ N(0,1), Q = W W^T/n + I, R = W W^T/m + I, M ~ 0.1 N(0,1)."""
    bt = tuple(batch) + (T,)
    def spd(shape, k):
        W = rng.standard_normal(shape + (k, k))
        return W @ _T(W) / k + np.eye(k)
    G = rng.standard_normal(bt + (n, n))
    Qm, _ = np.linalg.qr(G)
    lqr = LQR(Q=spd(bt, n), q=rng.standard_normal(bt + (n,)), Qf=spd(tuple(batch), n),
              qf=rng.standard_normal(tuple(batch) + (n,)), M=0.1 * rng.standard_normal(bt + (n, m)),
              R=spd(bt, m), r=rng.standard_normal(bt + (m,)), A=0.5 * Qm,
              B=rng.standard_normal(bt + (n, m)), d=rng.standard_normal(bt + (n,)))
    p = Params(x0=rng.standard_normal(tuple(batch) + (n,)), lqr=lqr)
    return Params(x0=p.x0.astype(dtype), lqr=LQR(*[z.astype(dtype) for z in p.lqr]))
