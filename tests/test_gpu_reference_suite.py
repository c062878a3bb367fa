"""The reference's own tests, re-run against the CUDA path through the reference API (idoc_b200.lqr / tilqr / blqr `build`):
same parameterisations, sizes, loss and metrics as /root/reference/tests/test_lqr.py, test_tilqr.py, test_blqr.py —
`utils.check_kkt` (mean |KKT| < 1e-5, idoc/utils.py:16-27) for `direct` and `implicit`, and the gradient of the tests' loss
through `implicit` for ALL leaves with `utils.relative_difference` < 1e-3 (idoc/utils.py:30-39).  The reference compares
`jax.grad` through `direct` (unrolled autodiff) with `jax.grad` through `implicit`; there is no unrolled autodiff here, so
the comparison partner is the central finite difference of the same loss through `direct` (what the reference's
`utils.finite_difference_grad` / tests/test_qp.py do), evaluated for every parameter at once as one batched solve."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ib():
    import idoc_b200
    idoc_b200._lib.require_device()
    return idoc_b200


def _init_lqr(ib, rng, n, m, T):  # tests/test_lqr.py:9-23 (NumPy rng instead of jax.random)
    A0, B0 = ib.utils.init_stable(rng, n), rng.standard_normal((n, m))
    rep = lambda z: np.stack(T * (z,))
    return ib.lqr.LQR(Q=rep(np.eye(n)), q=0.2 * rep(np.ones(n)), Qf=np.eye(n), qf=0.2 * np.ones(n), M=1e-4 * rep(np.ones((n, m))),
                      R=1e-4 * rep(np.eye(m)), r=1e-4 * rep(np.ones(m)), A=rep(A0), B=rep(B0), d=rep(np.ones(n)))


def _flatten(params, ib):
    leaves = [params.x0] + [getattr(params.lqr, k) for k in ib.lqr.LQR._fields]
    return np.concatenate([np.ravel(z) for z in leaves]), [z.shape for z in leaves]


def _unflatten(flat, shapes, ib, batch=()):
    out, o = [], 0
    for sh in shapes:
        k = int(np.prod(sh))
        out.append(flat[..., o:o + k].reshape(tuple(batch) + tuple(sh)))
        o += k
    return ib.lqr.Params(out[0], ib.lqr.LQR(*out[1:]))


def test_lqr_like_the_reference(ib):
    n, m, T = 3, 2, 5  # tests/test_lqr.py:38
    rng = np.random.default_rng(42)
    params = ib.lqr.Params(rng.standard_normal(n), _init_lqr(ib, rng, n, m, T))
    solver = ib.lqr.build(T)
    for solve in (solver.direct, solver.implicit):  # tests/test_lqr.py:46-49
        s = solve(params)
        ib.utils.check_kkt(solver.kkt, s, params)

    def loss_batched(p):  # tests/test_lqr.py:52-58 on a batch of parameter sets (symm() applied first, :61)
        p = ib.lqr.Params(p.x0, p.lqr.symm())
        s = solver.direct(p)
        ax = lambda z, k: z.reshape(z.shape[0], -1).sum(axis=1) if k else z
        return (0.5 * (s.X ** 2).reshape(s.X.shape[0], -1).sum(1) + 0.5 * (s.U ** 2).reshape(s.U.shape[0], -1).sum(1)
                + (p.x0 ** 2).sum(1) + (p.lqr.A ** 2).reshape(p.lqr.A.shape[0], -1).sum(1))

    # implicit gradient of the loss: pullback of dL/ds = (X, U, 0) + the explicit terms 2 x0, 2 A; symm() chain: sym(g) for Q, Qf, R
    psym = ib.lqr.Params(params.x0, params.lqr.symm())
    s, pull = solver.vjp(psym)
    g = pull(ib.typs.State(s.X, s.U, None))
    sym = lambda z: 0.5 * (z + np.swapaxes(z, -1, -2))
    g = ib.lqr.Params(g.x0 + 2 * params.x0, g.lqr._replace(A=g.lqr.A + 2 * params.lqr.A, Q=sym(g.lqr.Q), Qf=sym(g.lqr.Qf), R=sym(g.lqr.R)))
    flat, shapes = _flatten(params, ib)
    eps = 1e-6
    P = flat.size
    batch = np.concatenate((flat[None] + eps * np.eye(P), flat[None] - eps * np.eye(P)))
    L = loss_batched(_unflatten(batch, shapes, ib, batch=(2 * P,)))
    fd = _unflatten((L[:P] - L[P:]) / (2 * eps), shapes, ib)
    rd, pc = ib.utils.relative_difference, ib.utils.print_and_check
    pc(rd(fd.x0, g.x0))  # tests/test_lqr.py:79-90: every leaf
    for k in ("A", "B", "d", "M", "Q", "q", "R", "r", "Qf", "qf"):
        pc(rd(getattr(fd.lqr, k), getattr(g.lqr, k)))


def test_tilqr_like_the_reference(ib):
    n, m, T = 3, 2, 40  # tests/test_tilqr.py:40
    rng = np.random.default_rng(42)
    lq = ib.tilqr.TILQR(Q=np.eye(n), Qf=np.eye(n), R=0.01 * np.eye(m), A=ib.utils.init_stable(rng, n), B=rng.standard_normal((n, m)))
    theta = ib.tilqr.Params(rng.standard_normal(n), lq)  # tests/test_tilqr.py:9-34
    solver = ib.tilqr.build(T)
    for solve in (solver.direct, solver.implicit):
        ib.utils.check_kkt(solver.kkt, solve(theta), theta)  # tests/test_tilqr.py:47-50
    # check_grads(implicit_loss, order 1, rev) (tests/test_tilqr.py:70): a random-direction derivative against the VJP
    s, pull = solver.vjp(theta)
    g = pull(ib.typs.State(s.X, s.U, None))
    leaves = ("Q", "Qf", "R", "A", "B")
    d = {k: rng.standard_normal(getattr(lq, k).shape) for k in leaves}
    for k in ("Q", "Qf", "R"):
        d[k] = 0.5 * (d[k] + d[k].T)
    dx0 = rng.standard_normal(n)
    loss = lambda st: 0.5 * np.sum(st.X ** 2) + 0.5 * np.sum(st.U ** 2)
    eps = 1e-6
    shift = lambda sg: ib.tilqr.Params(theta.x0 + sg * eps * dx0, ib.tilqr.TILQR(**{k: getattr(lq, k) + sg * eps * d[k] for k in leaves}))
    fd = (loss(solver.direct(shift(+1))) - loss(solver.direct(shift(-1)))) / (2 * eps)
    an = float(np.sum(g.x0 * dx0) + sum(np.sum(getattr(g.lqr, k) * d[k]) for k in leaves))
    assert abs(fd - an) <= 1e-6 * max(1.0, abs(an)), (fd, an)


def test_blqr_like_the_reference(ib):
    bs, n, m, T = 30, 3, 2, 5  # tests/test_blqr.py:42-44
    rng = np.random.default_rng(42)
    rep = lambda z: np.stack(T * (z,))
    A0 = np.stack([ib.utils.init_stable(rng, n) for _ in range(bs)])
    B0 = rng.standard_normal((bs, n, m))
    l = ib.blqr.BLQR(Q=rep(np.stack(bs * (np.eye(n),))), q=0.2 * rep(np.ones((bs, n))), Qf=np.stack(bs * (np.eye(n),)), qf=0.2 * np.ones((bs, n)),
                     M=1e-4 * rep(np.ones((bs, n, m))), R=1e-4 * rep(np.eye(m)), r=1e-4 * rep(np.ones(m)), A=rep(A0), B=rep(B0),
                     d=rep(np.ones((bs, n))))  # tests/test_blqr.py:11-37
    params = ib.blqr.Params(rng.standard_normal((bs, n)), l)
    solver = ib.blqr.build(T)
    for solve in (solver.direct, solver.implicit):
        ib.utils.check_kkt(solver.kkt, solve(params), params)  # tests/test_blqr.py:52-55
    # gradient of the tests' loss along a random direction of every leaf against central differences
    s, pull = solver.vjp(params)
    g = pull(ib.typs.State(s.X, s.U, None))
    d = {k: rng.standard_normal(np.shape(getattr(l, k))) for k in ib.blqr.BLQR._fields}
    for k in ("Q", "Qf", "R"):
        d[k] = 0.5 * (d[k] + np.swapaxes(d[k], -1, -2))
    dx0 = rng.standard_normal((bs, n))
    loss = lambda st: 0.5 * np.sum(st.X ** 2) + 0.5 * np.sum(st.U ** 2)
    eps = 1e-6
    shift = lambda sg: ib.blqr.Params(params.x0 + sg * eps * dx0, ib.blqr.BLQR(**{k: getattr(l, k) + sg * eps * d[k] for k in d}))
    fd = (loss(solver.direct(shift(+1))) - loss(solver.direct(shift(-1)))) / (2 * eps)
    an = float(np.sum(g.x0 * dx0) + sum(np.sum(getattr(g.blqr, k) * d[k]) for k in d))
    assert abs(fd - an) <= 1e-5 * max(1.0, abs(an)), (fd, an)
