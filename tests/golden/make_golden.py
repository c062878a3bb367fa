"""Generate golden vectors by executing the UNMODIFIED reference sources
(/root/reference/idoc/*.py) on the NumPy stand-ins in tests/golden/refshim.

Run in the build container only (it needs /root/reference, which the GPU box lacks):

    python tests/golden/make_golden.py

Outputs: tests/golden/*.npz (small, committed).  For every case the file holds the inputs
(NumPy-seeded, fp64), the reference's `direct` outputs (X, U, Nu), its `backward` gains
(K, k, dC, dc), its `kkt` residual at a perturbed state, and the exact implicit VJP
theta_bar = -(dF/dtheta)^T J^{-T} w  assembled from the reference's own `kkt` function
(F = kkt is affine in s and in theta, so J and dF/dtheta are obtained exactly by evaluating
it on basis vectors) — i.e. what `jaxopt.implicit_diff.custom_root(kkt)` defines, with the
linear solve done densely instead of by CG.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
import idoc  # noqa: E402  (the reference)


# ------------------------------------------------------------------ pytree <-> flat
def flat(tree):
    out = []

    def rec(t):
        if isinstance(t, tuple):
            for c in t:
                rec(c)
        else:
            out.append(np.asarray(t, dtype=np.float64).ravel())

    rec(tree)
    return np.concatenate(out)


def unflat(tree, v):
    pos = [0]

    def rec(t):
        if isinstance(t, tuple):
            cols = [rec(c) for c in t]
            return type(t)(*cols) if hasattr(t, "_fields") else tuple(cols)
        a = np.asarray(t)
        z = v[pos[0]:pos[0] + a.size].reshape(a.shape)
        pos[0] += a.size
        return z

    return rec(tree)


def dense_implicit_vjp(kkt, s, params, w):
    """theta_bar = -(dF/dtheta)^T u,  J^T u = w,  F = kkt (affine in s and in theta)."""
    s0 = flat(s)
    th0 = flat(params)
    F = lambda sv, tv: flat(kkt(unflat(s, sv), unflat(params, tv)))
    base = F(np.zeros_like(s0), th0)
    J = np.stack([F(e, th0) - base for e in np.eye(s0.size)], axis=1)
    Fs = F(s0, th0)
    G = np.stack([F(s0, th0 + e) - Fs for e in np.eye(th0.size)], axis=1)
    u = np.linalg.solve(J.T, flat(w))
    return unflat(params, -G.T @ u), J


def lqr_case(params, T, rng, with_vjp=True):
    solver = idoc.lqr.build(T)
    s = solver.direct(params)
    gains, ec = idoc.lqr.backward(params.lqr, T, return_expected_change=True)
    dc = ec(1.0) - 0.0
    # expected_change(alpha) = alpha^2 dC + alpha dc  ->  recover (dC, dc) from two evaluations
    e1, e2 = ec(1.0), ec(2.0)
    dC = 0.5 * (e2 - 2.0 * e1)
    dc = e1 - dC
    sp = idoc.typs.State(*[z + 0.1 * rng.standard_normal(z.shape) for z in s])
    res = solver.kkt(sp, params)
    res0 = solver.kkt(s, params)
    w = idoc.typs.State(X=rng.standard_normal(s.X.shape), U=rng.standard_normal(s.U.shape),
                        Nu=rng.standard_normal(s.Nu.shape))
    out = {"x0": params.x0}
    for k, v in params.lqr._asdict().items():
        out[k] = v
    out.update(X=s.X, U=s.U, Nu=s.Nu, K=gains.K, k=gains.k, dC=dC, dc=dc,
               sp_X=sp.X, sp_U=sp.U, sp_Nu=sp.Nu, kkt_X=res.X, kkt_U=res.U, kkt_Nu=res.Nu,
               kkt0_max=max(np.abs(res0.X).max(), np.abs(res0.U).max(), np.abs(res0.Nu).max()),
               w_X=w.X, w_U=w.U, w_Nu=w.Nu)
    if not with_vjp:
        # eigen-shift cases: the KKT matrix is singular / ~1e-9-conditioned there, the implicit
        # gradient is ill-posed (jaxopt's CG would return noise); only the forward solve is pinned.
        return out
    g, _ = dense_implicit_vjp(solver.kkt, s, params, w)
    w0 = idoc.typs.State(X=s.X, U=s.U, Nu=np.zeros_like(s.Nu))  # the reference tests' loss: 0.5|X|^2+0.5|U|^2
    g0, _ = dense_implicit_vjp(solver.kkt, s, params, w0)
    out.update(g_x0=g.x0, g0_x0=g0.x0)
    for k, v in g.lqr._asdict().items():
        out["g_" + k] = v
    for k, v in g0.lqr._asdict().items():
        out["g0_" + k] = v
    return out


def ref_test_params(rng, n, m, T):
    """tests/test_lqr.py:9-32 with NumPy draws."""
    A0 = 0.5 * np.linalg.qr(rng.standard_normal((n, n)))[0]
    B0 = rng.standard_normal((n, m))
    st = lambda z: np.stack(T * (z,))
    lqr = idoc.lqr.LQR(Q=st(np.eye(n)), q=0.2 * st(np.ones(n)), Qf=np.eye(n), qf=0.2 * np.ones(n),
                       R=1e-4 * st(np.eye(m)), r=1e-4 * st(np.ones(m)), M=1e-4 * st(np.ones((n, m))),
                       A=st(A0), B=st(B0), d=st(np.ones(n)))
    return idoc.lqr.Params(rng.standard_normal(n), lqr)


def random_tv_params(rng, n, m, T, r_scale=1.0, degenerate_steps=()):
    """SURVEY.md 8(d) synthetic time-varying problem (fresh draw per timestep)."""
    def spd(k, lead):
        W = rng.standard_normal(lead + (k, k))
        return W @ np.swapaxes(W, -1, -2) / k + np.eye(k)
    A = 0.5 * np.linalg.qr(rng.standard_normal((T, n, n)))[0]
    R = r_scale * spd(m, (T,))
    Bm = rng.standard_normal((T, n, m))
    Mm = 0.1 * rng.standard_normal((T, n, m))
    rv = rng.standard_normal((T, m))
    # Eigen-shift branch (lqr.py:86-88): the last control has no effect and (almost) no cost at these
    # steps, so lambda_min(Guu) = R[-1,-1] in {0, -5e-9, +3e-9} < 1e-8 and the shift is taken; the first
    # control is cheap (1e-3) so the unshifted-Guu value update (lqr.py:91) differs visibly from a Schur form.
    for j, t in enumerate(degenerate_steps):
        R[t] = 0.0
        R[t, 0, 0] = 1e-3
        for i in range(1, m - 1):
            R[t, i, i] = 1.0
        R[t, -1, -1] = (0.0, -5e-9, 3e-9)[j % 3]
        Bm[t, :, -1] = 0.0
        Bm[t, :, 0] *= 0.05
        Mm[t] = 0.0  # keep the stage cost [Q M; M^T R] positive semidefinite
        rv[t, -1] = 0.0
    lqr = idoc.lqr.LQR(Q=spd(n, (T,)), q=rng.standard_normal((T, n)), Qf=spd(n, ()), qf=rng.standard_normal(n),
                       M=Mm, R=R, r=rv, A=A, B=Bm, d=rng.standard_normal((T, n)))
    return idoc.lqr.Params(rng.standard_normal(n), lqr)


def tilqr_case(rng, n, m, T):
    A = 0.5 * np.linalg.qr(rng.standard_normal((n, n)))[0]
    Bm = rng.standard_normal((n, m))
    W = rng.standard_normal((n, n))
    lq = idoc.tilqr.TILQR(Q=np.eye(n) + 0.3 * W @ W.T / n, Qf=np.eye(n), R=0.01 * np.eye(m), A=A, B=Bm)
    theta = idoc.tilqr.Params(rng.standard_normal(n), lq)
    solver = idoc.tilqr.build(T)
    s = solver.direct(theta)
    sp = idoc.typs.State(*[z + 0.1 * rng.standard_normal(z.shape) for z in s])
    res = solver.kkt(sp, theta)
    w = idoc.typs.State(X=rng.standard_normal(s.X.shape), U=rng.standard_normal(s.U.shape),
                        Nu=rng.standard_normal(s.Nu.shape))
    g, _ = dense_implicit_vjp(solver.kkt, s, theta, w)
    out = dict(x0=theta.x0, Q=lq.Q, Qf=lq.Qf, R=lq.R, A=lq.A, B=lq.B, X=s.X, U=s.U, Nu=s.Nu,
               sp_X=sp.X, sp_U=sp.U, sp_Nu=sp.Nu, kkt_X=res.X, kkt_U=res.U, kkt_Nu=res.Nu,
               w_X=w.X, w_U=w.U, w_Nu=w.Nu, g_x0=g.x0, g_Q=g.lqr.Q, g_Qf=g.lqr.Qf, g_R=g.lqr.R,
               g_A=g.lqr.A, g_B=g.lqr.B)
    return out


def blqr_case(rng, bs, n, m, T):
    """tests/test_blqr.py:10-38 parameterisation, with per-system variation added."""
    A = 0.5 * np.linalg.qr(rng.standard_normal((T, bs, n, n)))[0]
    Bm = rng.standard_normal((T, bs, n, m))
    lq = idoc.blqr.BLQR(Q=np.tile(np.eye(n), (T, bs, 1, 1)), q=0.2 * np.ones((T, bs, n)),
                        Qf=np.tile(np.eye(n), (bs, 1, 1)), qf=0.2 * np.ones((bs, n)),
                        R=1e-2 * np.tile(np.eye(m), (T, 1, 1)), r=1e-4 * np.ones((T, m)),
                        M=1e-4 * np.ones((T, bs, n, m)), A=A, B=Bm, d=np.ones((T, bs, n)))
    params = idoc.blqr.Params(rng.standard_normal((bs, n)), lq)
    solver = idoc.blqr.build(T)
    s = solver.direct(params)
    gains = idoc.blqr.backward(params.blqr, T)
    res0 = solver.kkt(s, params)
    w = idoc.typs.State(X=rng.standard_normal(s.X.shape), U=rng.standard_normal(s.U.shape),
                        Nu=rng.standard_normal(s.Nu.shape))
    g, _ = dense_implicit_vjp(solver.kkt, s, params, w)
    out = {"x0": params.x0}
    for k, v in lq._asdict().items():
        out[k] = v
    out.update(X=s.X, U=s.U, Nu=s.Nu, K=gains.K, k=gains.k,
               kkt0_max=max(np.abs(res0.X).max(), np.abs(res0.U).max(), np.abs(res0.Nu).max()),
               w_X=w.X, w_U=w.U, w_Nu=w.Nu, g_x0=g.x0)
    for k, v in g.blqr._asdict().items():
        out["g_" + k] = v
    return out


def ilqr_case(rng, n, m, T, maxiter, with_vjp, use_ls=True, a_scale=1.6):
    """The reference's own iLQR test system (tests/test_ilqr.py:14-64), run through the reference's ilqr.build
    (line search = idoc.make_line_search(), thres 1e-8) with refshim's forward-mode duals standing in for jax.grad."""
    import jax
    import jax.numpy as jnp
    from typing import NamedTuple

    class Theta(NamedTuple):
        Q: np.ndarray
        q: np.ndarray
        Qf: np.ndarray
        R: np.ndarray
        r: np.ndarray
        A: np.ndarray
        B: np.ndarray

    phi = lambda x: jax.nn.relu(x)

    def dynamics(_, x, u, th):
        return th.A @ phi(x) + th.B @ u + 0.5

    def cost(_, x, u, th):
        lQ = 0.5 * jnp.dot(jnp.dot(th.Q, x), x)
        lq = jnp.dot(th.q, x)
        lR = 1e-4 * jnp.dot(jnp.dot(th.R, u), u)
        lM = -1e-4 * jnp.dot(jnp.dot(jnp.ones((n, m)), u), x)
        lr = 1e-4 * jnp.dot(th.r, u)
        return lQ + lq + lR + lM + lr

    def costf(xf, th):
        return 0.5 * jnp.dot(jnp.dot(th.Qf, xf), xf)

    cs = idoc.ilqr.Problem(cost=cost, costf=costf, dynamics=dynamics, horizon=T, state_dim=n, control_dim=m)
    W = rng.standard_normal((n, n))
    th = Theta(Q=np.eye(n) + 0.1 * W @ W.T / n, q=0.01 * np.ones(n) + 0.01 * rng.standard_normal(n), Qf=np.eye(n),
               R=np.eye(m) + 0.05 * rng.standard_normal((m, m)), r=np.ones(m),
               A=0.5 * np.linalg.qr(rng.standard_normal((n, n)))[0] * a_scale, B=rng.standard_normal((n, m)))
    params = idoc.ilqr.Params(rng.standard_normal(n), theta=th)
    solver = idoc.ilqr.build(cs, thres=1e-8, maxiter=maxiter,
                             line_search=idoc.make_line_search() if use_ls else None, unroll=True, jit=False)
    U0 = np.zeros((T, m))
    X0, c0 = idoc.ilqr.simulate(cs, U0, params)
    s = solver.direct(idoc.typs.State(X=X0, U=U0, Nu=np.zeros_like(X0)), params)
    res = solver.kkt(s, params)
    out = dict(x0=params.x0, X0=X0, U0=U0, c0=c0, X=s.X, U=s.U, Nu=s.Nu, use_ls=int(use_ls), maxiter=maxiter,
               kkt_max=max(np.abs(res.X).max(), np.abs(res.U).max(), np.abs(res.Nu).max()))
    for k, v in th._asdict().items():
        out["th_" + k] = v
    sp = idoc.typs.State(*[z + 0.05 * rng.standard_normal(z.shape) for z in s])
    rp = solver.kkt(sp, params)
    out.update(sp_X=sp.X, sp_U=sp.U, sp_Nu=sp.Nu, kkt_X=rp.X, kkt_U=rp.U, kkt_Nu=rp.Nu)
    if with_vjp:
        # exact implicit VJP: J = d kkt / d s and G = d kkt / d (x0, theta) by forward-mode through the reference's kkt
        from jax._ad import jacfwd
        s0, p0 = flat(s), flat(params)
        Fs = lambda sv: _dflat(solver.kkt(_dunflat(s, sv), params))
        Fp = lambda pv: _dflat(solver.kkt(s, _dunflat(params, pv)))
        J = np.asarray(jacfwd(Fs)(s0))
        G = np.asarray(jacfwd(Fp)(p0))
        w = idoc.typs.State(X=rng.standard_normal(s.X.shape), U=rng.standard_normal(s.U.shape),
                            Nu=rng.standard_normal(s.Nu.shape))
        g = unflat(params, -G.T @ np.linalg.solve(J.T, flat(w)))
        out.update(w_X=w.X, w_U=w.U, w_Nu=w.Nu, g_x0=g.x0, J_sym=float(np.abs(J - J.T).max()))
        for k, v in g.theta._asdict().items():
            out["g_" + k] = v
    return out


def bilqr_case(rng, bs, n, m, T, maxiter, with_vjp):
    """The reference's bilqr test system (tests/test_bilqr.py:39-83: f = relu(A x) + B u + 0.5, weights 1e-6) through
    the reference's bilqr.build; bs systems share the controls and theta."""
    import jax
    import jax.numpy as jnp
    from typing import NamedTuple
    from idoc import bilqr

    class Theta(NamedTuple):
        Q: np.ndarray
        q: np.ndarray
        Qf: np.ndarray
        R: np.ndarray
        r: np.ndarray
        A: np.ndarray
        B: np.ndarray

    def dynamics(_, x, u, th):
        return jax.nn.relu(th.A @ x) + th.B @ u + 0.5

    def cost(_, x, u, th):
        lQ = 0.5 * jnp.dot(jnp.dot(th.Q, x), x)
        lq = jnp.dot(th.q, x)
        lR = 1e-6 * jnp.dot(jnp.dot(th.R, u), u)
        lM = -1e-6 * jnp.dot(jnp.dot(jnp.ones((n, m)), u), x)
        lr = 1e-6 * jnp.dot(th.r, u)
        return lQ + lq + lR + lr + lM

    def costf(xf, th):
        return 0.5 * jnp.dot(jnp.dot(th.Qf, xf), xf)

    cs = bilqr.Problem(cost=cost, costf=costf, dynamics=dynamics, horizon=T, state_dim=n, control_dim=m)
    W = rng.standard_normal((n, n))
    th = Theta(Q=np.eye(n) + 0.1 * W @ W.T / n, q=0.01 * np.ones(n), Qf=np.eye(n), R=np.eye(m), r=np.ones(m),
               A=0.5 * np.linalg.qr(rng.standard_normal((n, n)))[0], B=rng.standard_normal((n, m)))
    params = bilqr.Params(rng.standard_normal((bs, n)), theta=th)
    solver = bilqr.build(cs, maxiter=maxiter, thres=1e-8, line_search=idoc.make_line_search(), jit=False, unroll=True)
    U0 = np.zeros((T, m))
    X0, c0 = bilqr.simulate(cs, U0, params)
    s = solver.direct(idoc.typs.State(X=X0, U=U0, Nu=np.zeros_like(X0)), params)
    res = solver.kkt(s, params)
    out = dict(x0=params.x0, X0=X0, U0=U0, c0=c0, X=s.X, U=s.U, Nu=s.Nu, maxiter=maxiter,
               kkt_max=max(np.abs(res.X).max(), np.abs(res.U).max(), np.abs(res.Nu).max()))
    for k, v in th._asdict().items():
        out["th_" + k] = v
    sp = idoc.typs.State(*[z + 0.05 * rng.standard_normal(z.shape) for z in s])
    rp = solver.kkt(sp, params)
    out.update(sp_X=sp.X, sp_U=sp.U, sp_Nu=sp.Nu, kkt_X=rp.X, kkt_U=rp.U, kkt_Nu=rp.Nu)
    if with_vjp:
        from jax._ad import jacfwd
        s0, p0 = flat(s), flat(params)
        Fs = lambda sv: _dflat(solver.kkt(_dunflat(s, sv), params))
        Fp = lambda pv: _dflat(solver.kkt(s, _dunflat(params, pv)))
        J = np.asarray(jacfwd(Fs)(s0))
        G = np.asarray(jacfwd(Fp)(p0))
        w = idoc.typs.State(X=rng.standard_normal(s.X.shape), U=rng.standard_normal(s.U.shape),
                            Nu=rng.standard_normal(s.Nu.shape))
        g = unflat(params, -G.T @ np.linalg.solve(J.T, flat(w)))
        out.update(w_X=w.X, w_U=w.U, w_Nu=w.Nu, g_x0=g.x0)
        for k, v in g.theta._asdict().items():
            out["g_" + k] = v
    return out


def _dflat(tree):
    """flatten a pytree whose leaves may be refshim Duals"""
    import jax.numpy as jnp
    leaves = []

    def rec(t):
        if isinstance(t, tuple):
            for c in t:
                rec(c)
        else:
            leaves.append(t.reshape(-1) if hasattr(t, "reshape") else np.asarray(t, dtype=float).reshape(-1))

    rec(tree)
    return jnp.concatenate(leaves)


def _dunflat(tree, v):
    pos = [0]

    def rec(t):
        if isinstance(t, tuple):
            cols = [rec(c) for c in t]
            return type(t)(*cols) if hasattr(t, "_fields") else tuple(cols)
        a = np.asarray(t)
        z = v[pos[0]:pos[0] + a.size].reshape(a.shape)
        pos[0] += a.size
        return z

    return rec(tree)


def main():
    rng = np.random.default_rng(20261019)
    cases = {
        "lqr_ref_n3m2T5": lqr_case(ref_test_params(rng, 3, 2, 5), 5, rng),            # tests/test_lqr.py:38
        "lqr_ref_n4m2T50": lqr_case(ref_test_params(rng, 4, 2, 50), 50, rng),         # BASELINE config 1
        "lqr_tv_n8m4T12": lqr_case(random_tv_params(rng, 8, 4, 12), 12, rng),          # config-2 shape, short T
        "lqr_tv_n4m2T16_harshR": lqr_case(random_tv_params(rng, 4, 2, 16, r_scale=1e-4), 16, rng),
        "lqr_shift_n4m2T10": lqr_case(random_tv_params(rng, 4, 2, 10, degenerate_steps=(2, 5, 8)), 10, rng, with_vjp=False),
        "lqr_tv_n2m1T8": lqr_case(random_tv_params(rng, 2, 1, 8), 8, rng),
        "lqr_tv_n5m3T7": lqr_case(random_tv_params(rng, 5, 3, 7), 7, rng),
        "lqr_shift_n3m3T6": lqr_case(random_tv_params(rng, 3, 3, 6, degenerate_steps=(1, 2, 4)), 6, rng, with_vjp=False),
    }
    for name, c in cases.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **c)
        print(name, "kkt0_max", float(c["kkt0_max"]))
    t = tilqr_case(rng, 3, 2, 40)                                                         # tests/test_tilqr.py:40
    np.savez_compressed(os.path.join(HERE, "tilqr_n3m2T40.npz"), **t)
    t = tilqr_case(rng, 4, 2, 12)
    np.savez_compressed(os.path.join(HERE, "tilqr_n4m2T12.npz"), **t)
    b = blqr_case(rng, 4, 3, 2, 5)                                                        # tests/test_blqr.py:42-44 (bs 30 -> 4)
    np.savez_compressed(os.path.join(HERE, "blqr_b4n3m2T5.npz"), **b)
    print("blqr kkt0_max", float(b["kkt0_max"]))
    b = blqr_case(np.random.default_rng(30), 30, 3, 2, 5)                                  # tests/test_blqr.py:42-44 size
    np.savez_compressed(os.path.join(HERE, "blqr_b30n3m2T5.npz"), **b)
    print("blqr30 kkt0_max", float(b["kkt0_max"]))
    for name, seed, args in (("ilqr_relu_n4m2T8", 101, (4, 2, 8, 30, True)),
                             ("ilqr_relu_n3m1T6_nols", 102, (3, 1, 6, 30, True, False)),
                             ("ilqr_relu_n10m3T30", 6, (10, 3, 30, 30, False, True, 1.0)),       # tests/test_ilqr.py:80, converges
                             ("ilqr_relu_n10m3T30_cycle", 5, (10, 3, 30, 30, False, True, 1.0))):  # reference ends in a 2-cycle
        c = ilqr_case(np.random.default_rng(seed), *args)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **c)
        print(name, "kkt_max", float(c["kkt_max"]), "J_sym", c.get("J_sym"))
    for name, seed, args in (("bilqr_b1n10m10T20", 11, (1, 10, 10, 20, 30, False)),        # tests/test_bilqr.py:26-34
                             ("bilqr_b3n3m2T6", 12, (3, 3, 2, 6, 30, True))):
        c = bilqr_case(np.random.default_rng(seed), *args)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **c)
        print(name, "kkt_max", float(c["kkt_max"]))


if __name__ == "__main__":
    main()
