#!/usr/bin/env python
"""Small end-to-end exercise of every kernel family (for compute-sanitizer runs on the GPU box)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import idoc_b200 as ib  # noqa: E402
from oracle import lqr_np as O  # noqa: E402

rng = np.random.default_rng(0)
for dtype in (np.float64, np.float32):
    for (n, m, T, B) in ((8, 4, 6, 5), (4, 2, 5, 3), (3, 2, 4, 2), (7, 5, 4, 2)):
        p = O.random_params(rng, n, m, T, batch=(B,), dtype=dtype)
        s, pull = ib.lqr.build(T).vjp(ib.lqr.Params(p.x0, ib.lqr.LQR(*p.lqr)))
        g = pull(ib.typs.State(s.X, s.U, s.Nu))
        g = pull(ib.typs.State(s.X, s.U, None))
        r = ib.lqr.kkt(s, ib.lqr.Params(p.x0, ib.lqr.LQR(*p.lqr)))
        assert np.isfinite(g.lqr.A).all() and np.abs(r.X).max() < (1e-9 if dtype == np.float64 else 1e-2)
n, m, T, B = 4, 2, 6, 3
W = rng.standard_normal((B, n, n))
th = ib.tilqr.Params(rng.standard_normal((B, n)), ib.tilqr.TILQR(Q=np.eye(n) + W @ np.swapaxes(W, -1, -2) / n, Qf=np.tile(np.eye(n), (B, 1, 1)),
                                                                   R=np.tile(0.1 * np.eye(m), (B, 1, 1)), A=0.4 * W, B=rng.standard_normal((B, n, m))))
s, pull = ib.tilqr.build(T).vjp(th)
pull(ib.typs.State(s.X, s.U, None))
n, m, T = 4, 2, 6
cs = ib.ilqr.Problem(family="relu_rnn", horizon=T, state_dim=n, control_dim=m)
solver = ib.ilqr.build(cs, maxiter=5, line_search=ib.make_line_search())
theta = (np.eye(n), 0.01 * np.ones(n), np.eye(n), np.eye(m), np.ones(m), 0.4 * rng.standard_normal((n, n)), rng.standard_normal((n, m)))
params = ib.ilqr.Params(rng.standard_normal(n), theta)
init = ib.typs.State(np.zeros((T, n)), np.zeros((T, m)), np.zeros((T, n)))
s, pull = solver.vjp(init, params)
pull(ib.typs.State(s.X, s.U, None))
# fused iLQR (pendulum) on a ragged batch
B, T = 37, 12
theta = np.tile(np.array([0.05, 9.81, 0.1, 1.0, 1.0, 0.1, 0.01, 100.0, 10.0]), (B, 1))
x0 = np.stack([0.3 * rng.standard_normal(B), 0.1 * rng.standard_normal(B)], axis=1)
cs = ib.ilqr.Problem(family="pendulum", horizon=T, state_dim=2, control_dim=1)
solver = ib.ilqr.build(cs, maxiter=20, line_search=ib.make_line_search())
s, pull = solver.vjp(ib.typs.State(np.zeros((B, T, 2)), np.zeros((B, T, 1)), np.zeros((B, T, 2))), ib.ilqr.Params(x0, theta))
pull(ib.typs.State(s.X, s.U, None))
# the (8,4) kernels on ragged batches around the warp / CTA granularity, with and without a Nu cotangent, batch-reduced too
for dtype in (np.float64, np.float32):
    for B in (1, 7, 8, 9, 23, 24, 25, 65):
        p = O.random_params(rng, 8, 4, 5, batch=(B,), dtype=dtype)
        s, pull = ib.lqr.build(5).vjp(ib.lqr.Params(p.x0, ib.lqr.LQR(*p.lqr)))
        pull(ib.typs.State(s.X, s.U, s.Nu))
        pull(ib.typs.State(s.X, s.U, None))
# round 2 paths: batch-summed gradients (vectors-only sweeps + grad_reduce.cuh) on tuned / mid / generic shapes incl. edge tiles,
# the tensor-core Riccati sweep (mid_mma.cuh), the block-diagonal blqr solver, the C-ABI host pipeline
L = ib._lib
for dtype in (np.float64, np.float32):
    for (n, m, T, B) in ((8, 4, 6, 21), (8, 4, 4, 300), (16, 8, 4, 3), (32, 16, 3, 2), (5, 3, 4, 7), (3, 2, 5, 9), (10, 3, 4, 2)):
        p = O.random_params(rng, n, m, T, batch=(B,), dtype=dtype)
        dp = ib.lqr.DeviceProblem.from_params(ib.lqr.Params(p.x0, ib.lqr.LQR(*p.lqr)))
        plan = ib.lqr.DevicePlan(B, T, n, m, dtype, reduce_batch=True)
        plan.fwd(dp)
        plan.vjp(dp, plan.X, plan.U, plan.Nu)
        plan.vjp(dp, plan.X, plan.U, None)
        assert np.isfinite(plan.grad_buffer.numpy()).all()
for mma in ("1", "0"):
    os.environ["IDOC_MID_MMA"] = mma
    for (n, m, T, B) in ((32, 16, 4, 2), (16, 8, 4, 3), (32, 8, 3, 2), (16, 16, 3, 2)):
        p = O.random_params(rng, n, m, T, batch=(B,), dtype=np.float32)
        s = ib.lqr.build(T).direct(ib.lqr.Params(p.x0, ib.lqr.LQR(*p.lqr)))
        assert np.isfinite(s.X).all()
os.environ.pop("IDOC_MID_MMA")
T, bs, n, m = 5, 37, 3, 2
spd = lambda shape, k: (lambda W_: W_ @ np.swapaxes(W_, -1, -2) / k + np.eye(k))(rng.standard_normal(shape + (k, k)))
bl = ib.blqr.BLQR(Q=spd((T, bs), n), q=rng.standard_normal((T, bs, n)), Qf=spd((bs,), n), qf=rng.standard_normal((bs, n)),
                  M=0.05 * rng.standard_normal((T, bs, n, m)), R=bs * spd((T,), m), r=rng.standard_normal((T, m)),
                  A=0.4 * rng.standard_normal((T, bs, n, n)), B=rng.standard_normal((T, bs, n, m)), d=rng.standard_normal((T, bs, n)))
bp = ib.blqr.Params(rng.standard_normal((bs, n)), bl)
s, pull = ib.blqr.build(T).vjp(bp)
pull(ib.typs.State(s.X, s.U, s.Nu))
pull(ib.typs.State(s.X, s.U, None))
from idoc_b200 import host  # noqa: E402
n, m, T, B = 8, 4, 4, 50
p = O.random_params(rng, n, m, T, batch=(B,))
for reduce in (False, True):
    hp, hs, hg = host.pinned_like_problem(B, T, n, m, np.float64, reduce_batch=reduce)
    for k in ib.lqr.LEAVES:
        hp[k].a[...] = getattr(p.lqr, k)
    hp["x0"].a[...] = p.x0
    nat = host.NativeHostPipeline(B, T, n, m, np.float64, chunk=16, nslots=2, reduce_batch=reduce)
    nat.solve_and_vjp(hp, hs, hg)
    nat.solve_and_vjp(hp, hs, hg, wait=False)
    nat.solve_and_vjp(hp, hs, hg, wait=False)
    nat.finish()
    nat.close()
print("sanitize_small: ok")
