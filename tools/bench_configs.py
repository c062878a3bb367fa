#!/usr/bin/env python
"""Measurement of the BASELINE.json configurations other than the headline (bench.py covers config 2):

  config 3  tilqr   time-invariant LQR  n=16 m=8  T=200   (FMA roofline: algorithmic FLOPs of SURVEY 8d)
  config 4  iLQR    pendulum swing-up   n=2  m=1  T=100   (device linearise / Riccati / line-search loop)
  config 5  LQR     large state         n=32 m=16 T=200   (fwd + VJP, batch-reduced gradients)

Each leg runs a bounded batch on cuda:0 (CUDA events, inputs resident in HBM, fwd + VJP per step), reports solves/s,
the roofline fraction (HBM for the byte-bound legs, FP32/FP64 FMA for tilqr, nominal 74.4 / 37.2 TFLOP/s), and times
the NumPy oracle on a small sample of the same problems on the host beside it.  One JSON line per leg.

    python tools/bench_configs.py [--legs tilqr,ilqr,lqr32] [--dtype f32] [--scale 1.0]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import idoc_b200 as ib  # noqa: E402

L = ib._lib


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


KERNELS = {}


def timeit(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    L.sync()
    e0, e1 = L.Event(), L.Event()
    L.profile_begin()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    e1.sync()
    L.sync()
    KERNELS.clear()
    KERNELS.update({k: round(float(np.sum(v)) / steps, 4) for k, v in L.profile_end().items()})
    return e0.elapsed_ms(e1) / steps


def stable(rng, shape, n):
    W = rng.standard_normal(shape + (n, n))
    q, _ = np.linalg.qr(W)
    return 0.5 * q


def leg_tilqr(dtype, B, steps, warmup):
    n, m, T = 16, 8, 200
    rng = np.random.default_rng(0)
    W = rng.standard_normal((B, n, n))
    Wr = rng.standard_normal((B, m, m))
    th = ib.tilqr.Params(rng.standard_normal((B, n)),
                         ib.tilqr.TILQR(Q=W @ W.transpose(0, 2, 1) / n + np.eye(n), Qf=np.tile(np.eye(n), (B, 1, 1)),
                                        R=Wr @ Wr.transpose(0, 2, 1) / m + np.eye(m), A=stable(rng, (B,), n),
                                        B=rng.standard_normal((B, n, m))))
    plan = ib.tilqr.DevicePlan(B, T, n, m, dtype, vjp=True)
    plan.load(th)

    def step():
        plan.fwd()
        plan.vjp(plan.X, plan.U, None)

    ms = timeit(step, steps, warmup)
    flop = (2 * n * n * m + 2 * n * m * m + 2 * n * n * m + m ** 3 / 3 + 2 * m * m * n + 4 * n ** 3 + 4 * n * n * m
            + 2 * m * n + 2 * n * n + 2 * n * m + 4 * n * n
            + 4 * n * n + 4 * n * m + 2 * m * m + 2 * m * n + 2 * n * n + 4 * n * m + 4 * n * n + 2 * n * m
            + 4 * n * n + 4 * n * m + 2 * n * n + 2 * m * m) * T
    peak = 74.4e12 if np.dtype(dtype) == np.float32 else 37.2e12
    sps = B / (ms * 1e-3)
    # CPU: NumPy oracle on a small sample
    from oracle import tilqr_np as TO
    nb = 8
    t0 = time.perf_counter()
    for b in range(nb):
        TO.direct(TO.Params(th.x0[b], TO.TILQR(*[np.asarray(z[b], np.float64) for z in th.lqr])), T)
    cpu = nb / (time.perf_counter() - t0)
    return dict(leg="config 3: tilqr n=16 m=8 T=200", dtype=np.dtype(dtype).name, batch=B, ms_per_step=ms, value=sps,
                unit="solves/s", roofline=dict(bound="fma", algorithmic_flop_per_solve=flop, achieved=sps * flop / 1e12,
                                               peak=peak / 1e12, unit="TFLOP/s", frac=sps * flop / peak,
                                               peak_source="nominal FP32/FP64 FMA (SURVEY 8d)"),
                cpu_baseline=dict(value=cpu, unit="solves/s (forward solve only)", cores=1, kind="port",
                                  sample=f"{nb} problems, NumPy oracle (oracle/tilqr_np.py)"),
                kernels="mid-size warp-per-problem kernels, time-invariant flag (csrc/mid.cuh)")


def leg_lqr32(dtype, B, steps, warmup):
    n, m, T = 32, 16, 200
    stream = L.Stream()
    prob = ib.lqr.DeviceProblem.generate(B, T, n, m, dtype, seed=7, stream=stream)
    plan = ib.lqr.DevicePlan(B, T, n, m, dtype, vjp=True, reduce_batch=True)

    def step():
        plan.fwd(prob, stream=stream)
        plan.vjp(prob, plan.X, plan.U, None, stream=stream)

    for _ in range(warmup):
        step()
    stream.sync()
    e0, e1 = L.Event(), L.Event()
    L.profile_begin()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    stream.sync()
    L.sync()
    KERNELS.clear()
    KERNELS.update({k: round(float(np.sum(v)) / steps, 4) for k, v in L.profile_end().items()})
    ms = e0.elapsed_ms(e1) / steps
    sz = np.dtype(dtype).itemsize
    per_step = 2 * n * n + 2 * n * m + m * m + 2 * n + m
    once = n * n + 2 * n
    sol = 2 * n + m
    scal = (2 * per_step + 3 * sol) * T + 2 * once  # batch-reduced-gradient variant
    sps = B / (ms * 1e-3)
    peak = hbm_peak()
    return dict(leg="config 5: LQR n=32 m=16 T=200, fwd + VJP, batch-reduced gradients", dtype=np.dtype(dtype).name, batch=B,
                ms_per_step=ms, value=sps, unit="solves/s",
                roofline=dict(bound="hbm", algorithmic_bytes_per_solve=scal * sz, achieved=sps * scal * sz / 1e9, peak=peak,
                              unit="GB/s", frac=sps * scal * sz / 1e9 / peak,
                              fma=dict(algorithmic_flop_per_solve=148e6, achieved_tflops=sps * 148e6 / 1e12,
                                       frac_of_nominal_fma=sps * 148e6 / (74.4e12 if sz == 4 else 37.2e12),
                                       note="SURVEY 8d: ~148 MFLOP per solve, 17.6 FLOP/B in fp32: above the FMA ridge")),
                kernels="mid-size warp-per-problem kernels (csrc/mid.cuh)")


def leg_ilqr(dtype, B, steps, warmup, umax=None):
    n, m, T = 2, 1, 100
    rng = np.random.default_rng(3)
    cs = ib.ilqr.Problem(family="pendulum", horizon=T, state_dim=n, control_dim=m)
    pid = ib.ilqr.FAMILIES["pendulum"]
    tdim = L.lib.idoc_ilqr_theta_dim(pid, n, m)
    # theta = [h, g/l, b, c, w1, w2, wu, f1, f2]
    base = np.array([0.05, 9.81, 0.1, 1.0, 1.0, 0.1, 0.01, 100.0, 10.0])[:tdim]
    theta = np.tile(base, (B, 1))
    theta[:, 1] *= 1.0 + 0.1 * rng.standard_normal(B)
    x0 = np.stack([0.3 * rng.standard_normal(B), 0.1 * rng.standard_normal(B)], -1)   # hanging down, swing up to pi
    code = L.dtype_code(dtype)
    th = L.DeviceArray.from_numpy(theta.astype(dtype))
    x0d = L.DeviceArray.from_numpy(x0.astype(dtype))
    Xi = np.zeros((B, T, n))
    x = x0.copy()
    for t in range(T):  # initial trajectory: the uncontrolled rollout
        h, gl, bb = theta[:, 0], theta[:, 1], theta[:, 2]
        x = np.stack([x[:, 0] + h * x[:, 1], x[:, 1] + h * (-gl * np.sin(x[:, 0]) - bb * x[:, 1])], axis=1)
        Xi[:, t] = x
    X0 = L.DeviceArray.from_numpy(Xi.astype(dtype))
    U0 = L.DeviceArray((B, T, m), dtype, zero=True)
    X, U, Nu = L.DeviceArray((B, T, n), dtype), L.DeviceArray((B, T, m), dtype), L.DeviceArray((B, T, n), dtype)
    wsb = L.lib.idoc_ilqr_ws_bytes(code, B, T, n, m, 1)
    ws = L.DeviceArray((wsb,), np.uint8)
    iters, lsf = L.DeviceArray((B,), np.int32), L.DeviceArray((B,), np.int32)
    vwsb = L.lib.idoc_ilqr_vjp_ws_bytes(code, B, T, n, m, 1)
    vws = L.DeviceArray((vwsb,), np.uint8)
    gth, gx0 = L.DeviceArray((B, tdim), dtype), L.DeviceArray((B, n), dtype)
    ls = ib.make_line_search()
    maxiter = 200
    if umax is not None:
        ulo = L.DeviceArray.from_numpy(np.full((B, m), -umax, dtype))
        uhi = L.DeviceArray.from_numpy(np.full((B, m), umax, dtype))

    def step():
        L.check(L.lib.idoc_memcpy_d2d(X.ptr, X0.ptr, X.nbytes, None), "d2d")
        L.check(L.lib.idoc_memcpy_d2d(U.ptr, U0.ptr, U.nbytes, None), "d2d")
        if umax is None:
            L.check(L.lib.idoc_ilqr_solve(None, code, pid, B, T, n, m, 1, th.ptr, tdim, x0d.ptr, X.ptr, U.ptr, Nu.ptr, maxiter, 1e-8, 1,
                                          float(ls.lower), float(ls.upper), float(ls.rho), int(ls.maxiter), iters.ptr, lsf.ptr,
                                          ws.ptr, wsb, 1), "ilqr_solve")
            L.check(L.lib.idoc_ilqr_vjp(None, code, pid, B, T, n, m, 1, th.ptr, tdim, x0d.ptr, X.ptr, U.ptr, Nu.ptr, X.ptr, U.ptr,
                                        None, gth.ptr, gx0.ptr, vws.ptr, vwsb), "ilqr_vjp")
        else:  # control-limited
            L.check(L.lib.idoc_ilqr_solve_box(None, code, pid, B, T, n, m, th.ptr, tdim, x0d.ptr, ulo.ptr, uhi.ptr, X.ptr, U.ptr,
                                              Nu.ptr, maxiter, 1e-8, int(os.environ.get("IDOC_BENCH_BOX_LS", "1")), float(ls.lower), float(ls.upper), float(ls.rho),
                                              int(ls.maxiter), iters.ptr, lsf.ptr, ws.ptr, wsb), "ilqr_solve_box")
            L.check(L.lib.idoc_ilqr_vjp_box(None, code, pid, B, T, n, m, th.ptr, tdim, x0d.ptr, ulo.ptr, uhi.ptr, X.ptr, U.ptr, Nu.ptr,
                                            X.ptr, U.ptr, None, gth.ptr, gx0.ptr, vws.ptr, vwsb), "ilqr_vjp_box")

    ms = timeit(step, steps, warmup)
    it = iters.numpy()
    hist = np.bincount(it, minlength=maxiter + 1)
    extra = {}
    if umax is not None:
        Uh = U.numpy()
        extra = dict(control_limit=umax, steps_at_a_bound=float((np.abs(Uh) >= umax).mean()),
                     within_limits=bool(np.abs(Uh).max() <= umax), finite=bool(np.isfinite(X.numpy()).all()))
    return dict(**extra, leg="config 4: iLQR pendulum swing-up n=2 m=1 T=100" + (" with box control limits" if umax else "")
                + " (solve to 1e-8 + implicit VJP)", dtype=np.dtype(dtype).name,
                batch=B, ms_per_step=ms, value=B / (ms * 1e-3), unit="solves/s",
                iterations=dict(mean=float(it.mean()), max=int(it.max()), hit_maxiter=int((it >= maxiter).sum()),
                                histogram={int(k): int(v) for k, v in enumerate(hist) if v}),
                line_search_failures=int((lsf.numpy() > 0).sum()),
                roofline=dict(bound="latency", note="data-dependent iteration count; n=2 blocks: launch- and latency-bound"))


def leg_bilqr(dtype, B, steps, warmup, umax=4.0, systems=4):
    """config 4, the bilqr half: `systems` pendulums (same parameters, different initial states around the upright position)
    balanced by ONE shared torque with a box limit -- B independent shared-control problems, one kernel for the solve and one
    for the implicit VJP (idoc_bilqr_solve_box / idoc_bilqr_vjp_box, csrc/bilqr_fused.cu).  A STRESS case, not part of bench.py's
    line: at T = 100 this non-convex compromise problem leaves a quarter of the batch at the iteration cap with plain iLQR steps
    (the reference's algorithm has no globalisation beyond its ratio line search), and the implicit gradient of a problem that
    did not reach a stationary point is meaningless (non-finite for some): measured 69 k shared-control solves/s in fp64, 128 k
    in fp32 at B = 4096 x 4 systems (profiles/r2_bilqr_box_stress.jsonl); the solutions themselves are finite and within limits."""
    n, m, T = 2, 1, 100
    rng = np.random.default_rng(5)
    pid = ib.ilqr.FAMILIES["pendulum"]
    tdim = L.lib.idoc_ilqr_theta_dim(pid, n, m)
    # one torque cannot balance several pendulums at different angles: the problem is a compromise (all start on the same side of
    # the upright position; control cost high enough that the optimum is not bang-bang: the set-up of tests/test_gpu_bilqr.py)
    base = np.array([0.05, 9.81, 0.1, 1.0, 1.0, 0.1, 0.2, 20.0, 2.0])
    theta = np.tile(base, (B, 1))
    theta[:, 1] *= 1.0 + 0.05 * rng.standard_normal(B)
    x0 = np.stack([np.pi + 0.3 + 0.1 * rng.standard_normal((B, systems)), 0.2 * rng.standard_normal((B, systems))], -1)  # [B, systems, 2]
    Xi = np.zeros((B, T, systems, n))
    x = x0.copy()
    for t in range(T):  # initial trajectory: the uncontrolled rollout of every system
        h, gl, bb = theta[:, 0, None], theta[:, 1, None], theta[:, 2, None]
        x = np.stack([x[..., 0] + h * x[..., 1], x[..., 1] + h * (-gl * np.sin(x[..., 0]) - bb * x[..., 1])], axis=-1)
        Xi[:, t] = x
    N = n * systems
    code = L.dtype_code(dtype)
    th = L.DeviceArray.from_numpy(theta.astype(dtype))
    x0d = L.DeviceArray.from_numpy(x0.reshape(B, N).astype(dtype))
    X0 = L.DeviceArray.from_numpy(Xi.reshape(B, T, N).astype(dtype))
    U0 = L.DeviceArray((B, T, m), dtype, zero=True)
    X, U, Nu = L.DeviceArray((B, T, N), dtype), L.DeviceArray((B, T, m), dtype), L.DeviceArray((B, T, N), dtype)
    wsb = L.lib.idoc_ilqr_ws_bytes(code, B, T, n, m, systems)
    ws = L.DeviceArray((wsb,), np.uint8)
    iters, lsf = L.DeviceArray((B,), np.int32), L.DeviceArray((B,), np.int32)
    vwsb = L.lib.idoc_ilqr_vjp_ws_bytes(code, B, T, n, m, systems)
    vws = L.DeviceArray((vwsb,), np.uint8)
    gth, gx0 = L.DeviceArray((B, tdim), dtype), L.DeviceArray((B, N), dtype)
    ls = ib.make_line_search()
    maxiter = 200
    ulo = L.DeviceArray.from_numpy(np.full((B, m), -umax, dtype))
    uhi = L.DeviceArray.from_numpy(np.full((B, m), umax, dtype))

    def step():
        L.check(L.lib.idoc_memcpy_d2d(X.ptr, X0.ptr, X.nbytes, None), "d2d")
        L.check(L.lib.idoc_memcpy_d2d(U.ptr, U0.ptr, U.nbytes, None), "d2d")
        L.check(L.lib.idoc_bilqr_solve_box(None, code, pid, B, T, n, m, systems, th.ptr, tdim, x0d.ptr, ulo.ptr, uhi.ptr, X.ptr, U.ptr,
                                           Nu.ptr, maxiter, 1e-8, int(os.environ.get("IDOC_BENCH_BILQR_LS", "0")), float(ls.lower), float(ls.upper),
                                           float(ls.rho), int(ls.maxiter), iters.ptr, lsf.ptr, ws.ptr, wsb), "bilqr_solve_box")  # full steps: DESIGN section 10
        L.check(L.lib.idoc_bilqr_vjp_box(None, code, pid, B, T, n, m, systems, th.ptr, tdim, x0d.ptr, ulo.ptr, uhi.ptr, X.ptr, U.ptr,
                                         Nu.ptr, X.ptr, U.ptr, None, gth.ptr, gx0.ptr, vws.ptr, vwsb), "bilqr_vjp_box")

    ms = timeit(step, steps, warmup)
    it = iters.numpy()
    Uh = U.numpy()
    conv = it < maxiter  # the implicit gradient is defined at a stationary point: judged on the problems that converged
    return dict(leg=f"config 4 (bilqr): {systems} pendulums under one shared torque |u| <= {umax}, T=100 (solve to 1e-8 + implicit VJP)",
                dtype=np.dtype(dtype).name, batch=B, systems=systems, ms_per_step=ms, value=B / (ms * 1e-3),
                unit="shared-control solves/s", system_solves_per_s=B * systems / (ms * 1e-3), control_limit=umax,
                steps_at_a_bound=float((np.abs(Uh) >= umax).mean()), within_limits=bool(np.abs(Uh).max() <= umax),
                finite=bool(np.isfinite(X.numpy()).all()), gradient_finite_where_converged=bool(np.isfinite(gth.numpy()[conv]).all()),
                iterations=dict(mean=float(it.mean()), max=int(it.max()), hit_maxiter=int((it >= maxiter).sum())),
                roofline=dict(bound="latency", note="data-dependent iteration count; thread-per-problem fused kernel"))


def leg_cartpole(dtype, B, steps, warmup, umax=None):
    """cart-pole (device-autodiff family, n=4, m=1): stabilisation around the upright position from perturbed states"""
    n, m, T = 4, 1, 100
    rng = np.random.default_rng(4)
    pid = ib.ilqr.FAMILIES["cartpole"]
    tdim = L.lib.idoc_ilqr_theta_dim(pid, n, m)
    base = np.array([0.02, 1.0, 0.1, 0.5, 9.81, 0.1, 1.0, 0.1, 0.1, 0.01, 10.0, 100.0, 10.0, 10.0])
    theta = np.tile(base, (B, 1))
    theta[:, 2] *= 1.0 + 0.1 * rng.standard_normal(B)
    x0 = np.stack([0.1 * rng.standard_normal(B), np.pi + 0.2 * rng.standard_normal(B), 0.1 * rng.standard_normal(B),
                   0.1 * rng.standard_normal(B)], -1)
    Xi = np.zeros((B, T, n))
    x = x0.copy()
    for t in range(T):  # initial trajectory: the uncontrolled rollout
        h, mc, mp, l, g = (theta[:, i] for i in range(5))
        sn, cs = np.sin(x[:, 1]), np.cos(x[:, 1])
        pdd = (mp * sn * (l * x[:, 3] ** 2 + g * cs)) / (mc + mp * sn ** 2)
        tdd = -(pdd * cs + g * sn) / l
        x = np.stack([x[:, 0] + h * x[:, 2], x[:, 1] + h * x[:, 3], x[:, 2] + h * pdd, x[:, 3] + h * tdd], axis=1)
        Xi[:, t] = x
    code = L.dtype_code(dtype)
    th = L.DeviceArray.from_numpy(theta.astype(dtype))
    x0d = L.DeviceArray.from_numpy(x0.astype(dtype))
    X0 = L.DeviceArray.from_numpy(Xi.astype(dtype))
    U0 = L.DeviceArray((B, T, m), dtype, zero=True)
    X, U, Nu = L.DeviceArray((B, T, n), dtype), L.DeviceArray((B, T, m), dtype), L.DeviceArray((B, T, n), dtype)
    wsb = L.lib.idoc_ilqr_ws_bytes(code, B, T, n, m, 1)
    ws = L.DeviceArray((wsb,), np.uint8)
    iters, lsf = L.DeviceArray((B,), np.int32), L.DeviceArray((B,), np.int32)
    vwsb = L.lib.idoc_ilqr_vjp_ws_bytes(code, B, T, n, m, 1)
    vws = L.DeviceArray((vwsb,), np.uint8)
    gth, gx0 = L.DeviceArray((B, tdim), dtype), L.DeviceArray((B, n), dtype)
    ls = ib.make_line_search()
    maxiter = 200
    if umax is not None:
        ulo = L.DeviceArray.from_numpy(np.full((B, m), -umax, dtype))
        uhi = L.DeviceArray.from_numpy(np.full((B, m), umax, dtype))

    def step():
        L.check(L.lib.idoc_memcpy_d2d(X.ptr, X0.ptr, X.nbytes, None), "d2d")
        L.check(L.lib.idoc_memcpy_d2d(U.ptr, U0.ptr, U.nbytes, None), "d2d")
        if umax is None:
            L.check(L.lib.idoc_ilqr_solve(None, code, pid, B, T, n, m, 1, th.ptr, tdim, x0d.ptr, X.ptr, U.ptr, Nu.ptr, maxiter, 1e-8, 1,
                                          float(ls.lower), float(ls.upper), float(ls.rho), int(ls.maxiter), iters.ptr, lsf.ptr, ws.ptr,
                                          wsb, 1), "ilqr_solve")
            L.check(L.lib.idoc_ilqr_vjp(None, code, pid, B, T, n, m, 1, th.ptr, tdim, x0d.ptr, X.ptr, U.ptr, Nu.ptr, X.ptr, U.ptr, None,
                                        gth.ptr, gx0.ptr, vws.ptr, vwsb), "ilqr_vjp")
        else:  # force limit |u| <= umax
            L.check(L.lib.idoc_ilqr_solve_box(None, code, pid, B, T, n, m, th.ptr, tdim, x0d.ptr, ulo.ptr, uhi.ptr, X.ptr, U.ptr,
                                              Nu.ptr, maxiter, 1e-8, 1, float(ls.lower), float(ls.upper), float(ls.rho),
                                              int(ls.maxiter), iters.ptr, lsf.ptr, ws.ptr, wsb), "ilqr_solve_box")
            L.check(L.lib.idoc_ilqr_vjp_box(None, code, pid, B, T, n, m, th.ptr, tdim, x0d.ptr, ulo.ptr, uhi.ptr, X.ptr, U.ptr, Nu.ptr,
                                            X.ptr, U.ptr, None, gth.ptr, gx0.ptr, vws.ptr, vwsb), "ilqr_vjp_box")

    ms = timeit(step, steps, warmup)
    it = iters.numpy()
    hist = np.bincount(it, minlength=1)
    Xh = X.numpy()
    Uh_ = U.numpy()
    # final cost of every problem (host restatement of CartpoleF's cost, csrc/ilqr.cuh): compares the optima different rules reach
    sX_ = np.concatenate((x0[:, None, :], Xh[:, :-1].astype(np.float64)), axis=1)
    w = theta
    run = 0.5 * (w[:, None, 5] * sX_[..., 0] ** 2 + w[:, None, 6] * (sX_[..., 1] - np.pi) ** 2 + w[:, None, 7] * sX_[..., 2] ** 2
                 + w[:, None, 8] * sX_[..., 3] ** 2 + w[:, None, 9] * Uh_[..., 0].astype(np.float64) ** 2).sum(axis=1)
    xT = Xh[:, -1].astype(np.float64)
    fin = 0.5 * (w[:, 10] * xT[:, 0] ** 2 + w[:, 11] * (xT[:, 1] - np.pi) ** 2 + w[:, 12] * xT[:, 2] ** 2 + w[:, 13] * xT[:, 3] ** 2)
    cost_final = run + fin
    finite = np.isfinite(Xh).all(axis=(1, 2))
    upright = finite & (np.abs(Xh[:, -1, 1] - np.pi) < 0.05)
    extra = {}
    if umax is not None:
        Uh = U.numpy()
        extra = dict(control_limit=umax, steps_at_a_bound=float((np.abs(Uh) >= umax).mean()), within_limits=bool(np.abs(Uh).max() <= umax))
    return dict(**extra, nonfinite=int((~finite).sum()), ends_upright=int(upright.sum()),
                final_cost=dict(mean=float(np.nanmean(cost_final)), median=float(np.nanmedian(cost_final))),
                leg="config 4: iLQR cart-pole n=4 m=1 T=100" + (" with a force limit" if umax else "") + " (device-autodiff family; solve to 1e-8 + implicit VJP)",
                dtype=np.dtype(dtype).name, batch=B, ms_per_step=ms, value=B / (ms * 1e-3), unit="solves/s",
                iterations=dict(mean=float(it.mean()), max=int(it.max()), hit_maxiter=int((it >= maxiter).sum()),
                                histogram={int(k): int(v) for k, v in enumerate(hist) if v}),
                line_search_failures=int((lsf.numpy() > 0).sum()),
                roofline=dict(bound="latency", note="data-dependent iteration count; thread per problem"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--legs", default="tilqr,ilqr,lqr32")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--scale", type=float, default=1.0, help="batch multiplier (1.0 = the bounded default batches)")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    dtype = np.float32 if args.dtype == "f32" else np.float64
    L.require_device()
    legs = dict(tilqr=(leg_tilqr, 16384), ilqr=(leg_ilqr, 16384), bilqr_box=(leg_bilqr, 4096), ilqr_box=(lambda *a: leg_ilqr(*a, umax=float(os.environ.get('IDOC_BENCH_UMAX', '8'))), 16384),
                cartpole=(leg_cartpole, 16384),
                cartpole_box=(lambda *a: leg_cartpole(*a, umax=float(os.environ.get('IDOC_BENCH_FMAX', '3'))), 16384), lqr32=(leg_lqr32, 1024))
    for name in args.legs.split(","):
        fn, B = legs[name]
        out = fn(dtype, max(1, int(B * args.scale)), args.steps, args.warmup)
        out["kernel_ms_per_step"] = dict(KERNELS)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
