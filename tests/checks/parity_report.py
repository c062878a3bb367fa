#!/usr/bin/env python
"""Measured parity of the CUDA path against the oracle, per output and per gradient leaf (GPU box).

    python tests/checks/parity_report.py [--out gpurun_out/parity.jsonl]

For every case prints one JSON line: max-norm relative error (|ours - oracle|_inf / |oracle|_inf) of X, U, Nu, K, k and
of every gradient leaf (cotangent with and without a Nu part).  For fp32 cases the oracle runs in fp64 on the
fp32-rounded inputs, and the line also carries `budget`: the error of the reference ALGORITHM itself executed in fp32
(oracle under `compute_dtype(float32)`) against that fp64 result — what the reference would lose on the same inputs in
single precision.  north_star's bars: 1e-10 (fp64), 1e-4 (fp32).  The numbers this prints are what DESIGN.md section 5
quotes and what tests/ assert.  (Test infrastructure: imports oracle/.)
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def rel(a, b):
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(np.asarray(a, np.float64) - b)) / max(np.max(np.abs(b)), 1e-300))


def lqr_case(ib, O, n, m, T, B, dtype, seed, harsh=False):
    rng = np.random.default_rng(seed)
    if harsh:
        p = O.reference_test_params(rng, n, m, T, batch=(B,))
        p = O.Params(p.x0.astype(dtype), O.LQR(*[z.astype(dtype) for z in p.lqr]))
    else:
        p = O.random_params(rng, n, m, T, batch=(B,), dtype=dtype)
    so, go = O.direct(p, return_gains=True)
    pp = ib.lqr.Params(p.x0, ib.lqr.LQR(*p.lqr))
    s, pull = ib.lqr.build(T).vjp(pp)
    gains = ib.lqr.backward(pp.lqr, T)
    out = dict(X=rel(s.X, so.X), U=rel(s.U, so.U), Nu=rel(s.Nu, so.Nu), K=rel(gains.K, go.K), k=rel(gains.k, go.k))
    ws = {}
    for tag, nub in (("g", True), ("g0", False)):
        w = O.State(rng.standard_normal(so.X.shape).astype(dtype), rng.standard_normal(so.U.shape).astype(dtype),
                    (rng.standard_normal(so.Nu.shape) if nub else np.zeros(so.Nu.shape)).astype(dtype))
        ws[tag] = w
        g_o = O.vjp_exact(p, so, w)
        g = pull(ib.typs.State(w.X, w.U, w.Nu if nub else None))
        out[tag + "_x0"] = rel(g.x0, g_o.x0)
        for name in ib.lqr.LQR._fields:
            out[tag + "_" + name] = rel(getattr(g.lqr, name), getattr(g_o.lqr, name))
    budget = None
    if dtype == np.float32:
        with O.compute_dtype(np.float32):
            s32, g32 = O.direct(p, return_gains=True)
            gg = O.vjp_exact(p, s32, ws["g"])
        g_o = O.vjp_exact(p, so, ws["g"])
        budget = dict(X=rel(s32.X, so.X), U=rel(s32.U, so.U), Nu=rel(s32.Nu, so.Nu), K=rel(g32.K, go.K), k=rel(g32.k, go.k),
                      g_max=max(rel(getattr(gg.lqr, f), getattr(g_o.lqr, f)) for f in ib.lqr.LQR._fields))
    return out, budget


def tilqr_case(ib, O, TO, n, m, T, B, dtype, seed):
    rng = np.random.default_rng(seed)
    W = rng.standard_normal((B, n, n))
    A = np.stack([O.init_stable(rng, n) for _ in range(B)])
    th = TO.Params(rng.standard_normal((B, n)),
                   TO.TILQR(Q=np.eye(n) + 0.2 * W @ np.swapaxes(W, -1, -2) / n, Qf=np.broadcast_to(np.eye(n), (B, n, n)).copy(),
                            R=np.broadcast_to(0.01 * np.eye(m), (B, m, m)).copy(), A=A, B=rng.standard_normal((B, n, m))))
    th = TO.Params(th.x0.astype(dtype), TO.TILQR(*[z.astype(dtype) for z in th.lqr]))
    so = TO.direct(th, T)
    s, pull = ib.tilqr.build(T).vjp(ib.tilqr.Params(th.x0, ib.tilqr.TILQR(*th.lqr)))
    out = dict(X=rel(s.X, so.X), U=rel(s.U, so.U), Nu=rel(s.Nu, so.Nu))
    w = TO.State(so.X.astype(dtype), so.U.astype(dtype), np.zeros_like(so.Nu).astype(dtype))
    g = pull(ib.typs.State(w.X, w.U, None))
    g_o = TO.vjp_exact(th, so, w, T)
    out["g0_x0"] = rel(g.x0, g_o.x0)
    for name in ib.tilqr.TILQR._fields:
        want = getattr(g_o.lqr, name)
        out["g0_" + name] = rel(getattr(g.lqr, name), want) if np.max(np.abs(want)) > 1e-30 else 0.0
    budget = None
    if dtype == np.float32:
        with O.compute_dtype(np.float32):
            s32 = TO.direct(th, T)
            gg = TO.vjp_exact(th, s32, w, T)
        budget = dict(X=rel(s32.X, so.X), U=rel(s32.U, so.U), Nu=rel(s32.Nu, so.Nu),
                      g_max=max(rel(getattr(gg.lqr, f), getattr(g_o.lqr, f)) for f in ("Q", "R", "A", "B")))
    return out, budget


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    import idoc_b200 as ib
    from oracle import lqr_np as O
    from oracle import tilqr_np as TO
    ib._lib.require_device()
    fh = open(args.out, "w") if args.out else None

    def emit(rec):
        line = json.dumps(rec)
        print(line, flush=True)
        if fh:
            fh.write(line + "\n")
            fh.flush()

    lqr_cases = [  # (n, m, T, B, kind)
        (8, 4, 100, 32, "tuned: BASELINE config 2 shape"), (8, 4, 25, 37, "tuned"), (4, 2, 50, 5, "tuned: BASELINE config 1 shape"),
        (3, 2, 5, 64, "tuned: reference test shape"), (2, 1, 9, 33, "tuned"), (5, 3, 7, 3, "tuned"), (3, 3, 6, 10, "tuned"),
        (4, 1, 11, 17, "tuned"),
        (8, 8, 50, 5, "mid"), (16, 8, 50, 6, "mid"), (16, 8, 200, 4, "mid: config 3 shape, time-varying"), (16, 16, 50, 3, "mid"),
        (24, 8, 50, 3, "mid"), (32, 8, 50, 3, "mid"), (32, 16, 50, 3, "mid"), (32, 16, 200, 4, "mid: BASELINE config 5 shape"),
        (7, 5, 30, 5, "generic"), (10, 3, 30, 3, "generic"), (10, 10, 20, 2, "generic"), (64, 8, 10, 1, "generic"),
        (90, 2, 5, 2, "generic: blqr B=30 lift"),
    ]
    if args.quick:
        lqr_cases = lqr_cases[:2] + [lqr_cases[15]]
    for dtype in (np.float64, np.float32):
        for (n, m, T, B, kind) in lqr_cases:
            try:
                out, budget = lqr_case(ib, O, n, m, T, B, dtype, seed=1000 + n * 100 + m * 10 + T)
                emit(dict(solver="lqr", kind=kind, n=n, m=m, T=T, B=B, dtype=np.dtype(dtype).name, worst=max(out.values()),
                          worst_fwd=max(out[k] for k in ("X", "U", "Nu", "K", "k")), err=out, budget=budget))
            except Exception as e:  # keep going: the report is diagnostic
                emit(dict(solver="lqr", n=n, m=m, T=T, B=B, dtype=np.dtype(dtype).name, error=repr(e)))
        # the reference's own (harsh: R = 1e-4 I) parameterisation, tests/test_lqr.py:9-32
        for (n, m, T, B) in ((3, 2, 5, 8), (4, 2, 50, 4), (8, 4, 100, 4)):
            out, budget = lqr_case(ib, O, n, m, T, B, dtype, seed=42, harsh=True)
            emit(dict(solver="lqr", kind="reference test parameterisation (R = 1e-4 I)", n=n, m=m, T=T, B=B,
                      dtype=np.dtype(dtype).name, worst=max(out.values()),
                      worst_fwd=max(out[k] for k in ("X", "U", "Nu", "K", "k")), err=out, budget=budget))
        for (n, m, T, B) in ((16, 8, 200, 64), (3, 2, 40, 8), (8, 8, 100, 8), (32, 16, 100, 4), (5, 2, 30, 4)):
            try:
                out, budget = tilqr_case(ib, O, TO, n, m, T, B, dtype, seed=16)
                emit(dict(solver="tilqr", n=n, m=m, T=T, B=B, dtype=np.dtype(dtype).name, worst=max(out.values()),
                          worst_fwd=max(out[k] for k in ("X", "U", "Nu")), err=out, budget=budget))
            except Exception as e:
                emit(dict(solver="tilqr", n=n, m=m, T=T, B=B, dtype=np.dtype(dtype).name, error=repr(e)))


if __name__ == "__main__":
    main()
