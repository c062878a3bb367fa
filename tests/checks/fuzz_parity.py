#!/usr/bin/env python
"""Randomised differential test of the CUDA path against the oracle over the corners the unit tests sample sparsely:
every compiled shape family (tuned, mid-size register-tile, mid-size tensor-core, generic) x {fp64, fp32} x tiny / odd batches
(B in 1, 2, 3, 5, 33) x short / odd horizons (T in 1, 2, 3, 7, 33) x {time-varying lqr, tilqr}, gains + solution + every
gradient leaf with a full cotangent.  Prints one JSON line per case and a summary; exit code 1 if any case misses north_star's
bar (1e-10 fp64; fp32: 1e-4 or 4x the fp32 error budget of the reference algorithm on the same inputs).

    python tests/checks/fuzz_parity.py [--cases 120] [--seed 0] [--out gpurun_out/fuzz.jsonl]        (test infrastructure: imports oracle/)"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from parity_report import lqr_case, rel  # noqa: E402

SHAPES = [(8, 4), (4, 2), (2, 1), (3, 2), (5, 3), (8, 8), (16, 8), (16, 16), (24, 8), (32, 8), (32, 16), (7, 3), (12, 5), (20, 6)]


def tilqr_case(ib, O, TO, n, m, T, B, dtype, seed):
    rng = np.random.default_rng(seed)
    W = rng.standard_normal((B, n, n))
    Wr = rng.standard_normal((B, m, m))
    A = np.stack([O.init_stable(rng, n) for _ in range(B)])
    th = TO.Params(rng.standard_normal((B, n)).astype(dtype),
                   TO.TILQR(*[z.astype(dtype) for z in (np.eye(n) + 0.2 * W @ np.swapaxes(W, -1, -2) / n, np.broadcast_to(np.eye(n), (B, n, n)).copy(),
                                                        np.eye(m) + 0.2 * Wr @ np.swapaxes(Wr, -1, -2) / m, A, rng.standard_normal((B, n, m)))]))
    so = TO.direct(th, T)
    w = TO.State(rng.standard_normal(so.X.shape).astype(dtype), rng.standard_normal(so.U.shape).astype(dtype),
                 rng.standard_normal(so.Nu.shape).astype(dtype))
    g_o = TO.vjp_exact(th, so, w, T)
    s, pull = ib.tilqr.build(T).vjp(ib.tilqr.Params(th.x0, ib.tilqr.TILQR(*th.lqr)))
    g = pull(ib.typs.State(w.X, w.U, w.Nu))
    out = dict(X=rel(s.X, so.X), U=rel(s.U, so.U), Nu=rel(s.Nu, so.Nu), g_x0=rel(g.x0, g_o.x0))
    for name in ("Q", "Qf", "R", "A", "B"):
        out["g_" + name] = rel(getattr(g.lqr, name), getattr(g_o.lqr, name))
    budget = None
    if dtype == np.float32:
        with O.compute_dtype(np.float32):
            s32 = TO.direct(th, T)
            g32 = TO.vjp_exact(th, s32, w, T)
        budget = dict(X=rel(s32.X, so.X), g_max=max(rel(getattr(g32.lqr, f), getattr(g_o.lqr, f)) for f in ("Q", "Qf", "R", "A", "B")))
    return out, budget


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=120)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import idoc_b200 as ib
    from oracle import lqr_np as O
    from oracle import tilqr_np as TO
    ib._lib.require_device()
    rng = np.random.default_rng(a.seed)
    fh = open(a.out, "w") if a.out else None
    worst = {"float64": 0.0, "float32": 0.0}
    bad = 0
    for c in range(a.cases):
        n, m = SHAPES[c % len(SHAPES)]
        T = int(rng.choice([1, 2, 3, 7, 33]))
        B = int(rng.choice([1, 2, 3, 5, 33]))
        dtype = np.float64 if rng.random() < 0.5 else np.float32
        ti = bool(rng.random() < 0.4)
        seed = int(rng.integers(1 << 30))
        if ti:
            out, budget = tilqr_case(ib, O, TO, n, m, T, B, dtype, seed)
        else:
            out, budget = lqr_case(ib, O, n, m, T, B, dtype, seed)
        err = max(out.values())
        bar = 1e-10 if dtype == np.float64 else max(1e-4, 4 * max(budget.values()))
        ok = bool(err < bar)
        bad += not ok
        worst[np.dtype(dtype).name] = max(worst[np.dtype(dtype).name], err)
        line = json.dumps(dict(case=c, n=n, m=m, T=T, B=B, dtype=np.dtype(dtype).name, tilqr=ti, max_rel_err=err, bar=bar, ok=ok,
                               worst_leaf=max(out, key=out.get)))
        print(line, flush=True)
        if fh:
            fh.write(line + "\n")
    summary = json.dumps(dict(summary=True, cases=a.cases, failed=bad, worst_rel_err=worst))
    print(summary)
    if fh:
        fh.write(summary + "\n")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
