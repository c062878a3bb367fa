#!/usr/bin/env python
"""Measured error of the CUDA path on every committed golden vector (tests/golden/*.npz: outputs of the unmodified
reference sources, tests/golden/make_golden.py), per solver: forward outputs and every gradient leaf.  One JSON line per
golden file.  The tests assert north_star's bars (1e-10 in fp64); this prints the achieved numbers they rest on."""
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
G = os.path.join(ROOT, "tests", "golden")
rel = lambda a, b: float(np.max(np.abs(np.asarray(a, np.float64) - b)) / max(np.max(np.abs(b)), 1e-300))
FIELDS = ("Q", "q", "Qf", "R", "r", "A", "B")


def main():
    import idoc_b200 as ib
    ib._lib.require_device()
    out = []
    for path in sorted(glob.glob(os.path.join(G, "*.npz"))):
        name = os.path.basename(path)[:-4]
        z = np.load(path)
        e = {}
        if name.startswith("lqr_"):
            p = ib.lqr.Params(z["x0"], ib.lqr.LQR(**{k: z[k] for k in ib.lqr.LQR._fields}))
            T = z["A"].shape[0]
            s, pull = ib.lqr.build(T).vjp(p)
            gains = ib.lqr.backward(p.lqr, T)
            e.update(X=rel(s.X, z["X"]), U=rel(s.U, z["U"]), Nu=rel(s.Nu, z["Nu"]), K=rel(gains.K, z["K"]), k=rel(gains.k, z["k"]))
            if "shift" not in name:
                for pre, ct in (("g_", ib.typs.State(z["w_X"], z["w_U"], z["w_Nu"])), ("g0_", ib.typs.State(z["X"], z["U"], None))):
                    g = pull(ct)
                    e[pre + "x0"] = rel(g.x0, z[pre + "x0"])
                    for f in ib.lqr.LQR._fields:
                        e[pre + f] = rel(getattr(g.lqr, f), z[pre + f])
        elif name.startswith("tilqr_"):
            T = z["X"].shape[0]
            th = ib.tilqr.Params(z["x0"], ib.tilqr.TILQR(Q=z["Q"], Qf=z["Qf"], R=z["R"], A=z["A"], B=z["B"]))
            s, pull = ib.tilqr.build(T).vjp(th)
            e.update(X=rel(s.X, z["X"]), U=rel(s.U, z["U"]), Nu=rel(s.Nu, z["Nu"]))
            g = pull(ib.typs.State(z["w_X"], z["w_U"], z["w_Nu"]))
            e["g_x0"] = rel(g.x0, z["g_x0"])
            for f in ib.tilqr.TILQR._fields:
                e["g_" + f] = rel(getattr(g.lqr, f), z["g_" + f])
        elif name.startswith("blqr_"):
            p = ib.blqr.Params(z["x0"], ib.blqr.BLQR(**{k: z[k] for k in ib.blqr.BLQR._fields}))
            T = z["A"].shape[0]
            s, pull = ib.blqr.build(T).vjp(p)
            gains = ib.blqr.backward(p.blqr, T)
            e.update(X=rel(s.X, z["X"]), U=rel(s.U, z["U"]), Nu=rel(s.Nu, z["Nu"]), K=rel(gains.K, z["K"]), k=rel(gains.k, z["k"]))
            g = pull(ib.typs.State(z["w_X"], z["w_U"], z["w_Nu"]))
            e["g_x0"] = rel(g.x0, z["g_x0"])
            for f in ib.blqr.BLQR._fields:
                e["g_" + f] = rel(getattr(g.blqr, f), z["g_" + f])
        elif name.startswith("ilqr_") or name.startswith("bilqr_"):
            bi = name.startswith("bilqr_")
            mod = ib.bilqr if bi else ib.ilqr
            if bi:
                T, bs, n = z["X"].shape
            else:
                T, n = z["X"].shape
            m = z["U"].shape[1]
            cs = mod.Problem(family="relu_ax" if bi else "relu_rnn", horizon=T, state_dim=n, control_dim=m)
            use_ls = bool(z["use_ls"]) if "use_ls" in z.files else True
            solver = mod.build(cs, maxiter=int(z["maxiter"]), thres=1e-8, line_search=ib.make_line_search() if use_ls else None)
            params = mod.Params(z["x0"], tuple(z["th_" + k] for k in FIELDS))
            s, pull = solver.vjp(ib.typs.State(z["X0"], z["U0"], np.zeros_like(z["X0"])), params)
            e.update(X=rel(s.X, z["X"]), U=rel(s.U, z["U"]), Nu=rel(s.Nu, z["Nu"]))
            if "g_x0" in z.files:
                g = pull(ib.typs.State(z["w_X"], z["w_U"], z["w_Nu"]))
                e["g_x0"] = rel(g.x0, z["g_x0"])
                for k, got in zip(FIELDS, g.theta):
                    e["g_" + k] = rel(got, z["g_" + k])
        rec = dict(golden=name, worst=max(e.values()), worst_key=max(e, key=e.get), err=e)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    return out


if __name__ == "__main__":
    main()
