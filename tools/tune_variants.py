#!/usr/bin/env python
"""Tuning sweep over the kernel variants of the (8,4) LQR kernels (library built with IDOC_BUILD_TUNABLE=1).

For each kernel, every (NBUF, BM) variant compiled into shape_inst.cuh is selected through the IDOC_* environment
overrides, timed with the library's per-kernel CUDA events at BASELINE config 2, and its outputs are compared
bit-for-bit (hash of every output array) against the default variant: the variants differ only in how bytes
move, never in arithmetic.

    python tools/tune_variants.py [--dtype f64] [--batch 65536] [--reps 3] [--kernels ric,roll,abwd,afwd]
"""
import argparse
import json
import os
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import idoc_b200 as ib  # noqa: E402

L = ib._lib

VARIANTS = {
    "ric": ("IDOC_NBUF_R", "IDOC_BM_RIC", [(1, 0), (1, 10), (1, 26), (1, 60), (1, 124), (1, 122), (1, 90)]
            + [(1, 124 + 256 * p) for p in range(1, 8)] + [(1, 10 + 256 * p) for p in range(1, 8)]),
    "roll": ("IDOC_NBUF_O", "IDOC_BM_ROLL", [(2, 0), (2, 1), (2, 2), (2, 64), (2, 65), (2, 66), (1, 0), (1, 64), (1, 65), (2, 128), (2, 129),
             (2, 130), (2, 131), (2, 192), (2, 194), (2, 195), (2, 258), (2, 259)]
             + [(2, 2 + 512 * p) for p in range(1, 8)] + [(2, 131 + 512 * p) for p in range(1, 8)]),
    "abwd": ("IDOC_NBUF_B", "IDOC_BM_ABWD", [(2, 0), (2, 1), (2, 2), (2, 3), (1, 0), (1, 2), (1, 3), (2, 128), (2, 129), (2, 130), (2, 131)]
             + [(2, 130 + 512 * p) for p in range(1, 8)]),
    "afwd": ("IDOC_NBUF_F", "IDOC_BM_AFWD", [(2, 0), (2, 2), (2, 64), (2, 66), (2, 96), (2, 98), (2, 82), (1, 98), (2, 224), (2, 225), (2, 226),
             (2, 228), (1, 224), (1, 225), (2, 352), (2, 353), (2, 356)]
             + [(2, 228 + 512 * p) for p in range(1, 8)] + [(2, 353 + 512 * p) for p in range(1, 8)]),
}
KNAME = {"ric": "lqr_riccati", "roll": "lqr_rollout", "abwd": "lqr_adj_bwd", "afwd": "lqr_adj_fwd"}


def digest(plan, sample):
    """(crc of the sampled outputs, the sampled outputs themselves)"""
    h, vals = 0, []
    for a in [plan.X, plan.U, plan.Nu] + [plan.grads[k] for k in sorted(plan.grads)]:
        v = np.ascontiguousarray(a.numpy().reshape(a.shape[0], -1)[sample])
        h = zlib.crc32(v.tobytes(), h)
        vals.append(v)
    return h, vals


def max_rel(vals, ref):
    return max(float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)) for a, b in zip(vals, ref))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--kernels", default="ric,roll,abwd,afwd")
    ap.add_argument("--out", default=None)
    ap.add_argument("--ab", default=None, help="A/B mode: comma list of kernel:nbuf:bm, timed alternately for --rounds rounds")
    ap.add_argument("--rounds", type=int, default=5)
    args = ap.parse_args()
    dtype = np.float64 if args.dtype == "f64" else np.float32
    B, T, n, m = args.batch, 100, 8, 4
    L.require_device()
    stream = L.Stream()
    prob = ib.lqr.DeviceProblem.generate(B, T, n, m, dtype, seed=1234, stream=stream)
    plan = ib.lqr.DevicePlan(B, T, n, m, dtype, vjp=True)
    sample = np.r_[0:64, B // 2:B // 2 + 64, B - 64:B]

    def run():
        plan.fwd(prob, stream=stream)
        plan.vjp(prob, plan.X, plan.U, None, stream=stream)

    def timed():
        run()
        stream.sync()
        L.profile_begin()
        for _ in range(args.reps):
            run()
        stream.sync()
        return {k: float(np.mean(v)) for k, v in L.profile_end().items()}

    for k in list(os.environ):
        if k.startswith("IDOC_NBUF") or k.startswith("IDOC_BM"):
            del os.environ[k]
    base_t = timed()
    base_h, base_v = digest(plan, sample)
    tot = sum(v for k, v in base_t.items() if k in KNAME.values())
    print("default:", {k: round(v, 3) for k, v in base_t.items()}, "sum %.3f ms" % tot, flush=True)
    results = {"dtype": args.dtype, "B": B, "default": base_t, "variants": {}}
    if args.ab:
        items = [x.split(":") for x in args.ab.split(",")]
        acc = {tuple(i): [] for i in items}
        for _ in range(args.rounds):
            for key, nb, bm in items:
                e_nb, e_bm, _ = VARIANTS[key]
                os.environ[e_nb], os.environ[e_bm] = nb, bm
                acc[(key, nb, bm)].append(timed()[KNAME[key]])
                del os.environ[e_nb], os.environ[e_bm]
        for k, v in acc.items():
            print("%-16s median %.3f  min %.3f  max %.3f ms  %s" % (":".join(k), np.median(v), min(v), max(v), " ".join("%.3f" % x for x in v)), flush=True)
        return
    for key in args.kernels.split(","):
        e_nb, e_bm, lst = VARIANTS[key]
        for nb, bm in lst:
            os.environ[e_nb], os.environ[e_bm] = str(nb), str(bm)
            try:
                t = timed()
                h, v = digest(plan, sample)
                ok = h == base_h
                rel = 0.0 if ok else max_rel(v, base_v)
            except Exception as ex:  # noqa: BLE001
                print(f"{key} NBUF={nb} BM={bm}: FAILED {ex}", flush=True)
                raise
            finally:
                del os.environ[e_nb], os.environ[e_bm]
            results["variants"][f"{key}:{nb}:{bm}"] = {"ms": t[KNAME[key]], "bit_identical": ok, "max_rel_diff": rel}
            print(f"{key:5s} NBUF={nb} BM={bm:2d}: {t[KNAME[key]]:.3f} ms  (default {base_t[KNAME[key]]:.3f})  identical={ok} max_rel_diff={rel:.2e}", flush=True)
    if args.out:
        json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
