#!/usr/bin/env python
"""Benchmark of the batched LQR hot path (fwd solve + implicit VJP) — BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f64|f32] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic problems: idoc_lqr_fwd
(Riccati sweep, rollout) followed by idoc_lqr_vjp (adjoint sweep, adjoint rollout + gradient writer)
with per-problem gradients, cotangent (X, U, 0) (the reference tests' loss).  At N=1 the workload is
BASELINE config 2: B=65536, n=8, m=4, T=100.  With N>1 (torchrun, one rank per GPU) every rank
solves its own B problems (weak scaling; the path has no data-path collective).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (CUDA events, inputs in HBM);
`e2e` goes through the host-buffer API (pinned host -> device -> pinned host, copies timed);
`roofline` is algorithmic bytes / measured time against the measured HBM copy peak;
`cpu_baseline` is the C restatement of the reference algorithm (oracle/lqr_c.c) on the host cores.
`--impl reference` times that CPU restatement alone (the reference itself is JAX and cannot run here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "batched LQR solves/sec (fwd+implicit VJP), n=8 m=4 T=100"
UNIT = "solves/s"
HBM_FALLBACK_GBS = 6650.0


def algorithmic_scalars(n, m, T, reduce_batch=False):
    """every input read once per pass, every output written once (batch-reduced gradients: no
    per-problem gradient write: (2 per_step + 3 sol) T + 2 once)."""
    per_step = 2 * n * n + 2 * n * m + m * m + 2 * n + m
    once = n * n + 2 * n
    sol = 2 * n + m
    fwd = per_step * T + once + sol * T
    vjp = (per_step * T + once) + sol * T + sol * T + (0 if reduce_batch else per_step * T + once)
    return dict(per_step=per_step, once=once, sol_step=sol, fwd=fwd, vjp=vjp, total=fwd + vjp)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev, self.rows, self.proc = dev, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU leg
def cpu_leg(n, m, T, dtype, target_s=12.0, steps=1, warmup=0):
    """Time the C restatement of the reference algorithm (fwd + exact VJP) on all host cores, on a
    bounded sample of the workload.  Returns (solves_per_s, cores, sample description, ms_per_step)."""
    from oracle import lqr_np as O
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    from oracle import lqr_c

    # every core this process may run on, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)
    lqr_c.set_threads(len(os.sched_getaffinity(0)))
    cores = lqr_c.threads()
    rng = np.random.default_rng(0)

    def run(B, reps):
        p = O.random_params(rng, n, m, T, batch=(B,), dtype=dtype)
        fo = None
        go = None
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            s, K, k = lqr_c.fwd(p, out=fo)
            fo = (s.X, s.U, s.Nu, K, k)
            go = lqr_c.vjp(p, s, O.State(s.X, s.U, None), out=go)
            ts.append(time.perf_counter() - t0)
        return ts

    t_probe = min(run(64 * cores, 2))
    rate = 64 * cores / t_probe
    B = int(max(cores * 64, min(65536, rate * target_s / max(1, steps + warmup))))
    B = (B // cores) * cores
    ts = run(B, warmup + steps)[warmup:]
    t = float(np.mean(ts))
    return B / t, cores, f"B={B} problems/step x {steps} step(s), same n,m,T, same input distribution (NumPy PCG64 seed 0)", t * 1e3


# ------------------------------------------------------------------------------------------- main
def _quiet_stdout():
    """Route fd 1 to stderr for the run (NCCL and friends print banners on stdout) and return a writer for the one
    JSON line."""
    saved = os.dup(1)
    sys.stdout.flush()
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    json_out = _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--m", type=int, default=4)
    ap.add_argument("--horizon", type=int, default=100)
    ap.add_argument("--shape", default=None, help="n,m,T in one option (torchrun's own parser trips over --n / --m)")
    ap.add_argument("--e2e-batch", type=int, default=16384)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-chunk", type=int, default=2048, help="problems per pipelined chunk of the host-buffer path")
    ap.add_argument("--e2e-ramp", type=int, default=0, help="split the first chunk of the host-buffer pipeline (1/8, 1/4, 5/8)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--nub", action="store_true", help="dense Nu cotangent (default: none, as in the reference tests)")
    ap.add_argument("--reduce-batch", action="store_true",
                    help="batch-summed gradients (reduce_batch=1) + NCCL all-reduce across ranks (the path's one exchange step)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.shape:
        args.n, args.m, args.horizon = [int(x) for x in args.shape.split(",")]
    n, m, T, B = args.n, args.m, args.horizon, args.batch
    dtype = np.float64 if args.dtype == "f64" else np.float32
    sz = np.dtype(dtype).itemsize
    steps, warmup = args.steps, max(args.warmup, 0)
    cfg_name = {(8, 4, 100): "BASELINE config 2", (32, 16, 200): "BASELINE config 5"}.get((n, m, T), "custom shape")
    config = {"workload": f"{cfg_name}: batched time-varying LQR B={B} n={n} m={m} T={T} {args.dtype}, fwd + implicit VJP, "
                          + ("batch-reduced gradients + NCCL all-reduce" if args.reduce_batch else "per-problem gradients")
                          + ", cotangent (X,U,0)",
              "batch_per_gpu": B, "n": n, "m": m, "T": T,
              "l2": "inputs larger than L2 (%.1f GB streamed per step)" % (algorithmic_scalars(n, m, T)["total"] * sz * B / 1e9),
              "parallelism": f"batch sharded over {world} GPU(s), no data-path collective"}

    if args.impl == "reference":
        if rank != 0:
            return
        v, cores, sample, ms = cpu_leg(n, m, T, dtype, target_s=20.0,
                                       steps=steps, warmup=warmup)
        out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
               "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": args.dtype, "data": "synthetic", "config": config,
               "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                "note": "C restatement of idoc/lqr.py + exact adjoint-LQR VJP (oracle/lqr_c.c); the "
                                        "reference itself is JAX and cannot be installed in this image; its CG-based "
                                        "VJP would be slower than this port"},
               "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        print(json.dumps(out), file=json_out, flush=True)
        return

    dist = None
    if world > 1:
        import torch.distributed as dist  # plumbing only: barrier + max-over-ranks of the timings
        dist.init_process_group("gloo")

    import idoc_b200 as ib
    L = ib._lib
    L.require_device()
    L.check(L.lib.idoc_set_device(local_rank), "set_device")

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    stream = L.Stream()
    prob = ib.lqr.DeviceProblem.generate(B, T, n, m, dtype, seed=1234 + rank, stream=stream)
    plan = ib.lqr.DevicePlan(B, T, n, m, dtype, vjp=True, reduce_batch=args.reduce_batch)
    reducer = None
    if args.reduce_batch and world > 1:
        from idoc_b200 import dist as D
        reducer = D.GradAllReducer(rank, world, D.exchange_unique_id(rank, world))
        config["parallelism"] = f"batch sharded over {world} GPU(s); batch-summed gradients all-reduced with NCCL"
    if args.reduce_batch:
        config["workload"] = config["workload"].replace("per-problem gradients", "batch-reduced gradients")
    Nub = None
    if args.nub:
        Nub = L.DeviceArray((B, T, n), dtype)
        L.check(L.lib.idoc_memcpy_d2d(Nub.ptr, prob.arr["q"].ptr, Nub.nbytes, stream.ptr), "d2d")  # N(0,1) values

    def step():
        plan.fwd(prob, stream=stream)
        plan.vjp(prob, plan.X, plan.U, Nub, stream=stream)
        if reducer is not None:
            reducer.allreduce_plan_grads(plan, stream=stream)

    for _ in range(warmup):
        step()
    stream.sync()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    ev0, ev1 = L.Event(), L.Event()
    l0 = L.lib.idoc_launch_count()
    L.profile_begin()
    L.sync()
    barrier()
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    stream.sync()
    L.sync()
    barrier()
    ms_total = ev0.elapsed_ms(ev1)
    kernels = L.profile_end()
    launches = int(L.lib.idoc_launch_count() - l0)
    clocks = sampler.stop()
    ms_step = max_over_ranks(ms_total / steps)
    value = world * B / (ms_step * 1e-3)

    # sanity: the timed batch solved the problems (KKT residual of the device solution, one launch, untimed)
    st = plan.status.numpy()
    alg = algorithmic_scalars(n, m, T, args.reduce_batch)
    peak, peak_src = measured_peaks()
    alg_bytes_step = alg["total"] * sz * B
    achieved = alg_bytes_step / (ms_step * 1e-3) / 1e9
    kstats = {}
    # per-kernel: mean duration, algorithmic bytes attributed to it, and the bytes it really moves
    ws_slot = (L.lib.idoc_lqr_ws_bytes(L.dtype_code(dtype), 1024, 1, n, m) - 4096) // (1024 * sz)
    ws2_slot = L.lib.idoc_lqr_vjp_ws_bytes(L.dtype_code(dtype), 1, 1, n, m) // sz
    ps, so = alg["per_step"], alg["sol_step"]
    p16 = lambda w: -(-w * sz // 16) * 16 // sz
    mp, npk = m * (m + 1) // 2, n * (n + 1) // 2
    ws_pre = p16(mp + 1) + p16(m * n)                     # Linv, shift, K          (adjoint backward sweep)
    ws_roll = ws_slot - p16(mp + 1)                        # K, k, v, Vp             (rollout)
    ws_kv = p16(m * n) + p16(npk)                          # K, Vp                   (adjoint forward sweep)
    nub = n if args.nub else 0
    moved = {  # scalars per problem-step actually read+written by each kernel
        "lqr_riccati": ps + ws_slot,
        "lqr_rollout": (n * n + n * m + n) + ws_roll + so,
        "lqr_adj_bwd": (n * n + n * m) + ws_pre + (p16(npk) if args.nub else 0) + (n + m) + nub + ws2_slot,
        "lqr_adj_fwd": (n * n + n * m) + ws_kv + ws2_slot + so + nub + (0 if args.reduce_batch else ps),
    }
    for name, arr in kernels.items():
        mean_ms = float(np.mean(arr))
        e = {"launches": len(arr), "mean_ms": mean_ms}
        if name in moved:
            e["moved_GBs"] = moved[name] * T * sz * B / (mean_ms * 1e-3) / 1e9
            e["moved_frac_of_peak"] = e["moved_GBs"] / peak
        kstats[name] = e
    dom = max((k for k in kstats if k in moved), key=lambda k: kstats[k]["mean_ms"])
    dom_alg = {"lqr_riccati": ps * T + alg["once"], "lqr_rollout": so * T, "lqr_adj_bwd": so * T,
               "lqr_adj_fwd": (1 if args.reduce_batch else 2) * (ps * T + alg["once"]) + so * T}[dom] * sz * B
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src,
                "what": "whole hot path: algorithmic bytes of fwd+VJP (SURVEY 8d: %d scalars/solve) / time of all kernels "
                        "of one step" % alg["total"],
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "dominant_kernel": {"name": dom, "mean_ms": kstats[dom]["mean_ms"],
                                    "algorithmic_GBs": dom_alg / (kstats[dom]["mean_ms"] * 1e-3) / 1e9,
                                    "frac": dom_alg / (kstats[dom]["mean_ms"] * 1e-3) / 1e9 / peak},
                "kernels": kstats}
    prof_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof_path) and (n, m, T) == (8, 4, 100) and not args.reduce_batch:
        try:  # ncu capture at B = 16384, linear in the batch
            tr = json.load(open(prof_path)).get(f"{args.dtype}", {})
            roofline["traffic"] = tr["bytes_per_step_profiled"] / tr["profiled_batch"] * B
        except Exception:
            pass

    metric = METRIC if (n, m, T) == (8, 4, 100) else METRIC.split(",")[0] + f", n={n} m={m} T={T}"
    out = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": args.dtype, "data": "synthetic (device-generated, SURVEY 8d distribution)", "config": config,
           "clocks": clocks, "gpu_launches": launches, "roofline": roofline,
           "status_flags_set": int((st != 0).sum())}

    # ---- e2e: host buffers -> device -> host buffers, every step
    if not args.no_e2e and not args.reduce_batch:
        from idoc_b200 import host
        Be = min(args.e2e_batch, B)
        chunk = min(args.e2e_chunk, Be)
        hp, hs, hg = host.pinned_like_problem(Be, T, n, m, dtype)
        for k in host.GRAD_KEYS:  # fill the host problem from the device-generated one (setup, untimed)
            row = hp[k].nbytes // Be
            L.check(L.lib.idoc_memcpy_d2h(hp[k].ptr, prob.arr[k].ptr, Be * row, None), "d2h setup")
        pipe = host.HostPipeline(Be, T, n, m, dtype, chunk=chunk, nslots=3, ramp=bool(args.e2e_ramp))
        pipe.solve_and_vjp(hp, hs, hg)
        L.sync()
        barrier()
        # the timed steps are enqueued back to back (a stream of batches): every step copies its inputs in and its results
        # out inside the timed region; the pipeline's fill and drain overlap with the neighbouring steps
        ee0, ee1 = L.Event(), L.Event()
        ee0.record(pipe.s_in)
        for _ in range(args.e2e_steps):
            pipe.solve_and_vjp(hp, hs, hg, wait=False)
        ee1.record(pipe.s_out)
        ee1.sync()
        L.sync()
        e2e_ms = max_over_ranks(ee0.elapsed_ms(ee1) / args.e2e_steps)
        e2e_single_ms = max_over_ranks(pipe.solve_and_vjp(hp, hs, hg))  # one isolated batch (fill + drain exposed)
        L.sync()
        # spot check: host results equal the device-resident ones for the same problems
        ok = bool(np.array_equal(hs["X"].a[:64], plan.X.numpy()[:64]))
        out["e2e"] = {"value": world * Be / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
                      "d2h_bytes_per_step": pipe.d2h_bytes, "ms_per_step": e2e_ms, "batch_per_gpu": Be, "chunk": chunk,
                      "steps": args.e2e_steps, "ms_single_batch": e2e_single_ms, "matches_device_path": ok,
                      "what": "pinned host problem -> H2D -> fwd+VJP -> D2H of X,U,Nu and all 11 gradient leaves, "
                              "3-stream chunk pipeline (idoc_b200.host.HostPipeline); the timed steps are enqueued back to "
                              "back, ms_single_batch is one isolated batch"}
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, sample, _ = cpu_leg(n, m, T, dtype, target_s=12.0)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(out), file=json_out, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
