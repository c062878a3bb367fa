#!/usr/bin/env python
"""Benchmark of the batched LQR hot path (fwd solve + implicit VJP) — BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f64|f32] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic problems: idoc_lqr_fwd (Riccati sweep, rollout) followed by
idoc_lqr_vjp (adjoint sweep, adjoint rollout + gradient writer) with per-problem gradients, cotangent (X, U, 0) (the reference
tests' loss).  At N=1 the workload is BASELINE config 2: B=65536, n=8, m=4, T=100.  With N>1 (torchrun, one rank per GPU) every
rank solves its own B problems (weak scaling; the per-problem-gradient path has no data-path collective).

Prints ONE JSON line (rank 0):
  value / ms_per_step   device-resident throughput of config 2 (CUDA events, inputs in HBM, max over ranks);
  roofline              algorithmic bytes / measured time against the measured HBM copy peak, per-kernel moved bytes;
  e2e                   the same metric through the host-buffer API (pinned host -> device -> pinned host, copies timed), plus
                        variants (batch-summed gradients: the training use case; fp32);
  configs  (N = 1)      the other configurations the metric names: config 2 in fp32, config 3 (tilqr, B = 262144, roofline
                        against the FMA peak measured in this run), config 4 (pendulum and cart-pole iLQR, iteration histogram),
                        config 5 (n = 32 slice with batch-summed gradients);
  exchange (N > 1)      the path's one exchange step: fwd + VJP with batch-summed gradients + ONE NCCL all-reduce per step at
                        config 2 and at config 5 (B = 32768 sharded over the N GPUs), and the all-reduced gradient verified on
                        the GPUs against the per-problem gradients of every rank;
  cpu_baseline (N = 1)  the C restatement of the reference algorithm (oracle/lqr_c.c) on the host cores.
`--impl reference` times that CPU restatement alone (the reference itself is JAX and cannot run in this image).
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
    bounded sample of the workload.  Returns (solves_per_s, cores, sample description, ms_per_step, B)."""
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
    return B / t, cores, (f"B={B} problems/step x {steps} step(s) (a bounded sample of the B=65536 workload: throughput is "
                          f"batch-independent once every core is busy), same n,m,T, same input distribution (NumPy PCG64 seed 0)"), t * 1e3, B


# ------------------------------------------------------------------------------------------- helpers
def _quiet_stdout():
    """Route fd 1 to stderr for the run (NCCL and friends print banners on stdout) and return a writer for the one
    JSON line."""
    saved = os.dup(1)
    sys.stdout.flush()
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


class Ctx:
    """rank / world plumbing: barrier and max-over-ranks through torch.distributed (gloo), nothing on the data path."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("gloo")
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max(self, x):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def timed_region(L, ctx, stream, step, steps, warmup, sample_clocks=False):
    """W untimed warm-up steps, then exactly K steps bracketed by barrier + synchronize on both sides, CUDA events on the
    launching stream, max over ranks.  Returns (ms_per_step, per-kernel event times, launches, clocks)."""
    for _ in range(warmup):
        step()
    stream.sync()
    ctx.barrier()
    sampler = None
    if sample_clocks:
        sampler = ClockSampler(ctx.local_rank)
        sampler.start()
        time.sleep(0.3)
    ev0, ev1 = L.Event(), L.Event()
    l0 = L.lib.idoc_launch_count()
    L.profile_begin()
    L.sync()
    ctx.barrier()
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    stream.sync()
    L.sync()
    ctx.barrier()
    ms_total = ev0.elapsed_ms(ev1)
    kernels = L.profile_end()
    launches = int(L.lib.idoc_launch_count() - l0)
    clocks = sampler.stop() if sampler is not None else None
    return ctx.max(ms_total / steps), kernels, launches, clocks


def kernel_stats(L, kernels, n, m, T, B, dtype, peak, reduce_batch=False, nub=False):
    """per kernel: mean duration and the bytes it really moves against the HBM peak"""
    sz = np.dtype(dtype).itemsize
    alg = algorithmic_scalars(n, m, T, reduce_batch)
    ws_slot = (L.lib.idoc_lqr_ws_bytes(L.dtype_code(dtype), 1024, 1, n, m) - 4096) // (1024 * sz)
    ws2_slot = L.lib.idoc_lqr_vjp_ws_bytes(L.dtype_code(dtype), 1, 1, n, m) // sz
    ps, so = alg["per_step"], alg["sol_step"]
    p16 = lambda w: -(-w * sz // 16) * 16 // sz
    mp, npk = m * (m + 1) // 2, n * (n + 1) // 2
    ws_pre = p16(mp + 1) + p16(m * n)                     # Linv, shift, K          (adjoint backward sweep)
    ws_roll = ws_slot - p16(mp + 1)                        # K, k, v, Vp             (rollout)
    ws_kv = p16(m * n) + p16(npk)                          # K, Vp                   (adjoint forward sweep)
    nb_ = n if nub else 0
    moved = {  # scalars per problem-step actually read+written by each kernel
        "lqr_riccati": ps + ws_slot,
        "lqr_rollout": (n * n + n * m + n) + ws_roll + so,
        "lqr_adj_bwd": (n * n + n * m) + ws_pre + (p16(npk) if nub else 0) + (n + m) + nb_ + ws2_slot,
        # batch-summed mode: the sweep stores the adjoint solution (sol scalars) instead of the per-problem leaves
        "lqr_adj_fwd": (n * n + n * m) + ws_kv + ws2_slot + so + nb_ + (so if reduce_batch else ps),
        # batch contraction: reads solution + adjoint solution once per problem-step
        "grad_reduce": 2 * so,
    }
    out = {}
    for name, arr in kernels.items():
        per_step_ms = float(np.sum(arr)) / max(1, len(kernels.get("lqr_riccati", arr)))
        e = {"launches": len(arr), "mean_ms": float(np.mean(arr)), "ms_per_step": per_step_ms}
        if name in moved:
            e["moved_GBs"] = moved[name] * T * sz * B / (per_step_ms * 1e-3) / 1e9
            e["moved_frac_of_peak"] = e["moved_GBs"] / peak
        out[name] = e
    return out, moved


# ------------------------------------------------------------------------------------------- legs
def lqr_leg(ib, ctx, n, m, T, B, dtype, steps, warmup, reduce_batch=False, reducer=None, nub=False, seed=1234,
            sample_clocks=False, keep=False):
    """fwd + implicit VJP of B device-resident problems per rank (optionally batch-summed gradients + NCCL all-reduce)."""
    L = ib._lib
    sz = np.dtype(dtype).itemsize
    stream = L.Stream()
    prob = ib.lqr.DeviceProblem.generate(B, T, n, m, dtype, seed=seed + ctx.rank, stream=stream)
    plan = ib.lqr.DevicePlan(B, T, n, m, dtype, vjp=True, reduce_batch=reduce_batch)
    Nub = None
    if nub:
        Nub = L.DeviceArray((B, T, n), dtype)
        L.check(L.lib.idoc_memcpy_d2d(Nub.ptr, prob.arr["q"].ptr, Nub.nbytes, stream.ptr), "d2d")  # N(0,1) values

    def step():
        plan.fwd(prob, stream=stream)
        plan.vjp(prob, plan.X, plan.U, Nub, stream=stream)
        if reducer is not None:
            reducer.allreduce_plan_grads(plan, stream=stream)

    ms_step, kernels, launches, clocks = timed_region(L, ctx, stream, step, steps, warmup, sample_clocks)
    peak, peak_src = measured_peaks()
    alg = algorithmic_scalars(n, m, T, reduce_batch)
    achieved = alg["total"] * sz * B / (ms_step * 1e-3) / 1e9
    kstats, _ = kernel_stats(L, kernels, n, m, T, B, dtype, peak, reduce_batch, nub)
    res = dict(ms_per_step=ms_step, value=ctx.world * B / (ms_step * 1e-3), launches=launches, clocks=clocks,
               status_flags_set=int((plan.status.numpy() != 0).sum()),
               roofline={"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "what": "whole hot path: algorithmic bytes of fwd+VJP (SURVEY 8d: %d scalars/solve%s) / time of all "
                                 "kernels of one step" % (alg["total"], ", batch-summed-gradient variant" if reduce_batch else ""),
                         "frac_of_nominal_8TBs": achieved / 8000.0, "kernels": kstats})
    if keep:
        res["_prob"], res["_plan"], res["_stream"] = prob, plan, stream
    return res


def e2e_leg(ib, ctx, prob, plan_ref, n, m, T, dtype, Be, chunk, steps, reduce_batch=False, ramp=False):
    """host buffers -> device -> host buffers, every step (the 3-stream chunk pipeline of idoc_b200.host)."""
    from idoc_b200 import host
    L = ib._lib
    hp, hs, hg = host.pinned_like_problem(Be, T, n, m, dtype, reduce_batch=reduce_batch)
    for k in host.GRAD_KEYS:  # fill the host problem from the device-generated one (setup, untimed)
        row = hp[k].nbytes // Be
        L.check(L.lib.idoc_memcpy_d2h(hp[k].ptr, prob.arr[k].ptr, Be * row, None), "d2h setup")
    # one C-ABI call per batch (idoc_lqr_host_solve_vjp, csrc/host_pipeline.cu); --e2e-ramp selects the Python-side pipeline
    pipe = (host.HostPipeline(Be, T, n, m, dtype, chunk=chunk, nslots=3, ramp=True, reduce_batch=reduce_batch) if ramp else
            host.NativeHostPipeline(Be, T, n, m, dtype, chunk=chunk, nslots=3, reduce_batch=reduce_batch))
    pipe.solve_and_vjp(hp, hs, hg)
    L.sync()
    ctx.barrier()
    # the timed steps are enqueued back to back (a stream of batches): every step copies its inputs in and its results
    # out inside the timed region; the pipeline's fill and drain overlap with the neighbouring steps
    ee0, ee1 = L.Event(), L.Event()
    ee0.record(pipe.s_in)
    for _ in range(steps):
        pipe.solve_and_vjp(hp, hs, hg, wait=False)
    ee1.record(pipe.s_out)
    ee1.sync()
    pipe.finish()
    L.sync()
    e2e_ms = ctx.max(ee0.elapsed_ms(ee1) / steps)
    e2e_single_ms = ctx.max(pipe.solve_and_vjp(hp, hs, hg))  # one isolated batch (fill + drain exposed)
    L.sync()
    ok = None
    if plan_ref is not None:  # spot check: host results equal the device-resident ones for the same problems
        ok = bool(np.array_equal(hs["X"].a[:64], plan_ref.X.numpy()[:64]))
    return {"value": ctx.world * Be / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
            "d2h_bytes_per_step": pipe.d2h_bytes, "ms_per_step": e2e_ms, "batch_per_gpu": Be, "chunk": chunk,
            "steps": steps, "ms_single_batch": e2e_single_ms, "matches_device_path": ok}


def extra_configs(ib, dtype_main):
    """N = 1: the other configurations BASELINE.json's metric names, each a short device-resident run (CUDA events)."""
    L = ib._lib
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_configs as BC
    out = {}

    def guard(name, fn):
        try:
            out[name] = fn()
        except Exception as e:  # a failed side leg must not lose the headline line
            out[name] = {"error": repr(e)}

    peaks = {}
    for dt in (np.float32, np.float64):
        try:
            peaks[np.dtype(dt).name] = L.fma_peak_tflops(dt)
        except Exception:
            peaks[np.dtype(dt).name] = None

    def tilqr(dt, B):
        r = BC.leg_tilqr(dt, B, 3, 1)
        r["kernel_ms_per_step"] = dict(BC.KERNELS)
        pk = peaks.get(np.dtype(dt).name)
        if pk:
            rf = r["roofline"]
            rf.update(peak=pk, frac=rf["achieved"] / pk, peak_source="measured in this run (idoc_fma_peak: register-resident FMA chains)",
                      frac_of_nominal=rf["achieved"] / (74.4 if np.dtype(dt) == np.float32 else 37.2))
        return r

    def with_kernels(fn, *a):
        r = fn(*a)
        r["kernel_ms_per_step"] = dict(BC.KERNELS)
        return r

    guard("config3_tilqr_f32", lambda: tilqr(np.float32, 262144))
    guard("config3_tilqr_f64", lambda: tilqr(np.float64, 131072))
    guard("config4_pendulum_f64", lambda: with_kernels(BC.leg_ilqr, np.float64, 16384, 3, 1))
    guard("config4_cartpole_f64", lambda: with_kernels(BC.leg_cartpole, np.float64, 16384, 2, 1))
    guard("config4_pendulum_box_f64", lambda: with_kernels(lambda *a: BC.leg_ilqr(*a, umax=8.0), np.float64, 16384, 2, 1))
    guard("config4_cartpole_box_f64", lambda: with_kernels(lambda *a: BC.leg_cartpole(*a, umax=6.0), np.float64, 16384, 2, 1))
    guard("config5_slice_f32", lambda: with_kernels(BC.leg_lqr32, np.float32, 4096, 3, 1))
    guard("config5_slice_f64", lambda: with_kernels(BC.leg_lqr32, np.float64, 2048, 3, 1))
    out["fma_peak_tflops_measured"] = peaks
    return out


def exchange_legs(ib, ctx, args, dtype):
    """N > 1: the one exchange step of the path under the clock.  fwd + VJP with batch-summed gradients + ONE ncclAllReduce
    of the contiguous gradient buffer per step, at config 2 (B per GPU as in the headline) and at BASELINE config 5
    (B = 32768 problems sharded over the N GPUs, n = 32, m = 16, T = 200, fp32), and a verification on the GPUs."""
    from idoc_b200 import dist as D
    if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
        os.environ["NCCL_DEBUG"] = "INFO"  # the communicator's init line (rank, nranks) goes to the log
    reducer = D.GradAllReducer(ctx.rank, ctx.world, D.exchange_unique_id(ctx.rank, ctx.world))
    out = {"collective": "NCCL all-reduce (sum) over NVLink, ONE call per step on the contiguous buffer of the ten batch-summed "
                         "gradient leaves", "nranks": ctx.world}
    ver = {}
    for (n, m, T, Bv, dt, tol) in ((8, 4, 100, 512, np.float64, 1e-10), (8, 4, 100, 512, np.float32, 1e-4),
                                   (32, 16, 200, 64, np.float32, 1e-4), (32, 16, 200, 64, np.float64, 1e-10)):
        e = D.verify_allreduced_gradients(reducer, n, m, T, Bv, dt)
        ver[f"n{n}m{m}T{T}_{np.dtype(dt).name}"] = {"max_rel_err": e, "tol": tol, "ok": bool(e < tol), "batch_per_rank": Bv}
    out["verified"] = {"what": "all-reduced batch-summed gradients vs the sum over ranks and problems of the per-problem "
                               "gradients (host float64 sum, gloo)", "cases": ver, "ok": all(v["ok"] for v in ver.values())}
    steps, warmup = max(3, min(args.steps, 10)), max(3, min(args.warmup, 5))
    r = lqr_leg(ib, ctx, 8, 4, 100, args.batch, dtype, steps, warmup, reduce_batch=True, reducer=reducer)
    grad_scalars = algorithmic_scalars(8, 4, 100)["per_step"] * 100 + 8 * 8 + 8
    out["config2"] = {"workload": f"BASELINE config 2 shape, B={args.batch} per GPU, {args.dtype}, fwd + VJP with batch-summed "
                                  f"gradients + NCCL all-reduce of {grad_scalars} scalars", "value": r["value"], "unit": UNIT,
                      "ms_per_step": r["ms_per_step"], "steps": steps, "warmup": warmup, "gpu_launches": r["launches"],
                      "roofline": r["roofline"]}
    Btot = 32768
    Bg = Btot // ctx.world
    free, _tot = ib._lib.mem_info()
    need = Bg * (algorithmic_scalars(32, 16, 200)["per_step"] * 200 * 4 * 1.6)
    if need < 0.85 * free:
        r5 = lqr_leg(ib, ctx, 32, 16, 200, Bg, np.float32, 3, 3, reduce_batch=True, reducer=reducer, seed=4321)
        out["config5"] = {"workload": f"BASELINE config 5: B={Btot} n=32 m=16 T=200 f32 sharded over {ctx.world} GPU(s) ({Bg} per "
                                      f"GPU), fwd + VJP with batch-summed gradients + NCCL all-reduce of "
                                      f"{algorithmic_scalars(32, 16, 200)['per_step'] * 200 + 32 * 32 + 32} scalars",
                          "value": r5["value"], "unit": UNIT, "ms_per_step": r5["ms_per_step"], "steps": 3, "warmup": 3,
                          "gpu_launches": r5["launches"], "roofline": r5["roofline"]}
    else:
        out["config5"] = {"skipped": f"{Bg} problems per GPU need {need / 1e9:.0f} GB, {free / 1e9:.0f} GB free"}
    reducer.close()
    return out


# ------------------------------------------------------------------------------------------- main
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
    ap.add_argument("--no-configs", action="store_true", help="skip the extra configurations (N = 1)")
    ap.add_argument("--no-exchange", action="store_true", help="skip the NCCL exchange legs (N > 1)")
    ap.add_argument("--nub", action="store_true", help="dense Nu cotangent (default: none, as in the reference tests)")
    ap.add_argument("--reduce-batch", action="store_true",
                    help="main leg with batch-summed gradients (reduce_batch=1) + NCCL all-reduce across ranks")
    args = ap.parse_args()

    if args.shape:
        args.n, args.m, args.horizon = [int(x) for x in args.shape.split(",")]
    n, m, T, B = args.n, args.m, args.horizon, args.batch
    dtype = np.float64 if args.dtype == "f64" else np.float32
    sz = np.dtype(dtype).itemsize
    steps, warmup = args.steps, max(args.warmup, 0)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    cfg_name = {(8, 4, 100): "BASELINE config 2", (32, 16, 200): "BASELINE config 5"}.get((n, m, T), "custom shape")
    config = {"workload": f"{cfg_name}: batched time-varying LQR B={B} n={n} m={m} T={T} {args.dtype}, fwd + implicit VJP, "
                          + ("batch-summed gradients + NCCL all-reduce" if args.reduce_batch else "per-problem gradients")
                          + ", cotangent (X,U,0)",
              "batch_per_gpu": B, "n": n, "m": m, "T": T,
              "l2": "inputs larger than L2 (%.1f GB streamed per step)" % (algorithmic_scalars(n, m, T)["total"] * sz * B / 1e9),
              "parallelism": f"batch sharded over {world} GPU(s), no data-path collective"}

    if args.impl == "reference":
        if rank != 0:
            return
        v, cores, sample, ms, Bs = cpu_leg(n, m, T, dtype, target_s=20.0, steps=steps, warmup=warmup)
        config["workload"] += f"; this arm: the reference algorithm on the host cores, B={Bs} per step (bounded sample)"
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

    ctx = Ctx()
    import idoc_b200 as ib
    L = ib._lib
    L.require_device()
    L.check(L.lib.idoc_set_device(ctx.local_rank), "set_device")

    reducer = None
    if args.reduce_batch and world > 1:
        from idoc_b200 import dist as D
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "INFO"
        reducer = D.GradAllReducer(ctx.rank, world, D.exchange_unique_id(ctx.rank, world))
        config["parallelism"] = f"batch sharded over {world} GPU(s); batch-summed gradients all-reduced with NCCL (one call per step)"

    main_leg = lqr_leg(ib, ctx, n, m, T, B, dtype, steps, warmup, reduce_batch=args.reduce_batch, reducer=reducer,
                       nub=args.nub, sample_clocks=True, keep=True)
    prob, plan = main_leg.pop("_prob"), main_leg.pop("_plan")
    main_leg.pop("_stream")
    roofline = main_leg["roofline"]
    prof_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof_path) and (n, m, T) == (8, 4, 100) and not args.reduce_batch:
        try:  # ncu capture (dram__bytes_read.sum + dram__bytes_write.sum of the four sweeps) at a profiled batch, linear in B
            tr = json.load(open(prof_path)).get(f"{args.dtype}", {})
            roofline["traffic"] = tr["bytes_per_step_profiled"] / tr["profiled_batch"] * B
            roofline["traffic_source"] = tr.get("source", "profiles/traffic.json (ncu --set full capture, scaled linearly in the batch)")
        except Exception:
            pass

    metric = METRIC if (n, m, T) == (8, 4, 100) else METRIC.split(",")[0] + f", n={n} m={m} T={T}"
    out = {"metric": metric, "value": main_leg["value"], "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": main_leg["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": args.dtype, "data": "synthetic (device-generated, SURVEY 8d distribution)", "config": config,
           "clocks": main_leg["clocks"], "gpu_launches": main_leg["launches"], "roofline": roofline,
           "status_flags_set": main_leg["status_flags_set"]}

    # ---- e2e: host buffers -> device -> host buffers, every step
    if not args.no_e2e and not args.reduce_batch:
        Be = min(args.e2e_batch, B)
        chunk = min(args.e2e_chunk, Be)
        e2e = e2e_leg(ib, ctx, prob, plan, n, m, T, dtype, Be, chunk, args.e2e_steps, ramp=bool(args.e2e_ramp))
        e2e["what"] = ("pinned host problem -> H2D -> fwd+VJP -> D2H of X,U,Nu and all 11 per-problem gradient leaves, 3-stream chunk "
                       f"pipeline behind one C-ABI call per batch (idoc_lqr_host_solve_vjp) over B={Be} problems per step (PCIe-bound: the rate does not depend "
                       "on the batch); the timed steps are enqueued back to back, ms_single_batch is one isolated batch")
        config["workload"] += f"; e2e leg: B={Be} per GPU per step"
        variants = {}
        try:
            v = e2e_leg(ib, ctx, prob, None, n, m, T, dtype, Be, chunk, args.e2e_steps, reduce_batch=True)
            v["what"] = "same pipeline with batch-summed gradients (the training use case): D2H carries X,U,Nu, gx0 and one summed buffer per chunk"
            variants["batch_summed_gradients"] = v
        except Exception as e:
            variants["batch_summed_gradients"] = {"error": repr(e)}
        out["e2e"] = e2e
        out["e2e"]["variants"] = variants
    del prob, plan

    if world == 1 and not args.no_configs and not args.reduce_batch and (n, m, T) == (8, 4, 100):
        cfgs = {}
        other = np.float32 if dtype == np.float64 else np.float64
        try:
            r = lqr_leg(ib, ctx, n, m, T, B, other, max(3, min(steps, 10)), 3)
            cfgs["config2_" + np.dtype(other).name] = {
                "workload": f"BASELINE config 2 in {np.dtype(other).name}: B={B} n={n} m={m} T={T}, fwd + implicit VJP, per-problem gradients",
                "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "roofline": r["roofline"], "gpu_launches": r["launches"]}
        except Exception as e:
            cfgs["config2_other_dtype"] = {"error": repr(e)}
        try:
            r = lqr_leg(ib, ctx, n, m, T, B, dtype, max(3, min(steps, 10)), 3, reduce_batch=True)
            cfgs["config2_batch_summed_" + args.dtype] = {
                "workload": f"config 2 with batch-summed gradients (reduce_batch = 1, no atomics): B={B}, {args.dtype}",
                "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "roofline": r["roofline"], "gpu_launches": r["launches"]}
        except Exception as e:
            cfgs["config2_batch_summed"] = {"error": repr(e)}
        cfgs.update(extra_configs(ib, dtype))
        out["configs"] = cfgs
    if world > 1 and not args.no_exchange and not args.reduce_batch:
        try:
            out["exchange"] = exchange_legs(ib, ctx, args, dtype)
            config["parallelism"] += "; the exchange step (batch-summed gradients + NCCL all-reduce) is timed and verified under `exchange`"
        except Exception as e:
            out["exchange"] = {"error": repr(e)}
    if reducer is not None:
        reducer.close()
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, sample, _, _ = cpu_leg(n, m, T, dtype, target_s=12.0)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(out), file=json_out, flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
