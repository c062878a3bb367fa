#!/usr/bin/env python
"""Does splitting the batch over two streams (so that one half's kernels fill the other half's tails) help?
Times fwd + VJP of B problems as one launch chain vs NS chains of B / NS problems on NS streams."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import idoc_b200 as ib  # noqa: E402

L = ib._lib
dtype = np.float64 if (len(sys.argv) < 2 or sys.argv[1] == "f64") else np.float32
B, T, n, m = 65536, 100, 8, 4
for NS in (1, 2, 4, 1, 2, 3):
    Bs = B // NS
    streams = [L.Stream() for _ in range(NS)]
    probs = [ib.lqr.DeviceProblem.generate(Bs, T, n, m, dtype, seed=1234 + i, stream=streams[i]) for i in range(NS)]
    plans = [ib.lqr.DevicePlan(Bs, T, n, m, dtype, vjp=True) for _ in range(NS)]
    main = L.Stream()
    ev_start, ev_done = L.Event(), [L.Event() for _ in range(NS)]

    def step():
        for i in range(NS):
            plans[i].fwd(probs[i], stream=streams[i])
            plans[i].vjp(probs[i], plans[i].X, plans[i].U, None, stream=streams[i])

    for _ in range(3):
        step()
    for s in streams:
        s.sync()
    L.sync()
    e0, e1 = L.Event(), L.Event()
    steps = 10
    e0.record(main)
    for s in streams:
        s.wait(e0)
    for _ in range(steps):
        step()
    for i, s in enumerate(streams):
        ev_done[i].record(s)
        main.wait(ev_done[i])
    e1.record(main)
    main.sync()
    ms = e0.elapsed_ms(e1) / steps
    print(f"{NS} stream(s) x {Bs} problems: {ms:.3f} ms per {Bs * NS} solves  ({Bs * NS / ms / 1e3:.3f} M solves/s)", flush=True)
    del plans, probs
