#!/usr/bin/env python
"""Pinned host <-> device copy bandwidth of the box (what bounds bench.py's `e2e` leg): H2D alone, D2H alone, both
directions at once on two streams, for a few transfer sizes.  One JSON line per size.

    python tools/pcie_probe.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import idoc_b200 as ib  # noqa: E402

L = ib._lib


def timed(fn, streams, reps=5):
    fn()
    for s in streams:
        s.sync()
    e0, e1 = L.Event(), L.Event()
    best = 1e30
    for _ in range(reps):
        L.sync()
        e0.record(streams[0])
        for s in streams[1:]:
            s.wait(e0)
        fn()
        ends = []
        for s in streams[1:]:
            e = L.Event()
            e.record(s)
            ends.append(e)
        for e in ends:
            streams[0].wait(e)
        e1.record(streams[0])
        e1.sync()
        best = min(best, e0.elapsed_ms(e1))
    return best


if __name__ == "__main__":
    for mb in (16, 64, 256, 1024):
        nbytes = mb << 20
        h_in, h_out = L.PinnedArray((nbytes,), np.uint8), L.PinnedArray((nbytes,), np.uint8)
        d_in, d_out = L.DeviceArray((nbytes,), np.uint8), L.DeviceArray((nbytes,), np.uint8, zero=True)
        h_in.a[:] = 1
        s0, s1 = L.Stream(), L.Stream()

        def h2d():
            L.check(L.lib.idoc_memcpy_h2d(d_in.ptr, h_in.ptr, nbytes, s0.ptr), "h2d")

        def d2h():
            L.check(L.lib.idoc_memcpy_d2h(h_out.ptr, d_out.ptr, nbytes, s0.ptr), "d2h")

        def both():
            L.check(L.lib.idoc_memcpy_h2d(d_in.ptr, h_in.ptr, nbytes, s0.ptr), "h2d")
            L.check(L.lib.idoc_memcpy_d2h(h_out.ptr, d_out.ptr, nbytes, s1.ptr), "d2h")

        t_in, t_out, t_both = timed(h2d, [s0]), timed(d2h, [s0]), timed(both, [s0, s1])
        gb = nbytes / 1e9
        print(json.dumps(dict(mb=mb, h2d_gbs=round(gb / t_in * 1e3, 1), d2h_gbs=round(gb / t_out * 1e3, 1),
                              duplex_each_gbs=round(gb / t_both * 1e3, 1), duplex_total_gbs=round(2 * gb / t_both * 1e3, 1))),
              flush=True)
