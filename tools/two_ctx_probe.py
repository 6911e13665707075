#!/usr/bin/env python
"""Probe: do two rivecuda contexts on ONE GPU, fed alternately with independent frames, reach a
higher aggregate frame rate than one context (overlap of one frame's latency-bound front half
with the other's issue-bound raster)? usage: two_ctx_probe.py [trace] [contexts]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rive_runtime_b200 import abi, replay as R, trace as T  # noqa: E402

abi.load()
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "c2_4k.rvct.xz")
n_ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 2
records = T.parse(path)
setup, trace_frames = R.split_frames(records)


def make():
    rp = R.Replayer(device=0)
    result = R.ReplayResult()
    for r in setup:
        rp.apply(r, result)
    frames = []
    for ups, fls in trace_frames:
        frames.append(([(u.fields["kind"], np.ascontiguousarray(u.data)) for u in ups], [rp.prepare_flush(f.fields["flush"]) for f in fls]))
    return rp, frames


ctxs = [make() for _ in range(n_ctx)]
resident = len(trace_frames) == 1
if resident:
    for rp, frames in ctxs:
        for kind, data in frames[0][0]:
            rp.upload_buffer(kind, data)


def run(active, steps):
    for i in range(steps):
        for rp, frames in active:
            for ups, fls in frames:
                if not resident:
                    for kind, data in ups:
                        rp.upload_buffer(kind, data)
                for pf in fls:
                    rp.flush(pf)
    for rp, _ in active:
        rp.sync()


for k in range(1, n_ctx + 1):
    active = ctxs[:k]
    run(active, 3)
    steps = 30
    t0 = time.perf_counter()
    run(active, steps)
    dt = time.perf_counter() - t0
    print(f"{k} context(s): {k * steps * len(trace_frames) / dt:.1f} frames/s aggregate ({dt / (k * steps * len(trace_frames)) * 1e3:.3f} ms per frame)")
