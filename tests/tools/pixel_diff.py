#!/usr/bin/env python
"""Dev tool: pixel deviation (CUDA vs oracle) of a golden trace. usage: pixel_diff.py <name>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refcpu  # noqa: E402
from rive_runtime_b200 import abi, replay, trace as T  # noqa: E402

abi.load()
for name in sys.argv[1:]:
    path = name if os.path.exists(name) else os.path.join(ROOT, "tests", "golden", name + ".rvct.xz")
    recs = T.parse(path)
    ref = refcpu.replay(recs, threads=os.cpu_count() or 1, keep_intermediates=False)
    got = replay.replay(recs)
    for a, b in zip(ref.frames, got.frames):
        d = np.abs(a.astype(int) - b.astype(int)).max(axis=-1)
        ys, xs = np.nonzero(d > 2)
        mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
        print(name, "max", int(d.max()), "px>2:", int((d > 2).sum()), "psnr", 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse))
        for y, x in list(zip(ys, xs))[:12]:
            print("   ", x, y, a[y, x], b[y, x])
