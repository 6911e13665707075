#!/usr/bin/env python
"""Dev tool: for every golden trace, the worst tessellation deviation (position, theta) and
pixel deviation between the CUDA path and the oracle. usage: parity_sweep.py [prefix ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refcpu  # noqa: E402
from rive_runtime_b200 import abi, replay, trace as T  # noqa: E402

abi.load()
golden = os.path.join(ROOT, "tests", "golden")
names = sorted(n for n in os.listdir(golden) if n.endswith(".rvct.xz"))
if len(sys.argv) > 1:
    names = [n for n in names if n.startswith(tuple(sys.argv[1:]))]
tot = {"pos": 0, "theta": 0, "px1": 0, "px2": 0}
for name in names:
    recs = T.parse(os.path.join(golden, name))
    ref = refcpu.replay(recs, threads=os.cpu_count() or 1)
    got = replay.replay(recs, keep_intermediates=True)
    pos = theta = 0.0
    nbad = 0
    for fr, fg in zip(ref.flushes, got.flushes):
        n = fr.desc.tess_data_height * 2048
        if not n:
            continue
        rt, gt = fr.tess[:n], fg.tess[:n]
        same = (rt[:, :3] == gt[:, :3]).all(axis=1) | (np.isnan(rt[:, :3].view(np.float32)) & np.isnan(gt[:, :3].view(np.float32))).all(axis=1)
        nbad += int((~same).sum())
        rxy, gxy = rt[:, :2].view(np.float32), gt[:, :2].view(np.float32)
        pos = max(pos, float(np.nanmax(np.nan_to_num(np.abs(rxy - gxy)))))
        packed = ((rt[:, 3] >> 26) & 7) == 1
        dth = np.abs(rt[~packed, 2].view(np.float32) - gt[~packed, 2].view(np.float32))
        if dth.size:
            theta = max(theta, float(np.nanmax(np.nan_to_num(dth))))
    dmax = 0
    over2 = 0
    over0 = 0
    for a, b in zip(ref.frames, got.frames):
        d = np.abs(a.astype(int) - b.astype(int)).max(axis=-1)
        dmax = max(dmax, int(d.max()))
        over2 += int((d > 2).sum())
        over0 += int((d > 0).sum())
    print(f"{name:44s} tess words differing {nbad:7d} pos {pos:.3g} theta {theta:.3g} | px max {dmax} >0: {over0} >2: {over2}", flush=True)
