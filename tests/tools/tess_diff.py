#!/usr/bin/env python
"""Dev tool: worst tessellation-vertex deviations (K2 vs oracle) of a golden trace, with the
span each vertex belongs to. usage: tess_diff.py <name> [top]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refcpu  # noqa: E402
from rive_runtime_b200 import abi, replay, trace as T  # noqa: E402

abi.load()
name = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 10
path = name if os.path.exists(name) else os.path.join(ROOT, "tests", "golden", name + ".rvct.xz")
recs = T.parse(path)
ref = refcpu.replay(recs, threads=os.cpu_count() or 1)
got = replay.replay(recs, keep_intermediates=True)
host = {}
for r in recs:
    if r.tag == T.BUFFER_UNMAP:
        host[r.fields["kind"]] = r.data
fr, fg = ref.flushes[0], got.flushes[0]
d = fr.desc
n = d.tess_data_height * 2048
rt, gt = fr.tess[:n], fg.tess[:n]
rxy, gxy = rt[:, :2].view(np.float32), gt[:, :2].view(np.float32)
err = np.nan_to_num(np.abs(rxy - gxy).max(axis=1))
spans = np.frombuffer(host[6].tobytes()[:d.tess_vertex_span_count * 64], dtype=np.uint32).reshape(-1, 16)
y = spans[:, 10].view(np.float32).astype(np.int64)
x0 = ((spans[:, 12].astype(np.int64) & 0xffff) ^ 0x8000) - 0x8000
x1 = spans[:, 12].astype(np.int32) >> 16
start = y * 2048 + x0
end = y * 2048 + x1
print("scale", np.nanmax(np.abs(rxy)), "max err", err.max(), "count above 1e-3:", int((err > 1e-3).sum()))
for v in np.argsort(-err)[:top]:
    s = np.nonzero((start <= v) & (v < end))[0]
    print(f"vertex {v}: ref {rxy[v]} got {gxy[v]} err {err[v]:.5f} theta ref {rt[v, 2:3].view(np.float32)} got {gt[v, 2:3].view(np.float32)}")
    for k in s[:1]:
        sp = spans[k]
        seg = sp[14]
        print(f"   span {k} idx {v - start[k]} of {end[k] - start[k]}: pts {sp[:8].view(np.float32)} jt {sp[8:10].view(np.float32)} "
              f"par {seg & 1023} polar {(seg >> 10) & 1023} join {seg >> 20} flags {hex(sp[15])}")
