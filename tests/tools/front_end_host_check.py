#!/usr/bin/env python
"""Dev tool: run the host build of the front-end core (oracle/front_end_host) on a --dump-paths
file and compare the generated buffers with the ones in the flush trace recorded from the same
frame. usage: front_end_host_check.py <trace.rvct[.xz]> <dump.paths[.xz]>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import front_end_host as H  # noqa: E402
from rive_runtime_b200 import front_end as F, trace as T  # noqa: E402


def main():
    recs = T.parse(sys.argv[1])
    dump = F.load_paths(sys.argv[2])
    print("paths", len(dump.paths), "complete", dump.complete)
    host = {r.fields["kind"]: r.data for r in recs if r.tag == T.BUFFER_UNMAP}
    fr = next(r.fields["flush"] for r in recs if r.tag == T.FLUSH)
    d = fr.desc
    tc = next(r for r in recs if r.tag == T.TARGET_CREATE)
    out = H.run(dump, tc.fields["width"], tc.fields["height"])
    res = out.result
    print("result", res.path_count, res.contour_count, res.tess_vertex_span_count, res.midpoint_fan_tess_vertex_count, res.tess_data_height,
          "want", d.path_count, d.contour_count, d.tess_vertex_span_count, d.tess_data_height)
    n = min(res.tess_vertex_span_count, d.tess_vertex_span_count)
    want = np.frombuffer(host[6].tobytes()[:d.tess_vertex_span_count * 64], dtype=np.uint32).reshape(-1, 16)
    got = out.spans[:res.tess_vertex_span_count]
    bad = np.nonzero((got[:n] != want[:n]).any(axis=1))[0]
    print("spans differing:", bad.size, "of", n)
    for b in bad[:int(os.environ.get("SHOW", "5"))]:
        print(" span", b)
        print("   got ", got[b, :10].view(np.float32), got[b, 10:12].view(np.float32), [hex(x) for x in got[b, 12:]])
        print("   want", want[b, :10].view(np.float32), want[b, 10:12].view(np.float32), [hex(x) for x in want[b, 12:]])
    nc = min(res.contour_count, d.contour_count)
    wc = np.frombuffer(host[4].tobytes()[:d.contour_count * 16], dtype=np.uint32).reshape(-1, 4)
    badc = np.nonzero((out.contours[:nc] != wc[:nc]).any(axis=1))[0]
    print("contours differing:", badc.size, "of", nc, badc[:5])
    for b in badc[:3]:
        print("   got", out.contours[b], "want", wc[b])
    npth = min(res.path_count, d.path_count)
    wp = np.frombuffer(host[1].tobytes()[:d.path_count * 64], dtype=np.uint32).reshape(-1, 16)
    badp = np.nonzero((out.path_data[1:npth, :8] != wp[1:npth, :8]).any(axis=1))[0]
    print("path records differing:", badp.size, badp[:5])
    wpt = np.frombuffer(host[2].tobytes()[:d.path_count * 8], dtype=np.uint32).reshape(-1, 2)
    badpt = np.nonzero((out.paint_data[1:npth] != wpt[1:npth]).any(axis=1))[0]
    print("paint records differing:", badpt.size, badpt[:5])
    if badpt.size:
        print("   got", [hex(x) for x in out.paint_data[1 + badpt[0]]], "want", [hex(x) for x in wpt[1 + badpt[0]]])


main()
