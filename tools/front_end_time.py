#!/usr/bin/env python
"""Times rivecuda_front_end_paths (H2D of the RawPaths + three passes + scans + two syncs) on
the committed path dumps. usage: front_end_time.py [name ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rive_runtime_b200 import abi, front_end as F, replay as R, trace as T  # noqa: E402

abi.load()
for name in sys.argv[1:] or ["c2_4k", "f1", "s1", "lots_of_tess_spans_stroke"]:
    golden = os.path.join(ROOT, "tests", "golden")
    dump = F.load_paths(os.path.join(golden, name + ".paths.xz"))
    recs = T.parse(os.path.join(golden, name + ".rvct.xz"))
    tc = next(r for r in recs if r.tag == T.TARGET_CREATE)
    with R.Replayer(0) as rp:
        result = R.ReplayResult()
        for r in recs:
            if r.tag in (T.CREATE, T.DESTROY, T.FLUSH, T.TARGET_READ, T.TARGET_DESTROY, T.POST_FLUSH):
                continue
            rp.apply(r, result)
        for _ in range(5):
            res = F.run(rp, dump, tc.fields["width"], tc.fields["height"])
        rp.sync()
        n = 50
        t0 = time.perf_counter()
        for _ in range(n):
            F.run(rp, dump, tc.fields["width"], tc.fields["height"])
        rp.sync()
        ms = (time.perf_counter() - t0) / n * 1e3
    print(f"{name}: {len(dump.paths)} paths, {res.tess_vertex_span_count} spans, {dump.verbs.size} verbs: {ms:.3f} ms per call")
