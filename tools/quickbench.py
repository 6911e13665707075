import sys, os, ctypes, time
sys.path.insert(0, ".")
import numpy as np
from rive_runtime_b200 import trace as T, replay as R
recs = T.parse(sys.argv[1] if len(sys.argv) > 1 else "tests/golden/c2_4k.rvct.xz")
rp = R.Replayer(0, profiling=True)
res = R.ReplayResult()
flushes = []
for r in recs:
    if r.tag in (T.CREATE, T.DESTROY, T.TARGET_READ, T.TARGET_DESTROY): continue
    if r.tag == T.FLUSH:
        flushes.append(rp.prepare_flush(r.fields["flush"])); continue
    rp.apply(r, res)
acc = {}
N = 10
for it in range(N + 3):
    for pf in flushes:
        rp.flush(pf)
        tm = rp.timings()
        if it >= 3:
            for k in ("tessellate_ms", "atlas_ms", "setup_bin_ms", "raster_ms", "total_ms"):
                acc[k] = acc.get(k, 0) + getattr(tm, k)
print(os.environ.get("RIVECUDA_LIB", "default"), {k: round(v / N, 4) for k, v in acc.items()})
rp.close()
