"""Dev tool (GPU): frames/s of one flush trace replayed back to back with K rivecuda contexts on one
GPU taking the frames in turn (K frames in flight, each context with its own target, streams and
scratch). usage: frames_in_flight.py [trace] [K ...]"""
import sys, time
sys.path.insert(0, ".")
from rive_runtime_b200 import trace as T, replay as R

recs = T.parse(sys.argv[1] if len(sys.argv) > 1 else "tests/golden/c2_4k.rvct.xz")
for K in [int(a) for a in sys.argv[2:]] or [1, 2, 3]:
    rps, flushes = [], []
    for k in range(K):
        rp = R.Replayer(0)
        res = R.ReplayResult()
        fl = []
        for r in recs:
            if r.tag in (T.CREATE, T.DESTROY, T.TARGET_READ, T.TARGET_DESTROY):
                continue
            if r.tag == T.FLUSH:
                fl.append(rp.prepare_flush(r.fields["flush"]))
                continue
            rp.apply(r, res)
        rps.append(rp)
        flushes.append(fl)
    N = 60
    for it in range(N + 6):
        if it == 6:
            for rp in rps:
                rp.sync()
            t0 = time.perf_counter()
        rp = rps[it % K]
        for pf in flushes[it % K]:
            rp.flush(pf)
    for rp in rps:
        rp.sync()
    dt = time.perf_counter() - t0
    print("contexts %d: %.3f ms/frame, %.1f frames/s" % (K, dt * 1e3 / N, N / dt), flush=True)
    for rp in rps:
        rp.close()
