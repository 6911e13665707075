"""Dev probe: event-timed render of one band of C5 in a single process (compare with the
per-stage sums of c5_profile.py and with the N-rank run of band_shard_check.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from rive_runtime_b200 import trace as T, sharding
import band_shard_check as B
path = "/tmp/band_c5.rvct"
if not os.path.exists(path):
    B.record_scene("c5", path, [])
t0 = time.time(); records = T.parse(path); print("parse s", time.time() - t0)
s = T.summarize(records); W, H = s["width"], s["height"]
n, rank = int(sys.argv[1]), int(sys.argv[2])
band = sharding.band_for_rank(H, rank, n)
frame = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda:0")
r = B.BandRenderer(records, 0, frame)
for i in range(4):
    t0 = time.perf_counter(); ms = r.render(band); wall = (time.perf_counter() - t0) * 1e3
    print(f"rep {i}: band {band} event ms {ms:.1f} wall ms {wall:.1f}", flush=True)

# where does the host time go?
from rive_runtime_b200 import replay as R
rp = r.rp
acc = {"prepare": 0.0, "flush": 0.0, "apply": 0.0, "restrict": 0.0}
res = R.ReplayResult()
t_all = time.perf_counter()
nfl = 0
for rec in records:
    if rec.tag in (T.CREATE, T.DESTROY, T.TARGET_READ, T.TARGET_DESTROY):
        continue
    if rec.tag == T.FLUSH:
        t0 = time.perf_counter(); pf = rp.prepare_flush(rec.fields["flush"]); t1 = time.perf_counter()
        pf.desc = sharding.restrict_to_band(pf.desc, band); t2 = time.perf_counter()
        rp.flush(pf); t3 = time.perf_counter()
        acc["prepare"] += t1 - t0; acc["restrict"] += t2 - t1; acc["flush"] += t3 - t2
        nfl += 1
        continue
    if rec.tag not in (T.BUFFER_UNMAP, T.PREPARE_TO_FLUSH, T.POST_FLUSH):
        continue
    t0 = time.perf_counter(); rp.apply(rec, res); acc["apply"] += time.perf_counter() - t0
rp.sync()
print("flushes", nfl, "host seconds", {k: round(v, 4) for k, v in acc.items()}, "total", round(time.perf_counter() - t_all, 4))
print("batches per flush", [len(rec.fields["flush"].batches) for rec in records if rec.tag == T.FLUSH][:8])
