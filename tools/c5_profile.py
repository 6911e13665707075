"""Per-stage timings of the C5 scene (16384x16384, 200k paths, 36 logical flushes), whole frame
and restricted to one of N screen bands (dev tool). usage: c5_profile.py [N [rank]]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rive_runtime_b200 import trace as T, replay as R, sharding
path = "/tmp/band_c5.rvct"
if not os.path.exists(path):
    build = os.path.join(ROOT, "rive-runtime_b200", "_build")
    env = dict(os.environ, RIVECUDA_LIB=os.path.join(build, "librivecuda_trace.so"), RIVECUDA_TRACE_OUT=path)
    subprocess.check_call([os.path.join(build, "rive_cuda_player"), "--scene", "c5"], env=env, stdout=subprocess.DEVNULL)
recs = T.parse(path)
n_bands = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else n_bands // 2
for mode in ("warm-up", "whole frame", f"band {rank} of {n_bands}"):
    rp = R.Replayer(0, profiling=True)
    res = R.ReplayResult()
    acc = {"tessellate_ms": 0.0, "setup_bin_ms": 0.0, "raster_ms": 0.0, "total_ms": 0.0}
    tris = entries = 0
    for r in recs:
        if r.tag in (T.CREATE, T.DESTROY, T.TARGET_READ, T.TARGET_DESTROY):
            continue
        if r.tag == T.FLUSH:
            fr = r.fields["flush"]
            pf = rp.prepare_flush(fr)
            if mode.startswith("band"):
                h = rp.target_shapes[fr.target_id][0]
                pf.desc = sharding.restrict_to_band(pf.desc, sharding.band_for_rank(h, rank, n_bands))
            rp.flush(pf)
            tm = rp.timings()
            for k in acc:
                acc[k] += getattr(tm, k)
            tris += tm.triangle_count
            entries += tm.tile_entry_count
            continue
        rp.apply(r, res)
    print(mode, {k: round(v, 2) for k, v in acc.items()}, "tris", tris, "entries", entries, flush=True)
    rp.close()
