"""Per-kernel timings of the first flushes of the C5 scene (dev tool)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rive_runtime_b200 import trace as T, replay as R
path = "/tmp/band_c5.rvct"
if not os.path.exists(path):
    build = os.path.join(ROOT, "rive-runtime_b200", "_build")
    env = dict(os.environ, RIVECUDA_LIB=os.path.join(build, "librivecuda_trace.so"), RIVECUDA_TRACE_OUT=path)
    subprocess.check_call([os.path.join(build, "rive_cuda_player"), "--scene", "c5"], env=env, stdout=subprocess.DEVNULL)
recs = T.parse(path)
rp = R.Replayer(0, profiling=True)
res = R.ReplayResult()
n = 0
for r in recs:
    if r.tag in (T.CREATE, T.DESTROY, T.TARGET_READ, T.TARGET_DESTROY):
        continue
    rp.apply(r, res)
    if r.tag == T.FLUSH:
        tm = rp.timings()
        print(f"flush {n}: tess {tm.tessellate_ms:.2f} setup+bin {tm.setup_bin_ms:.2f} raster {tm.raster_ms:.2f} total {tm.total_ms:.2f} ms; tris {tm.triangle_count} entries {tm.tile_entry_count}", flush=True)
        n += 1
        if n >= int(sys.argv[1]) if len(sys.argv) > 1 else 3:
            break
rp.close()
