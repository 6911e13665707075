"""Replays a few traces + the GPU front end; run under compute-sanitizer (memcheck / racecheck):
  compute-sanitizer --tool memcheck python tools/sanitizer_scenes.py beziers img c1 ..."""
import sys, os
sys.path.insert(0, ".")
from rive_runtime_b200 import trace as T, replay as R, front_end as F
for name in sys.argv[1:]:
    recs = T.parse(f"tests/golden/{name}.rvct.xz")
    if name == "anim_juice":
        keep, frames = [], 0
        for r in recs:
            keep.append(r)
            if r.tag == T.TARGET_READ:
                frames += 1
                if frames == 3: break
        recs = keep
    R.replay(recs)
    print("ok", name, flush=True)
# front end
recs = T.parse("tests/golden/f1.rvct.xz")
dump = F.load_paths("tests/golden/f1.paths.xz")
fr = next(r.fields["flush"] for r in recs if r.tag == T.FLUSH)
with R.Replayer(0) as rp:
    res = R.ReplayResult()
    for r in recs:
        if r.tag in (T.CREATE, T.DESTROY, T.FLUSH, T.TARGET_READ, T.TARGET_DESTROY, T.POST_FLUSH): continue
        if r.tag == T.BUFFER_UNMAP and r.fields["kind"] in (1, 2, 3, 4, 6): continue
        rp.apply(r, res)
    F.run(rp, dump)
    rp.flush(rp.prepare_flush(fr))
    rp.read_target(fr.target_id)
print("ok front end", flush=True)
