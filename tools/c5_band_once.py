"""One band (rank 3 of 8) of the C5 frame, once (for ncu launch lists). Needs /tmp/band_c5.rvct (tools/c5_profile.py records it)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rive_runtime_b200 import trace as T, replay as R, sharding
recs = T.parse("/tmp/band_c5.rvct")
n, rank = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8, 3)
for it in range(2):
    rp = R.Replayer(0); res = R.ReplayResult()
    for r in recs:
        if r.tag in (T.CREATE, T.DESTROY, T.TARGET_READ, T.TARGET_DESTROY): continue
        if r.tag == T.FLUSH:
            fr = r.fields["flush"]; pf = rp.prepare_flush(fr)
            if n > 1:
                pf.desc = sharding.restrict_to_band(pf.desc, sharding.band_for_rank(rp.target_shapes[fr.target_id][0], rank, n))
            rp.flush(pf)
            continue
        rp.apply(r, res)
    rp.lib.rivecuda_sync(rp.ctx)
    rp.close()
