"""Regenerates tests/golden/oracle_frames.json (sha256 of the oracle's frames for
each committed trace). Run after an intentional oracle change."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refcpu  # noqa: E402
from rive_runtime_b200 import trace as T  # noqa: E402

golden = os.path.join(ROOT, "tests", "golden")
out = {}
for name in sorted(os.listdir(golden)):
    if not name.endswith(".rvct.xz") or name.startswith("c2_4k"):
        continue
    res = refcpu.replay(T.parse(os.path.join(golden, name)), threads=os.cpu_count())
    out[name] = {"frames": [hashlib.sha256(np.ascontiguousarray(f).tobytes()).hexdigest() for f in res.frames],
                 "shape": list(res.frames[0].shape)}
json.dump(out, open(os.path.join(golden, "oracle_frames.json"), "w"), indent=1, sort_keys=True)
print("wrote", len(out), "entries")
