#!/usr/bin/env python
"""Dev tool (GPU): GMs drawn through both front ends (--budget-ms 0), frames compared.
usage: gm_front_end_diff.py gm:NAME ..."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
player = os.path.join(ROOT, "rive-runtime_b200", "_build", "rive_cuda_player")
env = dict(os.environ, RIVECUDA_LIB=os.path.join(ROOT, "rive-runtime_b200", "_build", "librivecuda.so"))
same = differ = failed = 0
with tempfile.TemporaryDirectory() as tmp:
    for scene in sys.argv[1:]:
        outs = []
        for extra in ([], ["--gpu-front-end"]):
            out = os.path.join(tmp, "f%d.rgba" % len(outs))
            if os.path.exists(out):
                os.remove(out)
            p = subprocess.run([player, "--scene", scene, "--budget-ms", "0", "--out", out, *extra], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=120)
            if p.returncode != 0 or not os.path.exists(out):
                print("FAILED", scene, extra, p.stderr.decode(errors="replace")[-200:].strip())
                break
            outs.append(np.fromfile(out, dtype=np.uint8))
        if len(outs) != 2:
            failed += 1
            continue
        if outs[0].size == outs[1].size and np.array_equal(outs[0], outs[1]):
            same += 1
        else:
            differ += 1
            d = np.abs(outs[0].astype(np.int16) - outs[1].astype(np.int16)) if outs[0].size == outs[1].size else np.array([999])
            print("DIFFERS", scene, "max", int(d.max()), "bytes", int((d > 0).sum()))
print("identical %d, differing %d, failed %d" % (same, differ, failed))
