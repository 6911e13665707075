#!/usr/bin/env python
"""Dev tool (GPU): every .riv in a directory is played twice through the scene player -- once
with the reference's CPU front end, once with --gpu-front-end (CudaPathRenderer +
rivecuda_front_end_paths) -- and the frames are compared. Assets whose frame holds something the
device front end refuses (gradients, clip paths, images, feathers) are counted as refused.
usage: riv_front_end_sweep.py <dir with .riv files> [frames] > report"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
player = os.path.join(ROOT, "rive-runtime_b200", "_build", "rive_cuda_player")
env = dict(os.environ, RIVECUDA_LIB=os.path.join(ROOT, "rive-runtime_b200", "_build", "librivecuda.so"))
frames = sys.argv[2] if len(sys.argv) > 2 else "20"
same = differ = refused = failed = 0
worst = []
with tempfile.TemporaryDirectory() as tmp:
    for name in sorted(os.listdir(sys.argv[1])):
        if not name.endswith(".riv"):
            continue
        outs = []
        ok = True
        for extra in ([], ["--gpu-front-end"]):
            out = os.path.join(tmp, "f%d.rgba" % len(outs))
            if os.path.exists(out):
                os.remove(out)
            p = subprocess.run([player, "--scene", "riv:" + os.path.join(sys.argv[1], name), "--frames", frames, "--budget-ms", "0", "--out", out, *extra],
                               env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=120)
            if p.returncode != 0 or not os.path.exists(out):
                if b"cannot draw this frame" in p.stderr:
                    refused += 1
                else:
                    failed += 1
                    print("FAILED", name, extra, p.stderr.decode(errors="replace")[-300:])
                ok = False
                break
            outs.append(np.fromfile(out, dtype=np.uint8))
        if not ok:
            continue
        if outs[0].size == outs[1].size and np.array_equal(outs[0], outs[1]):
            same += 1
        else:
            differ += 1
            d = np.abs(outs[0].astype(np.int16) - outs[1].astype(np.int16)) if outs[0].size == outs[1].size else np.array([999])
            worst.append((int(d.max()), int((d > 0).sum()), name))
            print("DIFFERS", name, "max", int(d.max()), "bytes", int((d > 0).sum()))
print("identical %d, differing %d, refused %d, failed %d" % (same, differ, refused, failed))
