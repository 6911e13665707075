"""Find the worst pixel of one stream (CUDA vs oracle) among a spread of frames and print both
sides' per-fragment history of that pixel (needs librivecuda_debug.so: make -C csrc debug)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from rive_runtime_b200 import trace as T, replay as R
from oracle import refcpu

path = sys.argv[1]
recs = T.parse(path)
n = sum(1 for r in recs if r.tag == T.TARGET_READ)
pick = sorted({0, min(1, n - 1), n // 4, n // 2, (3 * n) // 4, n - 1})
got = R.replay(recs).frames
ref = refcpu.replay(recs, threads=os.cpu_count(), keep_intermediates=False, only_frames=set(pick)).frames
worst = (0, None)
for k in pick:
    d = np.abs(got[k].astype(int) - ref[k].astype(int)).max(axis=-1)
    if d.max() > worst[0]:
        y, x = np.unravel_index(np.argmax(d), d.shape)
        worst = (int(d.max()), (k, int(x), int(y)))
print("worst", worst, "cuda", got[worst[1][0]][worst[1][2], worst[1][1]], "oracle", ref[worst[1][0]][worst[1][2], worst[1][1]])
k, x, y = worst[1]
# Keep only frame k's flushes (all uploads stay: the rings are frame-wide).
frame, mini = 0, []
for r in recs:
    if r.tag == T.FLUSH and frame != k:
        continue
    if r.tag == T.TARGET_READ:
        frame += 1
        if frame - 1 != k:
            continue
    mini.append(r)
os.environ["REFCPU_DEBUG_PIXEL"] = f"{x},{y}"
os.environ["RIVECUDA_DEBUG_PIXEL"] = f"{x},{y}"
refcpu.replay(mini, threads=1, keep_intermediates=False)
sys.stderr.flush()
lib = os.path.join(ROOT, "rive-runtime_b200", "_build", "librivecuda_debug.so")
R.replay(mini, lib_path=lib)
