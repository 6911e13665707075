#!/usr/bin/env python
"""Dev tool (GPU): draws a synthetic scene (f1, f1o, f1b, f1c, f1w, f1g, s1...) through the scene
player with --gpu-front-end and compares the frame with the replay of the committed trace the
reference front end produced for it. usage: f1_diff.py <scene>"""
import os, subprocess, sys
import numpy as np
sys.path.insert(0, '.')
from rive_runtime_b200 import abi, replay as R, trace as T
abi.load()
scene = sys.argv[1]
want = R.replay(T.parse(f'tests/golden/{scene}.rvct.xz')).frames[-1]
env = dict(os.environ, RIVECUDA_LIB='rive-runtime_b200/_build/librivecuda.so')
subprocess.check_call(['rive-runtime_b200/_build/rive_cuda_player', '--scene', scene, '--gpu-front-end', '--budget-ms', '0', '--out', '/tmp/frame.rgba'], env=env, stdout=subprocess.DEVNULL)
got = np.fromfile('/tmp/frame.rgba', dtype=np.uint8).reshape(want.shape)
d = np.abs(got.astype(int) - want.astype(int)).max(-1)
ys, xs = np.nonzero(d)
print('differing pixels', len(ys), 'max', d.max())
if len(ys):
    print('bbox', xs.min(), ys.min(), xs.max(), ys.max())
    for k in range(0, len(ys), max(1, len(ys) // 12)):
        print(xs[k], ys[k], got[ys[k], xs[k]], want[ys[k], xs[k]])
    np.save('gpurun_out/f1g_diff.npy', d.astype(np.uint8))
