"""Print the per-fragment history of one pixel from both the oracle and the CUDA path.
usage: python tests/tools/debug_pixel.py <trace> <x> <y>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
trace, x, y = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
os.environ["REFCPU_DEBUG_PIXEL"] = f"{x},{y}"
os.environ["RIVECUDA_DEBUG_PIXEL"] = f"{x},{y}"
from rive_runtime_b200 import trace as T, replay as R
from oracle import refcpu
recs = T.parse(trace)
a = refcpu.replay(recs, threads=1, keep_intermediates=False).frames[-1]
sys.stderr.flush()
b = R.replay(recs).frames[-1]
print("oracle", a[y, x], "cuda", b[y, x])
