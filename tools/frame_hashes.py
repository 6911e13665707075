"""sha256 of every frame the CUDA path renders for the committed traces (A/B checks of kernel changes)."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rive_runtime_b200 import trace as T, replay as R
golden = os.path.join(ROOT, "tests", "golden")
for name in sorted(os.listdir(golden)):
    if not name.endswith(".rvct.xz"):
        continue
    frames = R.replay(T.parse(os.path.join(golden, name))).frames
    h = hashlib.sha256()
    for f in frames:
        h.update(f.tobytes())
    print(name, h.hexdigest()[:16], flush=True)
