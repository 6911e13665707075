"""Parity sweep over flush traces of the reference's .sriv silvers (tests/unit_tests/silvers):
every frame is rendered by the CUDA path; a spread of frames per stream (first, second,
quartiles, last) is also rendered by the CPU oracle and compared (max channel delta, pixels
above 2/255, PSNR). Prints one line per stream and a JSON summary.

  tools/record_sriv_traces.sh <dir>          # here (needs /root/reference): 261 streams, ~21 MB xz
  python tests/tools/sriv_parity.py <dir> [out.json]
"""
import glob
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rive_runtime_b200 import trace as T, replay as R  # noqa: E402
from oracle import refcpu  # noqa: E402


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def main():
    src = sys.argv[1]
    out_path = sys.argv[2] if len(sys.argv) > 2 else None
    paths = sorted(glob.glob(os.path.join(src, "*.rvct.xz")) + glob.glob(os.path.join(src, "*.rvct")))
    summary = {"streams": 0, "frames_rendered": 0, "frames_compared": 0, "worst_delta": 0, "worst_stream": None,
               "streams_over_2": [], "min_psnr": 99.0, "failed": []}
    t_start = time.time()
    for p in paths:
        name = os.path.basename(p).split(".")[0]
        try:
            recs = T.parse(p)
            n = sum(1 for r in recs if r.tag == T.TARGET_READ)
            pick = sorted({0, min(1, n - 1), n // 4, n // 2, (3 * n) // 4, n - 1})
            got = R.replay(recs).frames
            ref = refcpu.replay(recs, threads=os.cpu_count() or 1, keep_intermediates=False, only_frames=set(pick)).frames
        except Exception as e:  # noqa: BLE001
            summary["failed"].append((name, str(e)[:200]))
            print(f"{name}: FAILED {e}", flush=True)
            continue
        worst, over, lo = 0, 0, 99.0
        for k in pick:
            d = np.abs(got[k].astype(int) - ref[k].astype(int)).max(axis=-1)
            worst = max(worst, int(d.max()))
            over += int((d > 2).sum())
            lo = min(lo, psnr(got[k], ref[k]))
        summary["streams"] += 1
        summary["frames_rendered"] += n
        summary["frames_compared"] += len(pick)
        summary["min_psnr"] = min(summary["min_psnr"], lo)
        if worst > summary["worst_delta"]:
            summary["worst_delta"], summary["worst_stream"] = worst, name
        if worst > 2:
            summary["streams_over_2"].append((name, worst, over))
        print(f"{name}: frames {n} compared {len(pick)} maxdelta {worst} px>2 {over} psnr {lo:.1f}", flush=True)
    summary["seconds"] = time.time() - t_start
    print(json.dumps(summary), flush=True)
    if out_path:
        json.dump(summary, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
