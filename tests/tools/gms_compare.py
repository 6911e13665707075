#!/usr/bin/env python
"""Frames as PNGs, compared with the reference's own comparator (north_star: "via the repo's
image_diff.py").

  gms_compare.py --write-cuda DIR [scene ...]   on a B200: render the scenes' committed flush traces
                                                with the CUDA path, one RGBA PNG per frame
                                                (tests/golden/cuda_png/ holds a committed set)
  gms_compare.py --write-oracle DIR [scene ...] the same with the CPU oracle
  gms_compare.py --diff CANDIDATE GOLDEN        run /root/reference/tests/image_diff.py on the two
                                                directories (the reference's tool, unmodified, in place)
                                                and print per image: status, max_diff, PSNR

tests/test_image_diff_cpu.py runs the comparator on (oracle PNGs rendered on the spot, committed CUDA
PNGs); tests/test_parity_gpu.py::test_committed_cuda_pngs_are_current keeps the committed set equal
to what the kernels render today.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
IMAGE_DIFF = "/root/reference/tests/image_diff.py"
# Small scenes that between them use strokes, feathers (direct and atlas), clips, clip rects,
# gradients, every blend mode and an image paint.
SCENES = ["beziers", "strokes_round", "xfermodes2", "cliprects", "parallelclips", "verycomplexgrad",
          "trickycubicstrokes", "img"]


def write_png(path, rgba):
    import cv2
    cv2.imwrite(path, np.ascontiguousarray(rgba[..., [2, 1, 0, 3]]))


def read_png(path):
    import cv2
    bgra = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    return np.ascontiguousarray(bgra[..., [2, 1, 0, 3]])


def render(scene, which):
    from rive_runtime_b200 import trace as T
    recs = T.parse(os.path.join(GOLDEN, scene + ".rvct.xz"))
    if which == "cuda":
        from rive_runtime_b200 import replay
        return replay.replay(recs).frames
    from oracle import refcpu
    return refcpu.replay(recs, threads=os.cpu_count() or 1, keep_intermediates=False).frames


def write_set(directory, scenes, which):
    os.makedirs(directory, exist_ok=True)
    for scene in scenes:
        for k, frame in enumerate(render(scene, which)):
            write_png(os.path.join(directory, f"{scene}.{k}.png" if k else f"{scene}.png"), frame)


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def image_diff(candidate_dir, golden_dir):
    """{name: (status, max_diff, differing_pixels, psnr)} from the reference's image_diff.py."""
    names = sorted(n for n in os.listdir(candidate_dir) if n.endswith(".png"))
    with tempfile.TemporaryDirectory() as tmp:
        status = os.path.join(tmp, "status.txt")
        subprocess.check_call([sys.executable, IMAGE_DIFF, "--names", *names, "--candidate", candidate_dir, "--golden", golden_dir,
                               "--status", status])
        out = {}
        for line in open(status):
            f = line.rstrip("\n").split("\t")
            a, b = read_png(os.path.join(candidate_dir, f[0] + ".png")), read_png(os.path.join(golden_dir, f[0] + ".png"))
            if f[1] == "identical":
                out[f[0]] = ("identical", 0, 0, psnr(a, b))
            elif f[1].isdigit():
                out[f[0]] = ("different", int(f[1]), int(f[3]), psnr(a, b))
            else:
                out[f[0]] = (f[1], None, None, None)
        return out


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] in ("--write-cuda", "--write-oracle"):
        write_set(sys.argv[2], sys.argv[3:] or SCENES, "cuda" if sys.argv[1] == "--write-cuda" else "oracle")
    elif len(sys.argv) == 4 and sys.argv[1] == "--diff":
        for name, (status, max_diff, count, p) in image_diff(sys.argv[2], sys.argv[3]).items():
            print(f"{name:32s} {status:10s} max_diff {max_diff} differing {count} psnr {p}")
    else:
        print(__doc__)
        sys.exit(2)
