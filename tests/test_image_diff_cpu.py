"""north_star: pixels "within a stated tolerance ... (max per-channel delta <= 2/255 and PSNR >= 45 dB
via the repo's image_diff.py)". The reference's comparator (/root/reference/tests/image_diff.py,
unmodified, run in place) is applied to PNG pairs: golden = the CPU oracle's frame rendered here,
candidate = the CUDA path's frame of the same scene as rendered on a B200 and committed under
tests/golden/cuda_png/ (test_parity_gpu.py::test_committed_cuda_pngs_are_current fails as soon as the
kernels render anything else). Skipped where the reference tree is absent."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
import gms_compare as G  # noqa: E402

CUDA_PNGS = os.path.join(ROOT, "tests", "golden", "cuda_png")

pytestmark = pytest.mark.skipif(not os.path.exists(G.IMAGE_DIFF), reason="needs /root/reference/tests/image_diff.py")


def test_reference_image_diff_accepts_the_cuda_frames(tmp_path):
    scenes = sorted({n.split(".")[0] for n in os.listdir(CUDA_PNGS) if n.endswith(".png")})
    assert len(scenes) >= 8
    oracle_dir = str(tmp_path / "oracle")
    G.write_set(oracle_dir, scenes, "oracle")
    results = G.image_diff(CUDA_PNGS, oracle_dir)
    assert sorted(results) == sorted(n[:-4] for n in os.listdir(CUDA_PNGS) if n.endswith(".png"))
    for name, (status, max_diff, _, psnr) in results.items():
        assert status in ("identical", "different"), (name, status)
        assert max_diff <= 2, f"{name}: image_diff.py max_diff {max_diff}"
        assert psnr >= 45.0, f"{name}: PSNR {psnr:.1f} dB"
