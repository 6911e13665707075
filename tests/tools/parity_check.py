"""Compare the CUDA path (through the C ABI) with the CPU oracle on recorded
traces. Prints per-scene max channel delta, PSNR and tessellation parity; writes
diff images to gpurun_out/ for inspection. Dev tool (the pytest version is
tests/test_parity_gpu.py)."""
import glob
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
    paths = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "tests/golden/*.rvct.xz")))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    worst = 0
    for p in paths:
        name = os.path.basename(p).split(".")[0]
        recs = T.parse(p)
        t0 = time.time()
        ref = refcpu.replay(recs, threads=os.cpu_count())
        t1 = time.time()
        try:
            got = R.replay(recs, keep_intermediates=True, profiling=True)
        except Exception as e:  # noqa: BLE001
            print(f"{name}: CUDA FAILED: {e}")
            worst = 999
            continue
        t2 = time.time()
        line = f"{name}: oracle {t1 - t0:.2f}s cuda {t2 - t1:.2f}s"
        for fi, (fr, fg) in enumerate(zip(ref.flushes, got.flushes)):
            if fg.tess is not None:
                n = fr.desc.tess_data_height * 2048
                rt, gt = fr.tess[:n], fg.tess[:n]
                flags_equal = np.array_equal(rt[:, 3], gt[:, 3])
                # xy and theta as float
                rxy = rt[:, :2].view(np.float32); gxy = gt[:, :2].view(np.float32)
                dxy = np.nanmax(np.abs(rxy - gxy)) if n else 0
                is_packed = ((rt[:, 3] >> 26) & 7) == 1
                rth = rt[:, 2].view(np.float32); gth = gt[:, 2].view(np.float32)
                dth = np.abs(rth - gth); dth = np.minimum(dth, np.abs(dth - 2 * np.pi))
                dth = np.where(is_packed, (rt[:, 2] != gt[:, 2]).astype(np.float32), dth)
                line += f" | flush{fi} tess flags_eq={flags_equal} dxy={dxy:.2e} dtheta={np.nanmax(dth) if n else 0:.2e}"
            if fg.grad is not None:
                dg = np.abs(fr.grad[:fr.desc.grad_data_height].astype(int) - fg.grad.astype(int)).max()
                line += f" grad_maxdiff={dg}"
            if fg.timings is not None:
                tm = fg.timings
                line += f" | ms: tess {tm.tessellate_ms:.3f} setup {tm.setup_bin_ms:.3f} raster {tm.raster_ms:.3f} tris {tm.triangle_count} entries {tm.tile_entry_count}"
        for k, (a, b) in enumerate(zip(ref.frames, got.frames)):
            d = np.abs(a.astype(int) - b.astype(int))
            mx = int(d.max())
            worst = max(worst, mx)
            cnt = int((d.max(axis=-1) > 2).sum())
            line += f" | frame{k} maxdiff={mx} n>2={cnt} psnr={psnr(a, b):.1f}"
            if mx > 2:
                from PIL import Image
                vis = np.clip(d.max(axis=-1) * 40, 0, 255).astype(np.uint8)
                Image.fromarray(vis).save(os.path.join(ROOT, "gpurun_out", f"diff_{name}_{k}.png"))
                refcpu.save_png(os.path.join(ROOT, "gpurun_out", f"cuda_{name}_{k}.png"), b)
        print(line, flush=True)
    print("WORST", worst)


if __name__ == "__main__":
    main()
