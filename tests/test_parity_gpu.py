"""Parity tests proper: the CUDA path, called through the C ABI with the host
buffers the reference front end produced, against the CPU oracle on the same
inputs. Tolerance (north_star): max per-channel delta <= 2/255 -- the `max_diff` the
reference's tests/image_diff.py reports (max over pixels and channels of the absolute
difference, image_diff.py:92-119) -- and PSNR >= 45 dB. The tessellation texture is
bit-exact (flags, ids, positions, angles); gradient ramps are bit-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_traces

pytestmark = pytest.mark.gpu

MAX_DELTA = 2          # /255, per channel
MIN_PSNR = 45.0        # dB

# No scene is exempt. Two rasterisers share the pipeline:
#   * raster_tiles_exact_kernel interpolates coverage and varyings operation for operation as
#     the oracle does (fp64 barycentrics of the snapped vertices, the path's last fragment
#     decides the paint varyings): its frames are BIT-IDENTICAL to the oracle's. It runs
#     whenever a flush uses advanced blend modes, clip rectangles, image paints or meshes --
#     wherever a one-LSB difference could be amplified -- and on request (RIVECUDA_EXACT=1);
#   * raster_tiles_kernel evaluates coverage planes in fp32: within 1/255 of the oracle.
# Every scene is rendered both ways: default selection within MAX_DELTA with zero outliers,
# exact rasteriser identical to the oracle.


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


@pytest.fixture(scope="module")
def libs(built):
    from rive_runtime_b200 import replay, trace, abi
    from oracle import refcpu
    abi.load()  # fail loudly if the CUDA extension is missing
    return replay, trace, refcpu


@pytest.mark.parametrize("name", golden_traces())
def test_scene_parity(libs, name):
    replay, T, refcpu = libs
    recs = T.parse(os.path.join(GOLDEN, name))
    ref = refcpu.replay(recs, threads=os.cpu_count() or 1)
    got = replay.replay(recs, keep_intermediates=True)
    assert len(got.frames) == len(ref.frames) >= 1
    for fr, fg in zip(ref.flushes, got.flushes):
        n = fr.desc.tess_data_height * 2048
        if n:
            rt, gt = fr.tess[:n], fg.tess[:n]
            assert np.array_equal(rt[:, 3], gt[:, 3]), "contourIDWithFlags must be bit-exact"
            packed = ((rt[:, 3] >> 26) & 7) == 1  # feather joins pack (segmentCount<<16 | vertexID)
            assert np.array_equal(rt[packed, 2], gt[packed, 2])
            # Positions and angles: the kernels and the oracle run the same float operations
            # un-contracted and round their transcendentals once from double, so the tessellation
            # texture is bit-identical (it is on every committed scene). A double-rounding tie in
            # one of those calls (~1e-8 per call) may move a vertex by an ulp: allow a handful,
            # within the old tolerances.
            rf, gf = rt[:, :3].view(np.float32), gt[:, :3].view(np.float32)
            differs = ((rt[:, :3] != gt[:, :3]) & ~(np.isnan(rf) & np.isnan(gf))).any(axis=1)
            differs[packed] = (rt[packed, :2] != gt[packed, :2]).any(axis=1)
            assert int(differs.sum()) <= 4, f"{int(differs.sum())} tessellated vertices differ"
            if differs.any():
                rxy, gxy = rf[differs, :2], gf[differs, :2]
                scale = max(1.0, float(np.nanmax(np.abs(rxy))))
                assert np.nanmax(np.abs(rxy - gxy)) <= 2e-6 * scale + 1e-4
                dth = np.abs(rf[differs & ~packed, 2] - gf[differs & ~packed, 2])
                dth = np.minimum(dth, np.abs(dth - 2 * np.pi))
                assert dth.size == 0 or np.nanmax(dth) <= 2.5e-4
        if fr.desc.grad_data_height:
            assert np.array_equal(fr.grad[:fr.desc.grad_data_height], fg.grad), "colour ramps must be bit-exact"
    for a, b in zip(ref.frames, got.frames):
        d = np.abs(a.astype(int) - b.astype(int)).max(axis=-1)
        assert int(d.max()) <= MAX_DELTA, f"{name}: max channel delta {int(d.max())}/255 on {int((d > MAX_DELTA).sum())} pixels"
        assert psnr(a, b) >= MIN_PSNR, f"{name}: PSNR {psnr(a, b):.1f} dB"
    os.environ["RIVECUDA_EXACT"] = "1"
    try:
        exact = replay.replay(recs)
    finally:
        del os.environ["RIVECUDA_EXACT"]
    for k, (a, b) in enumerate(zip(ref.frames, exact.frames)):
        differing = int((a != b).any(axis=-1).sum())
        assert differing == 0, f"{name}: frame {k}: {differing} pixels of the exact rasteriser's frame differ from the oracle's"



SPAN_SCENES = ["c1", "s1", "beziers", "parallelclips", "largeclippedpath_winding_nested", "verycomplexgrad", "negative_interior_triangles",
               "batchedtriangulations", "strokes_round", "poly_evenOdd", "retrofitcubictristrips", "interleavedfillrule", "overfill_transparent"]


@pytest.mark.parametrize("name", SPAN_SCENES)
def test_span_rasteriser_agrees_with_the_in_order_rasteriser(libs, name, monkeypatch):
    """raster_spans_kernel (per-row spans into shared-memory delta planes, any order inside a path)
    against raster_tiles_kernel (every fragment in API order, fp16 rounding per fragment) on fills,
    strokes, interior triangulation, clips, nested clips, gradients and both fill rules: the same
    frame to within 1/255, and the flush timings must name the kernel that ran -- a silent switch
    of rasteriser would otherwise go unnoticed."""
    replay, T, _ = libs
    recs = T.parse(os.path.join(GOLDEN, name + ".rvct.xz"))

    def render(spans):
        monkeypatch.setenv("RIVECUDA_EXACT", "0")
        monkeypatch.setenv("RIVECUDA_SPANS", "1" if spans else "0")
        kernels = set()
        result = replay.ReplayResult()
        with replay.Replayer(0, profiling=True) as rp:
            for r in recs:
                if r.tag in (T.CREATE, T.DESTROY):
                    continue
                rp.apply(r, result)
                if r.tag == T.FLUSH:
                    kernels.add(rp.timings().raster_kernel)
        return result.frames, kernels

    fast, fast_kernels = render(True)
    ordered, ordered_kernels = render(False)
    assert 2 in fast_kernels and ordered_kernels == {0}
    assert len(fast) == len(ordered) >= 1
    for a, b in zip(fast, ordered):
        assert int(np.abs(a.astype(int) - b.astype(int)).max()) <= 1

def test_c2_full_size_parity_and_properties(libs):
    """BASELINE.json configs[1] at full size (10k paths, 3840x2160): parity with the
    oracle, idempotence, and band decomposition (size-independent properties)."""
    replay, T, refcpu = libs
    from rive_runtime_b200 import sharding
    recs = T.parse(os.path.join(GOLDEN, "c2_4k.rvct.xz"))
    got = replay.replay(recs)
    again = replay.replay(recs)
    assert np.array_equal(got.frames[0], again.frames[0]), "rendering must be deterministic"
    ref = refcpu.replay(recs, threads=os.cpu_count() or 1, keep_intermediates=False)
    delta = int(np.abs(ref.frames[0].astype(int) - got.frames[0].astype(int)).max())
    assert delta <= MAX_DELTA and psnr(ref.frames[0], got.frames[0]) >= MIN_PSNR
    # Render the frame as 4 screen bands (what tile-band sharding does) into one target:
    # the composite must be bit-identical to the single-pass render.
    result = replay.ReplayResult()
    with replay.Replayer(0) as rp:
        for r in recs:
            if r.tag in (T.CREATE, T.DESTROY, T.TARGET_READ, T.TARGET_DESTROY):
                continue
            if r.tag == T.FLUSH:
                fr = r.fields["flush"]
                pf = rp.prepare_flush(fr)
                h = rp.target_shapes[fr.target_id][0]
                for band_rank in range(4):
                    band = sharding.band_for_rank(h, band_rank, 4)
                    pf.desc = sharding.restrict_to_band(rp.prepare_flush(fr).desc, band)
                    rp.flush(pf)
                continue
            rp.apply(r, result)
        banded = rp.read_target(1)
    assert np.array_equal(banded, got.frames[0])


@pytest.mark.parametrize("name", ["s1", "c3", "c1", "strokes_round", "trickycubicstrokes", "feather_strokes", "img", "riv_off_road_car"])
def test_band_decomposition_is_bit_identical(libs, name):
    """Screen-band sharding on strokes, joins, caps, feathers, clips, images and a real
    .riv frame: every flush rendered as 5 bands -- with whole patches outside a band dropped before
    their vertices are shaded -- must composite to exactly the single-pass frame."""
    replay, T, _ = libs
    from rive_runtime_b200 import sharding
    recs = T.parse(os.path.join(GOLDEN, name + ".rvct.xz"))
    want = replay.replay(recs).frames[-1]
    result = replay.ReplayResult()
    target_id = None
    with replay.Replayer(0) as rp:
        for r in recs:
            if r.tag in (T.CREATE, T.DESTROY, T.TARGET_READ, T.TARGET_DESTROY):
                continue
            if r.tag == T.FLUSH:
                fr = r.fields["flush"]
                target_id = fr.target_id
                pf = rp.prepare_flush(fr)
                full = pf.desc
                h = rp.target_shapes[fr.target_id][0]
                for band_rank in range(5):
                    pf.desc = sharding.restrict_to_band(full, sharding.band_for_rank(h, band_rank, 5))
                    rp.flush(pf)
                continue
            rp.apply(r, result)
        banded = rp.read_target(target_id)
    assert np.array_equal(banded, want)


def test_clear_only_and_preserve(libs):
    """Empty draw list: clear fills exactly the premultiplied clear colour inside the
    update bounds; preserveRenderTarget leaves pixels untouched."""
    import ctypes
    replay, T, _ = libs
    recs = T.parse(os.path.join(GOLDEN, "beziers.rvct.xz"))
    with replay.Replayer(0) as rp:
        result = replay.ReplayResult()
        for r in recs:
            if r.tag in (T.STATIC_TABLES, T.BUFFER_RESIZE, T.BUFFER_UNMAP, T.RESIZE_GRADIENT, T.RESIZE_TESSELLATION,
                         T.TARGET_CREATE):
                rp.apply(r, result)
        fr = next(r.fields["flush"] for r in recs if r.tag == T.FLUSH)
        pf = rp.prepare_flush(fr)
        pf.batch_count = 0
        pf.desc.color_clear_value = 0x80ff8040  # a=128 r=255 g=128 b=64
        pf.desc.update_bounds[:] = [16, 32, 200, 300]
        rp.flush(pf)
        px = rp.read_target(1)
        inside = px[32:300, 16:200].reshape(-1, 4)
        a = 128 / 255
        want = [int(255 / 255 * a * 255 + .5), int(128 / 255 * a * 255 + .5), int(64 / 255 * a * 255 + .5), 128]
        assert (inside == np.array(want, np.uint8)).all()
        assert px[:32].max() == 0 and px[:, :16].max() == 0 and px[300:].max() == 0 and px[:, 200:].max() == 0
        before = px.copy()
        pf.desc.color_load_action = 1  # preserveRenderTarget
        pf.desc.update_bounds[:] = [0, 0, 400, 800]
        rp.flush(pf)
        assert np.array_equal(rp.read_target(1), before)


def test_band_sharding_over_two_gpus_with_nccl_gather():
    """one frame as screen-tile bands on 2 GPUs, composited by one NCCL gather,
    must equal the single-GPU render bit for bit (skipped on a 1-GPU box; the CPU/gloo
    version of the partitioning logic is tests/test_sharding_cpu.py)."""
    import json
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(GOLDEN.rstrip("/")).rsplit("/tests", 1)[0]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(root, "tools", "band_shard_check.py"), os.path.join(GOLDEN, "c3.rvct.xz")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["composite_identical_to_single_pass"] and line["n_gpus"] == 2


@pytest.mark.parametrize("tiled,whole", [("c1_virtual_tiles", "c1"), ("feather_shapes_virtual_tiles", "feather_shapes")])
def test_virtual_tiles_render_the_same_pixels(libs, tiled, whole):
    """FrameDescriptor::virtualTileWidth / Height: RenderContextCUDAImpl draws the flush
    virtual tile by virtual tile (20 / 42 passes in these traces, recorded through the C++ host);
    the frame must be the single-pass frame bit for bit."""
    replay, T, _ = libs
    a = replay.replay(T.parse(os.path.join(GOLDEN, tiled + ".rvct.xz"))).frames
    b = replay.replay(T.parse(os.path.join(GOLDEN, whole + ".rvct.xz"))).frames
    assert len(a) == len(b) == 1 and np.array_equal(a[0], b[0])


def test_committed_cuda_pngs_are_current(libs, tmp_path):
    """tests/golden/cuda_png/ holds CUDA-rendered frames as PNGs; the CPU suite runs the reference's
    own image_diff.py on them against the oracle (tests/test_image_diff_cpu.py). They must be exactly
    what the kernels render now (regenerate: tests/tools/gms_compare.py --write-cuda tests/golden/cuda_png)."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests", "tools"))
    import gms_compare as G
    committed = os.path.join(root, "tests", "golden", "cuda_png")
    names = sorted(n for n in os.listdir(committed) if n.endswith(".png"))
    assert len(names) >= 8
    G.write_set(str(tmp_path), sorted({n.split(".")[0] for n in names}), "cuda")
    for n in names:
        assert np.array_equal(G.read_png(os.path.join(committed, n)), G.read_png(str(tmp_path / n))), n


def test_cxx_host_band_mode_over_two_gpus(libs):
    """Band sharding from the C++ host, no Python on the data path: two rive_cuda_player
    processes (RenderContextCUDAImpl::ContextOptions{bandRank, bandCount}), one per GPU, render
    the bands of the same frame and rivecuda_band_gather (NCCL send / recv straight into the root
    target's rows) composites it; rank 0's frame must equal the single-process render bit for bit."""
    import subprocess
    import tempfile
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    player = os.path.join(root, "rive-runtime_b200", "_build", "rive_cuda_player")
    if not os.path.exists(player):
        pytest.skip("scene player not built (needs /root/reference at build time)")
    with tempfile.TemporaryDirectory() as tmp:
        scene = ["--scene", "c3", "--width", "1920", "--height", "1080", "--paths", "600"]
        single = os.path.join(tmp, "single.rgba")
        subprocess.run([player, *scene, "--out", single], check=True, capture_output=True, timeout=300)
        idfile = os.path.join(tmp, "nccl.id")
        procs = [subprocess.Popen([player, *scene, "--device", str(r), "--band-rank", str(r), "--band-count", "2", "--band-id-file", idfile,
                                   "--out", os.path.join(tmp, f"band{r}.rgba")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
                 for r in range(2)]
        for p in procs:
            out, err = p.communicate(timeout=300)
            assert p.returncode == 0, err.decode()[-2000:]
        want = np.fromfile(single, dtype=np.uint8)
        got = np.fromfile(os.path.join(tmp, "band0.rgba"), dtype=np.uint8)
        assert want.size == 1920 * 1080 * 4 and np.array_equal(want, got)


@pytest.mark.parametrize("scene,trace_name,size", [("gm:beziers", "beziers", (800, 400)), ("c1", "c1", (1600, 1600)),
                                                   ("img", "img", (960, 1280)), ("gm:feather_shapes", "feather_shapes", None)])
def test_reference_front_end_drives_the_cuda_backend(libs, scene, trace_name, size):
    """The drop-in boundary end to end, in C++: the reference's own RiveRenderer ->
    RenderContext (built in place, oracle/_ref) -> RenderContextCUDAImpl -> C ABI ->
    librivecuda.so on the GPU (host/player). Its pixels must equal the replay of the
    recorded ABI trace bit for bit (same calls, same kernels) -- i.e. the Python replayer
    the other tests use is a faithful stand-in for the C++ host layer."""
    import subprocess
    import tempfile
    replay, T, _ = libs
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    player = os.path.join(root, "rive-runtime_b200", "_build", "rive_cuda_player")
    if not os.path.exists(player):
        pytest.skip("scene player not built (needs the reference tree at build time)")
    got = replay.replay(T.parse(os.path.join(GOLDEN, trace_name + ".rvct.xz"))).frames[-1]
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "frame.rgba")
        env = dict(os.environ, RIVECUDA_LIB=os.path.join(root, "rive-runtime_b200", "_build", "librivecuda.so"))
        subprocess.check_call([player, "--scene", scene, "--out", out], env=env, stdout=subprocess.DEVNULL, timeout=300)
        px = np.fromfile(out, dtype=np.uint8)
    assert px.size == got.size
    assert np.array_equal(px.reshape(got.shape), got)


def test_tile_list_overflow_is_rerun_transparently(libs, monkeypatch):
    """A flush never waits for the size of its tile lists: it runs scatter / sort / raster
    against the buffer it has, the device checks that the lists fit, and the host re-runs
    those kernels with a larger buffer at its next synchronisation point if they did not.
    Forcing a tiny first buffer must not change a single pixel, for a clearing flush, for
    a preserving multi-flush sequence and for an animation."""
    replay, T, _ = libs
    for name in ("beziers", "preserverendertarget", "c1", "anim_juice"):
        recs = T.parse(os.path.join(GOLDEN, name + ".rvct.xz"))
        monkeypatch.delenv("RIVECUDA_INITIAL_TILE_ENTRIES", raising=False)
        want = replay.replay(recs).frames
        monkeypatch.setenv("RIVECUDA_INITIAL_TILE_ENTRIES", "7")
        got = replay.replay(recs).frames
        assert len(got) == len(want)
        assert all(np.array_equal(a, b) for a, b in zip(got, want)), name


@pytest.mark.parametrize("name", ["off_road_car", "bullet_man"])
def test_riv_file_through_the_unmodified_runtime(libs, name):
    """SURVEY 8 f3 -- the whole north-star call chain on the GPU, in C++: a real .riv file
    imported by the reference's unmodified core runtime (built in place), its state machine
    advanced at 1/60 s, Artboard::draw -> RiveRenderer -> RenderContext::flush ->
    RenderContextCUDAImpl -> librivecuda.so. The 60th frame must equal, bit for bit, the
    replay of the flush trace recorded from the same run (tests/golden/riv_*.rvct.xz), which the
    parity tests above compare with the oracle. The .riv assets are the reference's own test
    assets (tests/unit_tests/assets); they are not committed here, so the test skips without
    them (tools/fetch_riv_assets.sh copies them from /root/reference)."""
    import subprocess
    import tempfile
    replay, T, _ = libs
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    player = os.path.join(root, "rive-runtime_b200", "_build", "rive_cuda_player")
    asset = os.path.join(root, "tests", "_riv_assets", name + ".riv")
    if not os.path.exists(player) or not os.path.exists(asset):
        pytest.skip("scene player or .riv asset not present")
    want = replay.replay(T.parse(os.path.join(GOLDEN, f"riv_{name}.rvct.xz"))).frames[-1]
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "frame.rgba")
        env = dict(os.environ, RIVECUDA_LIB=os.path.join(root, "rive-runtime_b200", "_build", "librivecuda.so"))
        subprocess.check_call([player, "--scene", "riv:" + asset, "--frames", "60", "--out", out], env=env,
                              stdout=subprocess.DEVNULL, timeout=300)
        px = np.fromfile(out, dtype=np.uint8)
    assert px.size == want.size and np.array_equal(px.reshape(want.shape), want)
