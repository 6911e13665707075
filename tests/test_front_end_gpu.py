"""GPU path front end (rivecuda_front_end_paths) against the reference's own
front end: for the same RawPaths, the device-generated TessVertexSpan / ContourData / PathData
/ PaintData buffers -- Wang's-formula and polar segment counts, stroke chops (inflections,
180-degree turns, cusps), joins, emulated caps, the frame cull, prefix-summed span offsets,
vertex counts, row wraps, contour midpoints -- must equal, byte for byte, what
PathDraw::initForMidpointFan + pushMidpointFanTessellationData wrote into the mapped buffers
(recorded in the committed flush traces), and the rendered frame must be identical."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["f1", "f1w", "s1", "c2_4k", "trickycubicstrokes", "emptystroke", "strokes3", "OverStroke", "zero_control_stroke", "lots_of_tess_spans_stroke"])
def test_gpu_front_end_matches_reference_front_end(built, name):
    from rive_runtime_b200 import abi, front_end as F, replay as R, trace as T
    abi.load()
    recs = T.parse(os.path.join(GOLDEN, name + ".rvct.xz"))
    dump = F.load_paths(os.path.join(GOLDEN, name + ".paths.xz"))
    assert dump.complete, "the dump must cover every draw of the frame"
    want_frame = R.replay(recs).frames[-1]
    host = {r.fields["kind"]: r.data for r in recs if r.tag == T.BUFFER_UNMAP}
    fr = next(r.fields["flush"] for r in recs if r.tag == T.FLUSH)
    d = fr.desc

    with R.Replayer(0) as rp:
        result = R.ReplayResult()
        for r in recs:
            if r.tag in (T.CREATE, T.DESTROY, T.FLUSH, T.TARGET_READ, T.TARGET_DESTROY, T.POST_FLUSH):
                continue
            if r.tag == T.BUFFER_UNMAP and r.fields["kind"] in (1, 2, 3, 4, 6):
                continue  # path, paint, paintAux, contour, tessSpan: produced on the GPU below
            rp.apply(r, result)
        tc = next(r for r in recs if r.tag == T.TARGET_CREATE)
        res = F.run(rp, dump, tc.fields["width"], tc.fields["height"])
        # a1 / a3: counts and allocation
        assert res.path_count == d.path_count
        assert res.contour_count == d.contour_count
        assert res.tess_vertex_span_count == d.tess_vertex_span_count
        assert res.tess_data_height == d.tess_data_height
        assert (len(fr.batches) == 1 or name == "f1w") and all(b.draw_type == 0 for b in fr.batches)
        assert (res.first_patch, res.patch_count) == (fr.batches[0].base_element, sum(b.element_count for b in fr.batches))
        if name == "f1w":
            # the batch boundaries the reference chose are path boundaries of rivecuda_front_end_path_patches
            first_patch = np.zeros(dump.paths.size + 1, dtype=np.uint32)
            rp._call("rivecuda_front_end_path_patches", first_patch.ctypes.data, dump.paths.size)
            assert first_patch[0] == res.first_patch and first_patch[-1] == res.first_patch + res.patch_count
            assert np.all(np.diff(first_patch.astype(np.int64)) >= 0)
            assert set(b.base_element for b in fr.batches) <= set(first_patch.tolist())
        # a2: spans (segment counts, x0x1 / y offsets, reflections, wraps) and contours, byte for byte
        n = res.tess_vertex_span_count * 64
        got = F.read_buffer(rp, 6, n).view(np.uint32).reshape(-1, 16)
        want = np.frombuffer(host[6].tobytes()[:n], dtype=np.uint32).reshape(-1, 16)
        bad = np.nonzero((got != want).any(axis=1))[0]
        assert bad.size == 0, f"{bad.size} spans differ, first {bad[:5]}: got {got[bad[0]]} want {want[bad[0]]}"
        n = res.contour_count * 16
        assert np.array_equal(F.read_buffer(rp, 4, n), np.frombuffer(host[4].tobytes()[:n], dtype=np.uint8))
        # a4: path matrices / paint params + colours (defined fields of each record)
        n = res.path_count
        got_path = F.read_buffer(rp, 1, n * 64).view(np.uint32).reshape(-1, 16)
        want_path = np.frombuffer(host[1].tobytes()[:n * 64], dtype=np.uint32).reshape(-1, 16)
        assert np.array_equal(got_path[1:, :8], want_path[1:, :8])
        got_paint = F.read_buffer(rp, 2, n * 8).view(np.uint32).reshape(-1, 2)
        want_paint = np.frombuffer(host[2].tobytes()[:n * 8], dtype=np.uint32).reshape(-1, 2)
        assert np.array_equal(got_paint[1:], want_paint[1:])
        # ... and the frame rendered from the GPU-generated buffers is the same frame.
        pf = rp.prepare_flush(fr)
        rp.flush(pf)
        frame = rp.read_target(fr.target_id)
    assert np.array_equal(frame, want_frame)


@pytest.mark.parametrize("seed,n_paths", [(1, 2000), (2, 12000)])
def test_device_front_end_equals_its_host_build_on_random_paths(built, seed, n_paths):
    """Size-independent property: for arbitrary finite RawPaths the kernels and the host build of
    the same per-contour core (which the CPU suite pins against the reference) write the same
    bytes -- spans, contours, path and paint records, counts -- including the frame cull."""
    from oracle import front_end_host
    from rive_runtime_b200 import abi, front_end as F, replay as R
    from path_fuzz import prune_empty_segments, random_paths
    abi.load()
    dump, _ = prune_empty_segments(*random_paths(seed, n_paths))
    want = front_end_host.run(dump, 3840, 2160)
    with R.Replayer(0) as rp:
        res = F.run(rp, dump, 3840, 2160)
        for field in ("path_count", "contour_count", "tess_vertex_span_count", "midpoint_fan_tess_vertex_count",
                      "tess_data_height", "first_patch", "patch_count"):
            assert getattr(res, field) == getattr(want.result, field), field
        assert res.tess_vertex_span_count > n_paths and res.tess_data_height <= 2048
        n = res.tess_vertex_span_count
        got = F.read_buffer(rp, 6, n * 64).view(np.uint32).reshape(-1, 16)
        bad = np.nonzero((got != want.spans[:n]).any(axis=1))[0]
        assert bad.size == 0, f"{bad.size} spans differ, first {bad[:5]}: got {got[bad[0]]} want {want.spans[bad[0]]}"
        n = res.contour_count
        assert np.array_equal(F.read_buffer(rp, 4, n * 16).view(np.uint32).reshape(-1, 4), want.contours[:n])
        n = res.path_count
        assert np.array_equal(F.read_buffer(rp, 1, n * 64).view(np.uint32).reshape(-1, 16)[1:, :8], want.path_data[1:n, :8])
        assert np.array_equal(F.read_buffer(rp, 2, n * 8).view(np.uint32).reshape(-1, 2)[1:], want.paint_data[1:n])


@pytest.mark.parametrize("scene,golden", [("c2", "c2_4k"), ("f1", "f1"), ("f1o", "f1o"), ("f1b", "f1b"), ("f1c", "f1c"), ("f1w", "f1w"), ("f1g", "f1g"), ("f1p", "f1p"), ("f1i", "f1i"), ("s1", "s1"), ("gm:trickycubicstrokes", "trickycubicstrokes"),
                                          ("gm:strokes3", "strokes3")])
def test_cpp_path_renderer_draws_the_same_frame(built, scene, golden):
    """SURVEY 8 f1 in the compiled host: the scene player with --gpu-front-end draws through
    CudaPathRenderer (host/cuda_path_renderer.hpp: the RawPaths go to rivecuda_front_end_paths,
    no PathDraw / LogicalFlush on the CPU). The frame must equal, bit for bit, the replay of the
    flush trace the reference front end produced for the same scene."""
    import subprocess
    import tempfile
    from rive_runtime_b200 import abi, replay as R, trace as T
    abi.load()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    player = os.path.join(root, "rive-runtime_b200", "_build", "rive_cuda_player")
    if not os.path.exists(player):
        pytest.skip("scene player not built (needs the reference tree at build time)")
    want = R.replay(T.parse(os.path.join(GOLDEN, golden + ".rvct.xz"))).frames[-1]
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "frame.rgba")
        env = dict(os.environ, RIVECUDA_LIB=os.path.join(root, "rive-runtime_b200", "_build", "librivecuda.so"))
        subprocess.check_call([player, "--scene", scene, "--gpu-front-end", "--budget-ms", "0", "--out", out], env=env,
                              stdout=subprocess.DEVNULL, timeout=300)
        px = np.fromfile(out, dtype=np.uint8)
    assert px.size == want.size and np.array_equal(px.reshape(want.shape), want)


@pytest.mark.parametrize("name", ["off_road_car", "bullet_man", "shapetest"])
def test_riv_file_through_the_device_front_end(built, name):
    """Real .riv content (clockwise and nonZero fills, strokes, opacity, artboard clip rectangles;
    off_road_car and bullet_man: gradients and nested clip paths too) through --gpu-front-end: frame 20 must equal, bit for bit, the frame the reference's CPU front
    end produces for the same file through the same backend (midpoint fans only: --budget-ms 0
    switches the reference's interior triangulation of large paths off, which the device front end
    does not implement). The player's `rivs:DIR` scene runs the same comparison over a whole
    directory in one process: 321 of the reference's 342 assets are drawn by the device front end,
    all 321 identical (profiles/r02_riv_front_end_sweep.md).
    The assets are the reference's (tools/fetch_riv_assets.sh); the test skips without them."""
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    player = os.path.join(root, "rive-runtime_b200", "_build", "rive_cuda_player")
    asset = os.path.join(root, "tests", "_riv_assets", name + ".riv")
    if not os.path.exists(player) or not os.path.exists(asset):
        pytest.skip("scene player or .riv asset not present")
    env = dict(os.environ, RIVECUDA_LIB=os.path.join(root, "rive-runtime_b200", "_build", "librivecuda.so"))
    frames = []
    with tempfile.TemporaryDirectory() as tmp:
        for extra in ([], ["--gpu-front-end"]):
            out = os.path.join(tmp, "frame%d.rgba" % len(frames))
            subprocess.check_call([player, "--scene", "riv:" + asset, "--frames", "20", "--budget-ms", "0", "--out", out, *extra], env=env,
                                  stdout=subprocess.DEVNULL, timeout=300)
            frames.append(np.fromfile(out, dtype=np.uint8))
    assert frames[0].size == 1920 * 1080 * 4 and np.array_equal(frames[0], frames[1])
    assert len(np.unique(frames[0].reshape(-1, 4), axis=0)) > 1  # not an empty frame


def test_cpp_path_renderer_draws_the_image_scene(built):
    """The `img` scene -- a gradient background, 24 drawImage calls (three images, every wrap and
    filter, blend modes, under clip rectangles and an oval clip path) and four warped image MESHES,
    which CudaPathRenderer passes through as batches of their own between the paths' -- drawn through
    both front ends (midpoint fans only: --budget-ms 0): the frames must be identical."""
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    player = os.path.join(root, "rive-runtime_b200", "_build", "rive_cuda_player")
    if not os.path.exists(player):
        pytest.skip("scene player not built (needs the reference tree at build time)")
    env = dict(os.environ, RIVECUDA_LIB=os.path.join(root, "rive-runtime_b200", "_build", "librivecuda.so"))
    frames = []
    with tempfile.TemporaryDirectory() as tmp:
        for extra in ([], ["--gpu-front-end"]):
            out = os.path.join(tmp, "frame%d.rgba" % len(frames))
            subprocess.check_call([player, "--scene", "img", "--budget-ms", "0", "--out", out, *extra], env=env,
                                  stdout=subprocess.DEVNULL, timeout=300)
            frames.append(np.fromfile(out, dtype=np.uint8))
    assert frames[0].size > 0 and np.array_equal(frames[0], frames[1])
    assert len(np.unique(frames[0].reshape(-1, 4), axis=0)) > 1000


def test_riv_assets_sweep_both_front_ends_in_one_process(built):
    """`--scene rivs:DIR`: every asset is imported, advanced 20 frames and drawn through RiveRenderer and
    through CudaPathRenderer on the SAME RenderContextCUDAImpl, alternating -- which also checks that the
    two front ends can share a context (the plain path only grows the shared rings and hands the
    gradient texture back at the height the RenderContext allocated)."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    player = os.path.join(root, "rive-runtime_b200", "_build", "rive_cuda_player")
    assets = os.path.join(root, "tests", "_riv_assets")
    if not os.path.exists(player) or not os.path.isdir(assets) or len([n for n in os.listdir(assets) if n.endswith(".riv")]) < 10:
        pytest.skip("scene player or .riv assets not present")
    env = dict(os.environ, RIVECUDA_LIB=os.path.join(root, "rive-runtime_b200", "_build", "librivecuda.so"))
    out = subprocess.run([player, "--scene", "rivs:" + assets, "--frames", "20", "--budget-ms", "0"], env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, timeout=600)
    report = json.loads(out.stdout.decode().strip().splitlines()[-1])
    assert out.returncode == 0 and report["differing"] == 0 and report["failed"] == 0 and report["refused"] == 0
    assert report["identical"] == report["assets"] >= 10


def test_large_fills_can_be_delegated_to_the_reference_triangulator(built):
    """With the reference's own (deterministic) triangulation thresholds, fills of 512 x 512 px and more
    are interior-triangulated by the reference front end; the device front end draws midpoint fans,
    which is the same shape to within the tessellation tolerance but not the same pixels (11 of these
    15 assets differ at frame 45). `--delegate-large-fills` hands exactly those fills to the reference
    front end (CudaPathRenderer::setLargeFillDelegation): every asset within 2/255 of the reference's
    frame again (the frame is cut into flushes: see the feather test)."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    player = os.path.join(root, "rive-runtime_b200", "_build", "rive_cuda_player")
    assets = os.path.join(root, "tests", "_riv_assets")
    if not os.path.exists(player) or not os.path.isdir(assets) or len([n for n in os.listdir(assets) if n.endswith(".riv")]) < 10:
        pytest.skip("scene player or .riv assets not present")
    env = dict(os.environ, RIVECUDA_LIB=os.path.join(root, "rive-runtime_b200", "_build", "librivecuda.so"), RIVECUDA_SWEEP_TOLERANCE="2")
    out = subprocess.run([player, "--scene", "rivs:" + assets, "--frames", "45", "--delegate-large-fills"], env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, timeout=600)
    report = json.loads(out.stdout.decode().strip().splitlines()[-1])
    assert out.returncode == 0 and report["differing"] == 0 and report["failed"] == 0 and report["refused"] == 0, out.stdout.decode()[-1500:]
    assert report["identical"] + report["within_tolerance"] == report["assets"] >= 10


def test_riv_assets_with_feathers_are_drawn_by_delegation(built):
    """Feathers are the one thing the device front end does not tessellate: CudaPathRenderer flushes
    what it has, hands the feather draws to the reference's own front end on the same context and
    target (a flush of their own), and goes on. The frame is then cut into several flushes -- up to
    ~85 per frame in hunter_x_demo.riv, many of them empty: the run of empty flushes is what exposed
    that ring-slot reuse was only paced implicitly -- each of which picks its rasteriser and packs its
    own feather atlas, so the result may differ from the reference's single flush by an LSB or two:
    every asset within 2/255, none refused."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    player = os.path.join(root, "rive-runtime_b200", "_build", "rive_cuda_player")
    assets = os.path.join(root, "tests", "_riv_assets", "feathers")
    if not os.path.exists(player) or not os.path.isdir(assets) or len(os.listdir(assets)) < 4:
        pytest.skip("scene player or .riv assets not present")
    env = dict(os.environ, RIVECUDA_LIB=os.path.join(root, "rive-runtime_b200", "_build", "librivecuda.so"), RIVECUDA_SWEEP_TOLERANCE="2")
    for _ in range(2):  # (the slot-reuse race was intermittent)
        out = subprocess.run([player, "--scene", "rivs:" + assets, "--frames", "20", "--budget-ms", "0"], env=env, stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, timeout=600)
        report = json.loads(out.stdout.decode().strip().splitlines()[-1])
        assert out.returncode == 0 and report["differing"] == 0 and report["failed"] == 0 and report["refused"] == 0, out.stdout.decode()[-1500:]
        assert report["identical"] + report["within_tolerance"] == report["assets"] >= 4


def test_front_end_refuses_what_one_flush_cannot_hold(built):
    """Error behaviour: more paths / contours / tessellation vertices than one logical flush admits
    (RenderContext::LogicalFlush::pushDraws, render_context.cpp:528-536) is an error with a
    message, not a truncated frame; bad arguments likewise."""
    import ctypes
    from path_fuzz import random_paths
    from rive_runtime_b200 import abi, front_end as F, replay as R
    lib = abi.load()
    dump, _ = random_paths(5, 200)
    big = F.PathDump(np.tile(dump.paths, 200), dump.verbs, dump.points, True)  # 40 000 paths > 30 720 path ids
    with R.Replayer(0) as rp:
        with pytest.raises(RuntimeError, match="exceed one flush"):
            F.run(rp, big, 3840, 2160)
        res = F.FrontEndResult()
        assert lib.rivecuda_front_end_paths(rp.ctx, None, 1, None, 1, None, 1, 0, 0, ctypes.byref(res)) != 0
        assert b"bad arguments" in lib.rivecuda_last_error()
        # Verbs that need more points than the caller passed: caught on the device, nothing is read.
        short = F.PathDump(dump.paths, dump.verbs, dump.points[:-1], True)
        with pytest.raises(RuntimeError, match="more points than the point array holds"):
            F.run(rp, short, 3840, 2160)
        outside = dump.paths.copy()
        outside["first_verb"][3] = dump.verbs.size
        with pytest.raises(RuntimeError, match="outside the arrays"):
            F.run(rp, F.PathDump(outside, dump.verbs, dump.points, True), 3840, 2160)
        # ... and the context is still usable afterwards.
        ok = F.run(rp, dump, 3840, 2160)
        assert ok.path_count > 1


def test_cpp_path_renderer_splits_frames_that_exceed_one_flush(built):
    """40 000 paths (more path ids and 2.5x more tessellation vertices than one logical flush
    admits): rivecuda_front_end_paths answers RIVECUDA_STATUS_EXCEEDS_FLUSH, the C++ host halves
    the chunk until it fits and draws the frame in several flushes. The result must equal the
    frame the reference front end (which splits on its own) produces through the same backend."""
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    player = os.path.join(root, "rive-runtime_b200", "_build", "rive_cuda_player")
    if not os.path.exists(player):
        pytest.skip("scene player not built (needs the reference tree at build time)")
    env = dict(os.environ, RIVECUDA_LIB=os.path.join(root, "rive-runtime_b200", "_build", "librivecuda.so"))
    frames = []
    with tempfile.TemporaryDirectory() as tmp:
        for extra in ([], ["--gpu-front-end"]):
            out = os.path.join(tmp, "frame%d.rgba" % len(frames))
            subprocess.check_call([player, "--scene", "c2", "--paths", "40000", "--budget-ms", "0", "--out", out, *extra], env=env,
                                  stdout=subprocess.DEVNULL, timeout=300)
            frames.append(np.fromfile(out, dtype=np.uint8))
    assert frames[0].size == 3840 * 2160 * 4 and np.array_equal(frames[0], frames[1])
