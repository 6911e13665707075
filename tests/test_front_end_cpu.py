"""The front end's per-curve arithmetic, pinned against the reference's own output: every
TessVertexSpan in the committed traces carries the curve's control points and the parametric
segment count PathDraw::initForMidpointFan computed for it; the numpy restatement of Wang's
formula (oracle/front_end_ref.py) must reproduce all of them bit for bit. Also checks the
`--dump-paths` reader against the traces' path / contour counts."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import front_end_host, front_end_ref
from rive_runtime_b200 import front_end as F, trace as T


def spans_and_matrices(name):
    recs = T.parse(os.path.join(GOLDEN, name))
    bufs = {}
    out = []
    for r in recs:
        if r.tag == T.BUFFER_UNMAP:
            bufs[r.fields["kind"]] = r.data
        elif r.tag == T.FLUSH:
            d = r.fields["flush"].desc
            if d.tess_vertex_span_count == 0:
                continue
            spans = np.frombuffer(bufs[6].tobytes(), dtype=np.uint32).reshape(-1, 16)[
                d.first_tess_vertex_span:d.first_tess_vertex_span + d.tess_vertex_span_count]
            contours = np.frombuffer(bufs[4].tobytes(), dtype=np.uint32).reshape(-1, 4)[d.first_contour:]
            paths = np.frombuffer(bufs[1].tobytes(), dtype=np.float32).reshape(-1, 16)[d.first_path:]
            # Outer-curve patches (interior triangulation) always have 16 segments; their spans
            # live behind the midpoint-fan region of the tessellation buffer.
            outer = [b.base_element * 17 for b in r.fields["flush"].batches if b.draw_type == 2]
            out.append((spans, contours, paths, min(outer) if outer else 1 << 30))
    return out


@pytest.mark.parametrize("name", ["c2_4k.rvct.xz", "f1.rvct.xz", "c1.rvct.xz", "beziers.rvct.xz", "strokes_round.rvct.xz",
                                  "cubicpath.rvct.xz", "anim_db_health_tracker.rvct.xz"])
def test_wang_segment_counts_match_the_reference(name):
    checked = 0
    for spans, contours, paths, outer_start in spans_and_matrices(name):
        contour_id = spans[:, 15] & 0xffff
        x0 = ((spans[:, 12].astype(np.int64) & 0xffff) ^ 0x8000) - 0x8000
        location = spans[:, 10].view(np.float32).astype(np.int64) * 2048 + x0
        real = spans[(contour_id > 0) & (location < outer_start)]
        cid = (real[:, 15] & 0xffff).astype(np.int64) - 1
        path_id = contours[cid, 2] & 0xffff
        rec = paths[path_id]
        matrix = rec[:, :6]
        feather = rec[:, 7]
        retrofit = (real[:, 15] >> 31) != 0  # retrofitted triangle strips carry no curve
        keep = (feather == 0) & ~retrofit
        pts = real[keep][:, :8].view(np.float32).reshape(-1, 4, 2)
        want = real[keep][:, 14] & 0x3ff
        # Lines are stored as the cubic convert_line_to_cubic() makes and always get 1 segment;
        # zero-length "curves" carrying only a cap/join get 0. Everything else is Wang's formula.
        got = front_end_ref.wang_cubic_segments(pts, matrix[keep])
        is_curve = want > 1
        assert np.array_equal(got[is_curve], want[is_curve]), name
        # where the reference says 0 or 1 the formula may only say 1 (lines) -- never more for
        # a real cubic that the reference counted as 1
        assert np.all(got[~is_curve & (want == 1)] >= 1)
        checked += int(is_curve.sum())
    assert checked > 0


@pytest.mark.parametrize("name", ["c2_4k", "f1"])
def test_path_dump_reader(name):
    dump = F.load_paths(os.path.join(GOLDEN, name + ".paths.xz"))
    recs = T.parse(os.path.join(GOLDEN, name + ".rvct.xz"))
    d = next(r.fields["flush"].desc for r in recs if r.tag == T.FLUSH)
    assert dump.complete
    assert dump.paths.size == d.path_count - 1           # path 0 is the flush's reserved record
    assert int((dump.verbs == 0).sum()) == d.contour_count  # one contour per moveTo
    assert int(dump.paths["verb_count"].sum()) == dump.verbs.size
    assert dump.points.shape[0] >= int((dump.verbs == 4).sum()) * 3


@pytest.mark.parametrize("name", ["beziers.rvct.xz", "strokes_round.rvct.xz", "c1.rvct.xz", "strokes_poly.rvct.xz",
                                  "trickycubicstrokes.rvct.xz", "c3.rvct.xz", "roundjoinstrokes.rvct.xz",
                                  "anim_db_health_tracker.rvct.xz"])
def test_polar_segment_counts_match_the_reference(name):
    """Stroked curves: the polar segment count in every span equals the restatement of
    fast_acos(rotation between end tangents) * calc_polar_segments_per_radian<8>(radius * maxScale)."""
    checked = 0
    for spans, contours, paths, _ in spans_and_matrices(name):
        real = spans[(spans[:, 15] & 0xffff) > 0]
        cid = (real[:, 15] & 0xffff).astype(np.int64) - 1
        rec = paths[contours[cid, 2] & 0xffff]
        stroke, feather = rec[:, 6], rec[:, 7]
        keep = (stroke != 0) & (feather == 0) & ((real[:, 14] & 0x3ff) > 0)  # real curves, not cap/join carriers
        pts = real[keep][:, :8].view(np.float32).reshape(-1, 4, 2)
        want = (real[keep][:, 14] >> 10) & 0x3ff
        got = front_end_ref.polar_segments(pts, rec[keep][:, :6], stroke[keep])
        assert np.array_equal(got, want), name
        checked += int(keep.sum())
    assert checked > 0


FRONT_END_SCENES = ["f1", "f1w", "s1", "c2_4k", "trickycubicstrokes", "trickycubicstrokes_roundcaps", "emptystroke", "strokes3",
                    "labyrinth_round", "labyrinth_square", "zero_control_stroke", "zerolinestroke", "OverStroke",
                    "bevel180strokes", "roundjoinstrokes", "widebuttcaps", "beziers", "CubicStroke", "inner_join_geometry",
                    "teenyStrokes", "quadcap", "strokefill", "zeroPath", "lots_of_tess_spans_stroke"]


@pytest.mark.parametrize("name", FRONT_END_SCENES)
def test_front_end_core_matches_the_reference_front_end(name):
    """The per-contour core the F1 kernels run (csrc/front_end_core.h), built for the host: for
    the RawPaths of a frame it must produce, byte for byte, the TessVertexSpans / ContourData /
    PathData / PaintData the reference front end wrote for that frame -- stroke chops at
    inflections, 180-degree turns and cusps, round / miter / bevel joins, emulated caps, empty
    contours, the frame cull, padding, row wraps. tests/test_front_end_gpu.py checks the device
    build of the same code the same way."""
    recs = T.parse(os.path.join(GOLDEN, name + ".rvct.xz"))
    dump = F.load_paths(os.path.join(GOLDEN, name + ".paths.xz"))
    assert dump.complete
    host = {r.fields["kind"]: r.data for r in recs if r.tag == T.BUFFER_UNMAP}
    flushes = [r.fields["flush"] for r in recs if r.tag == T.FLUSH]
    assert len(flushes) == 1
    d = flushes[0].desc
    tc = next(r for r in recs if r.tag == T.TARGET_CREATE)
    out = front_end_host.run(dump, tc.fields["width"], tc.fields["height"])
    res = out.result
    assert (res.path_count, res.contour_count, res.tess_vertex_span_count, res.tess_data_height) == (
        d.path_count, d.contour_count, d.tess_vertex_span_count, d.tess_data_height)
    # One midpointFanPatches batch -- or, when clockwise fills alternate with nonZero / evenOdd ones
    # (f1w), contiguous batches whose ShaderMiscFlags::clockwiseFill alternates.
    batches = flushes[0].batches
    assert all(b.draw_type == 0 for b in batches) and (len(batches) == 1 or name == "f1w")
    assert all(a.base_element + a.element_count == b.base_element and a.shader_misc_flags != b.shader_misc_flags
               for a, b in zip(batches, batches[1:]))
    assert (res.first_patch, res.patch_count) == (batches[0].base_element, sum(b.element_count for b in batches))
    n = res.tess_vertex_span_count
    want = np.frombuffer(host[6].tobytes()[:n * 64], dtype=np.uint32).reshape(-1, 16)
    bad = np.nonzero((out.spans[:n] != want).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} spans differ, first {bad[:5]}"
    if name == "f1w":
        # clockwise fills under a left-handed matrix: forward copy first, coverage negated
        # (ContourDirections::forwardThenReverse + NEGATE_PATH_FILL_COVERAGE_FLAG, draw.cpp:657-680)
        negated = (want[:, 15] & (1 << 24)) != 0
        assert 0 < negated.sum() < n and len(batches) > 50
    want = np.frombuffer(host[4].tobytes()[:res.contour_count * 16], dtype=np.uint32).reshape(-1, 4)
    assert np.array_equal(out.contours[:res.contour_count], want)
    n = res.path_count
    want = np.frombuffer(host[1].tobytes()[:n * 64], dtype=np.uint32).reshape(-1, 16)
    assert np.array_equal(out.path_data[1:n, :8], want[1:, :8])
    want = np.frombuffer(host[2].tobytes()[:n * 8], dtype=np.uint32).reshape(-1, 2)
    assert np.array_equal(out.paint_data[1:n], want[1:])


def test_stroke_scenes_exercise_every_stroke_feature():
    """The stress scene is only a test if it reaches the hard cases."""
    recs = T.parse(os.path.join(GOLDEN, "s1.rvct.xz"))
    host = {r.fields["kind"]: r.data for r in recs if r.tag == T.BUFFER_UNMAP}
    d = next(r.fields["flush"].desc for r in recs if r.tag == T.FLUSH)
    w = np.frombuffer(host[6].tobytes()[:d.tess_vertex_span_count * 64], dtype=np.uint32).reshape(-1, 16)
    seg = w[:, 14]
    join, polar, parametric = seg >> 20, (seg >> 10) & 1023, seg & 1023
    flags = w[:, 15] >> 26 & 7
    assert set(flags.tolist()) >= {2, 3, 4, 5}                      # round, bevel, miter-revert, miter-clip (square caps)
    assert int(((w[:, 15] >> 25) & 1).sum()) > 1000                 # emulated caps
    assert int(((parametric == 0) & (polar == 0)).sum()) > 1000     # cap-only spans
    pivot = (w[:, 0] == w[:, 6]) & (w[:, 1] == w[:, 7]) & (w[:, 2] == w[:, 4]) & (w[:, 3] == w[:, 5]) & (
        (w[:, 0] != w[:, 2]) | (w[:, 1] != w[:, 3]))
    assert int(pivot.sum()) > 100                                   # cusp pivots (chop_cubic_around_cusps)
    assert int(((w[:, 12] >> 16).astype(np.int32) > 2048).sum()) > 100  # spans wrapping a 2048-texel row
    assert polar.max() > 32 and join.max() > 32


@pytest.mark.parametrize("scene,golden", [("s1", "s1"), ("f1", "f1"), ("f1w", "f1w"), ("gm:trickycubicstrokes", "trickycubicstrokes"),
                                          ("gm:strokes3", "strokes3")])
def test_cpp_path_renderer_hands_over_what_the_reference_front_end_saw(scene, golden, tmp_path):
    """host/cuda_path_renderer.hpp (the rive::Renderer that feeds the device front end from C++),
    run on the call recorder: the RawPaths, matrices, paints and the two per-stroke scalars it
    passes to rivecuda_front_end_paths must equal the --dump-paths record of the same scene (whose
    scalars come from front_end.stroke_scalars, the Python mirror) -- and therefore produce the
    reference's buffers through the front-end core, which the test above checks."""
    import subprocess
    from conftest import ROOT
    player = os.path.join(ROOT, "rive-runtime_b200", "_build", "rive_cuda_player")
    recorder = os.path.join(ROOT, "rive-runtime_b200", "_build", "librivecuda_trace.so")
    if not os.path.exists(player) or not os.path.exists(recorder):
        pytest.skip("scene player not built (needs the reference tree at build time)")
    out = str(tmp_path / "call.rpf")
    env = dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=str(tmp_path / "unused.rvct"), RIVECUDA_TRACE_FRONT_END_OUT=out)
    subprocess.check_call([player, "--scene", scene, "--gpu-front-end", "--budget-ms", "0"], env=env, stdout=subprocess.DEVNULL, timeout=120)
    got, width, height = F.load_front_end_call(out)
    want = F.load_paths(os.path.join(GOLDEN, golden + ".paths.xz"))
    recs = T.parse(os.path.join(GOLDEN, golden + ".rvct.xz"))
    tc = next(r for r in recs if r.tag == T.TARGET_CREATE)
    assert (width, height) == (tc.fields["width"], tc.fields["height"])
    assert np.array_equal(got.verbs, want.verbs)
    assert np.array_equal(got.points.view(np.uint32), want.points.view(np.uint32))
    assert got.paths.size == want.paths.size
    for field in want.paths.dtype.names:
        a, b = got.paths[field], want.paths[field]
        if a.dtype.kind == "f":
            a, b = a.view(np.uint32), b.view(np.uint32)
        assert np.array_equal(a, b), field


def _compare_device_front_end_call_with_reference_trace(call, trace, recs):
    """call / trace: what the call recorder wrote for a --gpu-front-end run; recs: the flush trace of
    the reference front end for the same frame. Returns how many paints of each kind were compared."""
    dump, width, height, tables = F.load_front_end_call(call, with_tables=True)
    out = front_end_host.run(dump, width, height, tables)
    res = out.result

    empty = np.zeros(0, np.uint8)
    host = {kind: empty for kind in range(9)}  # (an empty frame unmaps nothing)
    host.update({r.fields["kind"]: r.data for r in recs if r.tag == T.BUFFER_UNMAP})
    flush = [r.fields["flush"] for r in recs if r.tag == T.FLUSH][-1]
    d = flush.desc
    assert (res.path_count, res.contour_count, res.tess_vertex_span_count, res.tess_data_height) == (
        d.path_count, d.contour_count, d.tess_vertex_span_count, d.tess_data_height)
    assert res.patch_count == sum(b.element_count for b in flush.batches if b.draw_type == 0)
    n = res.tess_vertex_span_count
    # (a frame of several logical flushes shares its buffers: the flush's records start at first_*)
    sl = lambda kind, first, count, size: host[kind].tobytes()[first * size:(first + count) * size]
    want = np.frombuffer(sl(6, d.first_tess_vertex_span, n, 64), dtype=np.uint32).reshape(-1, 16)
    assert np.array_equal(out.spans[:n], want)
    want = np.frombuffer(sl(4, d.first_contour, res.contour_count, 16), dtype=np.uint32).reshape(-1, 4)
    assert np.array_equal(out.contours[:res.contour_count], want)
    n = res.path_count
    want = np.frombuffer(sl(1, d.first_path, n, 64), dtype=np.uint32).reshape(-1, 16)
    assert np.array_equal(out.path_data[1:n, :8], want[1:, :8])
    paint = np.frombuffer(sl(2, d.first_paint, n, 8), dtype=np.uint32).reshape(-1, 2)
    assert np.array_equal(out.paint_data[1:n], paint[1:])
    # PaintAuxData: the words a paint of each kind defines (the reference leaves the others unwritten)
    aux = np.frombuffer(sl(3, d.first_paint_aux, n, 128), dtype=np.uint32).reshape(-1, 32)
    kind, flags = paint[:, 0] & 0xf, paint[:, 0]
    gradient, clipped, image = np.isin(kind, (2, 3)), (flags & 0x400) != 0, (flags & 0x800) != 0
    gradient[0] = clipped[0] = image[0] = False
    assert np.array_equal(out.paint_aux[:n][gradient][:, 0:8], aux[gradient][:, 0:8])
    assert np.array_equal(out.paint_aux[:n][clipped][:, 8:16], aux[clipped][:, 8:16])
    assert np.array_equal(out.paint_aux[:n][image][:, 16:23], aux[image][:, 16:23])
    counts = {"gradient": int(gradient.sum()), "clip_rect": int(clipped.sum()), "image": int(image.sum()), "clip_update": int((kind[1:] == 0).sum()),
              "clipped_by_path": int(((flags[1:] >> 16) != 0).sum()), "paths": int(n - 1)}
    # what CudaPathRenderer wrote itself: the colour-ramp spans and the flush's gradient counts
    device_recs = T.parse(trace)
    device_host = {kind: empty for kind in range(9)}
    device_host.update({r.fields["kind"]: r.data for r in device_recs if r.tag == T.BUFFER_UNMAP})
    device_flush = [r.fields["flush"] for r in device_recs if r.tag == T.FLUSH][-1]
    dd = device_flush.desc
    assert (dd.grad_span_count, dd.grad_data_height) == (d.grad_span_count, d.grad_data_height)
    if d.grad_span_count:
        first = d.first_grad_span * 16
        assert device_host[5].tobytes()[:d.grad_span_count * 16] == host[5].tobytes()[first:first + d.grad_span_count * 16]
        resize = lambda rs: [r.fields["height"] for r in rs if r.tag == T.RESIZE_GRADIENT][-1]
        assert resize(device_recs) == resize(recs)  # the allocated height the paints are normalised by
    return counts


@pytest.mark.parametrize("scene", ["f1o", "f1b", "f1c", "f1w", "f1g", "f1p", "f1i"])
def test_cpp_path_renderer_and_core_reproduce_the_reference_records(scene, tmp_path):
    """The device front end's round-2 features -- modulated opacity, blend modes, clip rectangles,
    clockwise fills, gradients, clip PATHS, image paints -- pinned on the CPU: the scene is drawn
    through CudaPathRenderer on the call recorder; what it hands to rivecuda_front_end_paths (paths,
    incl. the clipUpdate paths it inserts, and the clip-rectangle / gradient / image tables) goes
    through the host build of the kernels' core; the spans, contours, path, paint and paint-aux
    records must equal, byte for byte, what the reference's own front end wrote for the same scene
    (tests/golden/<scene>.rvct.xz) -- clip IDs, ramp rows, paint matrices, image LODs and all -- and
    so must the GradientSpans CudaPathRenderer wrote and the rows they fill."""
    import subprocess
    from conftest import ROOT
    player = os.path.join(ROOT, "rive-runtime_b200", "_build", "rive_cuda_player")
    recorder = os.path.join(ROOT, "rive-runtime_b200", "_build", "librivecuda_trace.so")
    if not os.path.exists(player) or not os.path.exists(recorder):
        pytest.skip("scene player not built (needs the reference tree at build time)")
    call, trace = str(tmp_path / "call.rpf"), str(tmp_path / "device.rvct")
    env = dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=trace, RIVECUDA_TRACE_FRONT_END_OUT=call)
    subprocess.check_call([player, "--scene", scene, "--gpu-front-end", "--budget-ms", "0"], env=env, stdout=subprocess.DEVNULL, timeout=120)
    recs = T.parse(os.path.join(GOLDEN, scene + ".rvct.xz"))
    counts = _compare_device_front_end_call_with_reference_trace(call, trace, recs)
    expect = {"f1c": "clip_rect", "f1g": "gradient", "f1i": "image", "f1p": "clip_update"}.get(scene)
    if expect is not None:
        assert counts[expect] > 50  # the scene does exercise its feature


@pytest.mark.parametrize("seed", [3, 7, 21])
@pytest.mark.parametrize("scene", ["f1o", "f1b", "f1c", "f1w", "f1g", "f1p", "f1i"])
def test_feature_scenes_with_other_seeds_live_against_the_reference(scene, seed, tmp_path):
    """The feature scenes again with other random content (2000 paths), both front ends recorded on the
    spot: the reference front end and CudaPathRenderer + the host build of the core write the same bytes.
    (tests/tools/gm_records_sweep.py with RIVECUDA_SWEEP_EXTRA="--seed N" runs any number of seeds.)"""
    import subprocess
    from conftest import ROOT
    player = os.path.join(ROOT, "rive-runtime_b200", "_build", "rive_cuda_player")
    recorder = os.path.join(ROOT, "rive-runtime_b200", "_build", "librivecuda_trace.so")
    if not os.path.exists(player) or not os.path.exists(recorder):
        pytest.skip("scene player not built (needs the reference tree at build time)")
    reference, call, trace = str(tmp_path / "reference.rvct"), str(tmp_path / "call.rpf"), str(tmp_path / "device.rvct")
    common = [player, "--scene", scene, "--budget-ms", "0", "--seed", str(seed), "--paths", "2000"]
    subprocess.check_call(common, env=dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=reference), stdout=subprocess.DEVNULL, timeout=120)
    subprocess.check_call(common + ["--gpu-front-end"], env=dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=trace, RIVECUDA_TRACE_FRONT_END_OUT=call),
                          stdout=subprocess.DEVNULL, timeout=120)
    counts = _compare_device_front_end_call_with_reference_trace(call, trace, T.parse(reference))
    assert counts["paths"] > 1000


RIV_ASSETS = ["off_road_car", "bullet_man", "juice", "shapetest", "fix_rectangle", "follow_path_solos", "nested_artboard_opacity", "lock_icon_demo",
              "solos_collapse_tests", "group_effect", "tape", "image_fit_alignment_2"]


@pytest.mark.parametrize("name", RIV_ASSETS)
def test_riv_asset_records_of_both_front_ends_match(name, tmp_path):
    """The same comparison on real content, without a GPU: frame 20 of a .riv asset (the reference's
    own test assets, imported and animated by its unmodified core runtime) is drawn on the call
    recorder through the reference front end (midpoint fans only: --budget-ms 0) and through
    CudaPathRenderer + the host build of the kernels' core; the records must be the same bytes.
    off_road_car and bullet_man bring gradients and nested clip paths, juice gradients, tape and
    image_fit_alignment_2 image meshes between the paths. Skips without the assets
    (tools/fetch_riv_assets.sh) or the player."""
    import subprocess
    from conftest import ROOT
    player = os.path.join(ROOT, "rive-runtime_b200", "_build", "rive_cuda_player")
    recorder = os.path.join(ROOT, "rive-runtime_b200", "_build", "librivecuda_trace.so")
    asset = os.path.join(ROOT, "tests", "_riv_assets", name + ".riv")
    if not os.path.exists(player) or not os.path.exists(recorder) or not os.path.exists(asset):
        pytest.skip("scene player or .riv asset not present")
    reference, call, trace = str(tmp_path / "reference.rvct"), str(tmp_path / "call.rpf"), str(tmp_path / "device.rvct")
    common = [player, "--scene", "riv:" + asset, "--frames", "20", "--budget-ms", "0"]
    subprocess.check_call(common, env=dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=reference), stdout=subprocess.DEVNULL, timeout=120)
    subprocess.check_call(common + ["--gpu-front-end"], env=dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=trace, RIVECUDA_TRACE_FRONT_END_OUT=call),
                          stdout=subprocess.DEVNULL, timeout=120)
    counts = _compare_device_front_end_call_with_reference_trace(call, trace, T.parse(reference))
    assert counts["paths"] > 0
    if name in ("off_road_car", "bullet_man", "juice"):
        assert counts["gradient"] > 0
    if name in ("off_road_car", "bullet_man"):
        assert counts["clip_update"] > 0 and counts["clipped_by_path"] > 0


def _assert_last_flushes_identical(a, b):
    """a, b: parsed traces; their last flushes -- descriptor, batches, atlas batches and every mapped
    buffer -- must be the same bytes."""
    import ctypes
    fa = [r.fields["flush"] for r in a if r.tag == T.FLUSH][-1]
    fb = [r.fields["flush"] for r in b if r.tag == T.FLUSH][-1]
    raw = lambda st: bytes(ctypes.string_at(ctypes.addressof(st), ctypes.sizeof(st)))
    da, db = fa.desc, fb.desc
    for field, _ in type(da)._fields_:
        if field in ("render_target",):
            continue
        va, vb = getattr(da, field), getattr(db, field)
        va, vb = (list(va), list(vb)) if hasattr(va, "__len__") else (va, vb)
        assert va == vb, field
    assert len(fa.batches) == len(fb.batches) and all(raw(x) == raw(y) for x, y in zip(fa.batches, fb.batches))
    assert [raw(x) for x in fa.atlas_fills] == [raw(x) for x in fb.atlas_fills]
    assert [raw(x) for x in fa.atlas_strokes] == [raw(x) for x in fb.atlas_strokes]
    ha = {r.fields["kind"]: r.data for r in a if r.tag == T.BUFFER_UNMAP}
    hb = {r.fields["kind"]: r.data for r in b if r.tag == T.BUFFER_UNMAP}
    assert ha.keys() == hb.keys() and da.path_count > 1
    n = da.path_count
    for kind in ha:
        if kind == 3:
            continue  # PaintAuxData: the reference leaves the words a paint does not define unwritten
        if kind == 1:
            # PathData: matrix, stroke radius, feather radius, z index always; the feather-atlas transform
            # only where there is a feather (PathDraw leaves it uninitialised otherwise); the coverage
            # buffer range belongs to a mode this backend does not advertise
            pa = np.frombuffer(ha[1].tobytes()[:n * 64], dtype=np.uint32).reshape(-1, 16)[1:]
            pb = np.frombuffer(hb[1].tobytes()[:n * 64], dtype=np.uint32).reshape(-1, 16)[1:]
            assert np.array_equal(pa[:, :9], pb[:, :9])
            feathered = pa[:, 7] != 0
            assert np.array_equal(pa[feathered][:, 9:12], pb[feathered][:, 9:12])
            continue
        assert ha[kind].tobytes() == hb[kind].tobytes(), "buffer kind %d" % kind
    aux_a = np.frombuffer(ha[3].tobytes()[:n * 128], dtype=np.uint32).reshape(-1, 32)
    aux_b = np.frombuffer(hb[3].tobytes()[:n * 128], dtype=np.uint32).reshape(-1, 32)
    assert np.array_equal(aux_a[1:, 8:16], aux_b[1:, 8:16])  # clip rectangle words (always written)


@pytest.mark.parametrize("scene", ["gm:feather_strokes", "gm:feather_shapes", "gm:feather_polyshapes", "gm:feather_corner", "gm:feather_ellipse",
                                   "gm:feather_cusp", "gm:feather_roundcorner", "gm:trickycubicstrokes_feather", "gm:interleavedfeather"])
def test_delegated_feather_draws_reach_the_reference_front_end_unchanged(scene, tmp_path):
    """Feather delegation, pinned on the CPU: CudaPathRenderer rebuilds the state of the RiveRenderer
    it delegates to by replaying the calls of the scopes still open (save / transform / clipPath /
    modulateOpacity). These GMs draw feathers only, so the whole frame is delegated: the flush the
    reference front end writes BEHIND CudaPathRenderer must be, buffer for buffer and batch for batch,
    the flush it writes when it is driven directly -- same matrices bit for bit, same clips, same
    atlas allocations."""
    import subprocess
    from conftest import ROOT
    player = os.path.join(ROOT, "rive-runtime_b200", "_build", "rive_cuda_player")
    recorder = os.path.join(ROOT, "rive-runtime_b200", "_build", "librivecuda_trace.so")
    if not os.path.exists(player) or not os.path.exists(recorder):
        pytest.skip("scene player not built (needs the reference tree at build time)")
    direct, delegated = str(tmp_path / "direct.rvct"), str(tmp_path / "delegated.rvct")
    common = [player, "--scene", scene, "--budget-ms", "0"]
    subprocess.check_call(common, env=dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=direct), stdout=subprocess.DEVNULL, timeout=120)
    subprocess.check_call(common + ["--gpu-front-end"], env=dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=delegated,
                                                              RIVECUDA_TRACE_FRONT_END_OUT=str(tmp_path / "call.rpf")),
                          stdout=subprocess.DEVNULL, timeout=120)
    a, b = T.parse(direct), T.parse(delegated)
    assert sum(1 for r in a if r.tag == T.FLUSH) == 1 and sum(1 for r in b if r.tag == T.FLUSH) == 1
    assert sum(1 for r in b if r.tag == T.PREPARE_TO_FLUSH) == 1  # the one flush is the delegated one
    _assert_last_flushes_identical(a, b)


@pytest.mark.parametrize("name", ["shapetest", "fix_rectangle"])
def test_large_fills_delegated_to_the_reference_triangulator_are_its_own_draws(name, tmp_path):
    """--delegate-large-fills on the CPU: with the reference's deterministic thresholds every fill of
    these assets is interior-triangulated; delegated, the frame is the reference front end's own flush
    (outer-curve patches + interior triangles), byte for byte."""
    import subprocess
    from conftest import ROOT
    player = os.path.join(ROOT, "rive-runtime_b200", "_build", "rive_cuda_player")
    recorder = os.path.join(ROOT, "rive-runtime_b200", "_build", "librivecuda_trace.so")
    asset = os.path.join(ROOT, "tests", "_riv_assets", name + ".riv")
    if not os.path.exists(player) or not os.path.exists(recorder) or not os.path.exists(asset):
        pytest.skip("scene player or .riv asset not present")
    direct, delegated = str(tmp_path / "direct.rvct"), str(tmp_path / "delegated.rvct")
    common = [player, "--scene", "riv:" + asset, "--frames", "3"]
    subprocess.check_call(common, env=dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=direct), stdout=subprocess.DEVNULL, timeout=120)
    subprocess.check_call(common + ["--gpu-front-end", "--delegate-large-fills"],
                          env=dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=delegated, RIVECUDA_TRACE_FRONT_END_OUT=str(tmp_path / "call.rpf")),
                          stdout=subprocess.DEVNULL, timeout=120)
    a, b = T.parse(direct), T.parse(delegated)
    last = [r.fields["flush"] for r in a if r.tag == T.FLUSH][-1]
    assert any(batch.draw_type == 3 for batch in last.batches)  # the reference did triangulate
    if sum(1 for r in b if r.tag == T.FLUSH) == sum(1 for r in a if r.tag == T.FLUSH):
        _assert_last_flushes_identical(a, b)  # (every draw of the frame was delegated)
    else:
        # some draws were CudaPathRenderer's own: the delegated flushes together still hold the triangulated fills
        delegated_types = [batch.draw_type for r in b if r.tag == T.FLUSH for batch in r.fields["flush"].batches]
        assert delegated_types.count(3) == [batch.draw_type for batch in last.batches].count(3) * sum(1 for r in a if r.tag == T.FLUSH)


def test_null_images_are_ignored_like_the_reference_does():
    """Seven of the reference's .sriv silvers draw images that cannot be decoded here (null RenderImages):
    RiveRenderer returns from drawImage (LITE_RTTI_CAST_OR_RETURN), and so must CudaPathRenderer (it
    used to dereference them). image_fit_alignment.sriv is one of them; its records must match too."""
    import subprocess
    from conftest import ROOT
    player = os.path.join(ROOT, "rive-runtime_b200", "_build", "rive_cuda_player")
    recorder = os.path.join(ROOT, "rive-runtime_b200", "_build", "librivecuda_trace.so")
    stream = "/root/reference/tests/unit_tests/silvers/image_fit_alignment.sriv"
    if not os.path.exists(player) or not os.path.exists(recorder) or not os.path.exists(stream):
        pytest.skip("scene player or the reference's silvers not present")
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        reference, call, trace = os.path.join(tmp, "reference.rvct"), os.path.join(tmp, "call.rpf"), os.path.join(tmp, "device.rvct")
        common = [player, "--scene", "sriv:" + stream, "--frames", "0", "--budget-ms", "0"]
        subprocess.check_call(common, env=dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=reference), stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL, timeout=120)
        subprocess.check_call(common + ["--gpu-front-end"], env=dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=trace,
                                                                  RIVECUDA_TRACE_FRONT_END_OUT=call), stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL, timeout=120)
        counts = _compare_device_front_end_call_with_reference_trace(call, trace, T.parse(reference))
    assert counts["paths"] > 0


def test_gm_records_of_both_front_ends_match():
    """The same comparison over every GM the scene player holds (the reference's own tests/gm sources,
    compiled in place): 100 of them -- clips of every kind, blend modes, gradients, images, meshes,
    degenerate strokes, huge paths, frames of more paths (hittest_*) or more gradients (lots_of_grads_*)
    than one logical flush holds: both front ends split them at the same draw, the last flush is what
    is compared -- write identical records through both front ends; the others need feathers (18) or
    drive the RenderContext directly instead of a Renderer (preserverendertarget*, retrofitcubictristrips)."""
    import re
    import subprocess
    from conftest import ROOT
    player = os.path.join(ROOT, "rive-runtime_b200", "_build", "rive_cuda_player")
    if not os.path.exists(player):
        pytest.skip("scene player not built (needs the reference tree at build time)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "gm_records_sweep.py")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         timeout=600).stdout.decode()
    m = re.search(r"identical (\d+), differing (\d+), refused (\d+), skipped \(several flushes\) (\d+), failed (\d+)", out)
    assert m is not None, out[-2000:]
    identical, differing, refused, skipped, failed = (int(g) for g in m.groups())
    assert differing == 0 and failed == 0, out[-3000:]
    assert identical >= 98 and refused <= 20 and skipped <= 4


@pytest.mark.parametrize("seed", [11, 12, 13, 14, 15, 16, 17, 18, 111, 112, 113, 114])
def test_front_end_core_matches_the_reference_on_random_paths(seed, tmp_path):
    """Fuzz, pinned on the reference itself: random RawPaths (lines, generic / cusped / looping /
    degenerate cubics, closed, open and move-only contours, random matrices, joins, caps and
    thicknesses, many partly or wholly outside the frame) are drawn through the reference's own
    front end (scene player `--scene paths:FILE` on the call recorder); the host build of the
    F1 kernels' core must reproduce its spans, contours, path and paint records byte for byte."""
    import subprocess
    from conftest import ROOT
    from path_fuzz import prune_empty_segments, random_paths
    player = os.path.join(ROOT, "rive-runtime_b200", "_build", "rive_cuda_player")
    recorder = os.path.join(ROOT, "rive-runtime_b200", "_build", "librivecuda_trace.so")
    if not os.path.exists(player) or not os.path.exists(recorder):
        pytest.skip("scene player not built (needs the reference tree at build time)")
    width, height = 1920, 1080
    # (seeds > 100: with clockwise fills and left-handed matrices -- forwardThenReverse contour
    # directions, negated coverage; the reference then emits several batches)
    dump, thickness = prune_empty_segments(*random_paths(seed, 1500, width=width, height=height, clockwise=seed > 100))
    dump_file, trace_file = str(tmp_path / "fuzz.paths"), str(tmp_path / "fuzz.rvct")
    F.write_paths(dump_file, dump, thickness)
    env = dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=trace_file)
    subprocess.check_call([player, "--scene", "paths:" + dump_file, "--budget-ms", "0", "--width", str(width), "--height", str(height)],
                          env=env, stdout=subprocess.DEVNULL, timeout=120)
    recs = T.parse(trace_file)
    host = {r.fields["kind"]: r.data for r in recs if r.tag == T.BUFFER_UNMAP}
    flushes = [r.fields["flush"] for r in recs if r.tag == T.FLUSH]
    assert len(flushes) == 1 and all(b.draw_type == 0 for b in flushes[0].batches) and (len(flushes[0].batches) == 1 or seed > 100)
    d = flushes[0].desc
    out = front_end_host.run(dump, width, height)
    res = out.result
    assert (res.path_count, res.contour_count, res.tess_vertex_span_count, res.tess_data_height) == (
        d.path_count, d.contour_count, d.tess_vertex_span_count, d.tess_data_height)
    assert res.path_count > 1000  # most paths are on screen, some are culled
    n = res.tess_vertex_span_count
    want = np.frombuffer(host[6].tobytes()[:n * 64], dtype=np.uint32).reshape(-1, 16)
    bad = np.nonzero((out.spans[:n] != want).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} spans differ, first {bad[:5]}: got {out.spans[bad[0]]} want {want[bad[0]]}"
    want = np.frombuffer(host[4].tobytes()[:res.contour_count * 16], dtype=np.uint32).reshape(-1, 4)
    assert np.array_equal(out.contours[:res.contour_count], want)
    n = res.path_count
    want = np.frombuffer(host[1].tobytes()[:n * 64], dtype=np.uint32).reshape(-1, 16)
    assert np.array_equal(out.path_data[1:n, :8], want[1:, :8])
    want = np.frombuffer(host[2].tobytes()[:n * 8], dtype=np.uint32).reshape(-1, 2)
    assert np.array_equal(out.paint_data[1:n], want[1:])


def _core():
    import ctypes
    lib = ctypes.CDLL(front_end_host.build())
    lib.fe_find_cubic_convex_180_chops.restype = ctypes.c_int
    return lib, ctypes


def _chops(lib, ctypes, pts):
    p = (ctypes.c_float * 8)(*np.asarray(pts, np.float32).ravel())
    T = (ctypes.c_float * 2)()
    cusps = ctypes.c_int(-1)
    n = lib.fe_find_cubic_convex_180_chops(p, T, ctypes.byref(cusps))
    return n, bool(cusps.value), [T[0], T[1]][:n]


def test_convex_180_chops_known_answers_of_the_reference_unit_test():
    """The cases of tests/unit_tests/runtime/bezier_utils_test.cpp:451-556
    ("find_cubic_convex_180_chops", "..._lines") against the core's restatement."""
    lib, ctypes = _core()
    assert _chops(lib, ctypes, [(0, 0), (2, 2), (4, 2), (6, 0)])[0] == 0           # an exact quadratic
    n, cusps, _ = _chops(lib, ctypes, [(0, 0), (1, 1), (1, 0), (0, 1)])            # a cusp
    assert (n, cusps) == (1, True)
    epsilon = 1.0 / (1 << 11)
    h = (1 - epsilon * epsilon) / (3 * epsilon * epsilon + 1)
    dy = (1 - h) / 2
    n, cusps, T = _chops(lib, ctypes, [(0, 0), (1, np.float32(1 - 4 * dy)), (1, np.float32(4 * dy)), (0, 1)])
    assert (n, cusps) == (2, False) and 0 < T[0] < T[1] < 1                        # inflections > epsilon apart
    n, cusps, _ = _chops(lib, ctypes, [(0, 0), (1, np.float32(1 - .9 * dy)), (1, np.float32(.9 * dy)), (0, 1)])
    assert (n, cusps) == (1, True)                                                 # < epsilon apart: one cusp
    # "Ensure fast floatingpoint is not enabled" (bezier_utils_test.cpp:516-532)
    f = np.float32
    p3, p4, p5 = np.array([460, 460], f), np.array([667, 460], f), np.array([1060, 460], f)
    t = f(2 / f(3))
    c0, c1 = (p4 - p3) * t + p3, (p4 - p5) * t + p5
    assert _chops(lib, ctypes, [p3, c0, c1, p5])[0] == 0
    # Flat lines with ordered points are never cusps (:536-556).
    p0, p3 = np.array([123, 200], f), np.array([223, 432], f)
    t0 = f(1e-3)
    while t0 < 1:
        t1 = f(t0 + f(.097))
        while t1 < 1:
            line = [p0, p0 + (p3 - p0) * t0, p0 + (p3 - p0) * t1, p3]
            assert _chops(lib, ctypes, line)[1] is False
            t1 = f(t1 + f(.097))
        t0 = f(t0 + f(.12))
    # Every cubic on the corners of the unit square (:453-463): chop parameters are inside (0, 1)
    # and sorted.
    for i in range(256):
        pts = [((i >> 0) & 1, (i >> 1) & 1), ((i >> 2) & 1, (i >> 3) & 1), ((i >> 4) & 1, (i >> 5) & 1), ((i >> 6) & 1, (i >> 7) & 1)]
        n, cusps, T = _chops(lib, ctypes, pts)
        assert 0 <= n <= 2 and all(0 < t < 1 for t in T) and T == sorted(T)


def test_chop_cubic_at_known_answers_of_the_reference_unit_test():
    """tests/unit_tests/runtime/bezier_utils_test.cpp:34-128 ("chop_cubic_at")."""
    lib, ctypes = _core()
    f4 = ctypes.c_float
    # The diagonal 0..3 chopped at 1/2 gives points at multiples of 1/2 (:37-49).
    pts = (f4 * 8)(*[v for i in range(4) for v in (float(i), float(i))])
    dst = (f4 * 14)()
    lib.fe_chop_cubic_at(pts, dst, f4(.5))
    assert [dst[2 * i] for i in range(7)] == [i * .5 for i in range(7)] and all(dst[2 * i] == dst[2 * i + 1] for i in range(7))
    rng = np.random.default_rng(7)
    chop_ts = [0.0] + [3 / d for d in (83, 79, 73, 71, 67, 61, 59, 53, 47, 43, 41, 37, 31, 29, 23, 19, 17, 13, 11, 7, 5)] + [1.0]
    for _ in range(5):
        p = rng.uniform(0, 1, 8).astype(np.float32)
        pts = (f4 * 8)(*p)
        for t in chop_ts:
            t = np.float32(t)
            two = (f4 * 20)()
            lib.fe_chop_cubic_at2(pts, two, f4(t), f4(t))
            q = np.array(two[:], np.float32).reshape(10, 2)
            assert (q[3] == q[4]).all() and (q[3] == q[5]).all() and (q[3] == q[6]).all()  # the middle is exactly degenerate
            want = (f4 * 2)()
            lib.fe_eval_cubic_at(pts, f4(t), want)
            assert np.allclose(q[3], [want[0], want[1]], atol=1e-5)                       # ... at the right point
            if t == 0:
                assert (q[:4] == p[:2]).all()
            if t == 1:
                assert (q[6:] == p[6:]).all()


def test_polar_segment_count_of_degenerate_tangents():
    """simd::clamp returns lo for NaN (include/rive/math/simd.hpp:244-254): a zero tangent makes
    cosTheta NaN -> -1 -> half a turn of polar segments, as the reference computes it."""
    lib, ctypes = _core()
    lib.fe_polar_segments.restype = ctypes.c_uint32
    t = (ctypes.c_float * 2)
    assert lib.fe_polar_segments(t(1, 0), t(1, 0), ctypes.c_float(10)) == 1
    assert lib.fe_polar_segments(t(1, 0), t(-1, 0), ctypes.c_float(10)) == 32   # ceil(pi * 10)
    assert lib.fe_polar_segments(t(0, 0), t(1, 0), ctypes.c_float(10)) == 32    # NaN -> -1 -> pi


def test_core_scalar_functions_equal_the_numpy_restatement():
    """The core's fast_acos and Wang's-formula count against oracle/front_end_ref.py (which the tests
    at the top of this file pin on the reference's own spans): bit-equal on a grid / random cubics,
    and fast_acos within SIMD_FAST_ACOS_MAX_ERROR = 0.0167552 of acos, the bound the reference's
    simd_test.cpp checks ("fast_acos", include/rive/math/simd.hpp:495)."""
    lib, ctypes = _core()
    lib.fe_fast_acos.restype = ctypes.c_float
    lib.fe_fast_acos.argtypes = [ctypes.c_float]
    lib.fe_wang_cubic_segments.restype = ctypes.c_uint32
    xs = np.linspace(-1, 1, 4097).astype(np.float32)
    got = np.array([lib.fe_fast_acos(float(x)) for x in xs], np.float32)
    assert np.array_equal(got.view(np.uint32), front_end_ref.fast_acos(xs).view(np.uint32))
    assert np.abs(got - np.arccos(xs.astype(np.float64))).max() <= 0.0167552
    rng = np.random.default_rng(3)
    pts = rng.uniform(-500, 500, (2000, 4, 2)).astype(np.float32)
    mats = rng.uniform(-3, 3, (2000, 6)).astype(np.float32)
    want = front_end_ref.wang_cubic_segments(pts, mats)
    for i in range(2000):
        p = (ctypes.c_float * 8)(*pts[i].ravel())
        m = (ctypes.c_float * 6)(*mats[i])
        assert lib.fe_wang_cubic_segments(p, m) == want[i]
