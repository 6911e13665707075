"""Pins the pixel-stage oracle (oracle/refcpu, the hand-written restatement every GPU parity
test compares against) to REFERENCE-COMPILED code: oracle/glslref builds the reference's own
shader sources (/root/reference/renderer/src/shaders: tessellate.glsl, bezier_utils.glsl,
common.glsl, draw_path_common.glsl, advanced_blend.glsl, draw_path.vert,
draw_raster_order_path.frag) as C++ and runs their mains on the CPU. Both are fed the same
inputs and must agree BIT FOR BIT:

  * advanced_blend.glsl on the whole 6^6 colour grid x 15 modes and on 10^6 random colours
    (what the reference's advanced_blend_test.cpp samples with tolerances);
  * bezier_utils.glsl's feather helpers on 10^6 random cubics;
  * tessellate.glsl (vertex + fragment main) on every span of the committed flush traces;
  * draw_path.vert / unpack_tessellated_path_vertex on the patch vertices of every patch batch;
  * draw_raster_order_path.frag (+ find_paint_color, feather evaluation, clip, clip rect,
    advanced blend, dither) on randomised fragments and pixel-local-storage states.

What is outside the shader sources (triangle rasterisation, varying interpolation, texture
filtering, unorm8 / fp16 conversions) is fixed-function and is defined identically on both
sides; it is exercised by test_oracle_golden_cpu.py / test_oracle_known_answers.py.

The library needs /root/reference to build; where neither it nor a prebuilt
oracle/_ref/libglslref.so exists the tests skip.
"""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import glslref as G  # noqa: E402
from oracle import refcpu as R  # noqa: E402
from rive_runtime_b200 import trace as T  # noqa: E402

pytestmark = pytest.mark.skipif(not G.available(), reason="needs /root/reference (or a prebuilt oracle/_ref/libglslref.so)")

GOLDEN = os.path.join(ROOT, "tests", "golden")

# Scenes that between them reach every branch of the shaders: fills, every join / cap, cusps,
# feathered fills and strokes, gradients, clips, clip rects, all blend modes, even-odd / clockwise.
TRACES = ["beziers", "c1", "strokes_round", "trickycubicstrokes", "trickycubicstrokes_roundcaps", "trickycubicstrokes_feather",
          "feather_shapes", "feather_strokes", "feather_corner", "feather_cusp", "feather_roundcorner", "feather_polyshapes",
          "roundjoinstrokes", "bevel180strokes", "widebuttcaps", "labyrinth_square", "inner_join_geometry", "zero_control_stroke",
          "cliprects", "cliprectintersections", "clip_shapes_small_corners_feathered_blend", "parallelclips", "xfermodes2",
          "dstreadshuffle", "interleavedfeather", "interleavedfillrule", "poly_clockwise", "poly_evenOdd", "mutating_fill_rule",
          "retrofitcubictristrips", "batchedtriangulations", "largeclippedpath_winding_nested", "verycomplexgrad", "degengrad",
          "overstroke_blendmodes", "riv_off_road_car", "s1", "c3"]


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _same(a, b):
    """Bit equality, except that any NaN equals any NaN (payloads are not specified)."""
    return (_bits(a) == _bits(b)) | (np.isnan(a) & np.isnan(b))


def test_advanced_blend_is_bit_identical_on_the_colour_grid_and_random_colours():
    L, GL = R.lib(), G.lib()
    g = np.array([0, .2, .4, .6, .8, 1], dtype=np.float32)
    # src rgb x dst (rgb as a fraction of alpha) over the reference test's six levels, at
    # three alphas: the full 6^6 = 46 656 combinations each.
    sr, sg, sb, dr, dg, db = [m.ravel() for m in np.meshgrid(g, g, g, g, g, g, indexing="ij")]
    srcs, dsts = [], []
    for alpha in (1.0, .6, .2):
        a = np.full_like(sr, alpha)
        srcs.append(np.stack([sr, sg, sb], 1))
        dsts.append(np.stack([dr * a, dg * a, db * a, a], 1))
    rng = np.random.default_rng(7)
    n = 1_000_000
    a = rng.random((n, 1), dtype=np.float32)
    srcs.append(rng.random((n, 3), dtype=np.float32) * 1.25 - .125)  # includes out-of-gamut sources
    dsts.append(np.concatenate([rng.random((n, 3), dtype=np.float32) * a, a], 1))
    # destinations as they really occur: 8-bit premultiplied colours
    d8 = rng.integers(0, 256, (n, 4)).astype(np.float32)
    d8[:, :3] = np.minimum(d8[:, :3], d8[:, 3:4])
    srcs.append(rng.integers(0, 256, (n, 3)).astype(np.float32) / np.float32(255))
    dsts.append(d8 * np.float32(1 / 255))
    src = np.ascontiguousarray(np.concatenate(srcs).astype(np.float32))
    dst = np.ascontiguousarray(np.concatenate(dsts).astype(np.float32))
    total = 0
    for mode in range(1, 16):
        modes = np.full(len(src), mode, dtype=np.uint32)
        for coeffs_only in (1, 0):
            mine = np.zeros((len(src), 3), np.float32)
            ref = np.zeros_like(mine)
            L.refcpu_advanced_color_blend_n(len(src), src.ctypes.data, dst.ctypes.data, modes.ctypes.data, mine.ctypes.data, coeffs_only)
            GL.glslref_advanced_color_blend_n(len(src), src.ctypes.data, dst.ctypes.data, modes.ctypes.data, ref.ctypes.data, coeffs_only)
            ok = _same(mine, ref).all(1)
            assert ok.all(), (mode, coeffs_only, int((~ok).sum()), src[~ok][0], dst[~ok][0], mine[~ok][0], ref[~ok][0])
            total += len(src)
    assert total >= 15 * 2 * (3 * 6 ** 6 + 2_000_000)


def test_feather_cubic_helpers_are_bit_identical():
    L, GL = R.lib(), G.lib()
    rng = np.random.default_rng(11)
    n = 1_000_000
    pts = (rng.random((n, 8), dtype=np.float32) * 400 - 100).astype(np.float32)
    # degenerate families: lines, coincident control points, cusps
    pts[:1000, 2:4] = pts[:1000, 0:2]
    pts[1000:2000, 4:6] = pts[1000:2000, 6:8]
    pts[2000:3000, 2:8] = np.tile(pts[2000:3000, 0:2], 3)
    pts[3000:4000, 6:8] = pts[3000:4000, 0:2]
    spreads = (rng.random(n, dtype=np.float32) * 30 + .01).astype(np.float32)
    mine = np.zeros((n, 3), np.float32)
    ref = np.zeros_like(mine)
    L.refcpu_cubic_helpers_n(n, pts.ctypes.data, spreads.ctypes.data, mine.ctypes.data)
    GL.glslref_cubic_helpers_n(n, pts.ctypes.data, spreads.ctypes.data, ref.ctypes.data)
    ok = _same(mine, ref).all(1)
    assert ok.all(), (int((~ok).sum()), pts[~ok][0], spreads[~ok][0], mine[~ok][0], ref[~ok][0])


def _fragments_for(rng, verts, n):
    """Randomised fragment inputs built from real vertex-stage outputs of the batch."""
    rows = verts[rng.integers(0, len(verts), n)]
    other = verts[rng.integers(0, len(verts), n)]
    t = rng.random((n, 1), dtype=np.float32)
    frag = np.zeros((n, 24), np.float32)
    lerp = lambda a, b: a + (b - a) * t  # noqa: E731
    frag[:, 0:4] = rows[:, 4:8]
    # interpolate only what the rasteriser interpolates; gradients keep their row / span fields
    frag[:, 0:2] = lerp(rows[:, 4:6], other[:, 4:6])
    frag[:, 8:12] = lerp(rows[:, 8:12], other[:, 8:12])
    frag[:, 12:16] = rows[:, 12:16]
    frag[:, 16:20] = lerp(rows[:, 16:20], other[:, 16:20])
    frag[:, 20] = rng.integers(0, 4096, n)
    frag[:, 21] = rng.integers(0, 4096, n)
    L = R.lib()
    h = lambda x: L.refcpu_float_to_half(float(x))  # noqa: E731
    pls = np.zeros((n, 4), np.uint32)
    col = rng.integers(0, 256, (n, 4)).astype(np.uint32)
    col[:, :3] = np.minimum(col[:, :3], col[:, 3:4])
    pls[:, 0] = col[:, 0] | (col[:, 1] << 8) | (col[:, 2] << 16) | (col[:, 3] << 24)
    col = rng.integers(0, 256, (n, 4)).astype(np.uint32)
    pls[:, 2] = col[:, 0] | (col[:, 1] << 8) | (col[:, 2] << 16) | (col[:, 3] << 24)
    cov = rng.random(n) * 3 - 1
    for k in range(n):
        path_id = abs(float(frag[k, 12]))
        clip_id = abs(float(frag[k, 13]))
        outer_id = abs(float(frag[k, 14]))
        # coverage plane: this path's running count, another path's, or cleared
        which = rng.integers(0, 3)
        pls[k, 3] = (h(cov[k]) | (h(path_id) << 16)) if which == 0 else ((h(cov[k]) | (h(path_id + 1) << 16)) if which == 1 else 0)
        which = rng.integers(0, 4)
        content = clip_id if which == 0 else (outer_id if which == 1 else (clip_id + 2 if which == 2 else 0.))
        pls[k, 1] = h(rng.random()) | (h(content) << 16)
    return np.ascontiguousarray(frag), np.ascontiguousarray(pls)


@pytest.mark.parametrize("name", TRACES)
def test_shader_stages_are_bit_identical_on_committed_traces(name):
    L, GL = R.lib(), G.lib()
    records = T.parse(os.path.join(GOLDEN, name + ".rvct.xz"))
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 32))
    seen = {"texels": 0, "vertices": 0, "fragments": 0}

    def on_flush(rf, fo, fr):
        d = fo.desc
        # -- tessellate.glsl: every texel of the tessellation texture
        n_texels = d.tess_data_height * 2048
        if n_texels and seen["texels"] < 3_000_000:
            ref = np.zeros_like(fo.tess)
            saved = rf.tess_texture
            rf.tess_texture = ref.ctypes.data
            assert GL.glslref_tessellate(ctypes.byref(rf)) == 0
            rf.tess_texture = saved
            a, b = fo.tess[:n_texels], ref[:n_texels]
            same = _same(a.view(np.float32), b.view(np.float32)).all(1)
            assert same.all(), (name, "tessellate.glsl", int((~same).sum()), int(np.nonzero(~same)[0][0]), a[~same][0], b[~same][0])
            seen["texels"] += n_texels
        # -- draw_path.vert + draw_raster_order_path.frag per patch batch
        budget = 40
        for bi, batch in enumerate(fr.batches):
            if batch.draw_type not in (0, 1, 2) or budget == 0:
                continue
            budget -= 1
            vcount = ctypes.c_uint32()
            assert L.refcpu_path_vertices(ctypes.byref(rf), bi, 0, 0, None, ctypes.byref(vcount)) == 0
            n_inst = min(batch.element_count, 600)
            mine = np.zeros((n_inst * vcount.value, 24), np.float32)
            ref = np.zeros_like(mine)
            assert L.refcpu_path_vertices(ctypes.byref(rf), bi, 0, n_inst, mine.ctypes.data, None) == 0
            assert GL.glslref_path_vertices(ctypes.byref(rf), bi, 0, n_inst, ref.ctypes.data) == 0
            same = _same(mine, ref)
            discarded = mine[:, 2] != 0
            assert (mine[:, 2] == ref[:, 2]).all(), (name, bi, "discard flags differ")
            same[discarded, :] = True  # a discarded vertex (NaN position: its triangles are dropped) has no defined outputs
            assert same.all(), (name, bi, "draw_path.vert", np.nonzero(~same.all(1))[0][:4], mine[~same.all(1)][0], ref[~same.all(1)][0])
            seen["vertices"] += len(mine)
            live = mine[~discarded]
            if len(live) == 0:
                continue
            frag, pls = _fragments_for(rng, live, 400)
            out_mine = np.zeros_like(pls)
            out_ref = np.zeros_like(pls)
            assert L.refcpu_path_fragments(ctypes.byref(rf), bi, len(frag), frag.ctypes.data, pls.ctypes.data, out_mine.ctypes.data) == 0
            assert GL.glslref_path_fragments(ctypes.byref(rf), bi, len(frag), frag.ctypes.data, pls.ctypes.data, out_ref.ctypes.data) == 0
            same = (out_mine == out_ref)
            # the scratch plane is don't-care unless the shader stored to it on this fragment on both sides
            ok = same[:, [0, 1, 3]].all(1)
            assert ok.all(), (name, bi, "draw_raster_order_path.frag", int((~ok).sum()), frag[~ok][0], pls[~ok][0], out_mine[~ok][0], out_ref[~ok][0])
            seen["fragments"] += len(frag)

    R.replay(records, threads=os.cpu_count() or 1, keep_intermediates=False, max_frames=2, on_flush=on_flush)
    assert seen["texels"] > 0
