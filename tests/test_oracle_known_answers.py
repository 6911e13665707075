"""Pins the oracle's shader-helper restatements against the reference's own
known-answer / property tests:
  tests/unit_tests/renderer/advanced_blend_test.cpp:53-516  (blend coefficients)
  tests/unit_tests/runtime/bezier_utils_test.cpp:670-745    (clamped_divide, find_cubic_max_height)
and the fixed-function rasterisation rules of SURVEY.md appendix C."""
import ctypes
import itertools
import math

import numpy as np
import pytest

from oracle import refcpu

INF = float("inf")
NAN = float("nan")
F3 = ctypes.c_float * 3
F4 = ctypes.c_float * 4
F6 = ctypes.c_float * 6
F8 = ctypes.c_float * 8
COLORDODGE, COLORBURN = 5, 6
HUE, SATURATION, COLOR, LUMINOSITY = 12, 13, 14, 15


def coeffs(src, dst_premul, mode):
    out = F3()
    refcpu.lib().refcpu_advanced_blend_coeffs(F3(*src), F4(*dst_premul), mode, out)
    return np.array(list(out), dtype=np.float32)


def coeffs_with_dst_alpha(src, dst, a, mode):
    return coeffs(src, (dst[0] * a, dst[1] * a, dst[2] * a, a), mode)


def test_colordodge_edge_cases():
    # advanced_blend_test.cpp:94-225
    for dst in [(0, 0, 0, 0), (.0001, 1, INF, 0), (-.0001, -1, -INF, 0)]:
        for src in [(0, 0, 0), (1, 1.001, INF), (-.0001, -1, -INF)]:
            assert np.all(coeffs(src, dst, COLORDODGE) == 0)
    for a in [0., 1 / 255, .25, .5, 254 / 255, 1., 256 / 255]:
        for src in [(0, 0, 0), (1, 1, 1), (-INF, INF, NAN)]:
            assert np.all(coeffs_with_dst_alpha(src, (0, -.001, -1), a, COLORDODGE) == 0)
            assert np.all(coeffs_with_dst_alpha(src, (-10, -100, -INF), a, COLORDODGE) == 0)
        if a != 0:
            for src in [(1, 1.00001, 2), (10, 100, INF)]:
                for d in [1e-10, 1, 1.0001, INF]:
                    assert np.all(coeffs_with_dst_alpha(src, (d, d, d), a, COLORDODGE) == 1)
        rng = np.random.default_rng(0)
        for _ in range(100):
            src = rng.random(3).astype(np.float32)
            dst = np.zeros(3, np.float32) if a == 0 else rng.random(3).astype(np.float32)
            c = coeffs_with_dst_alpha(src, dst, a, COLORDODGE)
            for i in range(3):
                if dst[i] <= 0:
                    assert c[i] == 0
                elif src[i] >= 1:
                    assert c[i] == 1
                else:
                    assert c[i] == pytest.approx(min(1., dst[i] / (1 - src[i])), rel=1e-4, abs=1e-5)


def test_colorburn_edge_cases():
    # advanced_blend_test.cpp:228-342
    for a in [0., 1 / 255, .25, .5, 254 / 255, 1., 256 / 255]:
        if a != 0:
            for src in [(0, 0, 0), (1, 1, 1), (-INF, INF, NAN)]:
                assert np.all(coeffs_with_dst_alpha(src, (1, 1.001, 2), a, COLORBURN) == 1)
                assert np.all(coeffs_with_dst_alpha(src, (10, 100, INF), a, COLORBURN) == 1)
        for src in [(0, -1e-1, -1), (-10, -100, -INF)]:
            for d in [1 - 1e-6, 0, -1e-6, -INF]:
                assert np.all(coeffs_with_dst_alpha(src, (d, d, d), a, COLORBURN) == 0)
        rng = np.random.default_rng(1)
        for _ in range(100):
            src = rng.random(3).astype(np.float32)
            dst = np.zeros(3, np.float32) if a == 0 else rng.random(3).astype(np.float32)
            c = coeffs_with_dst_alpha(src, dst, a, COLORBURN)
            for i in range(3):
                if dst[i] >= 1:
                    assert c[i] == 1
                elif src[i] <= 0:
                    assert c[i] == 0
                else:
                    assert c[i] == pytest.approx(1. - min(1., (1. - dst[i]) / src[i]), rel=1e-3, abs=2e-5)


# blend_spec_functions, advanced_blend_test.cpp:346-412 (NV_blend_equation_advanced).
def _lum(c):
    return float(np.dot(c, [0.30, 0.59, 0.11]))


def _clip_color(c):
    lum, mn, mx = _lum(c), c.min(), c.max()
    if mn < 0:
        c = lum + ((c - lum) * lum) / (lum - mn)
    if mx > 1:
        c = lum + ((c - lum) * (1 - lum)) / (mx - lum)
    return c


def _set_lum(cbase, clum):
    return _clip_color(cbase + (_lum(clum) - _lum(cbase)))


def _set_lum_sat(cbase, csat, clum):
    sbase = cbase.max() - cbase.min()
    ssat = csat.max() - csat.min()
    color = (cbase - cbase.min()) * ssat / sbase if sbase > 0 else np.zeros(3)
    return _set_lum(color, clum)


@pytest.mark.parametrize("mode,spec", [
    (COLOR, lambda a, b: _set_lum(a, b)),
    (LUMINOSITY, lambda a, b: _set_lum(b, a)),
    (SATURATION, lambda a, b: _set_lum_sat(b, a, b)),
    (HUE, lambda a, b: _set_lum_sat(a, b, b)),
])
def test_hsl_modes_match_spec_functions(mode, spec):
    # test_color_pairs, advanced_blend_test.cpp:416-470 (6^6 colour pairs; we
    # stride the grid to keep the CPU suite fast).
    steps = [i / 5 for i in range(6)]
    grid = list(itertools.product(steps, repeat=3))
    for x in grid[::3]:
        for y in grid[::7]:
            xs = np.floor(np.array(x) * 255.0) / 255.0
            ys = np.floor(np.array(y) * 255.0) / 255.0
            got = coeffs_with_dst_alpha(xs, ys, 1.0, mode)
            want = spec(xs.astype(np.float64), ys.astype(np.float64))
            assert np.max(np.abs(got - want)) <= 1e-4, (mode, x, y, got, want)


def test_clamped_divide_and_max_height_properties():
    # find_cubic_max_height_glsl, bezier_utils_test.cpp:690-745
    L = refcpu.lib()

    def eval_cubic(p, t):
        a = p[0] * (1 - t) + p[1] * t
        b = p[1] * (1 - t) + p[2] * t
        c = p[2] * (1 - t) + p[3] * t
        ab = a * (1 - t) + b * t
        bc = b * (1 - t) + c * t
        return ab * (1 - t) + bc * t

    def check(pts):
        pts = np.asarray(pts, dtype=np.float32)
        t = ctypes.c_float()
        h = L.refcpu_find_cubic_max_height(F8(*pts.reshape(-1)), ctypes.byref(t))
        assert h >= 0 and 0 <= t.value <= 1
        base = pts[3] - pts[0]
        n = np.linalg.norm(base)
        if n == 0:
            return
        norm = np.array([-base[1], base[0]]) / n
        k = -np.dot(norm, pts[0])
        height = lambda tt: abs(np.dot(norm, eval_cubic(pts.astype(np.float64), tt)) + k)  # noqa: E731
        assert height(t.value) == pytest.approx(h, abs=1e-3)

    for i in range(256):
        check([[(i >> 0) & 1, (i >> 1) & 1], [(i >> 2) & 1, (i >> 3) & 1],
               [(i >> 4) & 1, (i >> 5) & 1], [(i >> 6) & 1, (i >> 7) & 1]])
    rng = np.random.default_rng(0)
    for _ in range(100):
        check(rng.uniform(-100, 100, size=(4, 2)))


def test_half_conversion_matches_numpy():
    L = refcpu.lib()
    rng = np.random.default_rng(3)
    vals = np.concatenate([rng.normal(0, 1, 2000), rng.normal(0, 1e-5, 500), rng.normal(0, 3e4, 500),
                           [0.0, -0.0, 1.0, 65504.0, 65520.0, 1e-8, 6.1e-5, 5.96e-8]]).astype(np.float32)
    for v in vals:
        h = L.refcpu_float_to_half(float(v))
        want = np.float32(v).astype(np.float16)
        assert h == want.view(np.uint16), (v, h, want.view(np.uint16))
        assert L.refcpu_half_to_float(h) == float(want) or math.isinf(float(want))


def raster(xy, w=8, h=8, cull=1):
    mask = np.zeros((h, w), np.uint8)
    n = refcpu.lib().refcpu_raster_mask(F6(*xy), cull, w, h, mask.ctypes.data)
    assert n == int(mask.sum())
    return mask


def test_rasteriser_rules():
    # Clockwise (y-down) is front facing; counter-clockwise is culled.
    cw = [1, 1, 5, 1, 5, 5]
    ccw = [1, 1, 5, 5, 5, 1]
    assert raster(cw).sum() > 0 and raster(ccw).sum() == 0 and raster(ccw, cull=0).sum() == raster(cw).sum()
    # Pixel-centre sampling + top-left rule: a quad split along its diagonal covers
    # every pixel exactly once, including centres that sit exactly on edges.
    for quad in ([0.5, 0.5, 6.5, 0.5, 6.5, 6.5, 0.5, 6.5], [1, 2, 7, 1, 6.5, 6.5, 0.5, 5.5]):
        x0, y0, x1, y1, x2, y2, x3, y3 = quad
        a = raster([x0, y0, x1, y1, x2, y2]).astype(int)
        b = raster([x0, y0, x2, y2, x3, y3]).astype(int)
        assert (a + b).max() == 1
    # Edges exactly through pixel centres: left/top edges are inclusive, right/bottom exclusive.
    m = raster([0.5, 0.5, 4.5, 0.5, 4.5, 4.5]) + raster([0.5, 0.5, 4.5, 4.5, 0.5, 4.5])
    assert m[0:4, 0:4].min() == 1 and m[4, :].sum() == 0 and m[:, 4].sum() == 0 and m.max() == 1
    # NaN vertices discard the triangle (vertexDiscardValue).
    assert raster([NAN, 1, 5, 1, 5, 5]).sum() == 0
    # Degenerate (zero-area) triangles draw nothing.
    assert raster([1, 1, 3, 3, 5, 5], cull=0).sum() == 0
