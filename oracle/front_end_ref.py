"""CPU restatement (numpy, float32) of the per-curve arithmetic of the reference's path front
end, used to check the GPU front end's inputs/outputs and pinned against the reference's own
output (the TessVertexSpans in the committed flush traces).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  wang_cubic_segments   wangs_formula::cubic_pow4 (include/rive/math/wangs_formula.hpp:157-168)
                        with VectorXform (:88-113), then ceil(sqrt(sqrt(n4))) clamped to
                        [1, 1023] (renderer/src/draw.cpp:1193-1197); kParametricPrecision = 4
                        (renderer/include/rive/renderer/gpu.hpp:56).
"""
from __future__ import annotations

import numpy as np

F = np.float32


def wang_cubic_segments(pts: np.ndarray, matrix: np.ndarray, precision: float = 4.0) -> np.ndarray:
    """pts: (n, 4, 2) float32 control points; matrix: (n, 6) float32 Mat2D values
    [xx, xy, yx, yy, tx, ty]. Returns uint32 parametric segment counts."""
    p = pts.astype(F)
    m = matrix.astype(F)
    with np.errstate(over="ignore", invalid="ignore"):
        a = (F(-2.0) * p[:, 1] + p[:, 0]) + p[:, 2]   # p0 - 2 p1 + p2
        b = (F(-2.0) * p[:, 2] + p[:, 1]) + p[:, 3]   # p1 - 2 p2 + p3
        # VectorXform: scale = (m0, m3), skew = (m2, m1): v' = scale * v + skew * v.yx

        def xform(v):
            return np.stack([m[:, 0] * v[:, 0] + m[:, 2] * v[:, 1], m[:, 3] * v[:, 1] + m[:, 1] * v[:, 0]], axis=1)

        ta, tb = xform(a), xform(b)
        term = F(9.0 * 4.0 / 64.0) * (F(precision) * F(precision))   # length_term_pow2<3>
        n4 = np.maximum(ta[:, 0] * ta[:, 0] + ta[:, 1] * ta[:, 1], tb[:, 0] * tb[:, 0] + tb[:, 1] * tb[:, 1]) * term
        n = np.ceil(np.sqrt(np.sqrt(n4.astype(F)).astype(F)).astype(F))
        n = np.clip(n, F(1), F(1023))
    return n.astype(np.uint32)


def find_max_scale(m: np.ndarray) -> np.ndarray:
    """Mat2D::findMaxScale (src/math/mat2d_find_max_scale.cpp:24-60), float32, vectorised."""
    m = m.astype(F)
    xx, xy, yx, yy = m[:, 0], m[:, 1], m[:, 2], m[:, 3]
    simple = (xy == 0) & (yx == 0)
    a = xx * xx + xy * xy
    b = xx * yx + yy * xy
    c = yx * yx + yy * yy
    b2 = b * b
    eps = F(1.0 / (1 << 12))  # math::EPSILON
    aminusc = a - c
    x = np.sqrt(aminusc * aminusc + F(4) * b2).astype(F) * F(.5)
    result = np.where(b2 <= eps * eps, np.maximum(a, c), (a + c) * F(.5) + x).astype(F)
    return np.where(simple, np.maximum(np.abs(xx), np.abs(yy)), np.sqrt(result).astype(F)).astype(F)


def fast_acos(x: np.ndarray) -> np.ndarray:
    """simd::fast_acos (include/rive/math/simd.hpp:496-507)."""
    x = x.astype(F)
    a, b, c, d = F(-0.939115566365855), F(0.9217841528914573), F(-1.2845906244690837), F(0.295624144969963174)
    xx = x * x
    numer = b * xx + a
    denom = xx * (d * xx + c) + F(1)
    return x * (numer / denom) + F(1.5707963267948966)


def polar_segments(pts: np.ndarray, matrix: np.ndarray, stroke_radius: np.ndarray, precision: int = 8) -> np.ndarray:
    """Polar segment count of a stroked (already chopped) cubic: the rotation between its end
    tangents, times calc_polar_segments_per_radian<kPolarPrecision = 8>(strokeRadius * maxScale)
    (bezier_utils.hpp:108-113; draw.cpp:776-813, 1203-1232), clamped to [1, 1023]."""
    p = pts.astype(F)

    def first_distinct(a, b, c, d):
        # (a != b ? b : b != c ? c : d)
        ne_ab = (a != b).any(axis=1)[:, None]
        ne_bc = (b != c).any(axis=1)[:, None]
        return np.where(ne_ab, b, np.where(ne_bc, c, d))

    t0 = first_distinct(p[:, 0], p[:, 1], p[:, 2], p[:, 3]) - p[:, 0]
    t1 = p[:, 3] - first_distinct(p[:, 3], p[:, 2], p[:, 1], p[:, 0])
    with np.errstate(invalid="ignore", divide="ignore"):
        numer = t0[:, 0] * t1[:, 0] + t0[:, 1] * t1[:, 1]
        denom2 = (t0[:, 0] * t0[:, 0] + t0[:, 1] * t0[:, 1]) * (t1[:, 0] * t1[:, 0] + t1[:, 1] * t1[:, 1])
        cos_theta = np.clip(numer / np.sqrt(denom2).astype(F), F(-1), F(1))
        theta = fast_acos(cos_theta)
        r = stroke_radius.astype(F) * find_max_scale(matrix)
        cos_step = F(1) - (F(1) / F(precision)) / r
        per_radian = F(.5) / np.arccos(np.maximum(cos_step, F(-1)).astype(F)).astype(F)
        n = np.ceil(theta * per_radian)
        n = np.clip(n, F(1), F(1023))
    return np.nan_to_num(n, nan=1.0).astype(np.uint32)
