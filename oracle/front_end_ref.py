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
