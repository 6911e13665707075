"""Python binding of the GPU path front end (rivecuda_front_end_paths) and
reader of the `--dump-paths` files the scene player writes (host/player/path_dump.hpp):
the RawPaths + matrices + paints of a frame's plain fill and stroke draws."""
from __future__ import annotations

import ctypes
import lzma
import struct
from dataclasses import dataclass

import numpy as np


class Path(ctypes.Structure):
    """ctypes mirror of rivecuda_path."""
    _fields_ = [("first_verb", ctypes.c_uint32), ("verb_count", ctypes.c_uint32), ("first_point", ctypes.c_uint32),
                ("fill_rule", ctypes.c_uint32), ("matrix", ctypes.c_float * 6), ("color", ctypes.c_uint32),
                ("stroke", ctypes.c_uint32), ("stroke_radius", ctypes.c_float), ("join", ctypes.c_uint32),
                ("cap", ctypes.c_uint32), ("polar_segments_per_radian", ctypes.c_float),
                ("matrix_max_scale", ctypes.c_float), ("blend_mode", ctypes.c_uint32)]


class FrontEndResult(ctypes.Structure):
    """ctypes mirror of rivecuda_front_end_result."""
    _fields_ = [("path_count", ctypes.c_uint32), ("contour_count", ctypes.c_uint32),
                ("tess_vertex_span_count", ctypes.c_uint32), ("midpoint_fan_tess_vertex_count", ctypes.c_uint32),
                ("tess_data_height", ctypes.c_uint32), ("first_patch", ctypes.c_uint32), ("patch_count", ctypes.c_uint32),
                ("reserved0", ctypes.c_uint32)]


assert ctypes.sizeof(Path) == 72 and ctypes.sizeof(FrontEndResult) == 32


@dataclass
class PathDump:
    paths: np.ndarray   # structured array with rivecuda_path's layout
    verbs: np.ndarray   # uint8
    points: np.ndarray  # float32 (n, 2)
    complete: bool      # every draw of the frame was a plain fill or stroke


PATH_DTYPE = np.dtype([("first_verb", "<u4"), ("verb_count", "<u4"), ("first_point", "<u4"), ("fill_rule", "<u4"),
                       ("matrix", "<f4", (6,)), ("color", "<u4"), ("stroke", "<u4"), ("stroke_radius", "<f4"),
                       ("join", "<u4"), ("cap", "<u4"), ("polar_segments_per_radian", "<f4"),
                       ("matrix_max_scale", "<f4"), ("blend_mode", "<u4")])
assert PATH_DTYPE.itemsize == 72


def find_max_scale(m) -> np.float32:
    """Mat2D::findMaxScale (src/math/mat2d_find_max_scale.cpp:24-60) in float32."""
    f = np.float32
    xx, xy, yx, yy = (f(v) for v in m[:4])
    if xy == 0 and yx == 0:
        return max(abs(xx), abs(yy))
    a = f(xx * xx + xy * xy)
    b = f(xx * yx + yy * xy)
    c = f(yx * yx + yy * yy)
    b2 = f(b * b)
    eps = f(1.0 / (1 << 12))
    if b2 <= f(eps * eps):
        result = max(a, c)
    else:
        amc = f(a - c)
        x = f(np.sqrt(f(f(amc * amc) + f(f(4) * b2))) * f(.5))
        result = f(f(f(a + c) * f(.5)) + x)
    return f(np.sqrt(result))


_LIBM = ctypes.CDLL("libm.so.6")
_LIBM.acosf.restype = ctypes.c_float
_LIBM.acosf.argtypes = [ctypes.c_float]


def stroke_scalars(matrix, thickness: float):
    """(stroke_radius, matrix_max_scale, polar_segments_per_radian) the way PathDraw does it on
    the host (draw.cpp:603-607, 776-813; bezier_utils.hpp:108-113), glibc acosf included."""
    f = np.float32
    radius = max(f(f(thickness) * f(.5)), np.finfo(np.float32).tiny)
    max_scale = find_max_scale(matrix)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        cos_theta = f(f(1) - f(f(f(1) / f(8)) / f(radius * max_scale)))
    psr = f(f(.5) / f(_LIBM.acosf(max(cos_theta, f(-1)))))
    return radius, max_scale, psr


def load_paths(path: str) -> PathDump:
    raw = lzma.open(path, "rb").read() if path.endswith(".xz") else open(path, "rb").read()
    magic, count, complete, _ = struct.unpack_from("<4I", raw, 0)
    if magic not in (0x48545052, 0x32545052):  # "RPTH" (fills only), "RPT2" (+ stroke fields)
        raise ValueError("not a path dump")
    v2 = magic == 0x32545052
    paths = np.zeros(count, dtype=PATH_DTYPE)
    verbs, points = [], []
    pos, nv, npnt = 16, 0, 0
    for i in range(count):
        m = struct.unpack_from("<6f", raw, pos)
        rule, color, n_verbs, n_pts = struct.unpack_from("<4I", raw, pos + 24)
        pos += 40
        stroke = (0, 0.0, 0, 0)
        if v2:
            stroke = struct.unpack_from("<IfII", raw, pos)
            pos += 16
        verbs.append(np.frombuffer(raw, dtype=np.uint8, count=n_verbs, offset=pos))
        pos += (n_verbs + 3) & ~3
        points.append(np.frombuffer(raw, dtype=np.float32, count=n_pts * 2, offset=pos))
        pos += n_pts * 8
        if stroke[0]:
            radius, max_scale, psr = stroke_scalars(m, stroke[1])
            paths[i] = (nv, n_verbs, npnt, 0, m, color, 1, radius, stroke[2], stroke[3], psr, max_scale, 0)
        else:
            paths[i] = (nv, n_verbs, npnt, rule, m, color, 0, 0.0, 0, 0, 0.0, 0.0, 0)
        nv += n_verbs
        npnt += n_pts
    return PathDump(paths, np.concatenate(verbs) if verbs else np.zeros(0, np.uint8),
                    (np.concatenate(points) if points else np.zeros(0, np.float32)).reshape(-1, 2), bool(complete))


def write_paths(path: str, dump: PathDump, thickness) -> None:
    """Writes a PathDump in the "RPT2" format of host/player/path_dump.hpp (the scene player draws
    such a file with `--scene paths:FILE`). thickness: stroke thickness per path (0 for fills)."""
    out = [struct.pack("<4I", 0x32545052, dump.paths.size, 1, 0)]
    for i, p in enumerate(dump.paths):
        out.append(np.asarray(p["matrix"], np.float32).tobytes())
        out.append(struct.pack("<4I", int(p["fill_rule"]), int(p["color"]), int(p["verb_count"]), 0))
        v = dump.verbs[int(p["first_verb"]):int(p["first_verb"]) + int(p["verb_count"])]
        n_pts = int((v == 0).sum() + (v == 1).sum() + 3 * (v == 4).sum())
        out[-1] = struct.pack("<4I", int(p["fill_rule"]), int(p["color"]), int(p["verb_count"]), n_pts)
        out.append(struct.pack("<IfII", int(p["stroke"]), float(thickness[i]), int(p["join"]), int(p["cap"])))
        out.append(v.tobytes() + b"\0" * ((4 - v.size % 4) % 4))
        out.append(np.ascontiguousarray(dump.points[int(p["first_point"]):int(p["first_point"]) + n_pts], np.float32).tobytes())
    with open(path, "wb") as f:
        f.write(b"".join(out))


CLIP_RECT_DTYPE = np.dtype([("inverse_matrix", "<f4", (6,)), ("inverse_fwidth", "<f4", (2,)), ("pixel_bounds", "<i4", (4,))])
GRADIENT_PAINT_DTYPE = np.dtype([("paint_type", "<u4"), ("grad_texture_y", "<f4"), ("paint_matrix", "<f4", (6,)), ("grad_horizontal_span", "<f4", (2,))])
IMAGE_PAINT_DTYPE = np.dtype([("image_matrix", "<f4", (6,)), ("image_texture_lod", "<f4"), ("reserved0", "<u4")])


@dataclass
class FrontEndTables:
    """The tables a rivecuda_front_end_paths call refers to (rivecuda.h): clip rectangles
    (path.stroke >> 8), gradient paints (path.fill_rule >> 8), image paints (path.cap >> 8)."""
    clip_rects: np.ndarray
    gradient_paints: np.ndarray
    image_paints: np.ndarray


def load_front_end_call(path: str, with_tables: bool = False):
    """Reads what the call recorder (librivecuda_trace.so with $RIVECUDA_TRACE_FRONT_END_OUT)
    saw in rivecuda_front_end_paths: (PathDump, frame_width, frame_height[, FrontEndTables])."""
    raw = open(path, "rb").read()
    magic, n_paths, n_points, n_verbs, width, height, n_clips, n_grads = struct.unpack_from("<8I", raw, 0)
    if magic != 0x32465052:
        raise ValueError("not a recorded front-end call")
    n_images = struct.unpack_from("<I", raw, 32)[0]
    pos = 48
    paths = np.frombuffer(raw, dtype=PATH_DTYPE, count=n_paths, offset=pos).copy()
    pos += n_paths * PATH_DTYPE.itemsize
    verbs = np.frombuffer(raw, dtype=np.uint8, count=n_verbs, offset=pos).copy()
    pos += (n_verbs + 3) & ~3
    points = np.frombuffer(raw, dtype=np.float32, count=n_points * 2, offset=pos).reshape(-1, 2).copy()
    pos += n_points * 8
    clips = np.frombuffer(raw, dtype=CLIP_RECT_DTYPE, count=n_clips, offset=pos).copy()
    pos += n_clips * CLIP_RECT_DTYPE.itemsize
    grads = np.frombuffer(raw, dtype=GRADIENT_PAINT_DTYPE, count=n_grads, offset=pos).copy()
    pos += n_grads * GRADIENT_PAINT_DTYPE.itemsize
    images = np.frombuffer(raw, dtype=IMAGE_PAINT_DTYPE, count=n_images, offset=pos).copy()
    dump = PathDump(paths, verbs, points, True)
    if with_tables:
        return dump, width, height, FrontEndTables(clips, grads, images)
    return dump, width, height


def run(replayer, dump: PathDump, frame_width: int = 0, frame_height: int = 0) -> FrontEndResult:
    """rivecuda_front_end_paths on a Replayer's context. A non-zero frame size enables the
    reference's frame cull (paths outside the render target draw nothing)."""
    res = FrontEndResult()
    pts = np.ascontiguousarray(dump.points, dtype=np.float32)
    verbs = np.ascontiguousarray(dump.verbs, dtype=np.uint8)
    paths = np.ascontiguousarray(dump.paths)
    replayer._call("rivecuda_front_end_paths", pts.ctypes.data, pts.shape[0], verbs.ctypes.data, verbs.size,
                   paths.ctypes.data, paths.size, frame_width, frame_height, ctypes.byref(res))
    return res


def read_buffer(replayer, kind: int, size: int) -> np.ndarray:
    out = np.empty(size, dtype=np.uint8)
    replayer._call("rivecuda_debug_read_buffer", kind, out.ctypes.data, 0, size)
    return out
