"""Python binding of the GPU path front end (rivecuda_front_end_fills) and
reader of the `--dump-paths` files the scene player writes (host/player/player_main.cpp,
PathDumpRenderer): the RawPaths + matrices + paints of a frame's fill draws."""
from __future__ import annotations

import ctypes
import lzma
import struct
from dataclasses import dataclass

import numpy as np


class FillPath(ctypes.Structure):
    """ctypes mirror of rivecuda_fill_path."""
    _fields_ = [("first_verb", ctypes.c_uint32), ("verb_count", ctypes.c_uint32), ("first_point", ctypes.c_uint32),
                ("fill_rule", ctypes.c_uint32), ("matrix", ctypes.c_float * 6), ("color", ctypes.c_uint32),
                ("reserved0", ctypes.c_uint32)]


class FrontEndResult(ctypes.Structure):
    """ctypes mirror of rivecuda_front_end_result."""
    _fields_ = [("path_count", ctypes.c_uint32), ("contour_count", ctypes.c_uint32),
                ("tess_vertex_span_count", ctypes.c_uint32), ("midpoint_fan_tess_vertex_count", ctypes.c_uint32),
                ("tess_data_height", ctypes.c_uint32), ("first_patch", ctypes.c_uint32), ("patch_count", ctypes.c_uint32),
                ("reserved0", ctypes.c_uint32)]


assert ctypes.sizeof(FillPath) == 48 and ctypes.sizeof(FrontEndResult) == 32


@dataclass
class PathDump:
    paths: np.ndarray   # structured array with FillPath's layout
    verbs: np.ndarray   # uint8
    points: np.ndarray  # float32 (n, 2)
    complete: bool      # every draw of the frame was a plain fill or stroke
    strokes: np.ndarray = None  # structured (is_stroke, thickness, join, cap) per path; zeros for "RPTH" dumps


PATH_DTYPE = np.dtype([("first_verb", "<u4"), ("verb_count", "<u4"), ("first_point", "<u4"), ("fill_rule", "<u4"),
                       ("matrix", "<f4", (6,)), ("color", "<u4"), ("reserved0", "<u4")])
assert PATH_DTYPE.itemsize == 48
STROKE_DTYPE = np.dtype([("is_stroke", "<u4"), ("thickness", "<f4"), ("join", "<u4"), ("cap", "<u4")])


def load_paths(path: str) -> PathDump:
    raw = lzma.open(path, "rb").read() if path.endswith(".xz") else open(path, "rb").read()
    magic, count, complete, _ = struct.unpack_from("<4I", raw, 0)
    if magic not in (0x48545052, 0x32545052):  # "RPTH" (fills only), "RPT2" (+ stroke fields)
        raise ValueError("not a path dump")
    v2 = magic == 0x32545052
    paths = np.zeros(count, dtype=PATH_DTYPE)
    strokes = np.zeros(count, dtype=STROKE_DTYPE)
    verbs, points = [], []
    pos, nv, npnt = 16, 0, 0
    for i in range(count):
        m = struct.unpack_from("<6f", raw, pos)
        rule, color, n_verbs, n_pts = struct.unpack_from("<4I", raw, pos + 24)
        pos += 40
        if v2:
            strokes[i] = struct.unpack_from("<IfII", raw, pos)
            pos += 16
        verbs.append(np.frombuffer(raw, dtype=np.uint8, count=n_verbs, offset=pos))
        pos += (n_verbs + 3) & ~3
        points.append(np.frombuffer(raw, dtype=np.float32, count=n_pts * 2, offset=pos))
        pos += n_pts * 8
        paths[i] = (nv, n_verbs, npnt, rule, m, color, 0)
        nv += n_verbs
        npnt += n_pts
    return PathDump(paths, np.concatenate(verbs) if verbs else np.zeros(0, np.uint8),
                    (np.concatenate(points) if points else np.zeros(0, np.float32)).reshape(-1, 2), bool(complete), strokes)


def run(replayer, dump: PathDump) -> FrontEndResult:
    """rivecuda_front_end_fills on a Replayer's context."""
    res = FrontEndResult()
    pts = np.ascontiguousarray(dump.points, dtype=np.float32)
    verbs = np.ascontiguousarray(dump.verbs, dtype=np.uint8)
    paths = np.ascontiguousarray(dump.paths)
    replayer._call("rivecuda_front_end_fills", pts.ctypes.data, pts.shape[0], verbs.ctypes.data, verbs.size,
                   paths.ctypes.data, paths.size, ctypes.byref(res))
    return res


def read_buffer(replayer, kind: int, size: int) -> np.ndarray:
    out = np.empty(size, dtype=np.uint8)
    replayer._call("rivecuda_debug_read_buffer", kind, out.ctypes.data, 0, size)
    return out
