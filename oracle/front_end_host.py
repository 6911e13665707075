"""ctypes binding of oracle/front_end_host/libfront_end_host.so: the GPU path front end's
per-contour core (rive-runtime_b200/csrc/front_end_core.h) built for the host.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/ to compare that code with the
reference front end's output without a GPU. The product never loads it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from rive_runtime_b200 import front_end as F  # noqa: E402

_LIB_PATH = os.path.join(_HERE, "front_end_host", "libfront_end_host.so")


def build(force: bool = False) -> str:
    # make decides (the library depends on csrc/front_end_core.h, which changes with the kernels)
    subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "front_end_host")] + (["-B"] if force else []))
    return _LIB_PATH


@dataclass
class HostFrontEndOutput:
    result: F.FrontEndResult
    spans: np.ndarray       # (n, 16) uint32
    contours: np.ndarray    # (n, 4) uint32
    path_data: np.ndarray   # (n, 16) uint32
    paint_data: np.ndarray  # (n, 2) uint32
    paint_aux: np.ndarray   # (n, 32) uint32


def run(dump: F.PathDump, frame_width: int = 0, frame_height: int = 0, tables: "F.FrontEndTables" = None) -> HostFrontEndOutput:
    lib = ctypes.CDLL(build())
    fn = lib.front_end_host_paths
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32,
                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    pts = np.ascontiguousarray(dump.points, dtype=np.float32)
    verbs = np.ascontiguousarray(dump.verbs, dtype=np.uint8)
    paths = np.ascontiguousarray(dump.paths)
    n_paths = paths.size
    cap = int(verbs.size) * 8 + 2 * 2048 + 3
    spans = np.zeros((cap, 16), np.uint32)
    contours = np.zeros((verbs.size + 1, 4), np.uint32)
    path_data = np.zeros((n_paths + 1, 16), np.uint32)
    paint_data = np.zeros((n_paths + 1, 2), np.uint32)
    paint_aux = np.zeros((n_paths + 1, 32), np.uint32)
    res = F.FrontEndResult()
    keep = [np.ascontiguousarray(t) for t in ((tables.clip_rects, tables.gradient_paints, tables.image_paints) if tables is not None else ())]
    table_ptrs = [t.ctypes.data if t.size else None for t in keep] if keep else [None, None, None]
    rc = fn(pts.ctypes.data, verbs.ctypes.data, paths.ctypes.data, n_paths, frame_width, frame_height, spans.ctypes.data, cap, contours.ctypes.data,
            path_data.ctypes.data, paint_data.ctypes.data, paint_aux.ctypes.data, ctypes.byref(res), *table_ptrs)
    if rc != 0:
        raise RuntimeError("front_end_host_paths: span capacity exceeded")
    return HostFrontEndOutput(res, spans, contours, path_data, paint_data, paint_aux)
