"""ctypes binding + trace replayer for the CPU restatement oracle
(oracle/refcpu/librefcpu.so).

TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product path
(rive_runtime_b200.replay -> librivecuda.so) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from rive_runtime_b200 import trace as T  # noqa: E402

_LIB_PATH = os.path.join(_HERE, "refcpu", "librefcpu.so")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (no GPU, no reference sources needed)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "refcpu")])
    return _LIB_PATH


class RefTexture(ctypes.Structure):
    _fields_ = [("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("level_count", ctypes.c_uint32),
                ("reserved0", ctypes.c_uint32), ("levels", ctypes.c_void_p * 16)]


class RefRenderBuffer(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("size_in_bytes", ctypes.c_uint64)]


class RefStaticTables(ctypes.Structure):
    _fields_ = [("patch_vertices", ctypes.c_void_p), ("patch_indices", ctypes.c_void_p),
                ("gaussian_f16", ctypes.c_void_p), ("inverse_gaussian_f16", ctypes.c_void_p)]


class RefFlush(ctypes.Structure):
    _fields_ = [
        ("desc", ctypes.POINTER(T.FlushDesc)),
        ("batches", ctypes.POINTER(T.DrawBatch)),
        ("batch_count", ctypes.c_uint32),
        ("atlas_fill_batch_count", ctypes.c_uint32),
        ("atlas_fill_batches", ctypes.POINTER(T.AtlasBatch)),
        ("atlas_stroke_batches", ctypes.POINTER(T.AtlasBatch)),
        ("atlas_stroke_batch_count", ctypes.c_uint32),
        ("threads", ctypes.c_uint32),
        ("buffers", ctypes.c_void_p * 9),
        ("tables", ctypes.POINTER(RefStaticTables)),
        ("target_width", ctypes.c_uint32),
        ("target_height", ctypes.c_uint32),
        ("target_pixels", ctypes.c_void_p),
        ("grad_texture", ctypes.c_void_p),
        ("grad_rows", ctypes.c_uint32),
        ("tess_rows", ctypes.c_uint32),
        ("tess_texture", ctypes.c_void_p),
        ("atlas", ctypes.c_void_p),
        ("atlas_width", ctypes.c_uint32),
        ("atlas_height", ctypes.c_uint32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        for name in ("refcpu_color_ramps", "refcpu_tessellate", "refcpu_render_atlas", "refcpu_draw", "refcpu_flush_run"):
            fn = getattr(_lib, name)
            fn.argtypes = [ctypes.POINTER(RefFlush)]
            fn.restype = ctypes.c_int
        _lib.refcpu_last_error.restype = ctypes.c_char_p
        _lib.refcpu_find_cubic_max_height.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
        _lib.refcpu_find_cubic_max_height.restype = ctypes.c_float
        _lib.refcpu_measure_cubic_local_curvature.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_float, ctypes.c_float]
        _lib.refcpu_measure_cubic_local_curvature.restype = ctypes.c_float
        for name in ("refcpu_advanced_color_blend", "refcpu_advanced_blend_coeffs"):
            fn = getattr(_lib, name)
            fn.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.c_uint32, ctypes.POINTER(ctypes.c_float)]
            fn.restype = None
        _lib.refcpu_float_to_half.argtypes = [ctypes.c_float]
        _lib.refcpu_float_to_half.restype = ctypes.c_uint16
        _lib.refcpu_half_to_float.argtypes = [ctypes.c_uint16]
        _lib.refcpu_half_to_float.restype = ctypes.c_float
        _lib.refcpu_raster_mask.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
        _lib.refcpu_raster_mask.restype = ctypes.c_int
        _bind_pin_exports(_lib, "refcpu")
        _lib.refcpu_set_atlas_mode.argtypes = [ctypes.c_int]
        _lib.refcpu_set_atlas_mode.restype = None
    return _lib


ATLAS_R16F_BLEND, ATLAS_R32I_ATOMIC = 0, 1


def set_atlas_mode(mode: int) -> None:
    """Which of render_atlas.glsl's accumulation variants the oracle restates (refcpu.h)."""
    lib().refcpu_set_atlas_mode(mode)


def _bind_pin_exports(L, prefix: str) -> None:
    """The stage-level exports oracle/refcpu and oracle/glslref share (refcpu.h)."""
    vp = ctypes.c_void_p
    fn = getattr(L, prefix + "_path_vertices")
    fn.argtypes = [ctypes.POINTER(RefFlush), ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, vp] + ([vp] if prefix == "refcpu" else [])
    fn.restype = ctypes.c_int
    fn = getattr(L, prefix + "_path_fragments")
    fn.argtypes = [ctypes.POINTER(RefFlush), ctypes.c_uint32, ctypes.c_uint32, vp, vp, vp]
    fn.restype = ctypes.c_int
    fn = getattr(L, prefix + "_advanced_color_blend_n")
    fn.argtypes = [ctypes.c_uint32, vp, vp, vp, vp, ctypes.c_int]
    fn.restype = None
    fn = getattr(L, prefix + "_cubic_helpers_n")
    fn.argtypes = [ctypes.c_uint32, vp, vp, vp]
    fn.restype = None


def _ptr(a: Optional[np.ndarray]):
    return None if a is None or a.size == 0 else a.ctypes.data


@dataclass
class FlushOutputs:
    desc: T.FlushDesc
    grad: Optional[np.ndarray] = None   # (rows, 512, 4) uint8
    tess: Optional[np.ndarray] = None   # (rows*2048, 4) uint32
    atlas: Optional[np.ndarray] = None  # (h, w) float32


@dataclass
class ReplayResult:
    frames: List[np.ndarray] = field(default_factory=list)       # (H, W, 4) uint8 per TARGET_READ
    flushes: List[FlushOutputs] = field(default_factory=list)


def replay(records: List[T.Record], threads: int = 1, keep_intermediates: bool = True,
           max_frames: Optional[int] = None, only_frames: Optional[set] = None, on_flush=None) -> ReplayResult:
    """Run every flush of a trace through the oracle; returns the frames read
    back (one per RVCT_TARGET_READ) and, per flush, the gradient / tessellation /
    atlas textures the oracle produced. on_flush(rf, outputs, flush_record), if given, runs
    after each flush with the RefFlush struct the oracle was called with (its pointers are
    alive for the duration of the call): the pinning tests use it to run oracle/glslref on
    the very same inputs."""
    L = lib()
    buffers: Dict[int, np.ndarray] = {}
    targets: Dict[int, np.ndarray] = {}
    textures: Dict[int, tuple] = {}
    renderbuffers: Dict[int, tuple] = {}
    tables = None
    keepalive = []
    atlas_size = (0, 0)
    grad_alloc_rows = 0  # height of the last resizeGradientTexture(): what gradTextureY is normalised by
    out = ReplayResult()
    for r in records:
        if r.tag == T.STATIC_TABLES:
            pv = np.ascontiguousarray(r.fields["patch_vertices"])
            pi = np.ascontiguousarray(r.fields["patch_indices"])
            g = np.ascontiguousarray(r.fields["gaussian"])
            ig = np.ascontiguousarray(r.fields["inverse_gaussian"])
            keepalive += [pv, pi, g, ig]
            tables = RefStaticTables(pv.ctypes.data, pi.ctypes.data, g.ctypes.data, ig.ctypes.data)
        elif r.tag == T.BUFFER_UNMAP:
            buffers[r.fields["kind"]] = np.ascontiguousarray(r.data)
        elif r.tag == T.RESIZE_ATLAS:
            atlas_size = (r.fields["width"], r.fields["height"])
        elif r.tag == T.RESIZE_GRADIENT:
            grad_alloc_rows = r.fields["height"]
        elif r.tag == T.TARGET_CREATE:
            targets[r.fields["id"]] = np.zeros((r.fields["height"], r.fields["width"], 4), dtype=np.uint8)
        elif r.tag == T.TARGET_WRITE:
            t = targets[r.fields["id"]]
            t[...] = r.data.reshape(t.shape)
        elif r.tag == T.TEXTURE_CREATE:
            w, h = r.fields["width"], r.fields["height"]
            levels = [np.ascontiguousarray(r.data[: w * h * 4]).reshape(h, w, 4)]
            if r.fields["generate_mips"]:
                n = max(r.fields["mip_level_count"], 1)
                while len(levels) < n:
                    levels.append(_box_downsample(levels[-1]))
            else:
                off = w * h * 4
                lw, lh = w, h
                for _ in range(1, max(r.fields["mip_level_count"], 1)):
                    lw, lh = max(lw // 2, 1), max(lh // 2, 1)
                    levels.append(np.ascontiguousarray(r.data[off: off + lw * lh * 4]).reshape(lh, lw, 4))
                    off += lw * lh * 4
            tex = RefTexture(w, h, len(levels), 0)
            for i, lv in enumerate(levels[:16]):
                tex.levels[i] = lv.ctypes.data
            textures[r.fields["id"]] = (tex, levels)
        elif r.tag == T.RENDERBUFFER_UNMAP:
            data = np.ascontiguousarray(r.data)
            renderbuffers[r.fields["id"]] = (RefRenderBuffer(data.ctypes.data, data.size), data)
        elif r.tag == T.FLUSH and only_frames is not None and len(out.frames) not in only_frames:
            continue  # a frame the caller does not want (frames are independent: each starts with its own load action)
        elif r.tag == T.FLUSH:
            fr: T.FlushRecord = r.fields["flush"]
            d = fr.desc
            target = targets[fr.target_id]
            nb = len(fr.batches)
            batches = (T.DrawBatch * max(nb, 1))()
            for i, b in enumerate(fr.batches):
                ctypes.memmove(ctypes.byref(batches[i]), ctypes.byref(b), ctypes.sizeof(T.DrawBatch))
                if b.image_texture:
                    batches[i].image_texture = ctypes.addressof(textures[b.image_texture][0])
                if b.vertex_buffer:
                    batches[i].vertex_buffer = ctypes.addressof(renderbuffers[b.vertex_buffer][0])
                if b.uv_buffer:
                    batches[i].uv_buffer = ctypes.addressof(renderbuffers[b.uv_buffer][0])
                if b.index_buffer:
                    batches[i].index_buffer = ctypes.addressof(renderbuffers[b.index_buffer][0])
            fills = (T.AtlasBatch * max(len(fr.atlas_fills), 1))(*fr.atlas_fills)
            strokes = (T.AtlasBatch * max(len(fr.atlas_strokes), 1))(*fr.atlas_strokes)
            fo = FlushOutputs(d)
            fo.grad = np.zeros((max(grad_alloc_rows, d.grad_data_height, 1), 512, 4), dtype=np.uint8)
            fo.tess = np.zeros((max(d.tess_data_height, 1) * 2048, 4), dtype=np.uint32)
            aw = max(atlas_size[0], d.feather_atlas_texture_width, 1)
            ah = max(atlas_size[1], d.feather_atlas_texture_height, 1)
            fo.atlas = np.zeros((ah, aw), dtype=np.float32)
            rf = RefFlush()
            rf.desc = ctypes.pointer(d)
            rf.batches = batches
            rf.batch_count = nb
            rf.atlas_fill_batches = fills
            rf.atlas_fill_batch_count = len(fr.atlas_fills)
            rf.atlas_stroke_batches = strokes
            rf.atlas_stroke_batch_count = len(fr.atlas_strokes)
            rf.threads = threads
            for k in range(9):
                rf.buffers[k] = _ptr(buffers.get(k))
            rf.tables = ctypes.pointer(tables)
            rf.target_width = target.shape[1]
            rf.target_height = target.shape[0]
            rf.target_pixels = target.ctypes.data
            rf.grad_texture = fo.grad.ctypes.data
            rf.grad_rows = fo.grad.shape[0]
            rf.tess_texture = fo.tess.ctypes.data
            rf.tess_rows = fo.tess.shape[0] // 2048
            rf.atlas = fo.atlas.ctypes.data
            rf.atlas_width = aw
            rf.atlas_height = ah
            if L.refcpu_flush_run(ctypes.byref(rf)) != 0:
                raise RuntimeError(L.refcpu_last_error().decode())
            if on_flush is not None:
                on_flush(rf, fo, fr)
            if keep_intermediates:
                out.flushes.append(fo)
            else:
                out.flushes.append(FlushOutputs(d))
        elif r.tag == T.TARGET_READ:
            skipped = only_frames is not None and len(out.frames) not in only_frames
            out.frames.append(None if skipped else targets[r.fields["id"]].copy())
            if max_frames is not None and len(out.frames) >= max_frames:
                break
    return out


def _box_downsample(img: np.ndarray) -> np.ndarray:
    """2x2 box filter (what a vkCmdBlitImage linear mip chain produces), with
    odd sizes handled by clamping."""
    h, w, _ = img.shape
    nh, nw = max(h // 2, 1), max(w // 2, 1)
    ys = np.minimum(np.arange(nh) * 2, h - 1)
    ys1 = np.minimum(ys + 1, h - 1)
    xs = np.minimum(np.arange(nw) * 2, w - 1)
    xs1 = np.minimum(xs + 1, w - 1)
    a = img[ys][:, xs].astype(np.uint32) + img[ys][:, xs1] + img[ys1][:, xs] + img[ys1][:, xs1]
    return np.ascontiguousarray(((a + 2) // 4).astype(np.uint8))


def save_png(path: str, rgba_premul: np.ndarray) -> None:
    """Write premultiplied RGBA8 as an (unpremultiplied) PNG for eyeballing."""
    from PIL import Image
    a = rgba_premul[..., 3:4].astype(np.float32)
    rgb = np.where(a > 0, rgba_premul[..., :3].astype(np.float32) * 255.0 / np.maximum(a, 1), 0)
    img = np.concatenate([np.clip(rgb + .5, 0, 255).astype(np.uint8), rgba_premul[..., 3:4]], axis=-1)
    Image.fromarray(img, "RGBA").save(path)


if __name__ == "__main__":
    import time
    recs = T.parse(sys.argv[1])
    t0 = time.time()
    res = replay(recs, threads=int(sys.argv[3]) if len(sys.argv) > 3 else os.cpu_count())
    print("oracle: %d frame(s) in %.3f s" % (len(res.frames), time.time() - t0))
    if len(sys.argv) > 2:
        save_png(sys.argv[2], res.frames[-1])
