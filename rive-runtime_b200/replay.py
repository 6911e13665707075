"""Drive recorded flush traces through the C ABI on a B200.

`Replayer` is the Python-level public entry point of the product path: it owns a
rivecuda context, feeds it the host buffers and FlushDescriptors exactly as
RenderContextCUDAImpl does (map -> write -> unmap -> flush), and reads frames
back. It is what the parity tests, smoke() and bench.py call.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import abi
from . import trace as T


@dataclass
class PreparedFlush:
    desc: T.FlushDesc
    batches: ctypes.Array
    batch_count: int
    fills: ctypes.Array
    fill_count: int
    strokes: ctypes.Array
    stroke_count: int


@dataclass
class FrameStats:
    flushes: int = 0
    timings: List[T.FlushTimings] = field(default_factory=list)


class Replayer:
    """One rivecuda context (one GPU, one stream)."""

    def __init__(self, device: int = 0, lib_path: Optional[str] = None, profiling: bool = False):
        self.lib = abi.load(lib_path)
        ctx = ctypes.c_void_p()
        abi.check(self.lib, self.lib.rivecuda_create(device, ctypes.byref(ctx)), "rivecuda_create")
        self.ctx = ctx
        self.device = device
        self.targets: Dict[int, ctypes.c_void_p] = {}
        self.target_shapes: Dict[int, tuple] = {}
        self.textures: Dict[int, ctypes.c_void_p] = {}
        self.renderbuffers: Dict[int, ctypes.c_void_p] = {}
        self.capacity = [0] * 9
        # Optional: id -> device pointer of caller-owned RGBA8 memory (e.g. a torch tensor
        # later handed to NCCL); TARGET_CREATE records then wrap it instead of allocating.
        self.external_targets: Dict[int, int] = {}
        self.profiling = profiling
        if profiling:
            self.lib.rivecuda_set_profiling(self.ctx, 1)

    def close(self) -> None:
        if self.ctx:
            for t in self.targets.values():
                self.lib.rivecuda_target_destroy(self.ctx, t)
            for t in self.textures.values():
                self.lib.rivecuda_texture_destroy(self.ctx, t)
            for t in self.renderbuffers.values():
                self.lib.rivecuda_renderbuffer_destroy(self.ctx, t)
            self.lib.rivecuda_destroy(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- single ABI calls ---------------------------------------------------

    def _call(self, name: str, *args) -> None:
        abi.check(self.lib, getattr(self.lib, name)(self.ctx, *args), name)

    def upload_buffer(self, kind: int, data: np.ndarray) -> None:
        """map + memcpy into pinned memory + unmap (async H2D), like
        RenderContext::mapResourceBuffers/unmapResourceBuffers."""
        size = int(data.size)
        if size == 0:
            return
        ptr = ctypes.c_void_p()
        self._call("rivecuda_buffer_map", kind, size, ctypes.byref(ptr))
        ctypes.memmove(ptr, data.ctypes.data, size)
        self._call("rivecuda_buffer_unmap", kind, size)

    def sync(self) -> None:
        self._call("rivecuda_sync")

    def read_target(self, target_id: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        h, w = self.target_shapes[target_id]
        if out is None:
            out = np.empty((h, w, 4), dtype=np.uint8)
        self._call("rivecuda_target_read_pixels", self.targets[target_id], out.ctypes.data, out.size)
        return out

    def timings(self) -> T.FlushTimings:
        t = T.FlushTimings()
        self._call("rivecuda_get_flush_timings", ctypes.byref(t))
        return t

    def read_tessellation(self, vertex_count: int) -> np.ndarray:
        out = np.empty((vertex_count, 4), dtype=np.uint32)
        self._call("rivecuda_debug_read_tessellation", out.ctypes.data, 0, vertex_count)
        return out

    def read_gradient(self, rows: int) -> np.ndarray:
        out = np.empty((rows, 512, 4), dtype=np.uint8)
        self._call("rivecuda_debug_read_gradient", out.ctypes.data, rows)
        return out

    # -- trace records --------------------------------------------------------

    def prepare_flush(self, fr: T.FlushRecord) -> PreparedFlush:
        d = T.FlushDesc.from_buffer_copy(fr.desc)
        d.render_target = self.targets[fr.target_id].value
        nb = len(fr.batches)
        batches = (T.DrawBatch * max(nb, 1))()
        for i, b in enumerate(fr.batches):
            ctypes.memmove(ctypes.byref(batches[i]), ctypes.byref(b), ctypes.sizeof(T.DrawBatch))
            batches[i].image_texture = self.textures[b.image_texture].value if b.image_texture else None
            batches[i].vertex_buffer = self.renderbuffers[b.vertex_buffer].value if b.vertex_buffer else None
            batches[i].uv_buffer = self.renderbuffers[b.uv_buffer].value if b.uv_buffer else None
            batches[i].index_buffer = self.renderbuffers[b.index_buffer].value if b.index_buffer else None
        fills = (T.AtlasBatch * max(len(fr.atlas_fills), 1))(*fr.atlas_fills)
        strokes = (T.AtlasBatch * max(len(fr.atlas_strokes), 1))(*fr.atlas_strokes)
        return PreparedFlush(d, batches, nb, fills, len(fr.atlas_fills), strokes, len(fr.atlas_strokes))

    def flush(self, pf: PreparedFlush) -> None:
        self._call("rivecuda_flush", ctypes.byref(pf.desc), pf.batches, pf.batch_count,
                   pf.fills, pf.fill_count, pf.strokes, pf.stroke_count)

    def apply(self, r: T.Record, result: "ReplayResult") -> None:
        """Execute one trace record."""
        tag = r.tag
        if tag == T.STATIC_TABLES:
            pv = np.ascontiguousarray(r.fields["patch_vertices"])
            pi = np.ascontiguousarray(r.fields["patch_indices"])
            g = np.ascontiguousarray(r.fields["gaussian"])
            ig = np.ascontiguousarray(r.fields["inverse_gaussian"])
            self._call("rivecuda_set_static_tables", pv.ctypes.data, pv.size // 32, pi.ctypes.data, pi.size,
                       g.ctypes.data, ig.ctypes.data, g.size)
        elif tag == T.BUFFER_RESIZE:
            self._call("rivecuda_buffer_resize", r.fields["kind"], r.fields["size"])
            self.capacity[r.fields["kind"]] = r.fields["size"]
        elif tag == T.BUFFER_UNMAP:
            self.upload_buffer(r.fields["kind"], r.data)
        elif tag == T.RESIZE_GRADIENT:
            self._call("rivecuda_resize_gradient_texture", r.fields["width"], r.fields["height"])
        elif tag == T.RESIZE_TESSELLATION:
            self._call("rivecuda_resize_tessellation_texture", r.fields["width"], r.fields["height"])
        elif tag == T.RESIZE_ATLAS:
            self._call("rivecuda_resize_feather_atlas_texture", r.fields["width"], r.fields["height"])
        elif tag == T.TARGET_CREATE:
            t = ctypes.c_void_p()
            ext = self.external_targets.get(r.fields["id"])
            if ext is not None:
                self._call("rivecuda_target_wrap", r.fields["width"], r.fields["height"], ctypes.c_void_p(ext), ctypes.byref(t))
            else:
                self._call("rivecuda_target_create", r.fields["width"], r.fields["height"], ctypes.byref(t))
            self.targets[r.fields["id"]] = t
            self.target_shapes[r.fields["id"]] = (r.fields["height"], r.fields["width"])
        elif tag == T.TARGET_DESTROY:
            t = self.targets.pop(r.fields["id"], None)
            if t is not None:
                self.lib.rivecuda_target_destroy(self.ctx, t)
        elif tag == T.TARGET_WRITE:
            data = np.ascontiguousarray(r.data)
            self._call("rivecuda_target_write_pixels", self.targets[r.fields["id"]], data.ctypes.data, data.size)
        elif tag == T.TEXTURE_CREATE:
            t = ctypes.c_void_p()
            data = np.ascontiguousarray(r.data)
            self._call("rivecuda_texture_create", r.fields["width"], r.fields["height"], r.fields["mip_level_count"],
                       data.ctypes.data, r.fields["generate_mips"], ctypes.byref(t))
            self.textures[r.fields["id"]] = t
        elif tag == T.TEXTURE_DESTROY:
            t = self.textures.pop(r.fields["id"], None)
            if t is not None:
                self.lib.rivecuda_texture_destroy(self.ctx, t)
        elif tag == T.RENDERBUFFER_CREATE:
            t = ctypes.c_void_p()
            self._call("rivecuda_renderbuffer_create", r.fields["type"], r.fields["flags"], r.fields["size"], ctypes.byref(t))
            self.renderbuffers[r.fields["id"]] = t
        elif tag == T.RENDERBUFFER_UNMAP:
            ptr = ctypes.c_void_p()
            rb = self.renderbuffers[r.fields["id"]]
            self._call("rivecuda_renderbuffer_map", rb, ctypes.byref(ptr))
            ctypes.memmove(ptr, np.ascontiguousarray(r.data).ctypes.data, r.fields["size"])
            self._call("rivecuda_renderbuffer_unmap", rb)
        elif tag == T.RENDERBUFFER_DESTROY:
            t = self.renderbuffers.pop(r.fields["id"], None)
            if t is not None:
                self.lib.rivecuda_renderbuffer_destroy(self.ctx, t)
        elif tag == T.PREPARE_TO_FLUSH:
            self._call("rivecuda_prepare_to_flush", r.fields["next_frame"], r.fields["safe_frame"])
        elif tag == T.FLUSH:
            pf = self.prepare_flush(r.fields["flush"])
            self.flush(pf)
            result.flush_count += 1
            if result.keep_intermediates:
                d = pf.desc
                fo = FlushOutputs(d)
                if d.tess_data_height:
                    fo.tess = self.read_tessellation(d.tess_data_height * 2048)
                if d.grad_data_height:
                    fo.grad = self.read_gradient(d.grad_data_height)
                if self.profiling:
                    fo.timings = self.timings()
                result.flushes.append(fo)
        elif tag == T.POST_FLUSH:
            self._call("rivecuda_post_flush")
        elif tag == T.TARGET_READ:
            result.frames.append(self.read_target(r.fields["id"]))


@dataclass
class FlushOutputs:
    desc: T.FlushDesc
    tess: Optional[np.ndarray] = None
    grad: Optional[np.ndarray] = None
    timings: Optional[T.FlushTimings] = None


@dataclass
class ReplayResult:
    frames: List[np.ndarray] = field(default_factory=list)
    flushes: List[FlushOutputs] = field(default_factory=list)
    flush_count: int = 0
    keep_intermediates: bool = False


def replay(records: List[T.Record], device: int = 0, keep_intermediates: bool = False,
           profiling: bool = False, lib_path: Optional[str] = None) -> ReplayResult:
    """Render every frame of a trace on the GPU through the C ABI."""
    result = ReplayResult(keep_intermediates=keep_intermediates)
    with Replayer(device, lib_path, profiling) as rp:
        for r in records:
            if r.tag in (T.CREATE, T.DESTROY):
                continue
            rp.apply(r, result)
    return result


def split_frames(records: List[T.Record]):
    """Split a trace into (setup records, frames). Setup = what persists across frames
    (static tables, the render target, textures, mesh buffers) plus ONE resize per buffer
    kind / per-flush texture at the largest size the trace ever asks for, so that looping
    over the frames never reallocates. A frame = (buffer uploads, flush records), ended by
    the trace's TARGET_READ."""
    setup: List[T.Record] = []
    frames = []
    uploads: List[T.Record] = []
    flushes: List[T.Record] = []
    buf_max: Dict[int, T.Record] = {}
    tex_max: Dict[int, T.Record] = {}
    for r in records:
        if r.tag in (T.CREATE, T.DESTROY, T.TARGET_DESTROY, T.TEXTURE_DESTROY, T.RENDERBUFFER_DESTROY,
                     T.PREPARE_TO_FLUSH, T.POST_FLUSH):
            continue
        if r.tag == T.BUFFER_RESIZE:
            k = r.fields["kind"]
            if k not in buf_max or r.fields["size"] > buf_max[k].fields["size"]:
                buf_max[k] = r
        elif r.tag in (T.RESIZE_GRADIENT, T.RESIZE_TESSELLATION, T.RESIZE_ATLAS):
            m = tex_max.get(r.tag)
            if m is None or r.fields["height"] * max(r.fields["width"], 1) > m.fields["height"] * max(m.fields["width"], 1):
                tex_max[r.tag] = r
        elif r.tag == T.BUFFER_UNMAP:
            uploads.append(r)
        elif r.tag == T.FLUSH:
            flushes.append(r)
        elif r.tag == T.TARGET_READ:
            frames.append((uploads, flushes))
            uploads, flushes = [], []
        else:
            setup.append(r)
    if flushes:
        frames.append((uploads, flushes))
    setup = setup + list(buf_max.values()) + list(tex_max.values())
    return setup, frames
