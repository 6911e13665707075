"""Reader for flush traces (csrc/rivecuda_trace_format.h).

A trace is the sequence of C-ABI calls (with payloads) that the reference's own
front end -- RiveRenderer -> RenderContext (renderer/src/render_context.cpp) ->
RenderContextCUDAImpl -- made for one or more frames, recorded by
librivecuda_trace.so. Parsing is pure numpy/struct; nothing here computes
pixels.
"""
from __future__ import annotations

import ctypes
import lzma
import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

MAGIC = 0x54435652
VERSION = 1

(CREATE, BUFFER_RESIZE, BUFFER_UNMAP, RESIZE_GRADIENT, RESIZE_TESSELLATION, RESIZE_ATLAS,
 TARGET_CREATE, TARGET_DESTROY, TARGET_READ, TARGET_WRITE, TEXTURE_CREATE, TEXTURE_DESTROY,
 RENDERBUFFER_CREATE, RENDERBUFFER_UNMAP, RENDERBUFFER_DESTROY, PREPARE_TO_FLUSH, FLUSH,
 POST_FLUSH, DESTROY, STATIC_TABLES) = range(1, 21)

TAG_NAMES = {
    1: "create", 2: "buffer_resize", 3: "buffer_unmap", 4: "resize_gradient",
    5: "resize_tessellation", 6: "resize_atlas", 7: "target_create", 8: "target_destroy",
    9: "target_read", 10: "target_write", 11: "texture_create", 12: "texture_destroy",
    13: "renderbuffer_create", 14: "renderbuffer_unmap", 15: "renderbuffer_destroy",
    16: "prepare_to_flush", 17: "flush", 18: "post_flush", 19: "destroy", 20: "static_tables",
}

BUFFER_KINDS = ["flush_uniform", "path", "paint", "paint_aux", "contour", "grad_span",
                "tess_span", "triangle", "image_draw"]
BUFFER_ELEMENT_SIZE = [256, 64, 8, 128, 16, 16, 64, 12, 64]


class FlushDesc(ctypes.Structure):
    """ctypes mirror of rivecuda_flush_desc (include/rivecuda.h)."""
    _fields_ = [
        ("abi_version", ctypes.c_uint32),
        ("interlock_mode", ctypes.c_uint32),
        ("render_target", ctypes.c_void_p),
        ("combined_shader_features", ctypes.c_uint32),
        ("color_load_action", ctypes.c_uint32),
        ("color_clear_value", ctypes.c_uint32),
        ("coverage_clear_value", ctypes.c_uint32),
        ("update_bounds", ctypes.c_int32 * 4),
        ("feather_atlas_texture_width", ctypes.c_uint32),
        ("feather_atlas_texture_height", ctypes.c_uint32),
        ("feather_atlas_content_width", ctypes.c_uint32),
        ("feather_atlas_content_height", ctypes.c_uint32),
        ("flush_uniform_data_offset_in_bytes", ctypes.c_uint64),
        ("path_count", ctypes.c_uint32),
        ("contour_count", ctypes.c_uint32),
        ("grad_span_count", ctypes.c_uint32),
        ("tess_vertex_span_count", ctypes.c_uint32),
        ("first_path", ctypes.c_uint64),
        ("first_paint", ctypes.c_uint64),
        ("first_paint_aux", ctypes.c_uint64),
        ("first_contour", ctypes.c_uint64),
        ("first_grad_span", ctypes.c_uint64),
        ("first_tess_vertex_span", ctypes.c_uint64),
        ("grad_data_height", ctypes.c_uint32),
        ("tess_data_height", ctypes.c_uint32),
        ("clockwise_fill_override", ctypes.c_uint8),
        ("has_triangle_vertices", ctypes.c_uint8),
        ("wireframe", ctypes.c_uint8),
        ("dither_mode", ctypes.c_uint8),
        ("reserved0", ctypes.c_uint32),
    ]


class DrawBatch(ctypes.Structure):
    """ctypes mirror of rivecuda_draw_batch."""
    _fields_ = [
        ("draw_type", ctypes.c_uint32),
        ("shader_misc_flags", ctypes.c_uint32),
        ("draw_contents", ctypes.c_uint32),
        ("shader_features", ctypes.c_uint32),
        ("element_count", ctypes.c_uint32),
        ("base_element", ctypes.c_uint32),
        ("index_count_per_instance", ctypes.c_uint32),
        ("base_index", ctypes.c_uint32),
        ("first_blend_mode", ctypes.c_uint32),
        ("barriers", ctypes.c_uint32),
        ("image_sampler", ctypes.c_uint32),
        ("reserved0", ctypes.c_uint32),
        ("image_texture", ctypes.c_void_p),
        ("vertex_buffer", ctypes.c_void_p),
        ("uv_buffer", ctypes.c_void_p),
        ("index_buffer", ctypes.c_void_p),
    ]


class AtlasBatch(ctypes.Structure):
    """ctypes mirror of rivecuda_atlas_batch."""
    _fields_ = [
        ("scissor_left", ctypes.c_uint16),
        ("scissor_top", ctypes.c_uint16),
        ("scissor_right", ctypes.c_uint16),
        ("scissor_bottom", ctypes.c_uint16),
        ("patch_count", ctypes.c_uint32),
        ("base_patch", ctypes.c_uint32),
    ]


class FlushTimings(ctypes.Structure):
    """ctypes mirror of rivecuda_flush_timings."""
    _fields_ = [
        ("color_ramp_ms", ctypes.c_float),
        ("tessellate_ms", ctypes.c_float),
        ("atlas_ms", ctypes.c_float),
        ("setup_bin_ms", ctypes.c_float),
        ("raster_ms", ctypes.c_float),
        ("total_ms", ctypes.c_float),
        ("kernel_launches", ctypes.c_uint32),
        ("triangle_count", ctypes.c_uint32),
        ("tile_entry_count", ctypes.c_uint32),
        ("raster_kernel", ctypes.c_uint32),
    ]


assert ctypes.sizeof(FlushDesc) == 152, ctypes.sizeof(FlushDesc)
assert ctypes.sizeof(DrawBatch) == 80, ctypes.sizeof(DrawBatch)
assert ctypes.sizeof(AtlasBatch) == 16


@dataclass
class Record:
    tag: int
    # Parsed fields; `data` (if any) is a numpy uint8 view of the payload blob.
    fields: dict = field(default_factory=dict)
    data: Optional[np.ndarray] = None

    @property
    def name(self) -> str:
        return TAG_NAMES.get(self.tag, str(self.tag))


@dataclass
class FlushRecord:
    desc: FlushDesc
    batches: List[DrawBatch]
    atlas_fills: List[AtlasBatch]
    atlas_strokes: List[AtlasBatch]
    target_id: int


def _read_bytes(path: str) -> bytes:
    if path.endswith(".xz"):
        with lzma.open(path, "rb") as f:
            return f.read()
    with open(path, "rb") as f:
        return f.read()


def parse(path_or_bytes) -> List[Record]:
    """Parse a trace file (optionally .xz compressed) into records."""
    raw = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else _read_bytes(path_or_bytes)
    magic, version = struct.unpack_from("<II", raw, 0)
    if magic != MAGIC or version != VERSION:
        raise ValueError("not a rivecuda flush trace (bad magic/version)")
    buf = np.frombuffer(raw, dtype=np.uint8)
    pos = 8
    out: List[Record] = []
    while pos < len(raw):
        tag, _, size = struct.unpack_from("<IIQ", raw, pos)
        pos += 16
        payload = raw[pos:pos + size]
        view = buf[pos:pos + size]
        pos += (size + 7) & ~7
        rec = Record(tag)
        if tag == CREATE:
            rec.fields["device"] = struct.unpack_from("<I", payload)[0]
        elif tag == STATIC_TABLES:
            nv, ni, ng, _ = struct.unpack_from("<IIII", payload)
            o = 16
            rec.fields["patch_vertices"] = view[o:o + nv * 32]
            o += nv * 32
            rec.fields["patch_indices"] = view[o:o + ni * 2].view(np.uint16)
            o += ni * 2 + (2 if ni & 1 else 0)
            rec.fields["gaussian"] = view[o:o + ng * 2].view(np.uint16)
            o += ng * 2
            rec.fields["inverse_gaussian"] = view[o:o + ng * 2].view(np.uint16)
        elif tag == BUFFER_RESIZE:
            kind, _, sz = struct.unpack_from("<IIQ", payload)
            rec.fields.update(kind=kind, size=sz)
        elif tag == BUFFER_UNMAP:
            kind, _, sz = struct.unpack_from("<IIQ", payload)
            rec.fields.update(kind=kind, size=sz)
            rec.data = view[16:16 + sz]
        elif tag in (RESIZE_GRADIENT, RESIZE_TESSELLATION, RESIZE_ATLAS):
            w, h = struct.unpack_from("<II", payload)
            rec.fields.update(width=w, height=h)
        elif tag == TARGET_CREATE:
            i, w, h, _ = struct.unpack_from("<IIII", payload)
            rec.fields.update(id=i, width=w, height=h)
        elif tag in (TARGET_DESTROY, TARGET_READ, TEXTURE_DESTROY, RENDERBUFFER_DESTROY):
            rec.fields["id"] = struct.unpack_from("<I", payload)[0]
        elif tag == TARGET_WRITE:
            i, _, sz = struct.unpack_from("<IIQ", payload)
            rec.fields.update(id=i, size=sz)
            rec.data = view[16:16 + sz]
        elif tag == TEXTURE_CREATE:
            i, w, h, mips, gen, _, sz = struct.unpack_from("<IIIIIIQ", payload)
            rec.fields.update(id=i, width=w, height=h, mip_level_count=mips, generate_mips=gen, size=sz)
            rec.data = view[32:32 + sz]
        elif tag == RENDERBUFFER_CREATE:
            i, t, fl, _, sz = struct.unpack_from("<IIIIQ", payload)
            rec.fields.update(id=i, type=t, flags=fl, size=sz)
        elif tag == RENDERBUFFER_UNMAP:
            i, _, sz = struct.unpack_from("<IIQ", payload)
            rec.fields.update(id=i, size=sz)
            rec.data = view[16:16 + sz]
        elif tag == PREPARE_TO_FLUSH:
            n, s = struct.unpack_from("<QQ", payload)
            rec.fields.update(next_frame=n, safe_frame=s)
        elif tag == FLUSH:
            dsz = ctypes.sizeof(FlushDesc)
            desc = FlushDesc.from_buffer_copy(payload[:dsz])
            nb, nf, ns, _ = struct.unpack_from("<IIII", payload, dsz)
            o = dsz + 16
            bsz = ctypes.sizeof(DrawBatch)
            batches = [DrawBatch.from_buffer_copy(payload[o + i * bsz:o + (i + 1) * bsz]) for i in range(nb)]
            o += nb * bsz
            asz = ctypes.sizeof(AtlasBatch)
            fills = [AtlasBatch.from_buffer_copy(payload[o + i * asz:o + (i + 1) * asz]) for i in range(nf)]
            o += nf * asz
            strokes = [AtlasBatch.from_buffer_copy(payload[o + i * asz:o + (i + 1) * asz]) for i in range(ns)]
            target_id = int(desc.render_target or 0)
            rec.fields["flush"] = FlushRecord(desc, batches, fills, strokes, target_id)
        out.append(rec)
    return out


def summarize(records: List[Record]) -> dict:
    """Counts used by the roofline accounting (SURVEY.md 8d / BASELINE.md 3)."""
    s = dict(flushes=0, paths=0, contours=0, tess_spans=0, grad_spans=0, tess_vertices=0,
             grad_rows=0, batches=0, triangle_vertices=0, frames=0, width=0, height=0)
    tri_bytes = 0
    for r in records:
        if r.tag == FLUSH:
            d = r.fields["flush"].desc
            s["flushes"] += 1
            s["paths"] += d.path_count
            s["contours"] += d.contour_count
            s["tess_spans"] += d.tess_vertex_span_count
            s["grad_spans"] += d.grad_span_count
            s["tess_vertices"] += d.tess_data_height * 2048
            s["grad_rows"] += d.grad_data_height
            s["batches"] += len(r.fields["flush"].batches)
        elif r.tag == BUFFER_UNMAP and r.fields["kind"] == 7:
            tri_bytes += r.fields["size"]
        elif r.tag == TARGET_READ:
            s["frames"] += 1
        elif r.tag == TARGET_CREATE:
            s["width"], s["height"] = r.fields["width"], r.fields["height"]
    s["triangle_vertices"] = tri_bytes // 12
    return s


def algorithmic_bytes(records: List[Record]) -> int:
    """B_alg of BASELINE.md section 3, per trace (all frames in it)."""
    total = 0
    w = h = 0
    for r in records:
        if r.tag == TARGET_CREATE:
            w, h = r.fields["width"], r.fields["height"]
        elif r.tag == FLUSH:
            d = r.fields["flush"].desc
            total += 4 * w * h
            if d.color_load_action == 1:
                total += 4 * w * h
            total += 64 * d.tess_vertex_span_count + 200 * d.path_count + 16 * d.contour_count
            total += 16 * d.grad_span_count + 2 * (512 * 4 * d.grad_data_height)
            total += 2 * 16 * d.tess_data_height * 2048
        elif r.tag == BUFFER_UNMAP and r.fields["kind"] == 7:
            total += r.fields["size"]
        elif r.tag == BUFFER_UNMAP and r.fields["kind"] == 8:
            total += r.fields["size"]
    return total
