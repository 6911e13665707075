"""ctypes binding of the C ABI (include/rivecuda.h) implemented by
_build/librivecuda.so -- the hand-written sm_100a kernels. This is the product
path: there is no CPU fallback; load() raises if the library is missing and
Context() raises if no B200-class device is usable."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

from . import trace as T

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "_build", "librivecuda.so")

_vp = ctypes.c_void_p
_u32 = ctypes.c_uint32
_u64 = ctypes.c_uint64
_sz = ctypes.c_size_t
_int = ctypes.c_int

# name -> (restype, argtypes). Every symbol include/rivecuda.h declares.
SIGNATURES = {
    "rivecuda_create": (_int, [_int, ctypes.POINTER(_vp)]),
    "rivecuda_destroy": (None, [_vp]),
    "rivecuda_last_error": (ctypes.c_char_p, []),
    "rivecuda_abi_version": (_u32, []),
    "rivecuda_set_static_tables": (_int, [_vp, _vp, _u32, _vp, _u32, _vp, _vp, _u32]),
    "rivecuda_buffer_resize": (_int, [_vp, _u32, _sz]),
    "rivecuda_buffer_map": (_int, [_vp, _u32, _sz, ctypes.POINTER(_vp)]),
    "rivecuda_buffer_unmap": (_int, [_vp, _u32, _sz]),
    "rivecuda_resize_gradient_texture": (_int, [_vp, _u32, _u32]),
    "rivecuda_resize_tessellation_texture": (_int, [_vp, _u32, _u32]),
    "rivecuda_resize_feather_atlas_texture": (_int, [_vp, _u32, _u32]),
    "rivecuda_target_create": (_int, [_vp, _u32, _u32, ctypes.POINTER(_vp)]),
    "rivecuda_target_wrap": (_int, [_vp, _u32, _u32, _vp, ctypes.POINTER(_vp)]),
    "rivecuda_target_destroy": (None, [_vp, _vp]),
    "rivecuda_target_read_pixels": (_int, [_vp, _vp, _vp, _sz]),
    "rivecuda_target_write_pixels": (_int, [_vp, _vp, _vp, _sz]),
    "rivecuda_target_read_pixels_async": (_int, [_vp, _vp, _vp, _sz]),
    "rivecuda_target_read_wait": (_int, [_vp, _vp]),
    "rivecuda_target_device_ptr": (_int, [_vp, _vp, ctypes.POINTER(_vp)]),
    "rivecuda_texture_create": (_int, [_vp, _u32, _u32, _u32, _vp, _int, ctypes.POINTER(_vp)]),
    "rivecuda_texture_destroy": (None, [_vp, _vp]),
    "rivecuda_renderbuffer_create": (_int, [_vp, _u32, _u32, _sz, ctypes.POINTER(_vp)]),
    "rivecuda_renderbuffer_destroy": (None, [_vp, _vp]),
    "rivecuda_renderbuffer_map": (_int, [_vp, _vp, ctypes.POINTER(_vp)]),
    "rivecuda_renderbuffer_unmap": (_int, [_vp, _vp]),
    "rivecuda_prepare_to_flush": (_int, [_vp, _u64, _u64]),
    "rivecuda_flush": (_int, [_vp, ctypes.POINTER(T.FlushDesc), ctypes.POINTER(T.DrawBatch), _u32,
                              ctypes.POINTER(T.AtlasBatch), _u32, ctypes.POINTER(T.AtlasBatch), _u32]),
    "rivecuda_post_flush": (_int, [_vp]),
    "rivecuda_sync": (_int, [_vp]),
    "rivecuda_stream": (_int, [_vp, ctypes.POINTER(_vp)]),
    "rivecuda_front_end_paths": (_int, [_vp, _vp, _u32, _vp, _u32, _vp, _u32, _u32, _u32, _vp]),
    "rivecuda_front_end_clip_rects": (_int, [_vp, _vp, _u32]),
    "rivecuda_front_end_path_patches": (_int, [_vp, _vp, _u32]),
    "rivecuda_front_end_gradient_paints": (_int, [_vp, _vp, _u32]),
    "rivecuda_front_end_image_paints": (_int, [_vp, _vp, _u32]),
    "rivecuda_band_unique_id": (_int, [_vp]),
    "rivecuda_band_init": (_int, [_vp, _u32, _u32, _vp]),
    "rivecuda_band_rows": (_int, [_u32, _u32, _u32, ctypes.POINTER(_u32), ctypes.POINTER(_u32)]),
    "rivecuda_band_gather": (_int, [_vp, _vp, _u32]),
    "rivecuda_debug_read_buffer": (_int, [_vp, _u32, _vp, _sz, _sz]),
    "rivecuda_set_profiling": (_int, [_vp, _int]),
    "rivecuda_get_flush_timings": (_int, [_vp, ctypes.POINTER(T.FlushTimings)]),
    "rivecuda_debug_read_tessellation": (_int, [_vp, _vp, _sz, _sz]),
    "rivecuda_debug_read_gradient": (_int, [_vp, _vp, _u32]),
    "rivecuda_debug_read_atlas": (_int, [_vp, _vp, _u32, _u32]),
}

_libs = {}


def load(path: Optional[str] = None) -> ctypes.CDLL:
    """dlopen the ABI library and bind every declared symbol. Raises (never
    falls back) if the library or a symbol is missing."""
    path = path or os.environ.get("RIVECUDA_LIB") or DEFAULT_LIB
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for the renderer.")
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.rivecuda_abi_version() != 1:
        raise RuntimeError("rivecuda ABI version mismatch")
    _libs[path] = lib
    return lib


class RiveCudaError(RuntimeError):
    pass


def check(lib, status: int, what: str) -> None:
    if status != 0:
        msg = lib.rivecuda_last_error()
        raise RiveCudaError(f"{what} failed ({status}): {msg.decode() if msg else '?'}")
