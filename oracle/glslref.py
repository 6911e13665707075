"""ctypes binding for oracle/_ref/libglslref.so: the reference's OWN shader sources
(/root/reference/renderer/src/shaders) compiled as C++ by oracle/glslref/Makefile.

TEST INFRASTRUCTURE: the pin that ties oracle/refcpu (the hand-written restatement the GPU
tests compare against) to reference-compiled code. The library can only be built where
/root/reference exists; the built .so is git-ignored and travels with oracle/_ref/.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libglslref.so")
REFERENCE = "/root/reference"

from oracle import refcpu as R  # noqa: E402

_lib = None


def available() -> bool:
    return os.path.exists(_LIB_PATH) or os.path.isdir(os.path.join(REFERENCE, "renderer", "src", "shaders"))


def build() -> str:
    if os.path.isdir(os.path.join(REFERENCE, "renderer", "src", "shaders")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "glslref")])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.glslref_tessellate.argtypes = [ctypes.POINTER(R.RefFlush)]
        _lib.glslref_tessellate.restype = ctypes.c_int
        R._bind_pin_exports(_lib, "glslref")
    return _lib
