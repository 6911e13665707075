"""The C ABI boundary without a GPU: the library loads, exports every symbol
include/rivecuda.h declares, struct layouts match the ctypes mirrors, the
product path fails loudly (no CPU fallback), and the call recorder round-trips
through the trace reader."""
import ctypes
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rivecuda.h")).read()
    return sorted(set(re.findall(r"\b(rivecuda_[a-z_0-9]+)\s*\(", text)))


def exported(lib_path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    return {line.split()[-1] for line in out.splitlines() if " T " in line}


def test_header_symbols_exported_by_cuda_library(built):
    from rive_runtime_b200 import abi
    syms = declared_symbols()
    assert len(syms) >= 30
    have = exported(abi.DEFAULT_LIB)
    missing = [s for s in syms if s not in have]
    assert not missing, f"librivecuda.so lacks {missing}"
    # and the ctypes table binds exactly the declared set
    assert sorted(abi.SIGNATURES) == syms


def test_header_symbols_exported_by_recorder(built):
    lib = os.path.join(ROOT, "rive-runtime_b200", "_build", "librivecuda_trace.so")
    have = exported(lib)
    missing = [s for s in declared_symbols() if s not in have]
    assert not missing


def test_struct_layouts_match_c(built, tmp_path):
    from rive_runtime_b200 import trace as T
    src = tmp_path / "sz.c"
    src.write_text('#include "rivecuda.h"\n#include <stdio.h>\n#include <stddef.h>\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",sizeof(rivecuda_flush_desc),'
                   'sizeof(rivecuda_draw_batch),sizeof(rivecuda_atlas_batch),sizeof(rivecuda_flush_timings),'
                   'offsetof(rivecuda_flush_desc,first_path),offsetof(rivecuda_flush_desc,tess_data_height),'
                   'offsetof(rivecuda_draw_batch,image_texture));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert sizes[0] == ctypes.sizeof(T.FlushDesc)
    assert sizes[1] == ctypes.sizeof(T.DrawBatch)
    assert sizes[2] == ctypes.sizeof(T.AtlasBatch)
    assert sizes[3] == ctypes.sizeof(T.FlushTimings)
    assert sizes[4] == T.FlushDesc.first_path.offset
    assert sizes[5] == T.FlushDesc.tess_data_height.offset
    assert sizes[6] == T.DrawBatch.image_texture.offset


def test_no_cpu_fallback_without_gpu(built):
    """rivecuda_create must fail loudly when there is no usable B200."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rive_runtime_b200 import abi, replay
    lib = abi.load()
    ctx = ctypes.c_void_p()
    status = lib.rivecuda_create(0, ctypes.byref(ctx))
    assert status != 0 and not ctx.value
    assert b"no CPU fallback" in lib.rivecuda_last_error() or b"CUDA" in lib.rivecuda_last_error()
    with pytest.raises(abi.RiveCudaError):
        replay.Replayer(0)


def test_missing_library_raises(tmp_path):
    from rive_runtime_b200 import abi
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        abi.load(str(tmp_path / "nope.so"))


def test_recorder_roundtrip(built, tmp_path):
    """Drive the recorder through ctypes like RenderContextCUDAImpl would and read
    the trace back."""
    from rive_runtime_b200 import abi, trace as T
    out = tmp_path / "t.rvct"
    os.environ["RIVECUDA_TRACE_OUT"] = str(out)
    lib = abi.load(os.path.join(ROOT, "rive-runtime_b200", "_build", "librivecuda_trace.so"))
    ctx = ctypes.c_void_p()
    assert lib.rivecuda_create(0, ctypes.byref(ctx)) == 0
    assert lib.rivecuda_buffer_resize(ctx, 1, 256) == 0
    p = ctypes.c_void_p()
    assert lib.rivecuda_buffer_map(ctx, 1, 128, ctypes.byref(p)) == 0
    ctypes.memmove(p, bytes(range(128)), 128)
    assert lib.rivecuda_buffer_unmap(ctx, 1, 128) == 0
    assert lib.rivecuda_buffer_map(ctx, 1, 512, ctypes.byref(p)) != 0  # beyond capacity
    t = ctypes.c_void_p()
    assert lib.rivecuda_target_create(ctx, 64, 32, ctypes.byref(t)) == 0
    d = T.FlushDesc()
    d.abi_version = 1
    d.render_target = t.value
    d.path_count = 7
    d.update_bounds[2] = 64
    d.update_bounds[3] = 32
    b = (T.DrawBatch * 1)()
    b[0].draw_type = 0
    b[0].element_count = 5
    assert lib.rivecuda_flush(ctx, ctypes.byref(d), b, 1, None, 0, None, 0) == 0
    px = np.empty((32, 64, 4), np.uint8)
    assert lib.rivecuda_target_read_pixels(ctx, t, px.ctypes.data, px.size) == 0
    lib.rivecuda_target_destroy(ctx, t)
    lib.rivecuda_destroy(ctx)
    recs = T.parse(str(out))
    names = [r.name for r in recs]
    assert names == ["create", "buffer_resize", "buffer_unmap", "target_create", "flush", "target_read",
                     "target_destroy", "destroy"]
    assert bytes(recs[2].data) == bytes(range(128))
    fr = recs[4].fields["flush"]
    assert fr.target_id == 1 and fr.desc.path_count == 7 and fr.batches[0].element_count == 5
    s = T.summarize(recs)
    assert s["frames"] == 1 and s["paths"] == 7 and (s["width"], s["height"]) == (64, 32)


def test_golden_traces_parse():
    from rive_runtime_b200 import trace as T
    from conftest import GOLDEN, golden_traces
    for name in golden_traces(include_large=True):
        recs = T.parse(os.path.join(GOLDEN, name))
        s = T.summarize(recs)
        assert s["flushes"] >= 1 and s["frames"] >= 1, name
        assert T.algorithmic_bytes(recs) > 4 * s["width"] * s["height"] * s["frames"] - 1
    s = T.summarize(T.parse(os.path.join(GOLDEN, "c2_4k.rvct.xz")))
    # BASELINE.json configs[1]: 10k filled cubic paths at 3840x2160.
    assert (s["width"], s["height"], s["paths"]) == (3840, 2160, 10001)
