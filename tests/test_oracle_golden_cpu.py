"""The oracle on the committed flush traces (which are the reference front end's
own output, recorded byte for byte): frame hashes are pinned, the front half is
re-derivable from the reference when /root/reference is present, and basic
invariants of the reference's data hold."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_traces
from oracle import refcpu
from rive_runtime_b200 import trace as T

HASHES = os.path.join(GOLDEN, "oracle_frames.json")


def frame_hash(frame):
    return hashlib.sha256(np.ascontiguousarray(frame).tobytes()).hexdigest()


@pytest.mark.parametrize("name", golden_traces())
def test_oracle_frames_are_pinned(name):
    pinned = json.load(open(HASHES))
    recs = T.parse(os.path.join(GOLDEN, name))
    big = name.startswith(("c3", "c1", "anim_", "f1", "s1", "riv_"))
    res = refcpu.replay(recs, threads=(os.cpu_count() or 2) if big else 2, keep_intermediates=False)
    assert [frame_hash(f) for f in res.frames] == pinned[name]["frames"]
    if not big:
        # threading only partitions rows: the result must not depend on it
        res1 = refcpu.replay(recs, threads=1, keep_intermediates=False)
        assert all(np.array_equal(a, b) for a, b in zip(res.frames, res1.frames))


@pytest.mark.parametrize("name", golden_traces())
def test_reference_data_invariants(name):
    """Segment counts / span offsets / vertex counts come from the reference
    itself; check the structural contracts the kernels rely on."""
    recs = T.parse(os.path.join(GOLDEN, name))
    bufs = {}
    for r in recs:
        if r.tag == T.BUFFER_UNMAP:
            bufs[r.fields["kind"]] = r.data
        elif r.tag == T.FLUSH:
            d = r.fields["flush"].desc
            assert d.interlock_mode == 0  # rasterOrdering is the only mode advertised
            if d.tess_vertex_span_count == 0:
                continue
            spans = np.frombuffer(bufs[6].tobytes(), dtype=np.uint32).reshape(-1, 16)[
                d.first_tess_vertex_span:d.first_tess_vertex_span + d.tess_vertex_span_count]
            contour_ids = spans[:, 15] & 0xffff
            assert contour_ids.max() <= d.contour_count
            real = spans[contour_ids > 0]  # contour id 0 = padding vertices
            seg = real[:, 14]
            parametric, polar, join = seg & 0x3ff, (seg >> 10) & 0x3ff, seg >> 20
            assert parametric.max() <= 1023 and polar.max() <= 1023  # kMaxParametric/PolarSegments
            x0x1 = real[:, 12].astype(np.int64)
            x0 = ((x0x1 & 0xffff) ^ 0x8000) - 0x8000
            x1 = x0x1.astype(np.uint32).astype(np.int32) >> 16
            total = parametric.astype(np.int64) + polar + join - 1
            # every span carries at least one vertex and fits a (wrapped) 2048-wide row
            assert np.all(total >= 0)
            assert np.all(np.abs(x1 - x0) <= 2048 + 3 * 1023)
            assert d.tess_data_height <= 2048
            for b in r.fields["flush"].batches:
                if b.draw_type in (0, 1, 2):
                    span = 17 if b.draw_type == 2 else 8
                    assert (b.base_element + b.element_count) * span <= d.tess_data_height * 2048
                    assert b.index_count_per_instance == {0: 72, 1: 120, 2: 249}[b.draw_type]


@pytest.mark.parametrize("name", ["riv_off_road_car.rvct.xz", "riv_bullet_man.rvct.xz", "lots_of_grads_mixed.rvct.xz",
                                  "degengrad.rvct.xz", "verycomplexgrad.rvct.xz"])
def test_gradient_rows_are_normalised_by_the_allocated_texture_height(name):
    """PaintData::set() writes gradTextureY = (row + .5) / gradTextureHeight with the height of the last
    resizeGradientTexture() (render_context.cpp:1442-1443, gpu.cpp:911), NOT this flush's
    gradDataHeight: the recorded paints only land on texel-row centres below gradDataHeight when the
    sampler (oracle and kernels) scales v by the allocated height."""
    recs = T.parse(os.path.join(GOLDEN, name))
    bufs, alloc_rows, seen, differs = {}, 0, 0, False
    for r in recs:
        if r.tag == T.BUFFER_UNMAP:
            bufs[r.fields["kind"]] = r.data
        elif r.tag == T.RESIZE_GRADIENT:
            alloc_rows = r.fields["height"]
        elif r.tag == T.FLUSH:
            d = r.fields["flush"].desc
            if d.grad_data_height == 0:
                continue
            assert d.grad_data_height <= alloc_rows
            differs |= d.grad_data_height != alloc_rows
            paint = np.frombuffer(bufs[2].tobytes(), dtype=np.uint32).reshape(-1, 2)[d.first_paint:d.first_paint + d.path_count + 1]
            grads = paint[np.isin(paint[:, 0] & 0xf, (2, 3))]  # PaintType::linearGradient / radialGradient
            v = grads[:, 1].copy().view(np.float32)
            rows = v * np.float32(alloc_rows) - np.float32(.5)
            assert np.all(np.abs(rows - np.round(rows)) < 1e-3)
            assert np.all((np.round(rows) >= 0) & (np.round(rows) < d.grad_data_height))
            seen += len(grads)
    assert seen > 0
    if name.startswith("riv_off_road_car"):
        assert differs  # the case that tells the two normalisers apart (10 rows allocated, 8 used)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference tree to re-derive the traces")
def test_committed_traces_are_what_the_reference_front_end_emits(built, tmp_path):
    """Bit-exact front half: re-run the reference's RiveRenderer/RenderContext
    (built in place) over RenderContextCUDAImpl + the recorder and compare with
    the committed traces byte for byte."""
    player = os.path.join(ROOT, "rive-runtime_b200", "_build", "rive_cuda_player")
    recorder = os.path.join(ROOT, "rive-runtime_b200", "_build", "librivecuda_trace.so")
    import lzma
    for name, scene, extra in [("beziers", "gm:beziers", []), ("poly_evenOdd", "gm:poly_evenOdd", []),
                               ("feather_shapes", "gm:feather_shapes", []), ("c1", "c1", []),
                               # image GMs: PNG assets decoded by RenderContextCUDAImpl::platformDecodeImageTexture
                               ("image_paint", "gm:image_paint", []), ("mesh", "gm:mesh", [])]:
        out = tmp_path / (name + ".rvct")
        env = dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=str(out))
        subprocess.check_call([player, "--scene", scene, *extra], env=env, stdout=subprocess.DEVNULL)
        with lzma.open(os.path.join(GOLDEN, name + ".rvct.xz")) as f:
            committed = f.read()
        assert out.read_bytes() == committed, name
