"""GPU path front end (rivecuda_front_end_paths) against the reference's own
front end: for the same RawPaths, the device-generated TessVertexSpan / ContourData / PathData
/ PaintData buffers -- Wang's-formula and polar segment counts, stroke chops (inflections,
180-degree turns, cusps), joins, emulated caps, the frame cull, prefix-summed span offsets,
vertex counts, row wraps, contour midpoints -- must equal, byte for byte, what
PathDraw::initForMidpointFan + pushMidpointFanTessellationData wrote into the mapped buffers
(recorded in the committed flush traces), and the rendered frame must be identical."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["f1", "s1", "c2_4k", "trickycubicstrokes", "emptystroke", "strokes3", "OverStroke", "zero_control_stroke"])
def test_gpu_front_end_matches_reference_front_end(built, name):
    from rive_runtime_b200 import abi, front_end as F, replay as R, trace as T
    abi.load()
    recs = T.parse(os.path.join(GOLDEN, name + ".rvct.xz"))
    dump = F.load_paths(os.path.join(GOLDEN, name + ".paths.xz"))
    assert dump.complete, "the dump must cover every draw of the frame"
    want_frame = R.replay(recs).frames[-1]
    host = {r.fields["kind"]: r.data for r in recs if r.tag == T.BUFFER_UNMAP}
    fr = next(r.fields["flush"] for r in recs if r.tag == T.FLUSH)
    d = fr.desc

    with R.Replayer(0) as rp:
        result = R.ReplayResult()
        for r in recs:
            if r.tag in (T.CREATE, T.DESTROY, T.FLUSH, T.TARGET_READ, T.TARGET_DESTROY, T.POST_FLUSH):
                continue
            if r.tag == T.BUFFER_UNMAP and r.fields["kind"] in (1, 2, 3, 4, 6):
                continue  # path, paint, paintAux, contour, tessSpan: produced on the GPU below
            rp.apply(r, result)
        tc = next(r for r in recs if r.tag == T.TARGET_CREATE)
        res = F.run(rp, dump, tc.fields["width"], tc.fields["height"])
        # a1 / a3: counts and allocation
        assert res.path_count == d.path_count
        assert res.contour_count == d.contour_count
        assert res.tess_vertex_span_count == d.tess_vertex_span_count
        assert res.tess_data_height == d.tess_data_height
        assert len(fr.batches) == 1 and fr.batches[0].draw_type == 0
        assert (res.first_patch, res.patch_count) == (fr.batches[0].base_element, fr.batches[0].element_count)
        # a2: spans (segment counts, x0x1 / y offsets, reflections, wraps) and contours, byte for byte
        n = res.tess_vertex_span_count * 64
        got = F.read_buffer(rp, 6, n).view(np.uint32).reshape(-1, 16)
        want = np.frombuffer(host[6].tobytes()[:n], dtype=np.uint32).reshape(-1, 16)
        bad = np.nonzero((got != want).any(axis=1))[0]
        assert bad.size == 0, f"{bad.size} spans differ, first {bad[:5]}: got {got[bad[0]]} want {want[bad[0]]}"
        n = res.contour_count * 16
        assert np.array_equal(F.read_buffer(rp, 4, n), np.frombuffer(host[4].tobytes()[:n], dtype=np.uint8))
        # a4: path matrices / paint params + colours (defined fields of each record)
        n = res.path_count
        got_path = F.read_buffer(rp, 1, n * 64).view(np.uint32).reshape(-1, 16)
        want_path = np.frombuffer(host[1].tobytes()[:n * 64], dtype=np.uint32).reshape(-1, 16)
        assert np.array_equal(got_path[1:, :8], want_path[1:, :8])
        got_paint = F.read_buffer(rp, 2, n * 8).view(np.uint32).reshape(-1, 2)
        want_paint = np.frombuffer(host[2].tobytes()[:n * 8], dtype=np.uint32).reshape(-1, 2)
        assert np.array_equal(got_paint[1:], want_paint[1:])
        # ... and the frame rendered from the GPU-generated buffers is the same frame.
        pf = rp.prepare_flush(fr)
        rp.flush(pf)
        frame = rp.read_target(fr.target_id)
    assert np.array_equal(frame, want_frame)
