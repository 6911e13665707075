#!/usr/bin/env python
"""Dev tool (no GPU): for every .riv of a directory, frame N is drawn on the call recorder through
the reference front end (--budget-ms 0) and through CudaPathRenderer + the host build of the
kernels' core, and the records (spans, contours, path / paint / paint-aux records, GradientSpans)
are compared byte for byte (frames with feathers are refused here: RIVECUDA_FRONT_END_NO_DELEGATE;
delegated draws are the reference front end's own) (the comparison of tests/test_front_end_cpu.py).
usage: riv_records_sweep.py <dir> [frame]"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from rive_runtime_b200 import trace as T  # noqa: E402
from test_front_end_cpu import _compare_device_front_end_call_with_reference_trace as compare  # noqa: E402

player = os.path.join(ROOT, "rive-runtime_b200", "_build", "rive_cuda_player")
recorder = os.path.join(ROOT, "rive-runtime_b200", "_build", "librivecuda_trace.so")
frame = sys.argv[2] if len(sys.argv) > 2 else "20"
same = differ = refused = failed = 0
totals = {}
with tempfile.TemporaryDirectory() as tmp:
    reference, call, trace = os.path.join(tmp, "reference.rvct"), os.path.join(tmp, "call.rpf"), os.path.join(tmp, "device.rvct")
    for name in sorted(os.listdir(sys.argv[1])):
        if not name.endswith(".riv"):
            continue
        common = [player, "--scene", "riv:" + os.path.join(sys.argv[1], name), "--frames", frame, "--budget-ms", "0"]
        for f in (reference, call, trace):
            if os.path.exists(f):
                os.remove(f)
        a = subprocess.run(common, env=dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=reference), stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        if a.returncode != 0:
            failed += 1
            print("FAILED", name, a.stderr.decode(errors="replace")[-120:].strip())
            continue
        b = subprocess.run(common + ["--gpu-front-end"], env=dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=trace, RIVECUDA_TRACE_FRONT_END_OUT=call, RIVECUDA_FRONT_END_NO_DELEGATE="1"),
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        if b.returncode != 0:
            refused += 1
            continue
        try:
            recs = T.parse(reference)
            if not any(r.tag == T.FLUSH for r in recs) or not os.path.exists(call):
                same += 1  # nothing drawn by either
                continue
            counts = compare(call, trace, recs)
            same += 1
            for k, v in counts.items():
                totals[k] = totals.get(k, 0) + v
        except AssertionError as e:
            differ += 1
            print("DIFFERS", name, str(e)[:200].replace("\n", " "))
        except Exception as e:  # noqa: BLE001
            differ += 1
            print("ERROR", name, type(e).__name__, str(e)[:200])
print("identical %d, differing %d, refused %d, failed %d; records compared: %s" % (same, differ, refused, failed, totals))
