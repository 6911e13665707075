#!/usr/bin/env python
"""Dev tool (no GPU): like riv_records_sweep.py, for the reference's .sriv silvers (serialized
Renderer call streams, several frames each): the last frame's records through the reference front
end (--budget-ms 0) and through CudaPathRenderer + the host build of the kernels' core, byte for
byte. Frames with feathers are refused here (RIVECUDA_FRONT_END_NO_DELEGATE).
usage: sriv_records_sweep.py <dir with .sriv files>"""
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
same = differ = refused = failed = mismatch = 0
totals = {}
with tempfile.TemporaryDirectory() as tmp:
    reference, call, trace = os.path.join(tmp, "reference.rvct"), os.path.join(tmp, "call.rpf"), os.path.join(tmp, "device.rvct")
    for name in sorted(os.listdir(sys.argv[1])):
        if not name.endswith(".sriv"):
            continue
        common = [player, "--scene", "sriv:" + os.path.join(sys.argv[1], name), "--frames", "0", "--budget-ms", "0"]
        for f in (reference, call, trace):
            if os.path.exists(f):
                os.remove(f)
        a = subprocess.run(common, env=dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=reference), stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        if a.returncode != 0:
            failed += 1
            print("FAILED", name, a.stderr.decode(errors="replace")[-120:].strip())
            continue
        b = subprocess.run(common + ["--gpu-front-end"], env=dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=trace, RIVECUDA_TRACE_FRONT_END_OUT=call,
                                                              RIVECUDA_FRONT_END_NO_DELEGATE="1"), stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        if b.returncode != 0:
            refused += 1
            continue
        try:
            recs = T.parse(reference)
            flushes = sum(1 for r in recs if r.tag == T.FLUSH)
            device_flushes = sum(1 for r in T.parse(trace) if r.tag == T.FLUSH)
            if flushes != device_flushes or not flushes or not os.path.exists(call):
                mismatch += 1
                print("SKIPPED", name, "flushes: reference %d, device %d" % (flushes, device_flushes))
                continue
            counts = compare(call, trace, recs)
            same += 1
            for k, v in counts.items():
                totals[k] = totals.get(k, 0) + v
        except AssertionError as e:
            import traceback
            differ += 1
            print("DIFFERS", name, "at:", traceback.extract_tb(e.__traceback__)[-1].line.strip()[:120])
        except Exception as e:  # noqa: BLE001
            differ += 1
            print("ERROR", name, type(e).__name__, str(e)[:200])
print("identical %d, differing %d, refused %d, skipped %d, failed %d; records compared (last frames): %s" % (same, differ, refused, mismatch, failed, totals))
