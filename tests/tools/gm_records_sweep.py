#!/usr/bin/env python
"""Dev tool (no GPU): like riv_records_sweep.py, for the scene player's GMs (the reference's own
tests/gm sources) and synthetic scenes: each is drawn on the call recorder through the reference
front end (--budget-ms 0) and through CudaPathRenderer + the host build of the kernels' core, and
the records are compared byte for byte (frames with feathers are refused here: RIVECUDA_FRONT_END_NO_DELEGATE;
delegated draws are the reference front end's own). Where both flush several times per frame (a frame
that needs more gradient rows than one texture holds), the last flush is compared; GMs that flush a
different number of times (they drive the RenderContext directly) are skipped. $RIVECUDA_SWEEP_EXTRA: more player arguments (e.g. "--seed 7 --paths 2000" for the f1* scenes).
usage: gm_records_sweep.py [scene ...]"""
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
scenes = sys.argv[1:] or [s for s in subprocess.run([player, "--list"], stdout=subprocess.PIPE).stdout.decode().split() if s.startswith("gm:")]
same = differ = refused = failed = multi = 0
with tempfile.TemporaryDirectory() as tmp:
    reference, call, trace = os.path.join(tmp, "reference.rvct"), os.path.join(tmp, "call.rpf"), os.path.join(tmp, "device.rvct")
    for scene in scenes:
        common = [player, "--scene", scene, "--budget-ms", "0"] + os.environ.get("RIVECUDA_SWEEP_EXTRA", "").split()
        for f in (reference, call, trace):
            if os.path.exists(f):
                os.remove(f)
        a = subprocess.run(common, env=dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=reference), stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        if a.returncode != 0:
            failed += 1
            print("FAILED", scene, a.stderr.decode(errors="replace")[-120:].strip())
            continue
        b = subprocess.run(common + ["--gpu-front-end"], env=dict(os.environ, RIVECUDA_LIB=recorder, RIVECUDA_TRACE_OUT=trace, RIVECUDA_TRACE_FRONT_END_OUT=call, RIVECUDA_FRONT_END_NO_DELEGATE="1"),
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        if b.returncode != 0:
            refused += 1
            why = [l for l in b.stderr.decode(errors="replace").splitlines() if "the frame contains" in l]
            print("REFUSED", scene, why[-1].split("contains")[-1].strip() if why else "")
            continue
        try:
            recs = T.parse(reference)
            flushes = [r for r in recs if r.tag == T.FLUSH]
            device_flushes = [r for r in T.parse(trace) if r.tag == T.FLUSH]
            if len(flushes) != len(device_flushes) or not flushes:
                multi += 1
                print("SKIPPED", scene, "flushes: reference %d, device %d" % (len(flushes), len(device_flushes)))
                continue
            counts = compare(call, trace, recs)
            same += 1
            print("identical", scene, counts)
        except AssertionError as e:
            import traceback
            differ += 1
            tb = traceback.extract_tb(e.__traceback__)[-1]
            draw_types = sorted(set(b.draw_type for b in flushes[0].fields["flush"].batches))
            print("DIFFERS", scene, "at:", tb.line.strip()[:110], "| reference draw types", draw_types)
        except Exception as e:  # noqa: BLE001
            differ += 1
            print("ERROR", scene, type(e).__name__, str(e)[:200])
print("identical %d, differing %d, refused %d, skipped (several flushes) %d, failed %d" % (same, differ, refused, multi, failed))
