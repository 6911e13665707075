#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 Rive back end.

A "step" is one pass of the hot path (RenderContextImpl::flush: colour ramps ->
tessellation -> feather atlas -> draw list) over one frame of the workload
BASELINE.json's metric is quoted on: configs[1], "synthetic 10k random filled
cubic paths, nonZero/evenOdd, 3840x2160" -- the flush trace the reference's own
front end produced for that scene (tests/golden/c2_4k.rvct.xz).

  value  frames/s with every input already resident in HBM (flush only)
  e2e    frames/s through the C ABI with HOST buffers: every step maps + fills the
         pinned ring buffers (H2D inside the timed region), flushes, and reads the
         4K RGBA8 frame back to pinned host memory (D2H inside the timed region)
  roofline  algorithmic bytes of the frame over the raster
         kernel's CUDA-event time, against the measured HBM peak
  cpu_baseline  the CPU oracle (a port of the reference's shaders) on the host cores

`--impl reference` times that CPU implementation alone (rank 0 only).
Multi-GPU (`torchrun ... bench.py --gpus N`): frames are independent units, so
each rank renders its own frames with no data-path collective (weak scaling).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name -> (description, trace, metric)
    "c2": ("c2: 10k random filled 4-cubic paths, nonZero/evenOdd, 3840x2160 (BASELINE.json configs[1])",
           os.path.join(ROOT, "tests", "golden", "c2_4k.rvct.xz"), "frames/sec at 4K (device-timed)"),
    # BASELINE.json configs[3]: an artboard state-machine animation at 1080p. The frames are what the
    # reference's unmodified core runtime produced for its own off_road_car.riv test asset (File::import ->
    # StateMachineInstance::advanceAndApply(1/60) -> Artboard::draw -> RiveRenderer; host/player `riv:`),
    # 60 frames per pass; every GPU renders its own instance of the animation.
    "c4": ("c4: off_road_car.riv artboard state-machine animation, 60 frames per pass at 1920x1080 "
           "(BASELINE.json configs[3])",
           os.path.join(ROOT, "tests", "golden", "riv_off_road_car.rvct.xz"), "frames/sec at 1080p (device-timed)"),
    # The same through a recorded renderer-call stream (.sriv silver) of a UI-style artboard with text.
    "c4sriv": ("c4sriv: db_health_tracker artboard animation (.sriv stream), 63 frames per pass at 1920x1080",
               os.path.join(ROOT, "tests", "golden", "anim_db_health_tracker.rvct.xz"), "frames/sec at 1080p (device-timed)"),
}
UNIT = "frames/s"


def kernel_sources_sha256():
    """Hash of the CUDA sources the tile rasterisers are built from (ties an ncu capture to a build)."""
    import hashlib
    h = hashlib.sha256()
    src = os.path.join(ROOT, "rive-runtime_b200", "csrc")
    for name in ("kernels_draw.cu", "raster_tiles.cuh", "raster_tiles_exact.cuh", "raster_tiles_span.cuh", "device_math.cuh", "rivecuda_internal.h"):
        with open(os.path.join(src, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.stop = threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                               "--format=csv,noheader,nounits"], text=True, timeout=5)
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:  # noqa: BLE001
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *exc):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = max(int(s[1]) for s in self.samples if s[1].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons}


PLAYER = os.path.join(ROOT, "rive-runtime_b200", "_build", "rive_cuda_player")


def reference_front_end(scene: str, frames: int = 12, extra=()):
    """The reference's OWN CPU front end on this host, single-threaded as the class is:
    RiveRenderer + RenderContext::flush over the reference's RenderContextNULL (what its
    tests/bench/draw_pls_path.cpp:25 measures), compiled in place from /root/reference into the
    scene player (`--null-backend`). None where the player binary was not built."""
    if not os.path.exists(PLAYER):
        return None
    try:
        out = subprocess.check_output([PLAYER, "--scene", scene, "--null-backend", "--frames", str(frames), *extra], text=True, timeout=300)
        j = json.loads(out.strip().splitlines()[-1])
        return {"front_end_ms": j["front_end_min_ms"], "scene_build_ms": j["scene_build_min_ms"], "frames": j["frames"], "cores": 1,
                "kind": "reference", "what": "RiveRenderer + RenderContext::flush on RenderContextNULL (reference code, built in place), "
                                             "min over frames, scene construction subtracted"}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:200]}


def riv_front_end_host_time(asset: str = "off_road_car", frames: int = 300):
    """SURVEY 8 f1 on real content, from the C++ host: host milliseconds per frame of the scene player
    animating a .riv file (the reference's unmodified core runtime) at 1080p and drawing it through
    RiveRenderer (the reference's CPU front end) and through CudaPathRenderer (--gpu-front-end: RawPaths
    to rivecuda_front_end_paths), both into the same CUDA backend, each frame read back (8 MB D2H).
    The frames of the two are bit-identical (tests/test_front_end_gpu.py). None without the player
    binary or the asset (tools/fetch_riv_assets.sh)."""
    path = os.path.join(ROOT, "tests", "_riv_assets", asset + ".riv")
    if not os.path.exists(PLAYER) or not os.path.exists(path):
        return None
    env = dict(os.environ)
    env.setdefault("RIVECUDA_LIB", os.path.join(ROOT, "rive-runtime_b200", "_build", "librivecuda.so"))
    out = {"asset": asset + ".riv", "frames": frames, "width": 1920, "height": 1080,
           "what": "rive_cuda_player --scene riv:... --budget-ms 0 [--gpu-front-end]: host ms per frame incl. the state machine's "
                   "advance, Artboard::draw, the flush and the read-back of the frame"}
    try:
        for key, extra in (("reference_front_end_host_ms_per_frame", []), ("device_front_end_host_ms_per_frame", ["--gpu-front-end"])):
            text = subprocess.check_output([PLAYER, "--scene", "riv:" + path, "--frames", str(frames), "--budget-ms", "0", *extra],
                                           text=True, timeout=120, env=env, stderr=subprocess.DEVNULL)
            out[key] = json.loads(text.strip().splitlines()[-1])["host_ms_per_frame"]
    except Exception as e:  # noqa: BLE001
        out["error"] = str(e)[:200]
    return out


def run_reference(args, rank: int) -> None:
    """--impl reference: the reference's pixel stage cannot be built here (GLSL ->
    SPIR-V -> Vulkan/SwiftShader), so this arm times its CPU port,
    the oracle, on all host threads. One step = one full 4K frame."""
    if rank != 0:
        return
    from oracle import refcpu
    from rive_runtime_b200 import trace as T
    workload, trace_path, metric = WORKLOADS[args.workload]
    records = T.parse(trace_path)
    summary = T.summarize(records)
    n_frames = max(summary["frames"], 1)
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        refcpu.replay(records, threads=cores, keep_intermediates=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        refcpu.replay(records, threads=cores, keep_intermediates=False)
    dt = time.perf_counter() - t0
    fps = args.steps * n_frames / dt
    line = {
        "impl": "reference", "metric": metric, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "width": summary["width"], "height": summary["height"], "frames_per_step_per_gpu": n_frames,
                   "paths": summary["paths"] // n_frames, "tess_vertices": summary["tess_vertices"] // n_frames,
                   "mpixels_per_s": fps * summary["width"] * summary["height"] / 1e6,
                   "parallelism": "host threads of rank 0 (the other ranks exit without work)"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} pass(es) over the workload's {n_frames} full-size frame(s), oracle (CPU "
                                   "restatement of the reference shaders, bit-identical to the reference's shader sources compiled "
                                   "as C++: tests/test_oracle_glslref_cpu.py), all host threads",
                         "front_end": reference_front_end("c2") if args.workload == "c2" else None},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def sharded_configs(args, rank: int, local_rank: int, world: int):
    """BASELINE.json configs[3] and configs[4], the two ways the path shards, measured
    with `world` ranks (device-timed, max over ranks):

    c4        1000 animation frames at 1080p, frame i -> rank i mod N (sharding.frames_for_rank):
              each rank uploads and renders only its own frames; no collective on the render path.
              Then ONE NCCL gather brings every rank's last frame to rank 0, which checks each
              against its own render of the same frame index (bit-identical).
    c5_bands  one 16384x16384 frame of 200k paths as N screen bands + one NCCL gather
              (rive_runtime_b200.band_render); rank 0 also renders the whole frame alone and checks
              the composite bit for bit."""
    import torch
    import torch.distributed as dist
    from rive_runtime_b200 import band_render, replay as R, sharding, trace as T
    dev = torch.device("cuda", local_rank)
    out = {}

    # ---- c4: frames sharded round-robin -------------------------------------------------
    records = T.parse(WORKLOADS["c4"][1])
    summary = T.summarize(records)
    width, height = summary["width"], summary["height"]
    setup, trace_frames = R.split_frames(records)
    total_frames = 1000
    rp = R.Replayer(device=local_rank)
    result = R.ReplayResult()
    for r in setup:
        rp.apply(r, result)
    frames = []
    target_id = None
    for ups, fls in trace_frames:
        frames.append(([(u.fields["kind"], np.ascontiguousarray(u.data)) for u in ups], [rp.prepare_flush(f.fields["flush"]) for f in fls]))
        target_id = fls[0].fields["flush"].target_id
    stream_ptr = ctypes.c_void_p()
    rp._call("rivecuda_stream", ctypes.byref(stream_ptr))
    stream = torch.cuda.ExternalStream(stream_ptr.value, device=dev)

    def render(index):
        ups, fls = frames[index % len(frames)]
        for kind, data in ups:
            rp.upload_buffer(kind, data)
        for pf in fls:
            rp.flush(pf)

    mine = sharding.frames_for_rank(total_frames, rank, world)
    for i in mine[:8]:
        render(i)
    rp.sync()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for i in mine:
        render(i)
    e1.record(stream)
    rp.sync()
    wall = time.perf_counter() - t0
    t = torch.tensor([e0.elapsed_time(e1) / 1e3, wall], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # The data-plane collective: each rank's last frame -> rank 0 over NVLink.
    last = torch.from_numpy(rp.read_target(target_id)).to(dev)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gathered = [torch.empty_like(last) for _ in range(world)] if rank == 0 else None
    dist.gather(last, gathered, dst=0)  # untimed: NCCL sets its peer connections up lazily
    torch.cuda.synchronize()
    g0.record()
    dist.gather(last, gathered, dst=0)
    g1.record()
    torch.cuda.synchronize()
    identical = None
    if rank == 0:
        identical = True
        for r in range(world):
            idx = sharding.frames_for_rank(total_frames, r, world)[-1]
            render(idx)
            rp.sync()
            identical = identical and bool(np.array_equal(rp.read_target(target_id), gathered[r].cpu().numpy()))
    rp.close()
    out["c4"] = {"workload": "c4: 1000 frames of the off_road_car.riv state-machine animation at 1920x1080, frame i -> rank i mod N, "
                             "each frame's inputs uploaded inside the timed region (BASELINE.json configs[3])",
                 "frames": total_frames, "frames_per_rank": len(mine), "value": total_frames / float(t[0]), "unit": UNIT,
                 "e2e_value": total_frames / float(t[1]), "device_s_max_over_ranks": float(t[0]),
                 "gather_ms": g0.elapsed_time(g1), "gather_bytes_per_rank": int(last.numel()), "identical": identical}

    # ---- c5: one huge frame as screen bands + NCCL gather ----------------------------------
    if not args.no_c5 and os.path.exists(PLAYER):
        try:
            line = band_render.run("scene:c5", 1, rank, local_rank, world)
            if rank == 0:
                out["c5_bands"] = line
        except Exception as e:  # noqa: BLE001
            out["c5_bands"] = {"error": str(e)[:300]}
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-sharded", action="store_true", help="world > 1: skip the c4 / c5_bands sharded configurations")
    ap.add_argument("--no-c5", action="store_true", help="world > 1: skip c5_bands (records a 370 MB trace on the box)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from rive_runtime_b200 import replay as R, trace as T

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the renderer has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    workload, trace_path, metric = WORKLOADS[args.workload]
    records = T.parse(trace_path)
    summary = T.summarize(records)
    width, height = summary["width"], summary["height"]
    setup, trace_frames = R.split_frames(records)
    n_frames = len(trace_frames)
    alg_bytes = T.algorithmic_bytes(records) / n_frames  # per frame

    rp = R.Replayer(device=local_rank)
    result = R.ReplayResult()
    for r in setup:
        rp.apply(r, result)
    frames = []  # per frame: ([(kind, host bytes)], [PreparedFlush])
    target_id = None
    for ups, fls in trace_frames:
        frames.append(([(u.fields["kind"], np.ascontiguousarray(u.data)) for u in ups],
                       [rp.prepare_flush(f.fields["flush"]) for f in fls]))
        target_id = fls[0].fields["flush"].target_id
    h2d_bytes = int(sum(d.size for ups, _ in frames for _, d in ups)) // n_frames
    d2h_bytes = width * height * 4
    # Two render targets + two pinned host frames: the read-back of frame k overlaps frame k+1.
    targets = [rp.targets[target_id], ctypes.c_void_p()]
    rp._call("rivecuda_target_create", width, height, ctypes.byref(targets[1]))
    rp.targets[-1] = targets[1]
    host_frames = [torch.empty((height, width, 4), dtype=torch.uint8, pin_memory=True) for _ in range(2)]

    stream_ptr = ctypes.c_void_p()
    rp._call("rivecuda_stream", ctypes.byref(stream_ptr))
    stream = torch.cuda.ExternalStream(stream_ptr.value, device=torch.device("cuda", local_rank))
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")
    resident = n_frames == 1  # one frame: its inputs stay in HBM; an animation uploads every frame's own inputs

    def render_frame(k, target=None):
        ups, fls = frames[k]
        if not resident:
            for kind, data in ups:
                rp.upload_buffer(kind, data)   # map + memcpy into the pinned ring + async H2D
        for pf in fls:
            if target is not None:
                pf.desc.render_target = target.value
            rp.flush(pf)

    def step_all_frames():
        for k in range(n_frames):
            render_frame(k)

    def barrier():
        rp.sync()
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    if resident:
        for kind, data in frames[0][0]:
            rp.upload_buffer(kind, data)

    # ---- value: device-timed; single-frame workloads have every input resident in HBM ----
    for _ in range(args.warmup):
        step_all_frames()
    barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
        barrier()
        for i in range(args.steps):
            with torch.cuda.stream(stream):
                l2_flush.fill_(i & 0xff)  # evict L2 between timed iterations (outside the event pair)
            starts[i].record(stream)
            step_all_frames()
            ends[i].record(stream)
        barrier()
    device_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    t = torch.tensor([device_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    device_ms = float(t.item())
    value = world * n_frames * args.steps / (device_ms / 1e3)

    # ---- the same with every step's inputs uploaded inside the device-timed region (BASELINE.md
    # section 3 counts the input H2D; `value` follows the bench contract: inputs resident) ----
    value_h2d = value
    if resident:
        def step_with_uploads():
            for kind, data in frames[0][0]:
                rp.upload_buffer(kind, data)
            step_all_frames()
        for _ in range(3):
            step_with_uploads()
        barrier()
        s2 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        e2 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        for i in range(args.steps):
            with torch.cuda.stream(stream):
                l2_flush.fill_(i & 0xff)
            rp.sync()  # the upload stream is ordered behind the render stream from here on
            s2[i].record(stream)
            step_with_uploads()
            e2[i].record(stream)
        barrier()
        ms2 = sum(a.elapsed_time(b) for a, b in zip(s2, e2))
        t = torch.tensor([ms2], dtype=torch.float64, device=f"cuda:{local_rank}")
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        value_h2d = world * n_frames * args.steps / (float(t.item()) / 1e3)

    # ---- roofline: the dominant kernel (tile raster), CUDA events on its stream --
    rp.lib.rivecuda_set_profiling(rp.ctx, 1)
    raster_ms, setup_ms, tess_ms, launches = [], [], [], 0
    raster_kernels = set()
    tri_count = entry_count = 0
    for _ in range(3):
        r_ms = s_ms = t_ms = 0.0
        launches = tri_count = entry_count = 0
        for k in range(n_frames):
            ups, fls = frames[k]
            if not resident:
                for kind, data in ups:
                    rp.upload_buffer(kind, data)
            for pf in fls:
                rp.flush(pf)
                tm = rp.timings()
                r_ms += tm.raster_ms
                s_ms += tm.setup_bin_ms
                t_ms += tm.tessellate_ms
                launches += tm.kernel_launches
                raster_kernels.add(("raster_tiles_kernel", "raster_tiles_exact_kernel", "raster_spans_kernel")[min(tm.raster_kernel, 2)])
                tri_count += tm.triangle_count
                entry_count += tm.tile_entry_count
        raster_ms.append(r_ms / n_frames)
        setup_ms.append(s_ms / n_frames)
        tess_ms.append(t_ms / n_frames)
    rp.lib.rivecuda_set_profiling(rp.ctx, 0)
    raster = float(np.mean(raster_ms))
    peak, peak_src = measured_peak_gbs()
    traffic = None
    if args.workload == "c2":
        # dram__bytes_read + dram__bytes_write of this kernel on this workload from this round's
        # `ncu --set full` capture (profiles/r02_raster_traffic.json, written by `tools/ncu_summary.py
        # --traffic-json` from the .ncu-rep). It only counts if it was captured from the kernel
        # sources being benchmarked (`sources_sha256`) and names the kernel that ran.
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_raster_traffic.json")))
            if tj["sources_sha256"] == kernel_sources_sha256() and {tj["kernel"]} == raster_kernels:
                traffic = int(tj["dram_bytes_read"]) + int(tj["dram_bytes_write"])
        except Exception:  # noqa: BLE001
            traffic = None
    achieved = alg_bytes / (raster / 1e3) / 1e9

    # ---- e2e: host buffers in, host frame out, through the C ABI ----------------
    # Every frame: map + fill the pinned rings (H2D), flush, read the RGBA8 frame back
    # to pinned host memory (D2H). Pipelined the way a presentation loop is: frame k's
    # read-back (copy stream) overlaps frame k+1's rendering into the other target;
    # every read-back completes inside the timed region.
    def e2e_pass(pipelined):
        n = 0
        for _ in range(args.steps):
            for k in range(n_frames):
                ups, fls = frames[k]
                for kind, data in ups:
                    rp.upload_buffer(kind, data)
                slot = n & 1
                render_frame_into = targets[slot]
                for pf in fls:
                    pf.desc.render_target = render_frame_into.value
                    rp.flush(pf)
                if pipelined:
                    rp._call("rivecuda_target_read_pixels_async", render_frame_into, host_frames[slot].data_ptr(), d2h_bytes)
                    if n > 0:
                        rp._call("rivecuda_target_read_wait", targets[slot ^ 1])
                else:
                    rp._call("rivecuda_target_read_pixels", render_frame_into, host_frames[slot].data_ptr(), d2h_bytes)
                n += 1
        if pipelined:
            rp._call("rivecuda_target_read_wait", targets[(n - 1) & 1])

    e2e = {}
    for mode in ("serial", "pipelined"):
        saved_steps = args.steps
        args.steps = min(args.warmup, 3)
        e2e_pass(mode == "pipelined")
        args.steps = saved_steps
        barrier()
        t0 = time.perf_counter()
        e2e_pass(mode == "pipelined")
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local_rank}")
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e[mode] = world * n_frames * args.steps / float(t.item())
    for pf in (p for _, fls in frames for p in fls):
        pf.desc.render_target = targets[0].value
    e2e_fps = e2e["pipelined"]

    # ---- raw paths in, host frame out: the GPU path front end ------
    # Same frame, but the host hands over only the RawPaths (verbs + points), matrices
    # and colours (1.5 MB); segment counts, span allocation and all per-path records are
    # produced on the device (rivecuda_front_end_paths), then the same flush runs.
    raw_paths = None
    dump_path = os.path.join(ROOT, "tests", "golden", "c2_4k.paths.xz")
    if args.workload == "c2" and os.path.exists(dump_path):
        from rive_runtime_b200 import front_end as F
        dump = F.load_paths(dump_path)
        pf = frames[0][1][0]

        def raw_pass(steps):
            for n in range(steps):
                slot = n & 1
                F.run(rp, dump, width, height)
                pf.desc.render_target = targets[slot].value
                rp.flush(pf)
                rp._call("rivecuda_target_read_pixels_async", targets[slot], host_frames[slot].data_ptr(), d2h_bytes)
                if n > 0:
                    rp._call("rivecuda_target_read_wait", targets[slot ^ 1])
            rp._call("rivecuda_target_read_wait", targets[(steps - 1) & 1])

        raw_pass(3)
        barrier()
        t0 = time.perf_counter()
        raw_pass(args.steps)
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local_rank}")
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item()) / world  # whole-job frames/s: every rank rendered args.steps frames
        t0 = time.perf_counter()
        for _ in range(args.steps):
            F.run(rp, dump, width, height)
        rp.sync()
        fe_ms = (time.perf_counter() - t0) / args.steps * 1e3
        pf.desc.render_target = targets[0].value
        for kind, data in frames[0][0]:
            rp.upload_buffer(kind, data)
        raw_paths = {"value": args.steps / dt, "unit": UNIT, "front_end_ms": fe_ms,
                     "h2d_bytes_per_step": int(dump.points.nbytes + dump.verbs.nbytes + dump.paths.nbytes),
                     "d2h_bytes_per_step": d2h_bytes,
                     "note": "RawPaths + matrices + colours in (host), RGBA8 frame out (host): rivecuda_front_end_paths (Wang's "
                             "formula counts, warp-scan span allocation, span/contour/path records on the device; byte-identical "
                             "to the reference front end, tests/test_front_end_gpu.py) + the same flush; front_end_ms includes "
                             "the H2D of the paths and one stream sync; cpu_baseline.front_end times the reference's own "
                             "CPU front end on the same frame."}

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload ---
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import refcpu
        cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        n = 0
        while n < 2 or (time.perf_counter() - t0 < 10 and n < 8):
            refcpu.replay(records, threads=cores, keep_intermediates=False)
            n += 1
        cpu_dt = time.perf_counter() - t0
        cpu = {"value": n * n_frames / cpu_dt, "unit": UNIT, "cores": cores, "kind": "port",
               "front_end": reference_front_end("c2") if args.workload == "c2" else None,
               "sample": f"{n} pass(es) over the workload's {n_frames} full-size frame(s) on the oracle (CPU restatement of "
                         "the reference shaders, bit-identical to the reference's shader sources compiled as C++; the reference's "
                         "own pixel stage needs Vulkan/SwiftShader, unbuildable here); front_end = the reference's own CPU front "
                         "end (RiveRenderer + RenderContext::flush, one thread) on the same frame"}

    riv_front_end = None
    if rank == 0 and world == 1 and args.workload == "c2":
        rp.sync()
        riv_front_end = riv_front_end_host_time()

    sharded = None
    if distributed and not args.no_sharded and args.workload == "c2":
        rp.sync()
        sharded = sharded_configs(args, rank, local_rank, world)

    if rank == 0:
        inputs = ("inputs resident in HBM" if resident else
                  "each frame's own inputs are uploaded (pinned ring -> H2D) inside the device-timed region")
        line = {
            "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": device_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+i32", "data": "synthetic",
            "config": {"workload": workload, "width": width, "height": height, "frames_per_step_per_gpu": n_frames,
                       "paths": summary["paths"] // n_frames, "tess_vertices": summary["tess_vertices"] // n_frames,
                       "raw_triangles": int(tri_count) // n_frames, "tile_entries": int(entry_count) // n_frames,
                       "mpixels_per_s": value * width * height / 1e6, "inputs": inputs,
                       "value_with_input_h2d": value_h2d,
                       "l2": "256 MiB written between timed iterations (outside the per-step event pairs); "
                             "the per-frame working set (triangle records + tile lists) also exceeds L2 on c2",
                       "parallelism": f"independent frames / artboard instances on {world} GPU(s) (every GPU renders the "
                                      "workload's frames), no data-path collective"},
            "e2e": {"value": e2e_fps, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes * n_frames,
                    "d2h_bytes_per_step": d2h_bytes * n_frames, "pipelined": True, "serial_value": e2e["serial"],
                    "note": "per frame: H2D of its inputs from pinned rings, flush, D2H of the RGBA8 frame to pinned memory; "
                            "frame k's read-back overlaps frame k+1's rendering (two targets); serial_value waits for each "
                            "read-back before the next frame"},
            "raw_paths_e2e": raw_paths,
            "riv_front_end": riv_front_end,
            "gpu_launches": int(launches) * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "+".join(sorted(raster_kernels)), "kernel_ms": raster,
                         "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                         "other_kernels_ms": {"tessellate": float(np.mean(tess_ms)), "setup_bin_sort": float(np.mean(setup_ms))},
                         "note": "per frame; the kernel is instruction-issue bound, not HBM bound (DESIGN.md section 5): "
                                 "ncu smsp__issue_active 70 % of peak, dram throughput 1.4 % of peak on c2 "
                                 "(profiles/r02_ncu_summary.txt); traffic = dram bytes of one launch from this round's ncu capture "
                                 "(profiles/r02_raster_traffic.json), null if that capture is of other sources"},
            "cpu_baseline": cpu,
            "clocks": clocks.summary(),
        }
        if sharded is not None:
            line["sharded"] = sharded
            line["config"]["parallelism"] += ("; `sharded` holds the configurations that partition work across the ranks: c4 (1000 frames, "
                                              "frame i -> rank i mod N) and c5_bands (one 16384^2 frame as screen bands + NCCL gather)")
        print(json.dumps(line), flush=True)
    rp.close()
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
