"""Screen-band sharding of one very large frame over N GPUs (BASELINE.json
configs[4]): every rank replays the same flush inputs with renderTargetUpdateBounds narrowed
to its band of whole tile rows; ONE NCCL gather of W x (H/N) x 4 bytes per rank composites the
frame on rank 0, which checks it bit for bit against its own single-GPU render of the whole
frame. Used by tools/band_shard_check.py and by bench.py (`c5_bands`, world > 1).
"""
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

from . import abi, replay as R, sharding, trace as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record_scene(scene: str, out: str, extra) -> None:
    build = os.path.join(ROOT, "rive-runtime_b200", "_build")
    env = dict(os.environ, RIVECUDA_LIB=os.path.join(build, "librivecuda_trace.so"), RIVECUDA_TRACE_OUT=out)
    subprocess.check_call([os.path.join(build, "rive_cuda_player"), "--scene", scene, *extra], env=env,
                          stdout=subprocess.DEVNULL)


class BandRenderer:
    """One context per rank, kept alive across repetitions so that device allocations
    (which only grow) happen in the warm-up repetition, not in the timed ones."""

    def __init__(self, records, device, frame_tensor):
        import ctypes
        self.records = records
        self.rp = R.Replayer(device)
        self.first = True
        self.prepared = {}
        sp = ctypes.c_void_p()
        self.rp._call("rivecuda_stream", ctypes.byref(sp))
        self.stream = torch.cuda.ExternalStream(sp.value, device=torch.device("cuda", device))
        self.frame = frame_tensor
        self.before_first_flush = None

    def init_band_gather(self, rank, world):
        """rivecuda_band_init on this rank's context: the 128-byte NCCL id is created by rank 0 and
        handed round with torch.distributed (any channel would do: rivecuda.h)."""
        import ctypes
        import torch.distributed as dist
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            abi.check(self.rp.lib, self.rp.lib.rivecuda_band_unique_id(buf), "rivecuda_band_unique_id")
            ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        ident = ident.to(self.frame.device)
        dist.broadcast(ident, src=0)
        raw = bytes(ident.cpu().numpy().tobytes())
        self.rp._call("rivecuda_band_init", rank, world, ctypes.create_string_buffer(raw, 128))
        self.abi_gather = True

    def render(self, band, gather_root=None):
        """Replay every flush restricted to `band` (rows); returns device ms of the flushes
        (and leaves the event recorded before the first one in self.started). With gather_root,
        rivecuda_band_gather is enqueued behind the flushes on the render stream -- the bands land in
        their rows of the root's target with no host synchronisation in between -- and self.finished
        is the event behind it."""
        rp, result = self.rp, R.ReplayResult()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        started = False
        for r in self.records:
            if r.tag in (T.CREATE, T.DESTROY, T.TARGET_READ, T.TARGET_DESTROY):
                continue
            if r.tag == T.TARGET_CREATE:
                rp.external_targets[r.fields["id"]] = self.frame.data_ptr()
            if r.tag == T.FLUSH:
                if not started:
                    # The frame's inputs are resident before the clock starts, and (bands) every rank
                    # starts its first flush together: the frame time is not to hold the ranks'
                    # host-side skew in getting here.
                    rp.sync()
                    if self.before_first_flush is not None:
                        self.before_first_flush()
                    ev0.record(self.stream)
                    started = True
                # The ctypes mirror of a flush (C5: ~9300 draw batches each) is built once; a host
                # written in C++ passes the reference's own arrays. Only the band changes per call.
                key = id(r)
                if key not in self.prepared:
                    prepared = rp.prepare_flush(r.fields["flush"])
                    self.prepared[key] = (prepared, prepared.desc)
                pf, full_desc = self.prepared[key]
                pf.desc = sharding.restrict_to_band(full_desc, band)
                rp.flush(pf)
                continue
            if not self.first and r.tag not in (T.BUFFER_UNMAP, T.PREPARE_TO_FLUSH, T.POST_FLUSH):
                continue  # targets, textures, tables and sizes persist across repetitions
            rp.apply(r, result)
        ev1.record(self.stream)
        self.finished = ev1
        if gather_root is not None:
            (target,) = rp.targets.values()
            rp._call("rivecuda_band_gather", target, gather_root)
            self.finished = torch.cuda.Event(enable_timing=True)
            self.finished.record(self.stream)
        rp.sync()
        self.first = False
        self.started = ev0
        return ev0.elapsed_time(ev1)

    def close(self):
        self.rp.close()



def run(src: str, reps: int, rank: int, local: int, world: int, scene_args=()):
    """Render `src` (a trace path, or scene:NAME recorded on the spot by the scene player) as
    `world` bands + one gather; returns the measurement dict on rank 0 (None elsewhere).
    Requires an initialised NCCL process group when world > 1."""
    path = src
    if src.startswith("scene:"):
        path = f"/tmp/band_{src[6:]}.rvct"
        if local == 0:
            record_scene(src[6:], path, list(scene_args))
        if world > 1:
            dist.barrier()
    records = T.parse(path)
    s = T.summarize(records)
    W, H = s["width"], s["height"]
    dev = torch.device("cuda", local)
    band = sharding.band_for_rank(H, rank, world)
    frame = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)

    band_ms, gather_ms, gather_min_ms, total_ms = [], [], [], []
    # The composite lives on rank 0 for the whole measurement: the gather's receive buffers are
    # its row ranges, and nothing is allocated inside the timed region.
    composite = None
    renderer = BandRenderer(records, local, frame)
    if world > 1:
        # The gather is the C ABI's (rivecuda_band_gather: NCCL send / recv straight into the rows of
        # the root's target, enqueued on the render stream), as the C++ host uses it.
        import ctypes
        r0, r1 = ctypes.c_uint32(), ctypes.c_uint32()
        abi.check(renderer.rp.lib, renderer.rp.lib.rivecuda_band_rows(H, rank, world, ctypes.byref(r0), ctypes.byref(r1)), "rivecuda_band_rows")
        assert (r0.value, r1.value) == tuple(band), "sharding.band_for_rank and rivecuda_band_rows must agree"
        renderer.init_band_gather(rank, world)
        renderer.before_first_flush = dist.barrier
    for _ in range(reps + 1):  # first repetition is the warm-up
        frame.zero_()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t_render = renderer.render(band, gather_root=0 if world > 1 else None)
        composite = frame  # on rank 0: its own band plus the gathered ones, in place
        # Device-timed on every rank, max over ranks: the band's flushes; the gather as this rank sees
        # it (a rank whose band is done early waits in it for the root, which joins after its own
        # band; the min over ranks is the exchange proper); and the whole frame, first flush to the
        # end of the gather.
        t_frame = renderer.started.elapsed_time(renderer.finished)
        t = torch.tensor([t_render, t_frame - t_render, t_frame], dtype=torch.float64, device=dev)
        tmin = torch.tensor([t_frame - t_render], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        band_ms.append(float(t[0]))
        gather_ms.append(float(t[1]))
        total_ms.append(float(t[2]))
        gather_min_ms.append(float(tmin[0]))
    band_ms, gather_ms, gather_min_ms, total_ms = band_ms[1:], gather_ms[1:], gather_min_ms[1:], total_ms[1:]
    renderer.close()

    line = None
    if rank == 0:
        full = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)
        single_ms = []
        single = BandRenderer(records, local, full)
        for _ in range(reps + 1):
            full.zero_()
            single_ms.append(single.render((0, H)))
        single.close()
        single_ms = single_ms[1:]
        identical = bool(torch.equal(full, composite))
        line = {
            "workload": src, "width": W, "height": H, "paths": s["paths"], "flushes": s["flushes"], "n_gpus": world,
            "bands": [sharding.band_for_rank(H, r, world) for r in range(world)],
            "single_gpu_ms": float(np.mean(single_ms)), "banded_render_ms_max_over_ranks": float(np.mean(band_ms)),
            "render_ms": float(np.mean(band_ms)),
            "gather_ms": float(np.mean(gather_min_ms)), "gather_bytes_per_rank": int(W * (band[1] - band[0]) * 4),
            "gather_ms_max_over_ranks_incl_wait_for_slowest_band": float(np.mean(gather_ms)),
            "frame_ms_max_over_ranks": float(np.mean(total_ms)),
            "speedup_vs_single": float(np.mean(single_ms)) / float(np.mean(total_ms)),
            "composite_identical_to_single_pass": identical, "identical": identical,
        }
        del full
    del frame, composite
    torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
    return line
