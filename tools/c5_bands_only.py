"""torchrun entry: only the c5_bands configuration of bench.py's `sharded` block (one 16384^2 frame
as N screen bands + rivecuda_band_gather), `reps` timed repetitions after a warm-up one.
usage: python -m torch.distributed.run --nproc-per-node N tools/c5_bands_only.py [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from rive_runtime_b200 import band_render

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
line = band_render.run("scene:c5", int(sys.argv[1]) if len(sys.argv) > 1 else 3, rank, local, world)
if rank == 0:
    print(json.dumps(line))
dist.destroy_process_group()
