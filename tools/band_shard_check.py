"""One very large frame sharded as screen-tile bands over N GPUs (BASELINE.json
configs[4]): every rank replays the same flush inputs with renderTargetUpdateBounds narrowed
to its band, then ONE NCCL gather composites the frame on rank 0, which checks it bit for bit
against its own single-GPU render of the whole frame and prints one JSON line with both
times (device-timed, max over ranks).

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/band_shard_check.py <trace | scene:NAME> [--reps K]

`scene:c5` records the trace on the spot with the scene player (the reference front end
built in place + the ABI recorder), since the 16384x16384 / 200k-path trace is 370 MB.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from rive_runtime_b200 import band_render  # noqa: E402


def main():
    src = sys.argv[1]
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 1
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = band_render.run(src, reps, rank, local, world)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and not line["identical"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
