#!/usr/bin/env python
"""N-rank vs 1-rank parity of the x-y decomposed hot path (launch with torchrun, one rank per GPU); the check itself lives in
cgfd3d_b200/nrank_check.py (bench.py runs the same check before every multi-GPU measurement).
  torchrun ... scripts/multi_gpu_check.py [x|y] [iso|vti|aniso|visco]"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cgfd3d_b200 import nrank_check  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    px, py = {2: (2, 1), 4: (2, 2), 8: (4, 2)}.get(world, (world, 1))
    args = sys.argv[1:]
    if args and args[0] == "y":
        px, py = py, px
    medium = args[1] if len(args) > 1 else "iso"
    res = nrank_check.check(rank, world, local, px, py, dist, medium=medium)
    if rank == 0:
        print("MULTI_GPU_CHECK " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not res["ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
