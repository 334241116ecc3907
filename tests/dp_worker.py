"""Worker for tests/test_dp_gpu.py (launched by torchrun, one process per GPU): batch-sharded training must equal
"the oracle per shard, gradients averaged" (local BN / latent / GP statistics).  The check itself
is oracle/dp_check.py (bench.py runs the same check after its timed regions when world > 1)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dp_check  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.set_num_threads(max(1, (os.cpu_count() or 8) // world))
    res = dp_check.run(rank, world, dev, verbose=True)
    print(f"rank {rank}: {res}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()
