"""Multi-GPU (needs >= 2 B200s; skipped otherwise): batch-sharded data-parallel training equals the oracle run per shard
with averaged gradients (within twice torch-bf16's own deviation), and all ranks hold bit-identical weights after every
optimiser step -- over NCCL and over the copy-engine peer exchange (parallel.PeerExchange, RG_DP_EXCHANGE=ce)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["nccl", "ce", "nvls"])
def test_data_parallel_two_gpus(cuda_dev, mode):
    """nccl: torch.distributed all_reduce; ce: copy-engine peer exchange (the default); nvls: in-switch multimem
    reduction through the multicast mapping (falls back to ce with a warning when the fabric has no multicast)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29577", os.path.join(ROOT, "tests", "dp_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, RG_DP_EXCHANGE=mode))
    sys.stdout.write(res.stdout[-3000:])
    sys.stderr.write(res.stderr[-3000:])
    assert res.returncode == 0


def test_slices_sum_matches_ordered_sum(cuda_dev):
    from rnagan_b200 import ops
    g = torch.Generator().manual_seed(1)
    for world, n, stride in ((2, 4096, 4096), (8, 1000 * 4, 4100), (3, 4, 8), (8, 1 << 20, 1 << 20)):
        stage = torch.randn(world, stride, generator=g).to(cuda_dev)
        out = torch.full((n,), float("nan"), device=cuda_dev)
        ops.slices_sum(stage, world, n, out)
        want = stage[0, :n].clone()
        for r in range(1, world):
            want += stage[r, :n]
        assert torch.equal(out, want)
