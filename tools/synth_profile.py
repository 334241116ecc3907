"""Per-kernel GPU time of tile synthesis (generate_tiles, chunk 1024) through CUPTI (torch.profiler)."""
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rnagan_b200 import gan_utils  # noqa: E402


def main(S=1024, reps=3):
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    tr = bench.build_job(dev, 64)
    G = tr.generator
    vae = tr.losses["WassersteinGeneratorLossVAE"]._encoder(dev)
    rows = torch.randn(S, bench.GENES, generator=torch.Generator().manual_seed(5)).to(dev)
    out = torch.empty(S, bench.SIZE, bench.SIZE, 3, dtype=torch.float32, device=dev)
    for _ in range(2):
        gan_utils.generate_tiles(G, vae, rows, S, chunk=S, device=dev, out=out)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            gan_utils.generate_tiles(G, vae, rows, S, chunk=S, device=dev, out=out)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = re.sub(r"\(.*", "", ev.name).replace("void ", "").replace("rg::", "")
        agg[name][0] += 1
        agg[name][1] += ev.device_time
    tot = sum(v[1] for v in agg.values())
    print(f"# generate_tiles chunk {S}: kernel time {tot / reps / 1e3:.3f} ms per chunk ({S / (tot / reps / 1e6):.0f} tiles/s kernel-bound)")
    print(f"{'share':>7} {'ms/chunk':>9} {'n':>5} {'avg us':>9}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / tot * 100:6.2f}% {v[1] / reps / 1e3:9.3f} {v[0] / reps:5.1f} {v[1] / v[0]:9.1f}  {k[:100]}")


if __name__ == "__main__":
    main()
