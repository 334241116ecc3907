"""Per-kernel GPU time of one betaVAE training step (BASELINE config 5 shapes, batch 128) through CUPTI."""
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rnagan_b200 import betaVAE as bv  # noqa: E402


def main(reps=5):
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    torch.manual_seed(7)
    vmod = bv.betaVAE(bench.GENES, bench.LATENT, [6000, 4000, 2048], [4000, 6000], beta=0.0005)
    for m in vmod.modules():
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.xavier_uniform_(m.weight)
            m.bias.data.fill_(0.01)
    vmod = vmod.to(dev).train()
    vopt = torch.optim.Adam(vmod.parameters(), lr=5e-5, weight_decay=0)
    xs = torch.randn(128, bench.GENES, generator=torch.Generator().manual_seed(11)).to(dev)
    for _ in range(3):
        bv.train_step(vmod, vopt, xs, 0.0005)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            bv.train_step(vmod, vopt, xs, 0.0005)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = re.sub(r"\(.*", "", ev.name).replace("void ", "").replace("rg::", "")
        agg[name][0] += 1
        agg[name][1] += ev.device_time
    tot = sum(v[1] for v in agg.values())
    print(f"# betaVAE train step, batch 128: kernel time {tot / reps / 1e3:.3f} ms per step")
    print(f"{'share':>7} {'ms/step':>9} {'n':>5} {'avg us':>9}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / tot * 100:6.2f}% {v[1] / reps / 1e3:9.3f} {v[0] / reps:5.1f} {v[1] / v[0]:9.1f}  {k[:100]}")


if __name__ == "__main__":
    main()
