"""Per-kernel GPU time of the device-resident training step through CUPTI (torch.profiler): warm, in-order, real clocks.
Usage: python tools/step_profile.py [steps] > profiles/...txt"""
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rnagan_b200 import steps  # noqa: E402

steps.use_graphs(False)      # per-kernel CUPTI records of eager launches (same kernels as the replayed graphs)


def main(n_steps=3):
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    B = 64
    tr = bench.build_job(dev, B)
    G, D = tr.generator, tr.discriminator
    vae = tr.losses["WassersteinGeneratorLossVAE"]._encoder(dev)
    hb = bench.host_batch(B, 1000)
    real, rna = hb["image"].to(dev), hb["rna_data"].to(dev)
    g = torch.Generator().manual_seed(7)
    noise = (torch.rand(3, B, bench.LATENT, generator=g) * 0.6 - 0.3).to(dev)
    eps = torch.rand(1, generator=g).to(dev)

    def step():
        z = vae.encode_mean(rna)
        steps.g_step(G, D, tr.optimizer_generator, noise[0], z)
        steps.critic_step(G, D, tr.optimizer_discriminator, noise[1], z, real)
        steps.gp_step(G, D, tr.optimizer_discriminator, noise[2], z, real, eps)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    base_ms = e0.elapsed_time(e1) / n_steps
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(n_steps):
            step()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    t_min, t_max = None, None
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = re.sub(r"\(.*", "", ev.name).replace("void ", "").replace("rg::", "")
        dur = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        agg[name][0] += 1
        agg[name][1] += dur
        s = ev.time_range.start
        t_min = s if t_min is None else min(t_min, s)
        t_max = max(t_max or 0, ev.time_range.end)
    tot = sum(v[1] for v in agg.values())
    print(f"# {n_steps} steps, {base_ms:.3f} ms/step unprofiled; kernel time {tot / n_steps / 1e3:.3f} ms/step, "
          f"span {(t_max - t_min) / n_steps / 1e3:.3f} ms/step (CUPTI via torch.profiler)")
    print(f"{'share':>7} {'ms/step':>9} {'n/step':>7} {'avg us':>9}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / tot * 100:6.2f}% {v[1] / n_steps / 1e3:9.3f} {v[0] / n_steps:7.1f} {v[1] / v[0]:9.1f}  {k[:100]}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
