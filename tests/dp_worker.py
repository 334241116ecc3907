"""Worker for tests/test_dp_gpu.py (launched by torchrun, one process per GPU): batch-sharded training must equal
"the oracle per shard, gradients averaged" (local BN / latent / GP statistics)."""
import os
import sys
import tempfile

import torch
import torch.distributed as dist
from torch.optim import Adam

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_oracle as O  # noqa: E402
from rnagan_b200 import dcgan, wgan_loss  # noqa: E402
from rnagan_b200.trainer import Trainer  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    size, batch, feats = 32, 8, 128
    lrelu, tanh = torch.nn.LeakyReLU(0.2), torch.nn.Tanh()
    oV = O.OracleVAE(feats, beta=0.005).eval()
    O.reinit_(oV, 13)
    ckpt = os.path.join(tempfile.mkdtemp(), f"vae{rank}.pt")
    torch.save(oV.state_dict(), ckpt)
    net = {
        "generator": {"name": dcgan.DCGANGenerator,
                      "args": {"encoding_dims": 2048, "out_channels": 3, "step_channels": 64, "out_size": size,
                               "nonlinearity": torch.nn.LeakyReLU(0.2), "last_nonlinearity": torch.nn.Tanh()},
                      "optimizer": {"name": Adam, "args": {"lr": 0.0001, "betas": (0.5, 0.999)}}},
        "discriminator": {"name": dcgan.DCGANDiscriminator,
                          "args": {"in_size": size, "in_channels": 3, "step_channels": 64,
                                   "nonlinearity": torch.nn.LeakyReLU(0.2),
                                   "last_nonlinearity": torch.nn.LeakyReLU(0.2)},
                          "optimizer": {"name": Adam, "args": {"lr": 0.0004, "betas": (0.5, 0.999)}}},
    }
    losses = [wgan_loss.WassersteinGeneratorLossVAE(ckpt, feats), wgan_loss.WassersteinDiscriminatorLossVAE(ckpt, feats),
              wgan_loss.WassersteinGradientPenaltyVAE(ckpt, feats)]
    tr = Trainer(net, losses, device=dev, sample_size=64, epochs=1, devices=[0])
    tr.generator.train(); tr.discriminator.train()
    names = list(tr.losses.keys())

    # per-shard oracle replicas (every rank computes ALL shards on CPU so it can form the averaged reference)
    def fresh():
        g = O.OracleGenerator(2048, size, 3, 64, nonlinearity=lrelu, last_nonlinearity=tanh).train()
        d = O.OracleCritic(size, 3, 64, nonlinearity=lrelu, last_nonlinearity=lrelu).train()
        O.reinit_(g, 11); O.reinit_(d, 12)
        return g, d

    base_g, base_d = fresh()
    tr.generator.load_state_dict(base_g.state_dict())
    tr.discriminator.load_state_dict(base_d.state_dict())
    shards = [O.make_batch(batch, feats, size, 100 + r) for r in range(world)]
    ok = True
    for which, (fn, mname) in enumerate([(O.g_step, "generator"), (O.critic_step, "discriminator"),
                                          (O.gp_step, "discriminator")]):
        # reference: oracle on each shard from the same weights, with that rank's RNG stream; average the gradients
        grads = None
        for r in range(world):
            g, d = fresh()
            g.load_state_dict(tr.generator.state_dict()); d.load_state_dict(tr.discriminator.state_dict())
            opt = Adam((g if which == 0 else d).parameters(), lr=0.0)
            torch.manual_seed(1000 * which + r)
            fn(g, d, opt, oV, shards[r])
            gs = [p.grad.clone() for p in (g if which == 0 else d).parameters()]
            grads = gs if grads is None else [a + b for a, b in zip(grads, gs)]
        grads = [x / world for x in grads]
        tr.real_inputs, tr.batch_size = shards[rank], batch
        torch.manual_seed(1000 * which + rank)
        tr._call(names[which])
        torch.cuda.synchronize()
        net_m = getattr(tr, mname)
        worst = 0.0
        for p, gref in zip(net_m.parameters(), grads):
            if gref.norm() == 0:
                continue
            worst = max(worst, rel(p.grad / world, gref))       # p.grad holds the SUM over ranks
        # every rank must end the step with identical weights
        flat = torch.cat([p.detach().flatten() for p in net_m.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        same = torch.equal(flat, ref)
        print(f"rank {rank} step {which}: worst grad rel-L2 vs averaged per-shard oracle {worst:.4f}, "
              f"weights identical across ranks: {same}", flush=True)
        ok = ok and same and worst < 0.35
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
