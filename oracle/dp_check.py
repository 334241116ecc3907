"""Data-parallel correctness check -- TEST INFRASTRUCTURE (uses the CPU oracle as the checker).

Used by tests/dp_worker.py (2-GPU parity test) and by bench.py AFTER its timed regions when world > 1 (the `dp_check`
key of the JSON line): batch-sharded training over N ranks must equal "the oracle run on each rank's shard from the
same weights, gradients averaged over ranks" (local BatchNorm / latent / gradient-norm
statistics, exactly what DDP around the reference would compute), and every rank must hold bit-identical weights after
every optimiser step.

Each rank evaluates the oracle on ITS shard only (CPU fp32, plus once under torch's own CPU bf16 autocast as the
yardstick) and the per-shard oracle gradients are averaged with one all-reduce, so the cost does not grow with N.
"""
import copy
import os
import tempfile

import torch
import torch.distributed as dist
from torch.optim import Adam

from oracle import ref_oracle as O


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()


def run(rank, world, dev, size=32, batch=8, feats=128, verbose=False):
    """Three train_ops on a small job (32x32 tiles, batch 8 per rank).  Returns a dict with
    weights_identical, worst_grad_rel / worst_grad_cos (ours vs the averaged per-shard oracle) and
    bf16_rel / bf16_cos (torch's own bf16-autocast deviation on the same averaged gradient, the yardstick)."""
    from rnagan_b200 import dcgan, steps, wgan_loss
    from rnagan_b200.trainer import Trainer

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
    os.unlink(ckpt)
    tr = Trainer(net, losses, device=dev, sample_size=64, epochs=1, devices=[0])
    tr.generator.train(); tr.discriminator.train()
    names = list(tr.losses.keys())

    def fresh():
        g = O.OracleGenerator(2048, size, 3, 64, nonlinearity=lrelu, last_nonlinearity=tanh).train()
        d = O.OracleCritic(size, 3, 64, nonlinearity=lrelu, last_nonlinearity=lrelu).train()
        O.reinit_(g, 11); O.reinit_(d, 12)
        return g, d

    base_g, base_d = fresh()
    tr.generator.load_state_dict(base_g.state_dict())
    tr.discriminator.load_state_dict(base_d.state_dict())
    shard = O.make_batch(batch, feats, size, 100 + rank)
    res = {"weights_identical": True, "worst_grad_rel": 0.0, "worst_grad_cos": 1.0, "bf16_rel": 0.0, "bf16_cos": 1.0}

    def averaged(grads):
        flat = torch.cat([g.flatten() for g in grads]).to(dev)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat = (flat / world).cpu()
        out, off = [], 0
        for g in grads:
            out.append(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        return out

    for which, (fn, mname) in enumerate([(O.g_step, "generator"), (O.critic_step, "discriminator"),
                                          (O.gp_step, "discriminator")]):
        # reference: the oracle on this rank's shard from the current (identical) weights, this rank's RNG stream
        sd_g = {k: v.cpu() for k, v in tr.generator.state_dict().items()}
        sd_d = {k: v.cpu() for k, v in tr.discriminator.state_dict().items()}
        per_shard = []
        for autocast in (False, True):
            g, d = fresh()
            g.load_state_dict(sd_g); d.load_state_dict(sd_d)
            opt = Adam((g if which == 0 else d).parameters(), lr=0.0)
            torch.manual_seed(1000 * which + rank)
            if autocast:
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    fn(g, d, opt, oV, shard)
            else:
                fn(g, d, opt, oV, shard)
            per_shard.append([p.grad.float().clone() for p in (g if which == 0 else d).parameters()])
        ref, ref_bf16 = averaged(per_shard[0]), averaged(per_shard[1])
        tr.real_inputs, tr.batch_size = shard, batch
        torch.manual_seed(1000 * which + rank)
        tr._call(names[which])
        net_m = getattr(tr, mname)
        steps.flush_updates(tr.generator, tr.discriminator)          # the data-parallel path defers its optimiser step
        torch.cuda.synchronize()
        for p, gref, gbf in zip(net_m.parameters(), ref, ref_bf16):
            if gref.norm() == 0:
                continue
            ours = p.grad.detach().cpu() / world                     # p.grad holds the SUM over ranks
            res["worst_grad_rel"] = max(res["worst_grad_rel"], _rel(ours, gref))
            res["worst_grad_cos"] = min(res["worst_grad_cos"], _cos(ours, gref))
            res["bf16_rel"] = max(res["bf16_rel"], _rel(gbf, gref))
            res["bf16_cos"] = min(res["bf16_cos"], _cos(gbf, gref))
        # every rank must end the step with bit-identical weights and BatchNorm-free optimiser inputs
        for m in (tr.generator, tr.discriminator):
            flat = torch.cat([p.detach().flatten() for p in m.parameters()])
            lo, hi = flat.clone(), flat.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            res["weights_identical"] = res["weights_identical"] and bool(torch.equal(lo, hi))
        if verbose:
            print(f"rank {rank} step {which}: {res}", flush=True)
    # tolerance: at most twice torch-bf16's own deviation from the fp32 oracle (the yardstick used by every bf16 test)
    res["ok"] = bool(res["weights_identical"]
                     and res["worst_grad_cos"] >= 1.0 - 2.0 * (1.0 - res["bf16_cos"]) - 0.01
                     and res["worst_grad_rel"] <= 2.0 * res["bf16_rel"] + 0.02)
    del tr
    return res
