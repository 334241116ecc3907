"""Diagnostic: run the three train_ops on the CUDA path next to the CPU oracle and print every deviation."""
import os
import sys
import tempfile

import numpy as np
import torch
from torch.optim import Adam

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_oracle as O  # noqa: E402
from rnagan_b200 import dcgan, wgan_loss  # noqa: E402
from rnagan_b200.trainer import Trainer  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()


def run(size, batch, feats, iters):
    dev = torch.device("cuda:0")
    lrelu, tanh = torch.nn.LeakyReLU(0.2), torch.nn.Tanh()
    oG = O.OracleGenerator(2048, size, 3, 64, nonlinearity=lrelu, last_nonlinearity=tanh)
    oD = O.OracleCritic(size, 3, 64, nonlinearity=lrelu, last_nonlinearity=lrelu)
    oV = O.OracleVAE(feats, beta=0.005)
    O.reinit_(oV, 13); O.reinit_(oG, 11); O.reinit_(oD, 12)
    oV.eval(); oG.train(); oD.train()
    ckpt = os.path.join(tempfile.mkdtemp(), "vae.pt")
    torch.save(oV.state_dict(), ckpt)
    net = {
        "generator": {"name": dcgan.DCGANGenerator,
                      "args": {"encoding_dims": 2048, "out_channels": 3, "step_channels": 64, "out_size": size,
                               "nonlinearity": torch.nn.LeakyReLU(0.2), "last_nonlinearity": torch.nn.Tanh()},
                      "optimizer": {"name": Adam, "args": {"lr": 0.0001, "betas": (0.5, 0.999)}}},
        "discriminator": {"name": dcgan.DCGANDiscriminator,
                          "args": {"in_size": size, "in_channels": 3, "step_channels": 64,
                                   "nonlinearity": torch.nn.LeakyReLU(0.2), "last_nonlinearity": torch.nn.LeakyReLU(0.2)},
                          "optimizer": {"name": Adam, "args": {"lr": 0.0004, "betas": (0.5, 0.999)}}},
    }
    losses = [wgan_loss.WassersteinGeneratorLossVAE(ckpt, feats), wgan_loss.WassersteinDiscriminatorLossVAE(ckpt, feats),
              wgan_loss.WassersteinGradientPenaltyVAE(ckpt, feats)]
    tr = Trainer(net, losses, device=dev, sample_size=64, epochs=1, devices=[0])
    tr.generator.load_state_dict(oG.state_dict())
    tr.discriminator.load_state_dict(oD.state_dict())
    tr.generator.train(); tr.discriminator.train()
    data = O.make_batch(batch, feats, size, 14)
    tr.real_inputs = data
    tr.batch_size = batch
    og = Adam(oG.parameters(), lr=1e-4, betas=(0.5, 0.999))
    od = Adam(oD.parameters(), lr=4e-4, betas=(0.5, 0.999))

    names = list(tr.losses.keys())
    import copy

    def autocast_ref(which):
        """torch's OWN bf16 (CPU autocast) deviation from fp32 for the same step, as a noise yardstick."""
        st = torch.get_rng_state()
        cG, cD = copy.deepcopy(oG), copy.deepcopy(oD)
        cg = Adam(cG.parameters(), lr=1e-4, betas=(0.5, 0.999))
        cd = Adam(cD.parameters(), lr=4e-4, betas=(0.5, 0.999))
        with torch.autocast("cpu", dtype=torch.bfloat16):
            if which == 0:
                v = O.g_step(cG, cD, cg, oV, data)
            elif which == 1:
                v = O.critic_step(cG, cD, cd, oV, data)
            else:
                v = O.gp_step(cG, cD, cd, oV, data)
        torch.set_rng_state(st)
        return v, cG, cD

    def cmp3(title, onet, mnet, anet):
        wc = [1.0, 1.0]
        for (n, po), (_, pm), (_, pa) in zip(onet.named_parameters(), mnet.named_parameters(), anet.named_parameters()):
            if po.grad is None or po.grad.norm() == 0:
                continue
            c1, r1 = cos(pm.grad, po.grad), rel(pm.grad, po.grad)
            c2, r2 = cos(pa.grad, po.grad), rel(pa.grad, po.grad)
            wc = [min(wc[0], c1), min(wc[1], c2)]
            print(f"    {title} {n:22s} cuda: cos {c1:.4f} rel {r1:.3f} | torch-bf16: cos {c2:.4f} rel {r2:.3f}")
        print(f"  {title}: worst cos cuda {wc[0]:.4f} torch-bf16 {wc[1]:.4f}")

    def cmp(title, onet, mnet, grads):
        worst_c, worst_r = 1.0, 0.0
        for (n, po), (_, pm) in zip(onet.named_parameters(), mnet.named_parameters()):
            a, b = (pm.grad, po.grad) if grads else (pm.data, po.data)
            c, r = cos(a, b), rel(a, b)
            worst_c, worst_r = min(worst_c, c), max(worst_r, r)
            print(f"    {title} {n:24s} cos {c:.5f} rel {r:.4f} |ref| {b.norm().item():.3e}")
        print(f"  {title}: worst cos {worst_c:.5f} worst rel {worst_r:.4f}")

    # oracle and product draw from the same CPU RNG stream: run them in lock-step with saved/restored RNG state
    def sync():
        """put the CUDA modules / optimizers in exactly the oracle's state so each step is compared in isolation"""
        tr.generator.load_state_dict(oG.state_dict())
        tr.discriminator.load_state_dict(oD.state_dict())
        if len(og.state_dict()["state"]):
            tr.optimizer_generator.load_state_dict(og.state_dict())
        if len(od.state_dict()["state"]):
            tr.optimizer_discriminator.load_state_dict(od.state_dict())

    torch.manual_seed(99)
    for it in range(iters):
        va, aG, aD = autocast_ref(0)
        sync()
        st = torch.get_rng_state()
        vg_o = O.g_step(oG, oD, og, oV, data)
        st_after = torch.get_rng_state()
        torch.set_rng_state(st)
        vg_m = tr._call(names[0])
        assert torch.equal(torch.get_rng_state(), st_after), "RNG stream diverged (G step)"
        print(f"it{it} G loss  oracle {vg_o:.6f} cuda {vg_m:.6f} torch-bf16 {va:.6f}")
        cmp3("G-step dG", oG, tr.generator, aG)
        va, aG, aD = autocast_ref(1)
        sync()
        st = torch.get_rng_state()
        vd_o = O.critic_step(oG, oD, od, oV, data)
        st_after = torch.get_rng_state()
        torch.set_rng_state(st)
        vd_m = tr._call(names[1])
        assert torch.equal(torch.get_rng_state(), st_after), "RNG stream diverged (critic step)"
        print(f"it{it} D loss  oracle {vd_o:.6f} cuda {vd_m:.6f} torch-bf16 {va:.6f}")
        cmp3("critic-step dD", oD, tr.discriminator, aD)
        va, aG, aD = autocast_ref(2)
        sync()
        st = torch.get_rng_state()
        vp_o = O.gp_step(oG, oD, od, oV, data)
        st_after = torch.get_rng_state()
        torch.set_rng_state(st)
        vp_m = tr._call(names[2])
        assert torch.equal(torch.get_rng_state(), st_after), "RNG stream diverged (GP step)"
        print(f"it{it} GP      oracle {vp_o:.6f} cuda {vp_m:.6f} torch-bf16 {va:.6f}")
        cmp3("gp-step dD", oD, tr.discriminator, aD)
        cmp("weights G", oG, tr.generator, False)
        cmp("weights D", oD, tr.discriminator, False)
        for (n, bo), (_, bm) in zip(oD.named_buffers(), tr.discriminator.named_buffers()):
            print(f"    D buffer {n:28s} rel {rel(bm.float(), bo.float()):.5f}")
        for (n, bo), (_, bm) in zip(oG.named_buffers(), tr.generator.named_buffers()):
            print(f"    G buffer {n:28s} rel {rel(bm.float(), bo.float()):.5f}")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 8)
    cfg = sys.argv[1] if len(sys.argv) > 1 else "mini32"
    run(*{"mini32": (32, 8, 256, 2), "mini64": (64, 4, 192, 1), "mid128": (128, 8, 512, 1)}[cfg])
