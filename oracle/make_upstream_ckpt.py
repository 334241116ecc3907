"""Write tests/golden/upstream_style_ckpt.model: a checkpoint in the layout torchgan's Trainer.save_model produces [tg],
holding instances of THE REFERENCE'S OWN loss classes pickled under their original top-level module names (build
container only -- needs /root/reference; only the fixture travels).

    python oracle/make_upstream_ckpt.py

The reference hard-codes the conditioning betaVAE's widths ([6000, 4000, 2048] / [4000, 6000], src/wgan_loss.py:67), which
would make the three pickled loss objects 1.8 GB; the name `betaVAE` inside the reference's `wgan_loss` module is therefore
bound to a constructor with small widths while the objects are built.  Classes, module paths, attribute names and the
checkpoint dictionary are untouched.  One reference iteration is run first so the Adam states are populated.
"""
import os
import sys
import tempfile

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/src"
sys.path.insert(0, os.path.join(HERE, "torchgan_shim"))
sys.path.insert(0, REF)

FEATS, Z, ENC, DEC = 24, 32, [48, 40, 32], [40, 48]
SIZE, STEP, BATCH = 16, 8, 4


def main():
    import betaVAE as ref_vae
    import wgan_loss as ref_loss
    from torch.optim import Adam
    from torchgan.models import DCGANDiscriminator, DCGANGenerator
    from torchgan.trainer import Trainer

    torch.manual_seed(5)
    small = lambda feats, z, enc, dec, beta=0.005: ref_vae.betaVAE(feats, Z, ENC, DEC, beta=beta)  # noqa: E731
    vae = small(FEATS, None, None, None)
    ckpt = os.path.join(tempfile.mkdtemp(), "vae.pt")
    torch.save(vae.state_dict(), ckpt)
    real_ctor, ref_loss.betaVAE = ref_loss.betaVAE, small
    try:
        losses = [ref_loss.WassersteinGeneratorLossVAE(ckpt, FEATS), ref_loss.WassersteinDiscriminatorLossVAE(ckpt, FEATS),
                  ref_loss.WassersteinGradientPenaltyVAE(ckpt, FEATS)]
    finally:
        ref_loss.betaVAE = real_ctor
    net = {
        "generator": {"name": DCGANGenerator,
                      "args": {"encoding_dims": Z, "out_channels": 3, "step_channels": STEP, "out_size": SIZE,
                               "nonlinearity": nn.LeakyReLU(0.2), "last_nonlinearity": nn.Tanh()},
                      "optimizer": {"name": Adam, "args": {"lr": 0.0001, "betas": (0.5, 0.999)}}},
        "discriminator": {"name": DCGANDiscriminator,
                          "args": {"in_size": SIZE, "in_channels": 3, "step_channels": STEP,
                                   "nonlinearity": nn.LeakyReLU(0.2), "last_nonlinearity": nn.LeakyReLU(0.2)},
                          "optimizer": {"name": Adam, "args": {"lr": 0.0004, "betas": (0.5, 0.999)}}},
    }
    tr = Trainer(net, losses, device=torch.device("cpu"), sample_size=8, epochs=1, devices=[0])
    tr.batch_size = BATCH
    tr.real_inputs = {"image": torch.rand(BATCH, 3, SIZE, SIZE) * 2 - 1, "rna_data": torch.randn(BATCH, FEATS),
                      "labels": torch.zeros(BATCH)}
    values = tr.train_iter()
    # the dictionary torchgan's Trainer.save_model writes [tg]
    model = {"epoch": 1, "loss_information": tr.loss_information, "loss_objects": tr.losses, "metric_objects": None,
             "loss_logs": {k: [float(v)] for k, v in values.items()}, "metric_logs": {}}
    for name in tr.model_names + tr.optimizer_names:
        model[name] = getattr(tr, name).state_dict()
    # fixture-only extras: what a loader must find again, computed from the reference objects
    model["_expected"] = {
        "loss_classes": {k: f"{type(v).__module__}.{type(v).__qualname__}" for k, v in tr.losses.items()},
        "vae_class": f"{type(losses[0].betavae).__module__}.{type(losses[0].betavae).__qualname__}",
        "vae_param_sum": float(sum(p.double().sum() for p in losses[0].betavae.state_dict().values())),
        "generator_param_sum": float(sum(p.double().sum() for p in tr.generator.state_dict().values())),
        "reduction_attr": losses[0].reduction, "override_attr": losses[0].override_train_ops,
        "dims": {"feats": FEATS, "z": Z, "enc": ENC, "dec": DEC, "size": SIZE, "step": STEP},
    }
    out = os.path.join(ROOT, "tests", "golden", "upstream_style_ckpt.model")
    torch.save(model, out)
    print("wrote", out, os.path.getsize(out), "bytes;", model["_expected"]["loss_classes"])


if __name__ == "__main__":
    main()
