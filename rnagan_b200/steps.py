"""The three optimiser steps of one RNA-GAN iteration as device-side schedules (no host synchronisation).

``wgan_loss.*.train_ops`` (the reference-facing API) and ``bench.py`` (device-resident timing) both call these; the
only difference is where the noise / eps / batch come from and whether the loss is read back with ``.item()``.
"""
import os

import torch

from . import ops
from .optim import adam_step

F32 = torch.float32
BF16 = torch.bfloat16


def latent(ge, noise_d, z):
    """standardise_0(noise + z) -> bf16 [B, E] (src/wgan_loss.py:105-106); z=None (the un-conditioned `wgan` losses,
    src/histopathology_gan.py:267-272): the N(0, I) noise itself, cast to bf16."""
    B, E = noise_d.shape
    lat = ge.bufs.get("lat", (B, E), BF16)
    if z is None:
        ops.cast_pad_bf16(noise_d, E, out=lat)
    else:
        ops.latent_prep(noise_d, z, lat_bf16=lat)
    return lat


def _loss_buf(ge, name):
    return ge.bufs.get(name, (1,), F32)


# ------------------------------------------------------------------------------------------------ optimiser updates
# Single process: Adam runs right after backward, like the reference.  Data-parallel: the all-reduce of the LAST layer
# a backward produces cannot overlap that backward (the generator's layer 0 is 60 % of its gradient bytes and is produced
# last), so the update is deferred to the first consumer of the new weights -- G's update to just before the next
# G forward (after D(real) of the critic step), D's to just before the next D forward (after the G forward of the GP /
# next G step) -- and the tail of the reduction hides behind that work.  Every path that reads the weights from outside
# these schedules (module.forward, state_dict, checkpoints, synthesis) goes through `module._engine()`, which applies
# what is pending first; observable values are those of the reference's order of operations.
DEFER_UPDATES = os.environ.get("RG_DP_DEFER", "1") != "0"


def _queue_update(eng, opt):
    if eng.sync.world() == 1 or not DEFER_UPDATES:
        adam_step(opt, grad_scale=eng.sync.finish())        # gradients averaged over ranks (no-op single-process)
        eng.pack(full=False)                                # Adam re-emitted the bf16 GEMM operands itself
        return
    eng.sync.launch_rest()
    eng.pending_update = opt


def apply_update(eng):
    """Apply the engine's deferred optimiser step, if any (waits for its gradient reductions on the stream)."""
    opt = getattr(eng, "pending_update", None)
    if opt is not None:
        eng.pending_update = None
        adam_step(opt, grad_scale=eng.sync.wait())
        eng.pack(full=False)


def _need_train_mode(what, *modules):
    """The hand-scheduled backward passes differentiate train-mode BatchNorm (batch statistics), which is what every
    reference driver runs (Trainer.train puts the models in train mode each epoch).  With a module in eval mode the
    forward would use running statistics while the backward subtracted batch-mean terms: refuse instead of returning
    silently wrong gradients."""
    for m in modules:
        if not m.training:
            raise NotImplementedError(f"{what}: {type(m).__name__} is in eval mode; the sm_100a backward of BatchNorm is "
                                      "implemented for train mode (batch statistics) only -- call .train() first")


def g_step(generator, discriminator, opt_g, noise_d, z):
    """WassersteinGeneratorLossVAE.train_ops body (src/wgan_loss.py:100-128). Returns the device loss tensor [1]."""
    _need_train_mode("generator step", generator, discriminator)      # both BatchNorm stacks are differentiated
    ge, de = generator._engine(flush=False), discriminator._engine(flush=False)
    B = noise_d.shape[0]
    lat = latent(ge, noise_d, z)
    apply_update(ge)
    fake = ge.forward(lat, tag="g", training=generator.training)
    apply_update(de)
    out = de.forward(fake, tag="gstep", training=discriminator.training)
    loss = _loss_buf(ge, "loss_g")
    ops.wgan_loss(out, -1.0, loss)                                   # mean(-D(G(z)))
    d_img = de.backward(B, -1.0 / B, tag="gstep", params=False, want_dimg=True)
    ge.backward(lat, d_img, fake, tag="g")
    _queue_update(ge, opt_g)
    return loss


def critic_step(generator, discriminator, opt_d, noise_d, z, real, clip=None):
    """WassersteinDiscriminatorLossVAE.train_ops body (src/wgan_loss.py:213-262)."""
    _need_train_mode("critic step", discriminator)                    # G is forward-only here: any mode
    ge, de = generator._engine(flush=False), discriminator._engine(flush=False)
    apply_update(de)
    if clip is not None:                                             # src/wgan_loss.py:213-215
        for p in discriminator.parameters():
            ops.clamp_(p.data, clip[0], clip[1])
        de.pack()
    B = noise_d.shape[0]
    lat = latent(ge, noise_d, z)
    out_real = de.forward(real, tag="real", training=discriminator.training)
    apply_update(ge)                                    # G's reduction tail overlapped D(real)
    fake = ge.forward(lat, tag="g", training=generator.training)
    out_fake = de.forward(fake, tag="fake", training=discriminator.training)
    loss = _loss_buf(ge, "loss_d")
    ops.wgan_loss(out_fake, 1.0, loss, b=out_real, sign_b=-1.0)      # mean(D(G(z)) - D(x))
    de.backward_pair(B, (("real", -1.0), ("fake", 1.0)))     # one wgrad / dgrad launch per layer over both passes
    _queue_update(de, opt_d)
    return loss


def gp_step(generator, discriminator, opt_d, noise_d, z, real, eps_d, lambd=10.0):
    """WassersteinGradientPenaltyVAE.train_ops body (src/wgan_loss.py:357-388). Returns device [P, seed, ||g||]."""
    _need_train_mode("gradient-penalty step", discriminator)
    ge, de = generator._engine(flush=False), discriminator._engine(flush=False)
    lat = latent(ge, noise_d, z)
    apply_update(ge)
    fake = ge.forward(lat, tag="g", training=generator.training)
    apply_update(de)                                    # D's critic-step reduction overlapped this G forward
    out3 = de.gradient_penalty(real, fake, eps_d, lambd=lambd)
    _queue_update(de, opt_d)
    return out3


def flush_updates(*modules):
    """Apply every deferred optimiser step of the given generator / critic modules (data-parallel runs)."""
    for m in modules:
        m._engine()
