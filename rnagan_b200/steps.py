"""The three optimiser steps of one RNA-GAN iteration as device-side schedules (no host synchronisation).

``wgan_loss.*.train_ops`` (the reference-facing API) and ``bench.py`` (device-resident timing) both call these; the
only difference is where the noise / eps / batch come from and whether the loss is read back with ``.item()``.
"""
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


def g_step(generator, discriminator, opt_g, noise_d, z):
    """WassersteinGeneratorLossVAE.train_ops body (src/wgan_loss.py:100-128). Returns the device loss tensor [1]."""
    ge, de = generator._engine(), discriminator._engine()
    B = noise_d.shape[0]
    lat = latent(ge, noise_d, z)
    fake = ge.forward(lat, tag="g", training=generator.training)
    out = de.forward(fake, tag="gstep", training=discriminator.training)
    loss = _loss_buf(ge, "loss_g")
    ops.wgan_loss(out, -1.0, loss)                                   # mean(-D(G(z)))
    d_img = de.backward(B, -1.0 / B, tag="gstep", params=False, want_dimg=True)
    ge.backward(lat, d_img, fake, tag="g")
    adam_step(opt_g, grad_scale=ge.sync.finish())       # gradients averaged over ranks (no-op single-process)
    ge.pack(full=False)                                 # Adam re-emitted the bf16 GEMM operands itself
    return loss


def critic_step(generator, discriminator, opt_d, noise_d, z, real, clip=None):
    """WassersteinDiscriminatorLossVAE.train_ops body (src/wgan_loss.py:213-262)."""
    ge, de = generator._engine(), discriminator._engine()
    if clip is not None:                                             # src/wgan_loss.py:213-215
        for p in discriminator.parameters():
            ops.clamp_(p.data, clip[0], clip[1])
        de.pack()
    B = noise_d.shape[0]
    lat = latent(ge, noise_d, z)
    out_real = de.forward(real, tag="real", training=discriminator.training)
    fake = ge.forward(lat, tag="g", training=generator.training)
    out_fake = de.forward(fake, tag="fake", training=discriminator.training)
    loss = _loss_buf(ge, "loss_d")
    ops.wgan_loss(out_fake, 1.0, loss, b=out_real, sign_b=-1.0)      # mean(D(G(z)) - D(x))
    de.backward_pair(B, (("real", -1.0), ("fake", 1.0)))     # one wgrad / dgrad launch per layer over both passes
    adam_step(opt_d, grad_scale=de.sync.finish())
    de.pack(full=False)
    return loss


def gp_step(generator, discriminator, opt_d, noise_d, z, real, eps_d, lambd=10.0):
    """WassersteinGradientPenaltyVAE.train_ops body (src/wgan_loss.py:357-388). Returns device [P, seed, ||g||]."""
    ge, de = generator._engine(), discriminator._engine()
    lat = latent(ge, noise_d, z)
    fake = ge.forward(lat, tag="g", training=generator.training)
    out3 = de.gradient_penalty(real, fake, eps_d, lambd=lambd)
    adam_step(opt_d, grad_scale=de.sync.finish())
    de.pack(full=False)
    return out3
