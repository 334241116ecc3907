"""Drop-in WGAN loss objects with betaVAE conditioning (src/wgan_loss.py) on the sm_100a kernels.

Same class names, constructor arguments, ``forward`` and -- because torchgan's Trainer binds them BY PARAMETER NAME
 -- exactly the same ``train_ops`` signatures:

  WassersteinGeneratorLossVAE.train_ops(generator, discriminator, optimizer_generator, device, batch_size,
                                        real_inputs, labels=None)                      src/wgan_loss.py:82-129
  WassersteinDiscriminatorLossVAE.train_ops(generator, discriminator, optimizer_discriminator, real_inputs, device,
                                            labels=None)                               src/wgan_loss.py:181-263
  WassersteinGradientPenaltyVAE.train_ops(generator, discriminator, optimizer_discriminator, real_inputs, device,
                                          labels=None)                                 src/wgan_loss.py:314-389

Semantics kept from the reference: three optimiser steps per batch, a fresh CPU-RNG
uniform(-0.3, 0.3) noise draw per step, additive conditioning + unbiased batch standardisation, train-mode
BatchNorm everywhere, one scalar eps per GP step drawn after that step's noise, whole-batch gradient norm,
lambda applied outside ``forward``, un-weighted penalty returned.  Work the reference wastes (encoder backward,
critic wgrad in the G step, generator backward in the GP step) is simply not scheduled.
"""
import os

import torch
import torch.nn as nn

from . import ops, steps
from .betaVAE import betaVAE

F32 = torch.float32
BF16 = torch.bfloat16


def reduce_vae(x, reduction=None):
    if reduction == "mean":
        return torch.mean(x)
    elif reduction == "sum":
        return torch.sum(x)
    return x


def wasserstein_generator_loss_vae(fgz, reduction="mean"):
    return reduce_vae(-1.0 * fgz, reduction="mean")


def wasserstein_discriminator_loss_vae(fx, fgz, reduction="mean"):
    return reduce_vae(fgz - fx, reduction="mean")


def wasserstein_gradient_penalty_vae(interpolate, d_interpolate, reduction="mean"):
    """The penalty VALUE (||d D(x_hat) / d x_hat||_F - 1)^2 over the whole batch tensor (src/wgan_loss.py:32-44) from an
    autograd graph `d_interpolate = D(interpolate)` through this package's critic (first-order autograd, dcgan._CriticFn).
    The result can be logged or compared; differentiating it AGAIN (`penalty.backward()`) needs the critic's double
    backward, which autograd cannot provide here and raises -- the trainable form is `train_ops`, whose hand-scheduled
    double backward (CriticEngine.gradient_penalty) produces the parameter gradients of lambda * penalty directly."""
    g, = torch.autograd.grad(d_interpolate, interpolate, grad_outputs=torch.ones_like(d_interpolate), create_graph=True,
                             retain_graph=True)
    return reduce_vae((g.norm(2) - 1) ** 2, reduction)


class GeneratorLoss(nn.Module):
    """torchgan.losses.GeneratorLoss contract (reduction / override_train_ops / arg_map)."""

    def __init__(self, reduction="mean", override_train_ops=None):
        super().__init__()
        self.reduction = reduction
        self.override_train_ops = override_train_ops
        self.arg_map = {}

    def set_arg_map(self, value):
        self.arg_map.update(value)


class DiscriminatorLoss(nn.Module):
    def __init__(self, reduction="mean", override_train_ops=None):
        super().__init__()
        self.reduction = reduction
        self.override_train_ops = override_train_ops
        self.arg_map = {}

    def set_arg_map(self, value):
        self.arg_map.update(value)


# ------------------------------------------------------------------------------------------------ shared plumbing
class _StepCache:
    """Per-iteration reuse of device copies: the three train_ops of one iteration receive the same ``real_inputs``
    (torchgan Trainer.train_iter), and the frozen eval-mode encoder is deterministic, so z and the H2D copies of the
    batch are computed once.  Keyed on tensor identity + version, so any new / modified batch misses."""

    def __init__(self):
        self.key = {}
        self.val = {}
        self.keep = {}

    def get(self, what, src, device, make, extra=()):
        key = (id(src), src._version, tuple(src.shape), str(device)) + tuple(extra)
        if self.key.get(what) != key:
            self.val[what] = make()
            self.key[what] = key
            self.keep[what] = src     # keep the source alive so its id cannot be recycled
        return self.val[what]


_CACHE = _StepCache()
_SHARED_VAE = {}


def strip_module_prefix(state):
    """`model_dict_best.pt` written from an nn.DataParallel-wrapped betaVAE carries `module.`-prefixed keys; the
    reference's `load_state_dict(torch.load(path))` (src/wgan_loss.py:68) only accepts the bare layout.  Accept both."""
    if state and all(k.startswith("module.") for k in state):
        return {k[len("module."):]: v for k, v in state.items()}
    return state


def _load_vae(checkpoint, rna_features, beta):
    vae = betaVAE(rna_features, 2048, [6000, 4000, 2048], [4000, 6000], beta=beta)
    vae.load_state_dict(strip_module_prefix(torch.load(checkpoint, map_location="cpu")))
    vae.eval()
    return vae


class _VAEConditioned:
    """Mixin: owns the frozen betaVAE like the reference's loss objects do (src/wgan_loss.py:67-69)."""

    def _init_vae(self, checkpoint, rna_features, beta):
        self.betavae = _load_vae(checkpoint, rna_features, beta)
        try:
            st = os.stat(checkpoint)
            stamp = (st.st_size, st.st_mtime_ns)
        except OSError:
            stamp = ()
        self._ckpt_key = (str(checkpoint), int(rna_features)) + stamp

    def __setstate__(self, state):
        """Instances pickled by the reference (inside a torchgan checkpoint's `loss_objects`, see compat.py) carry
        `betavae`, `reduction`, `override_train_ops`, `arg_map` but none of this package's bookkeeping."""
        super().__setstate__(state)
        if "_ckpt_key" not in self.__dict__:
            self._ckpt_key = ("unpickled", id(self))
        vae = self._modules.get("betavae")
        if vae is not None:
            vae.eval()

    def _encoder(self, device):
        """All three loss objects load the SAME checkpoint (src/histopathology_gan.py:275-277): share one device copy
        (keyed on path + size + mtime at load time, so a checkpoint re-saved at the same path is not confused)."""
        key = self._ckpt_key + (str(device),)
        vae = _SHARED_VAE.get(key)
        if vae is None:
            self.betavae = self.betavae.to(device)
            vae = self.betavae
            _SHARED_VAE[key] = vae
        return vae

    def _inputs(self, generator, real_inputs, device):
        """z = encode(rna) (once per batch) and this step's CPU uniform(-0.3, 0.3) noise draw, both on the device
        (src/wgan_loss.py:96-101).  The first step that sees a batch also starts the host->device copy of its images
        on a side stream, so the critic / GP steps (which need them, :238, :364) do not wait for PCIe."""
        B = real_inputs["image"].size(0)
        rna = real_inputs["rna_data"]
        vae = self._encoder(device)
        # keyed on the encoder too (its identity and parameter versions: another VAE / a reloaded checkpoint misses) and
        # cloned, because encode_mean returns the encoder engine's reusable output buffer
        vkey = (id(vae),) + tuple(p._version for p in vae.parameters())
        z = _CACHE.get("z", rna, device, lambda: vae.encode_mean(rna.to(device, non_blocking=True)).clone(), extra=vkey)
        self._prefetch_real(real_inputs, device)
        eng = generator._engine(flush=False)
        E = generator.encoding_dims
        # pinned staging PER LOSS OBJECT, two buffers used alternately: Trainer.train_iter defers the host
        # synchronisation (to the end of the iteration, or by one more iteration in `train`), so a step's asynchronous
        # copy may still be in flight when the same object draws its next noise; the event makes reuse safe at any lag
        stage = eng.bufs.__dict__.setdefault("_noise_pinned", {})
        ring = stage.get((id(self), B, E))
        if ring is None:
            ring = {"i": 0, "slots": [[torch.empty(B, E, dtype=F32).pin_memory(), None] for _ in range(2)]}
            stage[(id(self), B, E)] = ring
        slot = ring["slots"][ring["i"]]
        ring["i"] ^= 1
        if slot[1] is not None:
            slot[1].synchronize()
        pinned = slot[0]
        pinned.uniform_(-0.3, 0.3)                 # same CPU generator stream as torch.FloatTensor(B, E).uniform_()
        noise_d = eng.bufs.get(f"noise.{type(self).__name__}", (B, E), F32)
        noise_d.copy_(pinned, non_blocking=True)
        if slot[1] is None:
            slot[1] = torch.cuda.Event()
        slot[1].record(torch.cuda.current_stream(device))
        return noise_d, z

    @staticmethod
    def _prefetch_real(real_inputs, device):
        img = real_inputs["image"]

        def start():
            if img.device.type == "cuda":
                return (img.to(dtype=F32).contiguous(), None)
            side = _CACHE.__dict__.setdefault("_copy_stream", torch.cuda.Stream(device=device))
            with torch.cuda.stream(side):
                dst = img.to(device=device, dtype=F32, non_blocking=True).contiguous()
                ev = torch.cuda.Event()
                ev.record(side)
            return (dst, ev)

        return _CACHE.get("image", img, device, start)

    @staticmethod
    def _real(real_inputs, device):
        dst, ev = _VAEConditioned._prefetch_real(real_inputs, device)
        if ev is not None:
            cur = torch.cuda.current_stream(device)
            cur.wait_event(ev)
            dst.record_stream(cur)     # allocated on the copy stream, read here: its block is not reused before this
        return dst                     # stream is done with it (iterations are queued ahead of the host, Trainer.train)


def _eps_to_device(discriminator, eps):
    """The GP step's scalar eps (one CPU draw, src/wgan_loss.py:376) -> device buffer through a pinned slot: an
    asynchronous copy instead of the synchronous pageable one (the host stays ahead of the device, Trainer.train).
    Four slots guarded by events, like the noise staging."""
    eng = discriminator._engine(flush=False)
    ring = eng.bufs.__dict__.setdefault("_eps_pinned", None)
    if ring is None:
        ring = {"i": 0, "buf": torch.empty(4, 1, dtype=F32).pin_memory(), "ev": [None] * 4}
        eng.bufs.__dict__["_eps_pinned"] = ring
    k = ring["i"] = (ring["i"] + 1) % 4
    if ring["ev"][k] is not None:
        ring["ev"][k].synchronize()
    ring["buf"][k].copy_(eps)
    eps_d = eng.bufs.get("eps", (1,), F32)
    eps_d.copy_(ring["buf"][k], non_blocking=True)
    if ring["ev"][k] is None:
        ring["ev"][k] = torch.cuda.Event()
    ring["ev"][k].record(torch.cuda.current_stream(eng.device))
    return eps_d


def _check_labels(labels, *nets):
    if labels is None and any(n.label_type == "required" for n in nets):
        raise Exception("GAN model requires labels for training")
    for n in nets:
        if n.label_type != "none":
            raise NotImplementedError("only label_type='none' is on the sm_100a path (the only one the reference "
                                      "drivers exercise, SURVEY.md Appendix B.10)")


# ------------------------------------------------------------------------------------------------ loss objects
class WassersteinGeneratorLossVAE(_VAEConditioned, GeneratorLoss):
    def __init__(self, checkpoint, rna_features, beta=0.005):
        # the reference passes (checkpoint, rna_features) positionally into (reduction, override_train_ops)
        # (src/wgan_loss.py:64-66); keep the same observable attributes
        GeneratorLoss.__init__(self, checkpoint, rna_features)
        self._init_vae(checkpoint, rna_features, beta)

    def forward(self, fgz):
        return wasserstein_generator_loss_vae(fgz, self.reduction)

    def train_ops(self, generator, discriminator, optimizer_generator, device, batch_size, real_inputs, labels=None):
        return self.device_ops(generator, discriminator, optimizer_generator, device, batch_size, real_inputs,
                               labels).item()

    def device_ops(self, generator, discriminator, optimizer_generator, device, batch_size, real_inputs, labels=None):
        """train_ops without the host read-back: returns the loss as a device tensor [1] (Trainer.train_iter reads
        the three losses of an iteration with ONE synchronisation at its end)."""
        _check_labels(labels, generator)
        noise_d, z = self._inputs(generator, real_inputs, device)
        return steps.g_step(generator, discriminator, optimizer_generator, noise_d, z)


class WassersteinDiscriminatorLossVAE(_VAEConditioned, DiscriminatorLoss):
    def __init__(self, checkpoint, rna_features, beta=0.005, reduction="mean", clip=None, override_train_ops=None):
        DiscriminatorLoss.__init__(self, checkpoint, rna_features)
        self.clip = clip if isinstance(clip, (tuple, list)) and len(clip) > 1 else None
        self._init_vae(checkpoint, rna_features, beta)

    def forward(self, fx, fgz):
        return wasserstein_discriminator_loss_vae(fx, fgz, self.reduction)

    def train_ops(self, generator, discriminator, optimizer_discriminator, real_inputs, device, labels=None):
        return self.device_ops(generator, discriminator, optimizer_discriminator, real_inputs, device, labels).item()

    def device_ops(self, generator, discriminator, optimizer_discriminator, real_inputs, device, labels=None):
        _check_labels(labels, generator, discriminator)
        noise_d, z = self._inputs(generator, real_inputs, device)
        real = self._real(real_inputs, device)
        return steps.critic_step(generator, discriminator, optimizer_discriminator, noise_d, z, real, clip=self.clip)


class WassersteinGradientPenaltyVAE(_VAEConditioned, DiscriminatorLoss):
    def __init__(self, checkpoint, rna_features, reduction="mean", lambd=10.0, override_train_ops=None, beta=0.005):
        DiscriminatorLoss.__init__(self, checkpoint, rna_features)
        self.lambd = lambd
        self.override_train_ops = override_train_ops
        self._init_vae(checkpoint, rna_features, beta)

    def forward(self, interpolate, d_interpolate):
        return wasserstein_gradient_penalty_vae(interpolate, d_interpolate, self.reduction)

    def train_ops(self, generator, discriminator, optimizer_discriminator, real_inputs, device, labels=None):
        return self.device_ops(generator, discriminator, optimizer_discriminator, real_inputs, device, labels).item()

    def device_ops(self, generator, discriminator, optimizer_discriminator, real_inputs, device, labels=None):
        _check_labels(labels, generator, discriminator)
        noise_d, z = self._inputs(generator, real_inputs, device)
        real = self._real(real_inputs, device)
        eps = torch.rand(1)                                              # one scalar per step, src/wgan_loss.py:376
        eps_d = _eps_to_device(discriminator, eps)
        out3 = steps.gp_step(generator, discriminator, optimizer_discriminator, noise_d, z, real, eps_d,
                             lambd=self.lambd)
        return out3[0:1]




# ------------------------------------------------------------------------------------------------ un-conditioned WGAN
# The README's comparison model (`--loss_type wgan`, src/histopathology_gan.py:267-272) uses torchgan's own
# WassersteinGeneratorLoss / WassersteinDiscriminatorLoss(clip=(-0.01, 0.01)) / WassersteinGradientPenalty [tg]: the same
# three steps without the betaVAE conditioning -- latent = randn(batch, encoding_dims, device=device) as is -- plus the
# critic weight clamp before the critic step.  Same kernels (steps.py with z=None).
def _images_of(real_inputs, device):
    img = real_inputs["image"] if isinstance(real_inputs, dict) else real_inputs
    return img.to(device=device, dtype=F32, non_blocking=True).contiguous()


class WassersteinGeneratorLoss(GeneratorLoss):
    def forward(self, fgz):
        return reduce_vae(-1.0 * fgz, self.reduction)

    def train_ops(self, generator, discriminator, optimizer_generator, device, batch_size, labels=None):
        return self.device_ops(generator, discriminator, optimizer_generator, device, batch_size, labels).item()

    def device_ops(self, generator, discriminator, optimizer_generator, device, batch_size, labels=None):
        _check_labels(labels, generator)
        noise = torch.randn(batch_size, generator.encoding_dims, device=device)       # torchgan: device RNG [tg]
        return steps.g_step(generator, discriminator, optimizer_generator, noise, None)


class WassersteinDiscriminatorLoss(DiscriminatorLoss):
    def __init__(self, reduction="mean", clip=None, override_train_ops=None):
        super().__init__(reduction, override_train_ops)
        self.clip = clip if isinstance(clip, (tuple, list)) and len(clip) > 1 else None

    def forward(self, fx, fgz):
        return reduce_vae(fgz - fx, self.reduction)

    def train_ops(self, generator, discriminator, optimizer_discriminator, real_inputs, device, labels=None):
        return self.device_ops(generator, discriminator, optimizer_discriminator, real_inputs, device, labels).item()

    def device_ops(self, generator, discriminator, optimizer_discriminator, real_inputs, device, labels=None):
        _check_labels(labels, generator, discriminator)
        real = _images_of(real_inputs, device)
        noise = torch.randn(real.size(0), generator.encoding_dims, device=device)
        return steps.critic_step(generator, discriminator, optimizer_discriminator, noise, None, real, clip=self.clip)


class WassersteinGradientPenalty(DiscriminatorLoss):
    def __init__(self, reduction="mean", lambd=10.0, override_train_ops=None):
        super().__init__(reduction, override_train_ops)
        self.lambd = lambd

    def forward(self, interpolate, d_interpolate):
        return wasserstein_gradient_penalty_vae(interpolate, d_interpolate, self.reduction)

    def train_ops(self, generator, discriminator, optimizer_discriminator, real_inputs, device, labels=None):
        return self.device_ops(generator, discriminator, optimizer_discriminator, real_inputs, device, labels).item()

    def device_ops(self, generator, discriminator, optimizer_discriminator, real_inputs, device, labels=None):
        _check_labels(labels, generator, discriminator)
        real = _images_of(real_inputs, device)
        noise = torch.randn(real.size(0), generator.encoding_dims, device=device)
        eps = torch.rand(1)                                                # CPU draw, like torchgan's .item() [tg]
        eps_d = _eps_to_device(discriminator, eps)
        out3 = steps.gp_step(generator, discriminator, optimizer_discriminator, noise, None, real, eps_d,
                             lambd=self.lambd)
        return out3[0:1]
