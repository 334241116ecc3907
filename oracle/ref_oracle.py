"""CPU fp32 ORACLE for the RNA-GAN hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.
The product (rnagan_b200/) never does: it fails loudly when its CUDA extension is missing.

It restates, in plain PyTorch CPU ops, the algorithm of the reference path (file:line relative to /root/reference):

  * torchgan==0.1.0 DCGANGenerator / DCGANDiscriminator (third-party, pinned at requirements.txt:155, NOT vendored
    in the reference and not installed here; structure per SURVEY.md Appendix A, corroborated by the reference's
    edited copy src/dcgan.py:23-44,52,58-74,82 and the ctor dicts src/histopathology_gan.py:178-192)
  * DCGANUpGenerator                       src/dcgan.py:8-99
  * RNAEncoder / betaVAE / betaVAEloss     src/betaVAE.py:18-42, 63-143, 145-162
  * latent prep                            src/wgan_loss.py:100-106 (= :227-233, :357-363, src/gan_utils.py:211-216)
  * the three train_ops                    src/wgan_loss.py:82-129, 181-263, 314-389
  * Trainer.train_iter loss order          torchgan [tg]; call site src/histopathology_gan.py:298-314
  * generate_images                        src/gan_utils.py:197-244

PINNING: the reference ships no tests, golden vectors or fixtures for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/make_golden.py imports the reference's own
src/dcgan.py, src/wgan_loss.py, src/betaVAE.py (through oracle/torchgan_shim) in the build container, runs them
on seeded inputs and commits the results under tests/golden/; tests/test_oracle_cpu.py checks this restatement
against those fixtures.  At the torchgan boundary itself parity is UNPINNED (no upstream source or tests available
offline) -- stated in DESIGN.md.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

LATENT = 2048
ENC_DIMS = (6000, 4000, 2048)
DEC_DIMS = (4000, 6000)


# ------------------------------------------------------------------------------------------------ networks
def _seq(*mods):
    return nn.Sequential(*mods)


def _repeats(size, what):
    if size < 16 or math.ceil(math.log2(size)) != math.log2(size):
        raise Exception(f"{what} Image Size must be at least 16*16 and an exact power of 2")
    return size.bit_length() - 4


class OracleGenerator(nn.Module):
    """Transposed-conv generator (what histopathology_gan.py:176 instantiates)."""

    def __init__(self, encoding_dims=100, out_size=32, out_channels=3, step_channels=64, batchnorm=True,
                 nonlinearity=None, last_nonlinearity=None, label_type="none"):
        super().__init__()
        reps = _repeats(out_size, "Target")
        self.encoding_dims, self.label_type = encoding_dims, label_type
        act = nonlinearity if nonlinearity is not None else nn.LeakyReLU(0.2)
        last = last_nonlinearity if last_nonlinearity is not None else nn.Tanh()
        width = step_channels << reps
        stack = []
        cin, k, s, p = encoding_dims, 4, 1, 0
        for _ in range(reps + 1):
            parts = [nn.ConvTranspose2d(cin, width, k, s, p, bias=not batchnorm)]
            if batchnorm:
                parts.append(nn.BatchNorm2d(width))
            stack.append(_seq(*parts, act))
            cin, width, s, p = width, width // 2, 2, 1
        stack.append(_seq(nn.ConvTranspose2d(cin, out_channels, 4, 2, 1, bias=True), last))
        self.model = _seq(*stack)
        init_like_torchgan(self, nn.ConvTranspose2d)

    def forward(self, z, feature_matching=False):
        return self.model(z.view(-1, z.size(1), 1, 1))

    def sampler(self, sample_size, device):
        return [torch.randn(sample_size, self.encoding_dims, device=device)]


class OracleUpGenerator(nn.Module):
    """Resize-conv generator, src/dcgan.py:8-99 (bilinear x2 -> reflect pad 1 -> 3x3 conv; no final Tanh)."""

    def __init__(self, encoding_dims=100, out_size=32, out_channels=3, step_channels=64, batchnorm=True,
                 nonlinearity=None, last_nonlinearity=None, label_type="none"):
        super().__init__()
        reps = _repeats(out_size, "Target")
        self.encoding_dims, self.label_type = encoding_dims, label_type
        act = nonlinearity if nonlinearity is not None else nn.LeakyReLU(0.2)
        width = step_channels << reps
        first = [nn.ConvTranspose2d(encoding_dims, width, 4, 1, 0, bias=not batchnorm)]
        if batchnorm:
            first.append(nn.BatchNorm2d(width))
        stack = [_seq(*first, act)]
        for _ in range(reps):
            if batchnorm:
                stack.append(_seq(nn.Upsample(scale_factor=2, mode="bilinear"), nn.ReflectionPad2d(1),
                                  nn.Conv2d(width, width // 2, kernel_size=3, stride=1, padding=0),
                                  nn.BatchNorm2d(width // 2), act))
            else:
                stack.append(_seq(nn.ConvTranspose2d(width, width // 2, 4, 2, 1, bias=True), act))
            width //= 2
        stack.append(_seq(nn.Upsample(scale_factor=2, mode="bilinear"), nn.ReflectionPad2d(1),
                          nn.Conv2d(width, out_channels, kernel_size=3, stride=1, padding=0)))
        self.model = _seq(*stack)
        init_like_torchgan(self, nn.ConvTranspose2d)

    def forward(self, z, feature_matching=False):
        return self.model(z.view(-1, z.size(1), 1, 1))


class OracleCritic(nn.Module):
    """torchgan DCGANDiscriminator (histopathology_gan.py:177,186-192)."""

    def __init__(self, in_size=32, in_channels=3, step_channels=64, batchnorm=True, nonlinearity=None,
                 last_nonlinearity=None, label_type="none"):
        super().__init__()
        reps = _repeats(in_size, "Input")
        self.input_dims, self.label_type = in_channels, label_type
        act = nonlinearity if nonlinearity is not None else nn.LeakyReLU(0.2)
        last = last_nonlinearity if last_nonlinearity is not None else nn.LeakyReLU(0.2)
        width = step_channels
        stack = [_seq(nn.Conv2d(in_channels, width, 4, 2, 1, bias=True), act)]
        for _ in range(reps):
            parts = [nn.Conv2d(width, 2 * width, 4, 2, 1, bias=not batchnorm)]
            if batchnorm:
                parts.append(nn.BatchNorm2d(2 * width))
            stack.append(_seq(*parts, act))
            width *= 2
        self.disc = _seq(nn.Conv2d(width, 1, 4, 1, 0, bias=not batchnorm), last)
        self.model = _seq(*stack)
        init_like_torchgan(self, nn.Conv2d)

    def forward(self, x, feature_matching=False):
        feats = self.model(x)
        if feature_matching:
            return feats
        return self.disc(feats).view(feats.size(0))


def init_like_torchgan(net, conv_type):
    """torchgan's _weight_initializer [tg]: kaiming-normal conv/linear weights, zero biases, BN gamma=1 beta=0."""
    for m in net.modules():
        if isinstance(m, (conv_type, nn.Linear)):
            nn.init.kaiming_normal_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0.0)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1.0)
            nn.init.constant_(m.bias, 0.0)


class OracleVAE(nn.Module):
    """betaVAE (src/betaVAE.py:63-143): encoder = Dropout + 3x[Linear, BatchNorm1d, LeakyReLU(0.01)]."""

    def __init__(self, in_channels, z_dim=LATENT, encoder_dims=ENC_DIMS, hidden_dims_decoder=DEC_DIMS, beta=2):
        super().__init__()

        def block(i, o):
            return _seq(nn.Linear(i, o), nn.BatchNorm1d(o), nn.LeakyReLU())

        enc = [_seq(nn.Dropout())]
        width = in_channels
        for h in encoder_dims:
            enc.append(block(width, h))
            width = h
        self.encoder = nn.Module()
        self.encoder.encoder = _seq(*enc)
        self.z_mu = nn.Linear(z_dim, z_dim)
        self.z_logvar = nn.Linear(z_dim, z_dim)
        dec = []
        width = z_dim
        for h in hidden_dims_decoder:
            dec.append(block(width, h))
            width = h
        dec.append(_seq(nn.Linear(width, in_channels), nn.Tanh()))
        self.decoder = _seq(*dec)
        self.beta, self.z_dim = beta, z_dim

    def encode(self, x):
        h = self.encoder.encoder(x)
        return self.z_mu(h), self.z_logvar(h), h

    def reparametrize(self, mu, logvar):
        std = torch.exp(0.5 * logvar)
        return mu + torch.randn_like(std) * std

    def decode(self, z):
        return self.decoder(z)

    def forward(self, x):
        mu, logvar, _ = self.encode(x)
        return self.decoder(self.reparametrize(mu, logvar)), mu, logvar


def vae_loss(x, recon, mu, logvar, beta, training=True):
    """betaVAEloss, src/betaVAE.py:145-162."""
    rec = F.mse_loss(recon, x)
    kld = torch.mean(-0.5 * torch.sum(1 + logvar - mu ** 2 - logvar.exp(), dim=1), dim=0)
    total = rec + beta * kld if training else rec
    return {"total_loss": total, "reconstruction_loss": rec, "kl_loss": kld}


def vae_train_step(vae, optimizer, x, beta):
    """Inner step of train_betaVAE, src/betaVAE.py:221-235 (scheduler handled by the caller)."""
    optimizer.zero_grad(set_to_none=True)
    recon, mu, logvar = vae(x)
    losses = vae_loss(x, recon, mu, logvar, beta, training=True)
    losses["total_loss"].backward()
    optimizer.step()
    return {k: v.item() for k, v in losses.items()}


def vae_train_step_explicit(vae, optimizer, x, beta, keep_mask, eps):
    """Same step with the two random draws made explicit (Dropout keep mask, reparametrisation noise), so that a
    device implementation with its own RNG can be compared on identical randomness.  Equals vae_train_step when
    keep_mask / eps are the tensors F.dropout / randn_like would have drawn."""
    optimizer.zero_grad(set_to_none=True)
    p = vae.encoder.encoder[0][0].p
    h = x * keep_mask / (1.0 - p)
    for blk in list(vae.encoder.encoder)[1:]:
        h = blk(h)
    mu, logvar = vae.z_mu(h), vae.z_logvar(h)
    recon = vae.decoder(mu + eps * torch.exp(0.5 * logvar))
    losses = vae_loss(x, recon, mu, logvar, beta, training=True)
    losses["total_loss"].backward()
    optimizer.step()
    return {k: v.item() for k, v in losses.items()}


# ------------------------------------------------------------------------------------------------ step logic
def draw_noise(batch, dims):
    """CPU uniform(-0.3, 0.3) draw from the global generator, src/wgan_loss.py:100."""
    return torch.FloatTensor(batch, dims).uniform_(-0.3, 0.3)


def latent_prep(noise, z):
    """additive conditioning + per-feature batch standardisation with UNBIASED std, src/wgan_loss.py:105-106."""
    n = noise + z
    return (n - torch.mean(n, dim=0)) / torch.std(n, dim=0)


def _need_labels(labels, *nets):
    if labels is None and any(n.label_type == "required" for n in nets):
        raise Exception("GAN model requires labels for training")


def g_step(G, D, opt_g, vae, batch, labels=None):
    """WassersteinGeneratorLossVAE.train_ops, src/wgan_loss.py:82-129."""
    _need_labels(labels, G)
    B = batch["image"].size(0)
    z = vae.encode(batch["rna_data"])[0]
    lat = latent_prep(draw_noise(B, G.encoding_dims), z)
    opt_g.zero_grad()
    loss = torch.mean(-1.0 * D(G(lat)))
    loss.backward()
    opt_g.step()
    return loss.item()


def critic_step(G, D, opt_d, vae, batch, clip=None, labels=None):
    """WassersteinDiscriminatorLossVAE.train_ops, src/wgan_loss.py:181-263."""
    if clip is not None:
        for p in D.parameters():
            p.data.clamp_(clip[0], clip[1])
    _need_labels(labels, G, D)
    B = batch["image"].size(0)
    z = vae.encode(batch["rna_data"])[0]
    lat = latent_prep(draw_noise(B, G.encoding_dims), z)
    opt_d.zero_grad()
    dx = D(batch["image"])
    dgz = D(G(lat).detach())
    loss = torch.mean(dgz - dx)
    loss.backward()
    opt_d.step()
    return loss.item()


def gradient_penalty(interp, d_interp):
    """wasserstein_gradient_penalty_vae, src/wgan_loss.py:32-44: ONE norm over the whole batch tensor."""
    g = torch.autograd.grad(d_interp, interp, torch.ones_like(d_interp), create_graph=True, retain_graph=True,
                            only_inputs=True)[0]
    return (g.norm(2) - 1) ** 2


def gp_step(G, D, opt_d, vae, batch, lambd=10.0, labels=None):
    """WassersteinGradientPenaltyVAE.train_ops, src/wgan_loss.py:314-389 (scalar eps drawn after the noise)."""
    _need_labels(labels, G, D)
    B = batch["image"].size(0)
    z = vae.encode(batch["rna_data"])[0].detach()
    lat = latent_prep(draw_noise(B, G.encoding_dims), z)
    opt_d.zero_grad()
    fake = G(lat)
    eps = torch.rand(1).item()
    interp = eps * batch["image"] + (1 - eps) * fake
    loss = gradient_penalty(interp, D(interp))
    (lambd * loss).backward()
    opt_d.step()
    return loss.item()


def train_iter(G, D, opt_g, opt_d, vae, batch):
    """torchgan Trainer.train_iter with losses [G, critic, GP] and ncritic=1 (histopathology_gan.py:273-278)."""
    return (g_step(G, D, opt_g, vae, batch), critic_step(G, D, opt_d, vae, batch), gp_step(G, D, opt_d, vae, batch))


def synth_tiles(G, vae, gene_exp, sample_size, chunk=10):
    """generate_images, src/gan_utils.py:197-244: train-mode G on chunks of 10, (x+1)/2, NHWC float32 numpy."""
    noise = draw_noise(sample_size, G.encoding_dims)
    z = vae.encode(gene_exp)[0].detach()
    lat = latent_prep(noise, z)
    outs = [G(part).detach().cpu().numpy() for part in torch.split(lat, chunk)]
    img = torch.from_numpy(np.concatenate(outs, axis=0)).view(-1, outs[0].shape[1], outs[0].shape[2], outs[0].shape[3])
    img = (img + 1.0) / 2.0
    return img.permute(0, 2, 3, 1).contiguous().numpy()


# ------------------------------------------------------------------------------------------------ test helpers
def reinit_(module, seed):
    """Deterministic re-initialisation in state_dict order, independent of module construction RNG use, so the
    reference modules (golden generation) and the oracle/product modules (tests) start from identical weights."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, t in module.state_dict().items():
            if not t.is_floating_point():
                t.zero_()
                continue
            leaf = name.rsplit(".", 1)[-1]
            if leaf == "running_mean":
                t.copy_(torch.randn(t.shape, generator=g) * 0.05)
            elif leaf == "running_var":
                t.copy_(1.0 + 0.1 * torch.rand(t.shape, generator=g))
            elif t.dim() == 1 and leaf == "weight":       # BatchNorm gamma
                t.copy_(1.0 + 0.1 * torch.randn(t.shape, generator=g))
            elif t.dim() == 1:                            # biases / BatchNorm beta
                t.copy_(0.05 * torch.randn(t.shape, generator=g))
            else:
                fan_in = t[0].numel() if t.dim() > 1 else t.numel()
                t.copy_(torch.randn(t.shape, generator=g) * math.sqrt(2.0 / fan_in))
    return module


def make_batch(batch, rna_features, size, seed):
    g = torch.Generator().manual_seed(seed)
    return {
        "image": torch.rand(batch, 3, size, size, generator=g) * 2.0 - 1.0,
        "rna_data": torch.randn(batch, rna_features, generator=g),
        "labels": torch.zeros(batch),
    }


# ------------------------------------------------------------------------------------------------ un-conditioned WGAN
# torchgan's own WassersteinGeneratorLoss / WassersteinDiscriminatorLoss(clip) / WassersteinGradientPenalty default
# train_ops [tg] (used by `--loss_type wgan`, src/histopathology_gan.py:267-272).  torchgan is absent from the reference
# tree: restated from the package's published source; parity at this boundary is UNPINNED.  The device-RNG noise draw
# is an explicit argument here (the tests draw it on the GPU with a fixed seed and hand the same values to both sides).
def plain_g_step(G, D, opt_g, noise):
    opt_g.zero_grad()
    loss = torch.mean(-1.0 * D(G(noise)))
    loss.backward()
    opt_g.step()
    return loss.item()


def plain_critic_step(G, D, opt_d, noise, real, clip=None):
    if clip is not None:
        for p in D.parameters():
            p.data.clamp_(clip[0], clip[1])
    opt_d.zero_grad()
    dx = D(real)
    dgz = D(G(noise).detach())
    loss = torch.mean(dgz - dx)
    loss.backward()
    opt_d.step()
    return loss.item()


def plain_gp_step(G, D, opt_d, noise, real, lambd=10.0):
    fake = G(noise)
    opt_d.zero_grad()
    eps = torch.rand(1).item()
    interp = eps * real + (1 - eps) * fake
    loss = gradient_penalty(interp, D(interp))
    (lambd * loss).backward()
    opt_d.step()
    return loss.item()
