"""Drop-in betaVAE (src/betaVAE.py:18-162): same constructor, attribute / state_dict names, methods.

On the GAN path the model is frozen and in eval mode (src/wgan_loss.py:67-69) and only ``encode(x)[0]`` is used
(:97); that path runs on the tcgen05 GEMM engine with eval-BatchNorm1d + LeakyReLU(0.01) folded into the epilogues
(rnagan_b200.engine.EncoderEngine).  The layers below are parameter containers for the checkpoint layout.
"""
from typing import List

import torch
import torch.nn as nn
from torch.nn import functional as F

from . import engine as _engine


class RNAEncoder(nn.Module):
    def __init__(self, in_channels: int, hidden_dims: List):
        super().__init__()
        self.in_channels = in_channels
        mods = [nn.Sequential(nn.Dropout())]
        width = in_channels
        for h in hidden_dims:
            mods.append(nn.Sequential(nn.Linear(width, h), nn.BatchNorm1d(h), nn.LeakyReLU()))
            width = h
        self.encoder = nn.Sequential(*mods)

    def forward(self, x):
        raise NotImplementedError("use betaVAE.encode (CUDA engine); the container layers are never called")


class betaVAE(nn.Module):
    def __init__(self, in_channels: int, z_dim: int, encoder_dims: List, hidden_dims_decoder: List, beta: int = 2,
                 encoder_checkpoint=None):
        super().__init__()
        self.encoder = RNAEncoder(in_channels, encoder_dims)
        if encoder_checkpoint:
            self.encoder.load_state_dict(torch.load(encoder_checkpoint))
        self.z_mu = nn.Linear(z_dim, z_dim)
        self.z_logvar = nn.Linear(z_dim, z_dim)
        self.beta = beta
        mods = []
        width = z_dim
        for h in hidden_dims_decoder:
            mods.append(nn.Sequential(nn.Linear(width, h), nn.BatchNorm1d(h), nn.LeakyReLU()))
            width = h
        mods.append(nn.Sequential(nn.Linear(width, in_channels), nn.Tanh()))
        self.decoder = nn.Sequential(*mods)
        self.z_dim = z_dim

    # -- engine plumbing ---------------------------------------------------------------------------------------
    def __getstate__(self):
        """Pickling (torchgan checkpoints hold the loss objects and, through them, this module) and deepcopy carry the
        parameters and buffers only: the engines (`_rg_*`: device workspaces, bf16 operand copies) are rebuilt lazily."""
        return {k: v for k, v in self.__dict__.items() if not k.startswith("_rg_")}

    def _engine(self):
        p0 = next(self.parameters())
        if p0.device.type != "cuda":
            raise RuntimeError("betaVAE.encode runs only on a CUDA (sm_100a) device: there is no CPU fallback")
        if self.training:
            raise NotImplementedError("betaVAE on the sm_100a path is the frozen eval-mode conditioning encoder "
                                      "(src/wgan_loss.py:67-69); call .eval() first")
        vers = tuple((p.data_ptr(), p._version) for p in self.state_dict().values())
        eng = self.__dict__.get("_rg_engine")
        if eng is None or eng.device != p0.device:
            eng = _engine.EncoderEngine(self)
            self.__dict__["_rg_engine"] = eng
        elif vers != self.__dict__.get("_rg_versions") or self.__dict__.get("_rg_dirty", False):
            eng.refresh()            # weights changed (load_state_dict / a fused training step)
        self.__dict__["_rg_versions"] = vers
        self.__dict__["_rg_dirty"] = False
        return eng

    def _train_engine(self):
        p0 = next(self.parameters())
        if p0.device.type != "cuda":
            raise RuntimeError("betaVAE training runs only on a CUDA (sm_100a) device: there is no CPU fallback")
        eng = self.__dict__.get("_rg_train_engine")
        vers = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if eng is None or eng.device != p0.device:
            eng = _engine.VAETrainEngine(self)
            self.__dict__["_rg_train_engine"] = eng
        elif vers != self.__dict__.get("_rg_train_versions"):
            eng.pack()
        self.__dict__["_rg_train_versions"] = vers
        return eng

    def _dec_engine(self):
        p0 = next(self.parameters())
        if p0.device.type != "cuda":
            raise RuntimeError("betaVAE.decode runs only on a CUDA (sm_100a) device: there is no CPU fallback")
        if self.training:
            raise NotImplementedError("betaVAE.decode / forward / sample run in eval mode on the sm_100a path "
                                      "(running BatchNorm statistics); training goes through betaVAE.train_step")
        vers = tuple((p.data_ptr(), p._version) for p in self.decoder.state_dict().values())
        eng = self.__dict__.get("_rg_dec_engine")
        if eng is None or eng.device != p0.device:
            eng = _engine.DecoderEngine(self)
            self.__dict__["_rg_dec_engine"] = eng
        elif vers != self.__dict__.get("_rg_dec_versions") or self.__dict__.get("_rg_dec_dirty", False):
            eng.refresh()
        self.__dict__["_rg_dec_versions"] = vers
        self.__dict__["_rg_dec_dirty"] = False
        return eng

    def _apply(self, fn, *args, **kwargs):
        before = [p.data_ptr() for p in self.parameters()]
        out = super()._apply(fn, *args, **kwargs)
        if before != [p.data_ptr() for p in self.parameters()]:      # .to()/.cuda()/.float() really moved the storage
            self.__dict__.pop("_rg_engine", None)
            self.__dict__.pop("_rg_train_engine", None)
            self.__dict__.pop("_rg_dec_engine", None)
        return out

    # -- reference API -----------------------------------------------------------------------------------------
    @torch.no_grad()
    def encode(self, x):
        eng = self._engine()
        x = x.to(device=eng.device, dtype=torch.float32)
        z_mean, z_log_var, h = eng.encode(x, want_all=True)
        return z_mean.clone(), z_log_var.clone(), h

    @torch.no_grad()
    def encode_mean(self, x):
        """z_mean only (what the GAN path consumes); returns the engine's reusable buffer."""
        eng = self._engine()
        return eng.encode(x.to(device=eng.device, dtype=torch.float32))

    def reparametrize(self, z_mean, z_log_var):
        std = torch.exp(0.5 * z_log_var)
        return z_mean + torch.randn_like(std) * std

    @torch.no_grad()
    def decode(self, x):
        """src/betaVAE.py:142-143 (eval mode): tanh(decoder(x)) on the tcgen05 GEMMs."""
        eng = self._dec_engine()
        return eng.decode(x.to(device=eng.device, dtype=torch.float32))

    @torch.no_grad()
    def forward(self, x):
        """src/betaVAE.py:109-115 (eval mode): (decoder(reparametrize(z_mean, z_log_var)), z_mean, z_log_var)."""
        z_mean, z_log_var, _ = self.encode(x)
        z = self.reparametrize(z_mean, z_log_var)
        return self.decode(z), z_mean, z_log_var

    @torch.no_grad()
    def sample(self, num_samples, current_device, interpolation=None, alpha=1.0):
        """src/betaVAE.py:117-140: decode CPU-drawn N(0, I) latents (+ alpha * interpolation)."""
        z = torch.randn(num_samples, self.z_dim)
        z = z.to(current_device)
        if interpolation is not None:
            z = z + torch.from_numpy(alpha * interpolation).float().to(current_device)
        return self.decode(z)


def betaVAEloss(x, x_recons, z_mean, z_logvar, beta, kld_weight=0.005, training=True):
    """src/betaVAE.py:145-162 (plain tensor math on whatever device the inputs live on)."""
    recons_loss = F.mse_loss(x_recons, x)
    kld_loss = torch.mean(-0.5 * torch.sum(1 + z_logvar - z_mean ** 2 - z_logvar.exp(), dim=1), dim=0)
    total_loss = recons_loss + beta * kld_loss if training else recons_loss
    return {"total_loss": total_loss, "reconstruction_loss": recons_loss, "kl_loss": kld_loss}


def train_step(model, optimizer, x, beta, keep_mask=None, eps=None):
    """One betaVAE optimisation step on the sm_100a kernels: the body of the batch loop of ``train_betaVAE``
    (src/betaVAE.py:221-235: zero_grad, forward, betaVAEloss, backward, optimizer.step()).

    x: fp32 [B, in_channels].  keep_mask / eps: optional explicit Dropout keep mask (fp32 0/1 [B, in_channels]) and
    reparametrisation noise (fp32 [B, z_dim]); drawn on the device when omitted.
    Returns a device tensor [total_loss, reconstruction_loss, kl_loss] (no host synchronisation)."""
    from .optim import adam_step
    if not model.training:
        raise RuntimeError("train_step needs model.train() (BatchNorm batch statistics, Dropout)")
    eng = model._train_engine()
    x = x.to(device=eng.device, dtype=torch.float32).contiguous()
    out3 = eng.step(x, beta, keep_mask=keep_mask, eps=eps)
    adam_step(optimizer, grad_scale=eng.sync.finish())
    eng.pack(full=False)          # Adam re-emitted the unpadded bf16 operand copies itself
    model.__dict__["_rg_dirty"] = True
    model.__dict__["_rg_dec_dirty"] = True
    return out3


def _loss_means(acc, n):
    m = (acc / max(n, 1)).tolist()
    return {"total_loss": m[0], "reconstruction_loss": m[1], "kl_loss": m[2]}


def train_betaVAE(model, optimizer, dataloader, save_dir="checkpoints/models/", device=None, log_interval=100,
                  summary_writer=None, num_epochs=100, scheduler=None, verbose=True):
    """Drop-in for the reference's ``train_betaVAE`` (src/betaVAE.py:164-284), same arguments and return value:
    ``dataloader`` is ``{'train': ..., 'val': ...}`` of ``{'rna_data': [B, genes]}`` batches; every epoch runs the train
    phase (``train_step`` per batch, then ``scheduler.step()`` -- any torch scheduler works, the fused Adam reads
    ``param_groups[...]['lr']`` each step) and the val phase (eval-mode forward, total = reconstruction loss,
    src/betaVAE.py:159-160); the best validation epoch is written to ``{save_dir}/model_dict_best.pt``, the final weights
    to ``model_last.pt``, and the best ones are loaded back before returning ``(model, {'best_epoch', 'best_loss'})``.
    Losses stay on the device during an epoch (one read-back per phase instead of three ``.item()`` per batch);
    ``summary_writer.add_scalar`` is fed at ``log_interval`` like upstream when a writer is given."""
    import os

    import numpy as np
    os.makedirs(save_dir, exist_ok=True)
    dev = next(model.parameters()).device
    best_epoch, best_loss = 0, {"total_loss": np.inf}
    global_step = {"train": 0, "val": 0}
    for epoch in range(num_epochs):
        if verbose:
            print("Epoch {}/{}".format(epoch, num_epochs - 1))
            print("-" * 10)
        for phase in ("train", "val"):
            model.train(phase == "train")
            acc, n = torch.zeros(3, device=dev), 0
            last = torch.zeros(3, device=dev)
            step = global_step[phase]
            for batch in dataloader[phase]:
                x = (batch["rna_data"] if isinstance(batch, dict) else batch).to(dev)
                if phase == "train":
                    acc += train_step(model, optimizer, x, model.beta)
                    if scheduler:
                        scheduler.step()
                else:
                    with torch.no_grad():
                        out, z_mean, z_log_var = model(x)
                        ls = betaVAEloss(x, out, z_mean, z_log_var, model.beta, training=False)
                        acc += torch.stack([ls["total_loss"], ls["reconstruction_loss"], ls["kl_loss"]])
                n += 1
                step += 1
                if summary_writer is not None and step % log_interval == 0:
                    mean = acc / n                                  # upstream logs the change of the running mean
                    for k, key in enumerate(("total_loss", "reconstruction_loss", "kl_loss")):
                        summary_writer.add_scalar("{}/{}".format(phase, key), float(mean[k] - last[k]), step)
                    last = mean.clone()
            global_step[phase] = step
            epoch_loss = _loss_means(acc, n)
            if verbose:
                print("{} Total Loss: {:.4f} | Reconstruction Loss: {:.4f} | KL Loss: {:.4f}".format(
                    phase, epoch_loss["total_loss"], epoch_loss["reconstruction_loss"], epoch_loss["kl_loss"]))
            if phase == "val" and epoch_loss["total_loss"] < best_loss["total_loss"]:
                best_loss["total_loss"] = epoch_loss["total_loss"]
                torch.save(model.state_dict(), os.path.join(save_dir, "model_dict_best.pt"))
                best_epoch = epoch
    torch.save(model.state_dict(), os.path.join(save_dir, "model_last.pt"))
    model.load_state_dict(torch.load(os.path.join(save_dir, "model_dict_best.pt")))
    return model, {"best_epoch": best_epoch, "best_loss": best_loss}


def evaluate_betaVAE(model, dataloader, verbose=True):
    """Drop-in for src/betaVAE.py:286-331: eval-mode losses over a loader plus the predictions / inputs as nested lists."""
    import numpy as np
    model.eval()
    dev = next(model.parameters()).device
    running = {"total_loss": [], "reconstruction_loss": [], "kl_loss": []}
    predictions, real = [], []
    for batch in dataloader:
        x = (batch["rna_data"] if isinstance(batch, dict) else batch).to(dev)
        with torch.no_grad():
            out, z_mean, z_log_var = model(x)
            ls = betaVAEloss(x, out, z_mean, z_log_var, model.beta, training=False)
        predictions.append(out.detach().cpu().numpy().tolist())
        real.append(x.detach().cpu().numpy().tolist())
        for k in running:
            running[k].append(ls[k].item())
    test_loss = {k: np.mean(v) for k, v in running.items()}
    if verbose:
        print("Total Loss: {:.4f} | Reconstruction Loss: {:.4f} | KL Loss: {:.4f}".format(
            test_loss["total_loss"], test_loss["reconstruction_loss"], test_loss["kl_loss"]))
    return test_loss, predictions, real


class GradualWarmupScheduler(torch.optim.lr_scheduler._LRScheduler):
    """The warm-up wrapper the reference driver stacks on CosineAnnealingLR (src/betaVAE_training.py:14,165-166; the
    third-party `warmup_scheduler` package, not installed here): the learning rate rises linearly from base_lr to
    multiplier * base_lr (from 0 when multiplier == 1) over `total_epoch` steps, then `after_scheduler` takes over."""

    def __init__(self, optimizer, multiplier, total_epoch, after_scheduler=None):
        if multiplier < 1.0:
            raise ValueError("multiplier should be greater than or equal to 1.")
        self.multiplier, self.total_epoch, self.after_scheduler, self.finished = multiplier, total_epoch, after_scheduler, False
        super().__init__(optimizer)

    def get_lr(self):
        if self.last_epoch > self.total_epoch:
            if self.after_scheduler:
                if not self.finished:
                    self.after_scheduler.base_lrs = [b * self.multiplier for b in self.base_lrs]
                    self.finished = True
                return self.after_scheduler.get_last_lr()
            return [b * self.multiplier for b in self.base_lrs]
        if self.multiplier == 1.0:
            return [b * (float(self.last_epoch) / self.total_epoch) for b in self.base_lrs]
        return [b * ((self.multiplier - 1.0) * self.last_epoch / self.total_epoch + 1.0) for b in self.base_lrs]

    def step(self, epoch=None):
        if self.finished and self.after_scheduler:
            self.after_scheduler.step(None if epoch is None else epoch - self.total_epoch)
            self._last_lr = self.after_scheduler.get_last_lr()
            self.last_epoch += 1
        else:
            super().step(epoch)
