"""Minimal stand-in for the slice of ``torchgan.trainer.Trainer`` that the reference drives
(src/histopathology_gan.py:298-314, src/gan_utils.py:286-297; contract in SURVEY.md Appendix A):

  * builds each model / optimizer from the spec dict and exposes them as ``trainer.generator``,
    ``trainer.optimizer_generator`` ...;
  * binds every loss object's ``train_ops`` parameters BY NAME to same-named trainer attributes;
  * ``train_iter`` runs the losses in list order (G loss, critic loss, gradient penalty; ncritic=1);
  * ``save_model`` / ``load_model`` use the torchgan checkpoint dictionary layout
    (epoch, loss_information, loss_objects, metric_objects, loss_logs, metric_logs + one state_dict per model/optimizer);
  * at the end of every epoch ``train`` checkpoints, switches the models to eval mode and writes the generator's sample
    grid on a fixed ``test_noise`` to ``{recon}/epoch{N}_{model}.png`` (torchgan's end-of-epoch image log [tg]).

Console / tensorboard logging and metrics (torchgan Logger) are out of scope.
"""
import inspect
import os

import torch

from . import compat
from .image_grid import save_image_grid
from .wgan_loss import DiscriminatorLoss, GeneratorLoss


class Trainer:
    def __init__(self, models, losses_list, metrics_list=None, device=None, ncritic=1, epochs=5, sample_size=8,
                 checkpoints="./model/gan", retain_checkpoints=5, recon="./images", log_dir=None, test_noise=None,
                 nrow=8, **kwargs):
        self.device = torch.device("cuda:0") if device is None else torch.device(device)
        self.model_names, self.optimizer_names = [], []
        for key, spec in models.items():
            self.model_names.append(key)
            setattr(self, key, spec["name"](**spec.get("args", {})).to(self.device))
            opt = spec["optimizer"]
            name = "optimizer_" + key
            self.optimizer_names.append(name)
            setattr(self, name, opt["name"](getattr(self, key).parameters(), **opt.get("args", {})))
        self.losses = {type(l).__name__: l for l in losses_list}
        self.metrics = None
        self.ncritic = ncritic
        self.epochs = epochs
        self.start_epoch = 0
        self.sample_size = sample_size
        self.checkpoints = checkpoints
        self.retain_checkpoints = retain_checkpoints
        self.last_retained_checkpoint = 0
        self.recon = recon
        self.nrow = nrow
        # torchgan draws the fixed test noise of every generator-like model in __init__ [tg] (this consumes the device
        # RNG stream before training starts, like upstream)
        gens = [n for n in self.model_names if callable(getattr(getattr(self, n), "sampler", None))]
        self.test_noise = [getattr(self, n).sampler(sample_size, self.device) if test_noise is None else test_noise
                           for n in gens]
        self.batch_size = None
        self.real_inputs = None
        self.labels = None
        self._pending = []
        self.loss_information = {"generator_losses": 0.0, "discriminator_losses": 0.0, "generator_iters": 0,
                                 "discriminator_iters": 0}
        self.loss_logs, self.metric_logs = {}, {}
        self.loss_arg_maps = {n: [a for a in inspect.signature(l.train_ops).parameters if a != "self"]
                              for n, l in self.losses.items()}
        for k, v in kwargs.items():      # torchgan stores unknown kwargs (e.g. devices=[0]) as attributes
            setattr(self, k, v)

    def _call(self, name):
        """Run one loss object's optimiser step.  Loss objects of this package expose ``device_ops`` (train_ops minus
        the ``.item()``): their losses stay on the device and the whole iteration is read back with one
        synchronisation in train_iter; any other loss object goes through the reference's ``train_ops`` -> float."""
        loss = self.losses[name]
        kwargs = {a: getattr(self, a) for a in self.loss_arg_maps[name]}
        fn = getattr(loss, "device_ops", None)
        return fn(**kwargs) if fn is not None else loss.train_ops(**kwargs)

    # ---------------------------------------------------------------------------------------- loss read-back
    # The reference's train_ops return Python floats (three `.item()` synchronisations per iteration, [tg]); here the
    # losses stay on the device until the end of the iteration, are copied into a pinned ring slot asynchronously and
    # folded into loss_logs / loss_information either at once (`train_iter()`, one synchronisation) or one iteration
    # late (`train_iter(defer=True)`, what `train` uses): the host then waits on iteration i-1 while iteration i is
    # already queued, so the device never drains.  Reading `loss_logs` / `loss_information` flushes what is pending.
    _RING = 4

    def _enqueue(self, pending, kinds):
        dev = [(n, v) for n, v in pending.items() if torch.is_tensor(v)]
        entry = {"kinds": kinds, "host": {n: v for n, v in pending.items() if not torch.is_tensor(v)}, "dev": [],
                 "event": None, "slot": None}
        if dev:
            if getattr(self, "_loss_pinned", None) is None or self._loss_pinned.size(1) < len(dev):
                self.flush()
                self._loss_pinned = torch.empty(self._RING, max(8, len(dev)), dtype=torch.float32).pin_memory()
                self._loss_slot = 0
            if len(self._pending) >= self._RING - 1:
                self.flush()
            slot = self._loss_pinned[self._loss_slot]
            self._loss_slot = (self._loss_slot + 1) % self._RING
            for i, (_, v) in enumerate(dev):
                slot[i:i + 1].copy_(v.reshape(1), non_blocking=True)
            entry["event"] = torch.cuda.Event()
            entry["event"].record(torch.cuda.current_stream(dev[0][1].device))
            entry["slot"], entry["dev"] = slot, [n for n, _ in dev]
        self._pending.append(entry)
        return entry

    def _resolve(self, entry):
        """Wait for one iteration's losses and fold them into the logs; returns {loss name: float}."""
        values = dict(entry["host"])
        if entry["event"] is not None:
            entry["event"].synchronize()
            for i, n in enumerate(entry["dev"]):
                values[n] = float(entry["slot"][i])
        lgen = ldis = 0.0
        for name in self.losses:
            v = values.get(name)
            kind = entry["kinds"].get(name)
            if kind == "g":
                lgen += v
            elif kind == "d":
                ldis += v
            self._loss_logs.setdefault(name, []).append(v)
        self._loss_information["generator_losses"] += lgen
        self._loss_information["discriminator_losses"] += ldis
        return values

    def flush(self, keep=0):
        """Resolve pending iterations, oldest first, until at most `keep` remain."""
        values = None
        while len(self._pending) > keep:
            values = self._resolve(self._pending.pop(0))
        return values

    @property
    def loss_logs(self):
        self.flush()
        return self._loss_logs

    @loss_logs.setter
    def loss_logs(self, value):
        self._pending = []
        self._loss_logs = value

    @property
    def loss_information(self):
        self.flush()
        return self._loss_information

    @loss_information.setter
    def loss_information(self, value):
        self._pending = []
        self._loss_information = value

    def train_iter(self, defer=False):
        """One iteration = every loss object's optimiser step in list order (G loss, critic loss, gradient penalty).
        Returns this iteration's {loss name: float}; with defer=True returns None and the values reach the logs when
        the next iteration has been queued (or on the next read of the logs)."""
        info = self._loss_information                      # iteration counters do not depend on the loss values
        pending, kinds = {}, {}
        for name, loss in self.losses.items():
            if isinstance(loss, GeneratorLoss):
                if self.ncritic is None or info["discriminator_iters"] % self.ncritic == 0:
                    pending[name] = self._call(name)
                    kinds[name] = "g"
            elif isinstance(loss, DiscriminatorLoss):
                pending[name] = self._call(name)
                kinds[name] = "d"
        info["generator_iters"] += sum(1 for k in kinds.values() if k == "g")
        info["discriminator_iters"] += 1 if "d" in kinds.values() else 0
        self._enqueue(pending, kinds)
        if defer:
            self.flush(keep=1)
            return None
        return self.flush()

    def save_model(self, epoch, save_items=None):
        if self.last_retained_checkpoint == self.retain_checkpoints:
            self.last_retained_checkpoint = 0
        save_path = self.checkpoints + str(self.last_retained_checkpoint) + ".model"
        self.last_retained_checkpoint += 1
        os.makedirs(os.path.dirname(os.path.abspath(save_path)), exist_ok=True)
        model = {"epoch": epoch + 1, "loss_information": self.loss_information, "loss_objects": self.losses,
                 "metric_objects": self.metrics, "loss_logs": self.loss_logs, "metric_logs": self.metric_logs}
        for name in self.model_names + self.optimizer_names:
            model[name] = getattr(self, name).state_dict()
        torch.save(model, save_path)
        return save_path

    def load_model(self, load_path="", load_items=None):
        if load_path == "":
            load_path = self.checkpoints + str(self.last_retained_checkpoint) + ".model"
        ckpt = compat.load_checkpoint(load_path, map_location=self.device)   # also files the reference wrote
        self.start_epoch = ckpt["epoch"]
        self.loss_information = ckpt.get("loss_information", self.loss_information)
        self.loss_logs = ckpt.get("loss_logs", self.loss_logs)
        self.metric_logs = ckpt.get("metric_logs", self.metric_logs)
        for name in self.model_names + self.optimizer_names:
            if name in ckpt:
                getattr(self, name).load_state_dict(ckpt[name])
        return ckpt

    def train(self, data_loader, **kwargs):
        for epoch in range(self.start_epoch, self.epochs):
            for name in self.model_names:
                getattr(self, name).train()
            for data in data_loader:
                if isinstance(data, (tuple, list)):          # (images, labels) loaders, as torchgan unpacks them [tg]
                    self.real_inputs = data[0].to(self.device)
                    self.labels = data[1].to(self.device)
                    self.batch_size = self.real_inputs.size(0)
                elif torch.is_tensor(data):
                    self.real_inputs = data.to(self.device)
                    self.batch_size = data.size(0)
                else:
                    self.real_inputs = data
                    if isinstance(data, dict) and "image" in data:
                        self.batch_size = data["image"].size(0)
                self.train_iter(defer=True)
            self.flush()
            # data-parallel runs: parameters and optimiser state are bit-identical on every rank (averaged gradients);
            # BatchNorm running statistics are rank-LOCAL (each rank saw its own shard, like DDP without SyncBN), so the
            # checkpoint rank 0 writes carries rank 0's running statistics
            main = self._is_main_process()
            if main:
                self.save_model(epoch)
            for name in self.model_names:
                getattr(self, name).eval()
            if main:
                self.sample_grids(epoch)

    @staticmethod
    def _is_main_process():
        import torch.distributed as dist
        return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0

    def sample_grids(self, epoch):
        """Eval-mode sample grid of every generator-like model (one with a `sampler`) on the trainer's fixed test noise:
        `{recon}/epoch{epoch+1}_{model}.png`, `nrow` tiles per row, min-max normalised.  Returns the written paths."""
        if not self.recon:
            return []
        gens = [n for n in self.model_names if callable(getattr(getattr(self, n), "sampler", None))]
        paths = []
        for n, noise in zip(gens, self.test_noise):
            with torch.no_grad():
                images = getattr(self, n)(*noise) if isinstance(noise, (list, tuple)) else getattr(self, n)(noise)
            paths.append(save_image_grid(images, os.path.join(self.recon, f"epoch{epoch + 1}_{n}.png"), nrow=self.nrow))
        return paths

    def __call__(self, data_loader, **kwargs):
        self.batch_size = getattr(data_loader, "batch_size", None)
        self.train(data_loader, **kwargs)
