"""Minimal stand-in for the slice of ``torchgan.trainer.Trainer`` that the reference drives
(src/histopathology_gan.py:298-314, src/gan_utils.py:286-297; contract in SURVEY.md Appendix A):

  * builds each model / optimizer from the spec dict and exposes them as ``trainer.generator``,
    ``trainer.optimizer_generator`` ...;
  * binds every loss object's ``train_ops`` parameters BY NAME to same-named trainer attributes;
  * ``train_iter`` runs the losses in list order (G loss, critic loss, gradient penalty; ncritic=1);
  * ``save_model`` / ``load_model`` use the torchgan checkpoint dictionary layout
    (epoch, loss_information, loss_objects, metric_objects, loss_logs, metric_logs + one state_dict per model/optimizer).

Logging / visualisation (torchgan Logger, tensorboard, image grids) is out of scope.
"""
import inspect
import os

import torch

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
        self.batch_size = None
        self.real_inputs = None
        self.labels = None
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

    def _read_back(self, pending):
        """{name: device tensor [1] | float} -> {name: float} with a single device->host copy + synchronisation."""
        dev = [(n, v) for n, v in pending.items() if torch.is_tensor(v)]
        out = {n: v for n, v in pending.items() if not torch.is_tensor(v)}
        if dev:
            if getattr(self, "_loss_pinned", None) is None or self._loss_pinned.numel() < len(dev):
                self._loss_pinned = torch.empty(max(8, len(dev)), dtype=torch.float32).pin_memory()
            for i, (_, v) in enumerate(dev):
                self._loss_pinned[i:i + 1].copy_(v.reshape(1), non_blocking=True)
            torch.cuda.current_stream(dev[0][1].device).synchronize()
            for i, (n, _) in enumerate(dev):
                out[n] = float(self._loss_pinned[i])
        return out

    def train_iter(self):
        lgen = ldis = 0.0
        gen_iter = dis_iter = 0
        pending, kinds = {}, {}
        for name, loss in self.losses.items():
            if isinstance(loss, GeneratorLoss):
                if self.loss_information["discriminator_iters"] % self.ncritic == 0:
                    pending[name] = self._call(name)
                    kinds[name] = "g"
            elif isinstance(loss, DiscriminatorLoss):
                pending[name] = self._call(name)
                kinds[name] = "d"
        values = self._read_back(pending)
        for name in self.losses:
            v = values.get(name)
            if kinds.get(name) == "g":
                lgen += v
                gen_iter += 1
            elif kinds.get(name) == "d":
                ldis += v
                dis_iter += 1
            self.loss_logs.setdefault(name, []).append(v)
        self.loss_information["generator_losses"] += lgen
        self.loss_information["discriminator_losses"] += ldis
        self.loss_information["generator_iters"] += gen_iter
        self.loss_information["discriminator_iters"] += 1 if dis_iter else 0
        return values

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
        ckpt = torch.load(load_path, map_location=self.device, weights_only=False)
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
                self.real_inputs = data
                if isinstance(data, dict) and "image" in data:
                    self.batch_size = data["image"].size(0)
                self.train_iter()
            self.save_model(epoch)

    def __call__(self, data_loader, **kwargs):
        self.batch_size = getattr(data_loader, "batch_size", None)
        self.train(data_loader, **kwargs)
