"""torchgan.trainer stand-in: the slice of Trainer the reference drives (src/histopathology_gan.py:298-314):
model/optimizer construction from the spec dict, train_ops binding by parameter NAME, and train_iter's loss order."""
import inspect

import torch

from .losses import DiscriminatorLoss, GeneratorLoss


class Trainer:
    def __init__(self, models, losses_list, metrics_list=None, device=torch.device("cpu"), ncritic=1, epochs=5,
                 sample_size=8, checkpoints="./model/gan", retain_checkpoints=5, recon="./images", log_dir=None,
                 test_noise=None, nrow=8, **kwargs):
        self.device = device
        self.model_names = []
        self.optimizer_names = []
        for key, spec in models.items():
            self.model_names.append(key)
            setattr(self, key, spec["name"](**spec.get("args", {})).to(device))
            opt = spec["optimizer"]
            name = "optimizer_" + key
            self.optimizer_names.append(name)
            setattr(self, name, opt["name"](getattr(self, key).parameters(), **opt.get("args", {})))
        self.losses = {type(l).__name__: l for l in losses_list}
        self.ncritic = ncritic
        self.epochs = epochs
        self.sample_size = sample_size
        self.batch_size = None
        self.real_inputs = None
        self.labels = None
        self.loss_information = {"generator_losses": 0.0, "discriminator_losses": 0.0,
                                 "generator_iters": 0, "discriminator_iters": 0}
        for k, v in kwargs.items():
            setattr(self, k, v)

    def _call_train_ops(self, loss):
        names = [n for n in inspect.signature(loss.train_ops).parameters if n != "self"]
        return loss.train_ops(**{n: getattr(self, n) for n in names})

    def train_iter(self):
        out = {}
        for name, loss in self.losses.items():
            if isinstance(loss, GeneratorLoss):
                if self.loss_information["discriminator_iters"] % self.ncritic == 0:
                    v = self._call_train_ops(loss)
                    self.loss_information["generator_losses"] += v
                    self.loss_information["generator_iters"] += 1
                    out[name] = v
            elif isinstance(loss, DiscriminatorLoss):
                v = self._call_train_ops(loss)
                self.loss_information["discriminator_losses"] += v
                out[name] = v
        self.loss_information["discriminator_iters"] += 1
        return out
