"""Checkpoint interop with files written by the reference.

torchgan's `Trainer.save_model` [tg] pickles one dictionary: `epoch`, `loss_information`, `loss_logs`, `metric_logs`,
`metric_objects`, one `state_dict` per model / optimizer -- and `loss_objects`, the loss instances themselves.  The reference
runs its scripts from `src/`, so those instances are pickled under the top-level module names `wgan_loss` (and, through
their `betavae` attribute, `betaVAE`: src/wgan_loss.py:63-69).  `load_checkpoint` is `torch.load` with an unpickler that
resolves those names to this package's classes of the same name, which accept the reference's instance state
(`wgan_loss._VAEConditioned.__setstate__`).  `model_dict_best.pt` (a bare betaVAE state_dict, possibly with
`module.`-prefixed keys) is handled by `wgan_loss.strip_module_prefix`.

Upstream's released checkpoints are not reachable offline; the fixture `tests/golden/upstream_style_ckpt.model` is written
by the reference's OWN classes under their original module names (`oracle/make_upstream_ckpt.py`, torchgan names from the
oracle shim), so the module/class resolution and the instance layout are the real ones.
"""
import importlib
import pickle
from pickle import *  # noqa: F401,F403  (this module doubles as torch.load's `pickle_module`)

import torch

# top-level module the reference pickles under -> module of this package defining the same class names
MODULE_MAP = {"wgan_loss": "rnagan_b200.wgan_loss", "betaVAE": "rnagan_b200.betaVAE", "dcgan": "rnagan_b200.dcgan"}
# torchgan's own classes that can appear in a pickle (instances of the un-conditioned losses, `--loss_type wgan`)
CLASS_MAP = {
    ("torchgan.losses.wasserstein", "WassersteinGeneratorLoss"): ("rnagan_b200.wgan_loss", "WassersteinGeneratorLoss"),
    ("torchgan.losses.wasserstein", "WassersteinDiscriminatorLoss"): ("rnagan_b200.wgan_loss",
                                                                      "WassersteinDiscriminatorLoss"),
    ("torchgan.losses.wasserstein", "WassersteinGradientPenalty"): ("rnagan_b200.wgan_loss", "WassersteinGradientPenalty"),
    ("torchgan.losses.loss", "GeneratorLoss"): ("rnagan_b200.wgan_loss", "GeneratorLoss"),
    ("torchgan.losses.loss", "DiscriminatorLoss"): ("rnagan_b200.wgan_loss", "DiscriminatorLoss"),
}


class Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in CLASS_MAP:
            module, name = CLASS_MAP[(module, name)]
        elif module in MODULE_MAP:
            module = MODULE_MAP[module]
        elif module.startswith("torchgan.losses") and hasattr(importlib.import_module("rnagan_b200.wgan_loss"), name):
            module = "rnagan_b200.wgan_loss"
        return super().find_class(module, name)


def load(file, **kwargs):
    return Unpickler(file, **kwargs).load()


def load_checkpoint(path, map_location=None):
    """torch.load of a `{dir}{k}.model` written by this package OR by the reference's torchgan Trainer."""
    import sys
    return torch.load(path, map_location=map_location, weights_only=False, pickle_module=sys.modules[__name__])
