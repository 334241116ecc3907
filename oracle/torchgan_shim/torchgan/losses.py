"""torchgan.losses stand-in: only the two base classes the reference subclasses (src/wgan_loss.py:2,47,131,266)."""
import torch.nn as nn


class _LossBase(nn.Module):
    def __init__(self, reduction="mean", override_train_ops=None):
        super().__init__()
        self.reduction = reduction
        self.override_train_ops = override_train_ops
        self.arg_map = {}

    def set_arg_map(self, value):
        self.arg_map.update(value)


class GeneratorLoss(_LossBase):
    pass


class DiscriminatorLoss(_LossBase):
    pass
