"""torchgan.models stand-in: Generator / Discriminator bases, DCGANGenerator, DCGANDiscriminator.

Structure cross-checked against the reference's edited copy of the generator (src/dcgan.py:23-44,52,58-74,82)
and the constructor-argument dicts at src/histopathology_gan.py:178-192.
"""
from math import ceil, log2

import torch
import torch.nn as nn


class Generator(nn.Module):
    def __init__(self, encoding_dims, label_type="none"):
        super().__init__()
        self.encoding_dims = encoding_dims
        self.label_type = label_type

    def _weight_initializer(self):
        for m in self.modules():
            if isinstance(m, nn.ConvTranspose2d):
                nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0.0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1.0)
                nn.init.constant_(m.bias, 0.0)
            elif isinstance(m, nn.Linear):
                nn.init.kaiming_normal_(m.weight)
                nn.init.constant_(m.bias, 0.0)

    def sampler(self, sample_size, device):
        return [torch.randn(sample_size, self.encoding_dims, device=device)]


class Discriminator(nn.Module):
    def __init__(self, input_dims, label_type="none"):
        super().__init__()
        self.input_dims = input_dims
        self.label_type = label_type

    def _weight_initializer(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0.0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1.0)
                nn.init.constant_(m.bias, 0.0)
            elif isinstance(m, nn.Linear):
                nn.init.kaiming_normal_(m.weight)
                nn.init.constant_(m.bias, 0.0)


def _check_size(size, what):
    if size < 16 or ceil(log2(size)) != log2(size):
        raise Exception(f"{what} Image Size must be at least 16*16 and an exact power of 2")


class DCGANGenerator(Generator):
    def __init__(self, encoding_dims=100, out_size=32, out_channels=3, step_channels=64, batchnorm=True,
                 nonlinearity=None, last_nonlinearity=None, label_type="none"):
        super().__init__(encoding_dims, label_type)
        _check_size(out_size, "Target")
        num_repeats = out_size.bit_length() - 4
        self.ch = out_channels
        self.n = step_channels
        use_bias = not batchnorm
        nl = nn.LeakyReLU(0.2) if nonlinearity is None else nonlinearity
        last_nl = nn.Tanh() if last_nonlinearity is None else last_nonlinearity
        d = int(self.n * (2 ** num_repeats))
        blocks = []
        first = [nn.ConvTranspose2d(self.encoding_dims, d, 4, 1, 0, bias=use_bias)]
        if batchnorm:
            first.append(nn.BatchNorm2d(d))
        first.append(nl)
        blocks.append(nn.Sequential(*first))
        for _ in range(num_repeats):
            layer = [nn.ConvTranspose2d(d, d // 2, 4, 2, 1, bias=use_bias)]
            if batchnorm:
                layer.append(nn.BatchNorm2d(d // 2))
            layer.append(nl)
            blocks.append(nn.Sequential(*layer))
            d = d // 2
        blocks.append(nn.Sequential(nn.ConvTranspose2d(d, self.ch, 4, 2, 1, bias=True), last_nl))
        self.model = nn.Sequential(*blocks)
        self._weight_initializer()

    def forward(self, x, feature_matching=False):
        x = x.view(-1, x.size(1), 1, 1)
        return self.model(x)


class DCGANDiscriminator(Discriminator):
    def __init__(self, in_size=32, in_channels=3, step_channels=64, batchnorm=True, nonlinearity=None,
                 last_nonlinearity=None, label_type="none"):
        super().__init__(in_channels, label_type)
        _check_size(in_size, "Input")
        num_repeats = in_size.bit_length() - 4
        self.n = step_channels
        use_bias = not batchnorm
        nl = nn.LeakyReLU(0.2) if nonlinearity is None else nonlinearity
        last_nl = nn.LeakyReLU(0.2) if last_nonlinearity is None else last_nonlinearity
        d = self.n
        blocks = [nn.Sequential(nn.Conv2d(self.input_dims, d, 4, 2, 1, bias=True), nl)]
        for _ in range(num_repeats):
            layer = [nn.Conv2d(d, d * 2, 4, 2, 1, bias=use_bias)]
            if batchnorm:
                layer.append(nn.BatchNorm2d(d * 2))
            layer.append(nl)
            blocks.append(nn.Sequential(*layer))
            d *= 2
        self.disc = nn.Sequential(nn.Conv2d(d, 1, 4, 1, 0, bias=use_bias), last_nl)
        self.model = nn.Sequential(*blocks)
        self._weight_initializer()

    def forward(self, x, feature_matching=False):
        x = self.model(x)
        if feature_matching:
            return x
        x = self.disc(x)
        return x.view(x.size(0))
