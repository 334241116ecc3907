"""Drop-in DCGAN modules for the RNA-GAN hot path, running on the sm_100a kernels.

Same constructor keywords, attributes (``encoding_dims``, ``label_type``, ``sampler``, ``model`` / ``disc``) and
``state_dict`` keys / shapes / dtypes as

  * torchgan==0.1.0 ``DCGANGenerator`` / ``DCGANDiscriminator`` -- the classes the reference actually instantiates
    (src/histopathology_gan.py:176-192, src/gan_utils.py:255-271); structure per SURVEY.md Appendix A/D, and
  * the reference's own ``DCGANUpGenerator`` (src/dcgan.py:8-99), the resize-conv variant.

The ``nn`` layers inside ``self.model`` are PARAMETER CONTAINERS (fp32 masters, checkpoint layout); ``forward`` never
calls them -- it runs rnagan_b200.engine on the CUDA kernels and raises when the module is not on a B200.
Training goes through the loss objects' ``train_ops`` (rnagan_b200.wgan_loss), which drive the same engines with a
hand-scheduled backward.  ``forward`` is also differentiable ONCE through ``torch.autograd`` (``_GeneratorFn`` /
``_CriticFn`` below: a custom ``override_train_ops`` or a feature-matching style loss can call ``loss.backward()``);
double backward (``create_graph=True``, i.e. a gradient penalty written with autograd) is not -- that schedule exists
only inside ``WassersteinGradientPenalty*.train_ops``.
"""
from math import ceil, log2

import torch
import torch.nn as nn

from . import engine as _engine
from . import ops as _ops


class Generator(nn.Module):
    """torchgan.models.Generator contract [SURVEY.md Appendix A]."""

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
    """torchgan.models.Discriminator contract [SURVEY.md Appendix A]."""

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


# ------------------------------------------------------------------------------------------------ autograd bridge
# The engines write parameter gradients in place (into the flat buffers that back `p.grad`).  For autograd the gradients
# must instead be RETURNED so that the engine accumulates them (two critic passes in one graph, repeated backward calls):
# every parameter's `.grad` is detached for the duration of the engine's backward, the engine fills a scratch gradient,
# and the previous `.grad` is put back.
_AG_TAGS = 4


def _engine_param_grads(module, run):
    params = list(module.parameters())
    saved = [p.grad for p in params]
    for p in params:
        p.grad = None
    try:
        extra = run()
        grads = [p.grad for p in params]
    finally:
        for p, g in zip(params, saved):
            p.grad = g
    return grads, extra


def _need_train(module, what):
    if not module.training:
        raise NotImplementedError(f"{what}: backward through eval-mode BatchNorm is not implemented on the sm_100a path "
                                  "(call .train() first, or wrap the forward in torch.no_grad())")


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        eng = module._engine()
        lat = _ops.cast_pad_bf16(x, x.size(1))
        n = module.__dict__["_rg_ag"] = (module.__dict__.get("_rg_ag", -1) + 1) % _AG_TAGS
        ctx.module, ctx.tag, ctx.training = module, f"ag{n}", module.training
        img = eng.forward(lat, tag=ctx.tag, training=module.training).clone()
        ctx.save_for_backward(lat, img)
        return img

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_img):
        module = ctx.module
        _need_train(module, "generator")
        if ctx.training is not True:
            raise NotImplementedError("generator: the forward ran in eval mode; its backward is not implemented")
        lat, img = ctx.saved_tensors
        eng = module._engine()
        d_img = d_img.to(torch.float32).contiguous()

        def run():
            eng.backward(lat, d_img, img, tag=ctx.tag)
            return eng.input_grad(lat.shape[0]).clone() if ctx.needs_input_grad[1] else None

        grads, dx = _engine_param_grads(module, run)
        return (None, dx) + tuple(grads)


class _CriticFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        eng = module._engine()
        n = module.__dict__["_rg_ag"] = (module.__dict__.get("_rg_ag", -1) + 1) % _AG_TAGS
        ctx.module, ctx.tag, ctx.B, ctx.training = module, f"ag{n}", x.shape[0], module.training
        return eng.forward(x, tag=ctx.tag, training=module.training).clone()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        module = ctx.module
        _need_train(module, "critic")
        if ctx.training is not True:
            raise NotImplementedError("critic: the forward ran in eval mode; its backward is not implemented")
        eng = module._engine()
        dout = dout.to(torch.float32).contiguous()

        def run():
            dimg = eng.backward(ctx.B, 1.0, tag=ctx.tag, params=True, acc=0.0, want_dimg=ctx.needs_input_grad[1],
                                dout=dout)
            return dimg.clone() if dimg is not None else None

        grads, dx = _engine_param_grads(module, run)
        return (None, dx) + tuple(grads)


def _repeats(size, what):
    if size < 16 or ceil(log2(size)) != log2(size):
        raise Exception(f"{what} Image Size must be at least 16*16 and an exact power of 2")
    return size.bit_length() - 4


class _EngineMixin:
    """Lazy engine construction + re-packing of the bf16 operands when the fp32 masters changed under us."""

    _engine_cls = None

    def __getstate__(self):
        """pickle / deepcopy carry parameters and buffers only; the engine (`_rg_*`) is rebuilt lazily."""
        return {k: v for k, v in self.__dict__.items() if not k.startswith("_rg_")}

    def _engine(self, flush=True):
        """The kernel schedule of this module.  flush: first apply an optimiser step that a data-parallel train step
        deferred (steps.apply_update), so every outside reader sees the reference's post-step weights."""
        eng = self.__dict__.get("_rg_engine")
        if flush and eng is not None and getattr(eng, "pending_update", None) is not None:
            from .steps import apply_update
            apply_update(eng)
        p0 = next(self.parameters())
        if p0.device.type != "cuda":
            raise RuntimeError(f"{type(self).__name__} runs only on a CUDA (sm_100a) device: there is no CPU fallback; "
                               "move the module with .to('cuda') first")
        if eng is None or eng.device != p0.device:
            eng = self._engine_cls(self)
            self.__dict__["_rg_engine"] = eng
            self.__dict__["_rg_versions"] = self._versions()
        elif self._versions() != self.__dict__["_rg_versions"]:
            eng.pack()
            self.__dict__["_rg_versions"] = self._versions()
        return eng

    def _versions(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def state_dict(self, *args, **kwargs):
        if self.__dict__.get("_rg_engine") is not None and next(self.parameters()).device.type == "cuda":
            self._engine()                  # applies a deferred (data-parallel) optimiser step first
        return super().state_dict(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        if self.__dict__.get("_rg_engine") is not None and next(self.parameters()).device.type == "cuda":
            self._engine()                  # a deferred step must not land on top of the loaded weights
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):   # .to()/.cuda()/.float(): storage moves invalidate the engine
        before = [p.data_ptr() for p in self.parameters()]
        out = super()._apply(fn, *args, **kwargs)
        if before != [p.data_ptr() for p in self.parameters()]:
            self.__dict__.pop("_rg_engine", None)
        return out


class DCGANGenerator(_EngineMixin, Generator):
    r"""Transposed-convolution DCGAN generator (torchgan DCGANGenerator; the ConvTranspose2d lines the reference keeps
    as comments at src/dcgan.py:52,82)."""

    _engine_cls = _engine.GeneratorEngine

    def __init__(self, encoding_dims=100, out_size=32, out_channels=3, step_channels=64, batchnorm=True,
                 nonlinearity=None, last_nonlinearity=None, label_type="none"):
        super().__init__(encoding_dims, label_type)
        reps = _repeats(out_size, "Target")
        self.ch = out_channels
        self.n = step_channels
        act = nn.LeakyReLU(0.2) if nonlinearity is None else nonlinearity
        last = nn.Tanh() if last_nonlinearity is None else last_nonlinearity
        d = int(self.n * (2 ** reps))
        layers = []
        cin, stride, pad = self.encoding_dims, 1, 0
        for _ in range(reps + 1):
            blk = [nn.ConvTranspose2d(cin, d, 4, stride, pad, bias=not batchnorm)]
            if batchnorm:
                blk.append(nn.BatchNorm2d(d))
            blk.append(act)
            layers.append(nn.Sequential(*blk))
            cin, d, stride, pad = d, d // 2, 2, 1
        layers.append(nn.Sequential(nn.ConvTranspose2d(cin, self.ch, 4, 2, 1, bias=True), last))
        self.model = nn.Sequential(*layers)
        self._weight_initializer()

    def forward(self, x, feature_matching=False):
        """x: [B, encoding_dims] -> [B, out_channels, S, S] fp32 NCHW (batch statistics in train mode)."""
        eng = self._engine()
        x = x.view(-1, x.size(1)).to(device=eng.device, dtype=torch.float32).contiguous()
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return _GeneratorFn.apply(self, x, *self.parameters())
        with torch.no_grad():
            lat = _ops.cast_pad_bf16(x, x.size(1))
            return eng.forward(lat, tag="module", training=self.training, keep=False).clone()


class DCGANDiscriminator(_EngineMixin, Discriminator):
    r"""DCGAN critic (torchgan DCGANDiscriminator; ctor args at src/histopathology_gan.py:186-192)."""

    _engine_cls = _engine.CriticEngine

    def __init__(self, in_size=32, in_channels=3, step_channels=64, batchnorm=True, nonlinearity=None,
                 last_nonlinearity=None, label_type="none"):
        super().__init__(in_channels, label_type)
        reps = _repeats(in_size, "Input")
        self.n = step_channels
        act = nn.LeakyReLU(0.2) if nonlinearity is None else nonlinearity
        last = nn.LeakyReLU(0.2) if last_nonlinearity is None else last_nonlinearity
        d = self.n
        layers = [nn.Sequential(nn.Conv2d(self.input_dims, d, 4, 2, 1, bias=True), act)]
        for _ in range(reps):
            blk = [nn.Conv2d(d, d * 2, 4, 2, 1, bias=not batchnorm)]
            if batchnorm:
                blk.append(nn.BatchNorm2d(d * 2))
            blk.append(act)
            layers.append(nn.Sequential(*blk))
            d *= 2
        self.disc = nn.Sequential(nn.Conv2d(d, 1, 4, 1, 0, bias=not batchnorm), last)
        self.model = nn.Sequential(*layers)
        self._weight_initializer()

    def forward(self, x, feature_matching=False):
        """x: [B, C, S, S] fp32 NCHW -> critic score [B] (or the last feature map when feature_matching; that output is
        not differentiable)."""
        eng = self._engine()
        x = x.to(device=eng.device, dtype=torch.float32).contiguous()
        if (not feature_matching and torch.is_grad_enabled()
                and (x.requires_grad or any(p.requires_grad for p in self.parameters()))):
            return _CriticFn.apply(self, x, *self.parameters())
        with torch.no_grad():
            out = eng.forward(x, tag="module", training=self.training)
            if feature_matching:
                B = x.shape[0]
                feat = eng.bufs.get(f"module.h{eng.n}", (B, 4, 4, eng.Cn))
                return feat.float().permute(0, 3, 1, 2).contiguous()
            return out.clone()


class DCGANUpGenerator(_EngineMixin, Generator):
    """Resize-convolution generator of src/dcgan.py:8-99 (bilinear x2 -> ReflectionPad2d(1) -> Conv2d 3x3; the
    ``last_nonlinearity`` is built but NOT applied, src/dcgan.py:32,82).  Same constructor and state_dict layout as the
    reference class; ``forward`` runs rnagan_b200.engine.UpGeneratorEngine.  The reference imports this class
    (src/histopathology_gan.py:24) but its drivers instantiate the transposed-conv DCGANGenerator."""

    _engine_cls = _engine.UpGeneratorEngine

    def __init__(self, encoding_dims=100, out_size=32, out_channels=3, step_channels=64, batchnorm=True,
                 nonlinearity=None, last_nonlinearity=None, label_type="none"):
        super().__init__(encoding_dims, label_type)
        reps = _repeats(out_size, "Target")
        self.ch = out_channels
        self.n = step_channels
        act = nn.LeakyReLU(0.2) if nonlinearity is None else nonlinearity
        d = int(self.n * (2 ** reps))
        first = [nn.ConvTranspose2d(self.encoding_dims, d, 4, 1, 0, bias=not batchnorm)]
        if batchnorm:
            first.append(nn.BatchNorm2d(d))
        layers = [nn.Sequential(*first, act)]
        for _ in range(reps):
            if batchnorm:
                layers.append(nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear"), nn.ReflectionPad2d(1),
                                            nn.Conv2d(d, d // 2, kernel_size=3, stride=1, padding=0),
                                            nn.BatchNorm2d(d // 2), act))
            else:
                layers.append(nn.Sequential(nn.ConvTranspose2d(d, d // 2, 4, 2, 1, bias=True), act))
            d //= 2
        layers.append(nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear"), nn.ReflectionPad2d(1),
                                    nn.Conv2d(d, self.ch, kernel_size=3, stride=1, padding=0)))
        self.model = nn.Sequential(*layers)
        self._weight_initializer()

    @torch.no_grad()
    def forward(self, x, feature_matching=False):
        eng = self._engine()
        x = x.view(-1, x.size(1)).to(device=eng.device, dtype=torch.float32).contiguous()
        lat = _ops.cast_pad_bf16(x, x.size(1))
        return eng.forward(lat, tag="module", training=self.training).clone()
