"""Hand-scheduled forward / backward / double-backward of the DCGAN generator and critic on the C-ABI kernels.

This replaces what the reference gets from ATen + autograd inside the three ``train_ops``
(src/wgan_loss.py:82-129, 181-263, 314-389) with explicit sequences of kernels:

  generator  G.0  ConvTranspose2d(E,C0,4,1,0)      -> rg_gemm_nt (projection)            + BN + LeakyReLU
             G.l  ConvTranspose2d(d,d/2,4,2,1)     -> rg_conv_up                         + BN + LeakyReLU
             G.n  ConvTranspose2d(64,3,4,2,1)+Tanh -> rg_conv_up_img (bias+tanh epilogue)
  critic     D.0  Conv2d(3,64,4,2,1)+LeakyReLU     -> rg_im2col_img + rg_gemm_nt (bias+LeakyReLU epilogue)
             D.l  Conv2d(d,2d,4,2,1)               -> rg_conv_down                       + BN + LeakyReLU
             disc Conv2d(C,1,4,1,0)+LeakyReLU      -> rg_head_fwd

Backward uses the same engine with the roles swapped (dgrad of a conv is the transposed form and vice versa,
wgrad is rg_conv_wgrad); the gradient penalty's double backward follows SURVEY.md Appendix C and is checked
against autograd in tests/test_gp_math_cpu.py.  Activations are bf16 NHWC, parameters/gradients/statistics fp32.
"""
import os

import torch
import torch.nn as nn

from . import ops
from .parallel import GradSync

BF16 = torch.bfloat16
F32 = torch.float32
SLOPE = 0.2
# BatchNorm batch statistics accumulated in the epilogue of the producing convolution (RG_FUSED_STATS=0: separate pass)
FUSED_STATS = os.environ.get("RG_FUSED_STATS", "1") != "0"
# the Cs == 64 transposed convolutions contract all four output phases in one tile (RG_MERGED_UP=0: one phase per tile)
MERGED_UP = os.environ.get("RG_MERGED_UP", "1") != "0"
# the LeakyReLU / BatchNorm backward masks and the BatchNorm backward sums are computed in the epilogue of the
# input-gradient contraction that produces dh (RG_FUSED_BWD=0: separate rg_bn_bwd_reduce / rg_lrelu_bwd passes)
# Measured on B200 (profiles/r1_step_profile_fused_bwd.txt): OFF by default.  The fusion removes 22 of 30
# rg_bn_bwd_reduce and 3 of 5 rg_lrelu_bwd launches (-0.87 ms/step) but slows the 25 contractions that carry it by
# 1.07 ms/step (106 vs 67 us per launch) even with the aux tile TMA-staged two slabs ahead: four epilogue warps cannot
# absorb the mask + normalise + second reduction arithmetic of the narrow, short-K layers under their mainloop.
FUSED_BWD = os.environ.get("RG_FUSED_BWD", "0") != "0"
# the layer-0 LeakyReLU mask alone (rg_epilogue_aux mode 1: no statistics, no per-channel vectors) in the epilogue of the
# contraction that produces dh0: ON -- measured -0.18 ms/step (3 of 5 rg_lrelu_bwd passes over 3 x 134 MB disappear, the
# three merged-phase launches that carry the mask cost 0.04 ms more)
FUSED_LRELU = os.environ.get("RG_FUSED_LRELU", "1") != "0"
# the 3-channel image-side layers (critic layer 0, generator output layer) run as fused halo-tile kernels (csrc/rg_img.cu:
# no materialised im2col / col2im buffers) whenever the wide side has 64 channels (step_channels = 64, every reference
# config); RG_FUSED_IMG=0 keeps the im2col + GEMM + col2im path that other widths use
FUSED_IMG = os.environ.get("RG_FUSED_IMG", "1") != "0"


def _grad_of(p):
    if p.grad is None or p.grad.dtype != F32 or p.grad.stride() != p.stride():
        p.grad = torch.zeros_like(p, memory_format=torch.preserve_format)
    return p.grad


def _to_native(params):
    """Store 4x4 conv weights [Cp, Cs, 4, 4] channels_last, i.e. physically [Cp][kh][kw][Cs]: the element order of the
    packed w_down operand and of the wgrad accumulator rows.  Shapes, values, state_dict keys and every torch op on
    the parameter are unchanged; Adam then re-emits the bf16 GEMM operand while it updates the fp32 master (no pack
    kernel) and the weight gradient is stored in whole 128-byte lines without a layout-changing reduce pass.
    Returns the parameters whose storage was replaced."""
    changed = []
    for p in params:
        if not ops.is_native4(p):
            with torch.no_grad():
                p.data = p.data.contiguous(memory_format=torch.channels_last)
            changed.append(p)
    return changed


class _Bufs:
    """Named, shape-keyed device buffers reused across steps (static addresses: CUDA-graph friendly)."""

    def __init__(self, device):
        self.device = device
        self._d = {}
        self._alias = {}

    def pair(self, name_a, name_b):
        """Make `name_a` / `name_b` the two halves of ONE allocation [2, *shape]: two passes (real / fake, or the two
        operand sets of the gradient penalty) then feed a single weight-gradient or input-gradient launch over the
        concatenated batch instead of two launches with an accumulating epilogue."""
        group = ("pair", name_a, name_b)
        self._alias.setdefault(name_a, (group, 0))
        self._alias.setdefault(name_b, (group, 1))

    def joint(self, name_a, shape, dtype=BF16):
        """The whole [2*B, ...] tensor whose first half is `name_a` (which must have been paired)."""
        group, _ = self._alias[name_a]
        t = self._pair_tensor(group, shape, dtype)
        return t.view(2 * shape[0], *shape[1:])

    def _pair_tensor(self, group, shape, dtype):
        key = (group, tuple(shape), dtype)
        t = self._d.get(key)
        if t is None:
            t = torch.zeros((2,) + tuple(shape), dtype=dtype, device=self.device)
            self._d[key] = t
        return t

    def get(self, name, shape, dtype=BF16, zero=False):
        al = self._alias.get(name)
        if al is not None:
            return self._pair_tensor(al[0], shape, dtype)[al[1]]
        key = (name, tuple(shape), dtype)
        t = self._d.get(key)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(tuple(shape), dtype=dtype, device=self.device)
            self._d[key] = t
        return t


class _BNState:
    """Per-pass statistics of one BatchNorm layer (a critic step holds two passes -- real and fake -- at once)."""

    def __init__(self, C, device):
        f = lambda *s: torch.zeros(*s, dtype=F32, device=device)
        self.sums, self.bsums, self.q = f(2, C), f(2, C), f(3, C)
        self.mean, self.rstd, self.scale, self.shift = f(C), f(C), f(C), f(C)


class _BN:
    """One BatchNorm2d layer: parameter references + per-channel work vectors keyed by pass tag."""

    def __init__(self, mod, device, slope=SLOPE):
        self.mod = mod
        self.C = mod.num_features
        self.device = device
        self.slope = slope
        self._states = {}

    def st(self, tag):
        s = self._states.get(tag)
        if s is None:
            s = _BNState(self.C, self.device)
            self._states[tag] = s
        return s

    def stats_ws(self, training=True):
        """Scratch for the statistics the producing convolution accumulates in its epilogue (None: separate pass)."""
        if not training or self.C % 64 != 0 or not FUSED_STATS:
            return None
        return ops.stats_ws(self.C, self.device)

    def forward(self, a, h, M, training=True, tag="", stats=None, apply=True):
        """apply=False: only the statistics / scale / shift (and the running-statistics update); the caller's next kernel
        applies lrelu(scale * a + shift) itself (GeneratorEngine's last block, img_conv_up bn=...)."""
        m, s = self.mod, self.st(tag)
        if training and stats is not None:
            ops.bn_finalize_partials(stats, m.weight, m.bias, M, self.C, m.eps, m.momentum, m.running_mean,
                                     m.running_var, m.num_batches_tracked, s.sums, s.mean, s.rstd, s.scale, s.shift)
        elif training:
            ops.bn_stats(a, M, self.C, s.sums)
            ops.bn_finalize(s.sums, m.weight, m.bias, M, self.C, m.eps, m.momentum, m.running_mean, m.running_var,
                            m.num_batches_tracked, s.mean, s.rstd, s.scale, s.shift)
        else:   # eval: running statistics (torchgan Trainer samples in eval mode at epoch end)
            with torch.no_grad():
                s.mean.copy_(m.running_mean)
                s.rstd.copy_((m.running_var + m.eps).rsqrt())
                s.scale.copy_(m.weight * s.rstd)
                s.shift.copy_(m.bias - s.mean * s.scale)
        if apply:
            ops.bn_act(a, s.scale, s.shift, self.slope, h, M, self.C)

    def bwd_ws(self, slot=1):
        """Partial-sum scratch for the fused backward (None: unfused)."""
        if not FUSED_BWD or self.C % 64 != 0:
            return None
        return ops.stats_ws(self.C, self.device, slot=slot)

    def aux(self, a, tag=""):
        """rg_epilogue_aux mode 2 descriptor: the contraction producing dh for THIS layer stores du and sums it."""
        s = self.st(tag)
        return ("bn", a, s.mean, s.rstd, s.scale, s.shift, self.slope)

    def backward(self, dh, a, da, M, param_grads=False, acc_gamma=0.0, acc_beta=0.0, add=None, du_out=None, tag="",
                 fused=None):
        """da = d(loss)/d(a) from dh = d(loss)/d(h); optionally (accumulating) gamma/beta gradients.
        fused: per-CTA partial sums left by the contraction that produced `dh`, which then already holds
        du = dh * lrelu'(u) (see aux): only the small fixed-order reduce of the partials remains."""
        s = self.st(tag)
        slope = self.slope
        if fused is not None:
            ops.reduce_partials(fused, s.bsums)
            slope = 1.0                      # the mask is applied already: rg_bn_bwd_apply takes dh as du
        else:
            ops.bn_bwd_reduce(dh, a, s.mean, s.rstd, s.scale, s.shift, self.slope, M, self.C, s.bsums)
        if param_grads:
            ops.bn_param_grads(s.bsums, _grad_of(self.mod.weight), _grad_of(self.mod.bias), self.C, acc_gamma,
                               acc_beta)
        ops.bn_bwd_apply(dh, a, add, s.mean, s.rstd, s.scale, s.shift, slope, s.bsums, M, self.C, da, du_out)


def _up_operand(eng, l, npix):
    """B operand of rg_conv_up for link l: the merged-phase w_up9 (Cs == 64 and at least two 128-pixel M tiles), the
    K-major w_up copy of the narrow layers, else w_down itself read MN-major."""
    if eng.w_up9[l - 1] is not None and npix >= 256 and MERGED_UP:
        return eng.w_up9[l - 1]
    return eng.w_upk[l - 1] if eng.w_upk[l - 1] is not None else eng.w_down[l - 1]


def _check_act(mod, slope, what):
    if not isinstance(mod, nn.LeakyReLU) or abs(mod.negative_slope - slope) > 1e-12:
        raise NotImplementedError(f"{what}: only LeakyReLU({slope}) is implemented on the sm_100a path (got {mod!r})")


# ====================================================================================================== generator
class GeneratorEngine:
    """Kernel schedule for torchgan-style DCGANGenerator (batchnorm=True, LeakyReLU(0.2), Tanh)."""

    def __init__(self, module):
        blocks = list(module.model)
        self.module = module
        dev = next(module.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GeneratorEngine needs the module on a CUDA device (there is no CPU path)")
        self.device = dev
        self.bufs = _Bufs(dev)
        first = blocks[0]
        if len(first) != 3 or not isinstance(first[1], nn.BatchNorm2d):
            raise NotImplementedError("generator without BatchNorm is not implemented on the sm_100a path")
        self.conv0, self.bn0 = first[0], _BN(first[1], dev)
        _check_act(first[2], SLOPE, "generator")
        self.E, self.C0 = self.conv0.weight.shape[0], self.conv0.weight.shape[1]
        self.convs, self.bns = [], []
        for blk in blocks[1:-1]:
            if len(blk) != 3 or not isinstance(blk[0], nn.ConvTranspose2d) or not isinstance(blk[1], nn.BatchNorm2d):
                raise NotImplementedError("unexpected generator block layout")
            _check_act(blk[2], SLOPE, "generator")
            self.convs.append(blk[0])
            self.bns.append(_BN(blk[1], dev))
        last = blocks[-1]
        self.conv_last = last[0]
        if not isinstance(last[1], nn.Tanh):
            raise NotImplementedError("generator last_nonlinearity must be Tanh on the sm_100a path")
        self.Cimg = self.conv_last.weight.shape[1]
        self.Cn = self.conv_last.weight.shape[0]
        if self.E % 64 or self.C0 % 64 or self.Cn % 64:
            raise NotImplementedError("channel counts must be multiples of 64 on the sm_100a path")
        self.n = len(self.convs)
        self.size = 4 * (2 ** (self.n + 1))
        self._native_params = [self.conv0.weight] + [c.weight for c in self.convs]
        _to_native(self._native_params)
        # packed bf16 operands (derived, non-persistent).  G.0: [E][16*C0] = the bf16 image of the native weight, read
        # as an MN-major B operand by rg_gemm_nn (column n = tap*C0 + co)
        self.w_projkn = torch.empty(self.E, 16 * self.C0, dtype=BF16, device=dev)
        self.conv0.weight._rg_shadow = self.w_projkn
        # one packed copy per link (w_down: K-major B for rg_conv_down, MN-major B for rg_conv_up); layers with
        # Cs <= 128 also keep the tiny K-major w_up copy, which is faster for narrow N tiles
        self.w_down, self.w_upk, self.w_up9 = [], [], []
        for c in self.convs:
            Cp, Cs = c.weight.shape[0], c.weight.shape[1]
            self.w_down.append(torch.empty(Cp, 16 * Cs, dtype=BF16, device=dev))
            c.weight._rg_shadow = self.w_down[-1]        # re-emitted by the fused Adam step
            self.w_upk.append(torch.zeros(4, ops.up_pad(Cs), 4 * Cp, dtype=BF16, device=dev) if Cs <= 128 else None)
            self.w_up9.append(torch.zeros(9, Cp // 64, 2, 4, 32, 64, dtype=BF16, device=dev) if Cs == 64 else None)
        self.w_colT_last = torch.zeros(16 * self.Cimg, self.Cn, dtype=BF16, device=dev)
        self.w_col_last = torch.empty(self.Cn, 64, dtype=BF16, device=dev)
        self.fused_img = FUSED_IMG and self.Cn == 64 and self.Cimg <= 4 and self.size >= 16
        self.sync = GradSync(module)
        self.pack()

    def pack(self, full=True):
        """Refresh the bf16 operand copies.  full=True (load_state_dict, external weight edits): everything from the
        fp32 masters.  full=False (right after the fused Adam step, which already re-emitted w_down / w_projkn): only
        the derived copies Adam does not write."""
        if full:
            for p in _to_native(self._native_params):
                self.sync.rebind(p)
            ops.cast_pad_bf16(ops.phys2d(self.conv0.weight.detach()), out=self.w_projkn)
            for c, wd in zip(self.convs, self.w_down):
                ops.cast_pad_bf16(ops.phys2d(c.weight.detach()), out=wd)
        for c, wd, wu, w9 in zip(self.convs, self.w_down, self.w_upk, self.w_up9):
            if wu is not None:
                ops.pack_up_from_down(wd, wu, c.weight.shape[1])
            if w9 is not None:
                ops.pack_up9_from_down(wd, c.weight.shape[1], out=w9)
        if self.fused_img:                   # the other fused image-side kernels read the fp32 parameter itself
            self.w_up_img = ops.img_conv_up_pack(self.conv_last.weight.detach(), getattr(self, "w_up_img", None))
        else:
            ops.pack_edge_t(self.conv_last.weight.detach(), self.w_colT_last)
            ops.pack_edge(self.conv_last.weight.detach(), self.w_col_last)

    def forward(self, lat, tag="g", training=True, out=None, unit_nhwc=False, u8=False, bgr=False, keep=True):
        """lat: bf16 [B, E] -> fp32 NCHW image [B, Cimg, S, S]; keeps activations under `tag` for backward.
        keep=False (passes that are not differentiated: synthesis, the generator pass of the critic / penalty steps): the
        last hidden activation h_n is not written -- the output kernel applies its BatchNorm + LeakyReLU on the fly.
        unit_nhwc (synthesis): `out` [B, S, S, Cimg] receives (image + 1) / 2 in NHWC straight from the last kernel;
        u8: `out` is a uint8 [B, S, S, Cimg] tensor receiving trunc(255 * (image + 1) / 2) (bgr: reversed channels)."""
        B = lat.shape[0]
        g = self.bufs.get
        a = g(f"{tag}.a0", (B, 4, 4, self.C0))
        h = g(f"{tag}.h0", (B, 4, 4, self.C0))
        ops.gemm_nn(lat, self.w_projkn, out=a.view(B, 16 * self.C0))
        self.bn0.forward(a, h, B * 16, training, tag=tag)
        H = 4
        fuse_last = False
        for l, (c, bn) in enumerate(zip(self.convs, self.bns), start=1):
            Cs = c.weight.shape[1]
            a = g(f"{tag}.a{l}", (B, 2 * H, 2 * H, Cs))
            sws = bn.stats_ws(training)
            ops.conv_up(h, _up_operand(self, l, B * H * H), Cs, out=a, stats=sws)
            H *= 2
            h = g(f"{tag}.h{l}", (B, H, H, Cs))
            fuse_last = (not keep) and l == self.n and self.fused_img
            bn.forward(a, h, B * H * H, training, tag=tag, stats=sws, apply=not fuse_last)
        if out is None:
            out = g(f"{tag}.img", (B, 2 * H, 2 * H, self.Cimg) if (unit_nhwc or u8) else (B, self.Cimg, 2 * H, 2 * H),
                    torch.uint8 if u8 else F32)
        if self.fused_img:
            st = self.bns[self.n - 1].st(tag) if fuse_last else None
            ops.img_conv_up(a if fuse_last else h, self.w_up_img, out, bias=self.conv_last.bias.detach(), act_tanh=True,
                            unit_nhwc=unit_nhwc, u8=u8, bgr=bgr, Cimg=self.Cimg,
                            bn=(st.scale, st.shift, self.bns[self.n - 1].slope) if fuse_last else None)
            return out
        col = g("fwd.colimg", (B * H * H, 16 * self.Cimg), F32)
        ops.conv_up_img_col(h, self.w_colT_last, self.Cimg, col, out, bias=self.conv_last.bias.detach(), act_tanh=True,
                            unit_nhwc=unit_nhwc, u8=u8, bgr=bgr)
        return out

    def backward(self, lat, d_img, img, tag="g"):
        """Parameter gradients of the generator from d_img = dL/d(image) (fp32 NCHW); writes .grad (overwrites)."""
        B = lat.shape[0]
        g = self.bufs.get
        n = self.n
        H = self.size // 2
        npix = B * H * H
        ops.img_channel_sum(d_img, _grad_of(self.conv_last.bias), y=img, mode=2, acc=0.0)
        hn = g(f"{tag}.h{n}", (B, H, H, self.Cn))
        dh = g(f"bwd.dh{n}", (B, H, H, self.Cn))
        # every contraction that produces dh for a BatchNorm'd layer stores du = dh * lrelu'(u) and sums it (fused)
        fused = self.bns[n - 1].bwd_ws()
        if self.fused_img and fused is None:
            # d(pre-tanh) = d_img * (1 - img^2) is formed while the image patch is staged (mode 2)
            ops.img_conv_wgrad(hn, d_img, _grad_of(self.conv_last.weight), y=img, mode=2)
            self.sync.layer_done(self.conv_last.weight, self.conv_last.bias)
            ops.img_conv_down(d_img, self.conv_last.weight.detach(), dh, y=img, mode=2)
        else:
            if self.fused_img:               # opt-in RG_FUSED_BWD needs the packed operands of the GEMM path
                ops.pack_edge(self.conv_last.weight.detach(), self.w_col_last)
            col = g("bwd.col", (npix, 64))
            ops.im2col_img(d_img, col, y=img, mode=2)                       # d(pre-tanh) in im2col form
            dcol = g("bwd.dcol", (self.Cn, 64), F32)
            ops.gemm_tn(hn.view(npix, self.Cn), col, out=dcol)
            ops.unpack_edge_grad(dcol, _grad_of(self.conv_last.weight), acc=0.0)
            self.sync.layer_done(self.conv_last.weight, self.conv_last.bias)
            if fused is not None:
                ops.gemm_nt_bwd(col, self.w_col_last, dh.view(npix, self.Cn), stats=fused,
                                aux=self.bns[n - 1].aux(g(f"{tag}.a{n}", (B, H, H, self.Cn)), tag))
            else:
                ops.gemm_nt(col, self.w_col_last, out=dh.view(npix, self.Cn))
        for l in range(n, 0, -1):
            c, bn = self.convs[l - 1], self.bns[l - 1]
            Cp, Cs = c.weight.shape[0], c.weight.shape[1]
            a = g(f"{tag}.a{l}", (B, H, H, Cs))
            da = g(f"bwd.da{l}", (B, H, H, Cs))
            bn.backward(dh, a, da, B * H * H, param_grads=True, tag=tag, fused=fused)
            H //= 2
            hprev = g(f"{tag}.h{l - 1}", (B, H, H, Cp))
            ops.conv_wgrad(hprev, da, _grad_of(c.weight))
            self.sync.layer_done(c.weight, bn.mod.weight, bn.mod.bias)
            dh = g(f"bwd.dh{l - 1}", (B, H, H, Cp))
            bnp = self.bns[l - 2] if l >= 2 else self.bn0
            fused = bnp.bwd_ws()
            aux = bnp.aux(g(f"{tag}.a{l - 1}", (B, H, H, Cp)), tag) if fused is not None else None
            ops.conv_down(da, self.w_down[l - 1], out=dh, stats=fused, aux=aux)
        a0 = g(f"{tag}.a0", (B, 4, 4, self.C0))
        da0 = g("bwd.da0", (B, 4, 4, self.C0))
        self.bn0.backward(dh, a0, da0, B * 16, param_grads=True, tag=tag, fused=fused)
        ops.proj_wgrad(lat, da0, _grad_of(self.conv0.weight))
        self.sync.layer_done(self.conv0.weight, self.bn0.mod.weight, self.bn0.mod.bias)

    def input_grad(self, B):
        """dL/d(latent) fp32 [B, E] of the backward pass that just ran (autograd through the module only: the train_ops
        never need it -- the reference's encoder backward is discarded work)."""
        da0 = self.bufs.get("bwd.da0", (B, 4, 4, self.C0))
        return ops.gemm_nt(da0.view(B, 16 * self.C0), self.w_projkn, out_f32=True)


# ====================================================================================================== resize-conv G
class UpGeneratorEngine:
    """Kernel schedule for DCGANUpGenerator (src/dcgan.py:8-99): G.0 projection + BN + LeakyReLU, then per block
    bilinear x2 + reflect-pad (rg_upsample2x_reflectpad) -> 3x3 conv as a 9-tap implicit GEMM (rg_conv3x3, bias in the
    epilogue) -> BN + LeakyReLU; last block ends in a bias-only Conv2d(64, 3, 3) with NO Tanh (src/dcgan.py:76-84)."""

    def __init__(self, module):
        blocks = list(module.model)
        self.module = module
        dev = next(module.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("UpGeneratorEngine needs the module on a CUDA device (there is no CPU path)")
        self.device = dev
        self.bufs = _Bufs(dev)
        first = blocks[0]
        if len(first) != 3 or not isinstance(first[1], nn.BatchNorm2d):
            raise NotImplementedError("resize-conv generator without BatchNorm is not implemented on the sm_100a path")
        self.conv0, self.bn0 = first[0], _BN(first[1], dev)
        _check_act(first[2], SLOPE, "generator")
        self.E, self.C0 = self.conv0.weight.shape[0], self.conv0.weight.shape[1]
        self.convs, self.bns = [], []
        for blk in blocks[1:-1]:
            if len(blk) != 5 or not isinstance(blk[2], nn.Conv2d) or not isinstance(blk[3], nn.BatchNorm2d):
                raise NotImplementedError("unexpected resize-conv block layout")
            _check_act(blk[4], SLOPE, "generator")
            self.convs.append(blk[2])
            self.bns.append(_BN(blk[3], dev))
        self.conv_last = blocks[-1][2]
        self.Cimg, self.Cn = self.conv_last.weight.shape[0], self.conv_last.weight.shape[1]
        if self.Cn != 64 or self.Cimg > 3:
            raise NotImplementedError("resize-conv generator: last layer must be Conv2d(64, <=3, 3) (step_channels=64)")
        self.n = len(self.convs)
        self.size = 4 * (2 ** (self.n + 1))
        self.w_proj = torch.empty(16 * self.C0, self.E, dtype=BF16, device=dev)
        self.w3 = [torch.empty(max(16, c.weight.shape[0]), 9 * c.weight.shape[1], dtype=BF16, device=dev)
                   for c in self.convs]
        self.w3_last = torch.zeros(16, 9 * self.Cn, dtype=BF16, device=dev)
        self.tmpC = torch.zeros(max(c.weight.shape[0] for c in self.convs), dtype=F32, device=dev)
        self.sync = GradSync(module)
        self.pack()

    def pack(self, full=True):
        ops.pack_proj(self.conv0.weight.detach(), self.w_proj)
        for c, w in zip(self.convs, self.w3):
            ops.pack_conv3(c.weight.detach(), w)
        ops.pack_conv3(self.conv_last.weight.detach(), self.w3_last)

    def forward(self, lat, tag="g", training=True, out=None):
        B = lat.shape[0]
        g = self.bufs.get
        a = g(f"{tag}.a0", (B, 4, 4, self.C0))
        h = g(f"{tag}.h0", (B, 4, 4, self.C0))
        ops.gemm_nt(lat, self.w_proj, out=a.view(B, 16 * self.C0))
        self.bn0.forward(a, h, B * 16, training, tag=tag)
        H = 4
        for l, (c, bn) in enumerate(zip(self.convs, self.bns), start=1):
            Cout, Cin = c.weight.shape[0], c.weight.shape[1]
            u = g(f"{tag}.u{l}", (B, 2 * H + 2, 2 * H + 2, Cin))
            ops.upsample2x_reflectpad(h, u)
            H *= 2
            a = g(f"{tag}.a{l}", (B, H, H, Cout))
            sws = bn.stats_ws(training)
            ops.conv3x3(u, self.w3[l - 1], a, bias=c.bias.detach(), stats=sws)
            h = g(f"{tag}.h{l}", (B, H, H, Cout))
            bn.forward(a, h, B * H * H, training, tag=tag, stats=sws)
        u = g(f"{tag}.ulast", (B, 2 * H + 2, 2 * H + 2, self.Cn))
        ops.upsample2x_reflectpad(h, u)
        if out is None:
            out = g(f"{tag}.img", (B, self.Cimg, 2 * H, 2 * H), F32)
        ops.conv3x3(u, self.w3_last, out, bias=self.conv_last.bias.detach())
        return out

    def backward(self, lat, d_img, img, tag="g"):
        B = lat.shape[0]
        g = self.bufs.get
        n, S = self.n, self.size
        H = S // 2
        u = g(f"{tag}.ulast", (B, S + 2, S + 2, self.Cn))
        du = g("bwd.dulast", (B, S + 2, S + 2, self.Cn))
        ops.upg_last_bwd(u, d_img, self.conv_last.weight.detach(), _grad_of(self.conv_last.weight), du)
        ops.img_channel_sum(d_img, _grad_of(self.conv_last.bias), mode=0, acc=0.0)
        self.sync.layer_done(self.conv_last.weight, self.conv_last.bias)
        dh = g(f"bwd.dh{n}", (B, H, H, self.Cn))
        ops.upsample2x_reflectpad_bwd(du, dh)
        for l in range(n, 0, -1):
            c, bn = self.convs[l - 1], self.bns[l - 1]
            Cout, Cin = c.weight.shape[0], c.weight.shape[1]
            a = g(f"{tag}.a{l}", (B, H, H, Cout))
            da = g(f"bwd.da{l}", (B, H, H, Cout))
            bn.backward(dh, a, da, B * H * H, param_grads=True, tag=tag)
            u = g(f"{tag}.u{l}", (B, H + 2, H + 2, Cin))
            ops.conv3x3_wgrad(da, u, _grad_of(c.weight))
            ops.col_sum(da, B * H * H, Cout, self.tmpC, _grad_of(c.bias), 0.0)
            self.sync.layer_done(c.weight, c.bias, bn.mod.weight, bn.mod.bias)
            du = g(f"bwd.du{l}", (B, H + 2, H + 2, Cin))
            ops.conv3x3_dgrad(da, self.w3[l - 1], du)
            H //= 2
            dh = g(f"bwd.dh{l - 1}", (B, H, H, Cin))
            ops.upsample2x_reflectpad_bwd(du, dh)
        a0 = g(f"{tag}.a0", (B, 4, 4, self.C0))
        da0 = g("bwd.da0", (B, 4, 4, self.C0))
        self.bn0.backward(dh, a0, da0, B * 16, param_grads=True, tag=tag)
        ops.proj_wgrad(lat, da0, _grad_of(self.conv0.weight))
        self.sync.layer_done(self.conv0.weight, self.bn0.mod.weight, self.bn0.mod.bias)


# ====================================================================================================== critic
class CriticEngine:
    """Kernel schedule for torchgan-style DCGANDiscriminator (batchnorm=True, LeakyReLU(0.2) everywhere)."""

    def __init__(self, module):
        blocks = list(module.model)
        self.module = module
        dev = next(module.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("CriticEngine needs the module on a CUDA device (there is no CPU path)")
        self.device = dev
        self.bufs = _Bufs(dev)
        first = blocks[0]
        self.conv0 = first[0]
        _check_act(first[1], SLOPE, "critic")
        self.Cimg, self.C0 = self.conv0.weight.shape[1], self.conv0.weight.shape[0]
        if self.Cimg > 4:
            raise NotImplementedError("critic input with more than 4 channels is not implemented")
        self.convs, self.bns = [], []
        for blk in blocks[1:]:
            if len(blk) != 3 or not isinstance(blk[1], nn.BatchNorm2d):
                raise NotImplementedError("critic without BatchNorm is not implemented on the sm_100a path")
            _check_act(blk[2], SLOPE, "critic")
            self.convs.append(blk[0])
            self.bns.append(_BN(blk[1], dev))
        self.head = module.disc[0]
        _check_act(module.disc[1], SLOPE, "critic head")
        if self.head.bias is not None:
            raise NotImplementedError("critic head with bias is not implemented")
        self.n = len(self.convs)
        self.Cn = self.head.weight.shape[1]
        self.size = 4 * (2 ** (self.n + 1))
        if self.C0 % 64:
            raise NotImplementedError("channel counts must be multiples of 64 on the sm_100a path")
        self.w_col0 = torch.empty(self.C0, 64, dtype=BF16, device=dev)
        self.w_colT0 = torch.zeros(16 * self.Cimg, self.C0, dtype=BF16, device=dev)
        self.w_down, self.w_upk, self.w_up9 = [], [], []
        self._native_params = [c.weight for c in self.convs]
        _to_native(self._native_params)
        for c in self.convs:
            Cp, Cs = c.weight.shape[0], c.weight.shape[1]
            self.w_down.append(torch.empty(Cp, 16 * Cs, dtype=BF16, device=dev))
            c.weight._rg_shadow = self.w_down[-1]        # re-emitted by the fused Adam step
            self.w_upk.append(torch.zeros(4, ops.up_pad(Cs), 4 * Cp, dtype=BF16, device=dev) if Cs <= 128 else None)
            self.w_up9.append(torch.zeros(9, Cp // 64, 2, 4, 32, 64, dtype=BF16, device=dev) if Cs == 64 else None)
        self.w_head = torch.empty(16 * self.Cn, dtype=F32, device=dev)
        self.tmpC = torch.zeros(max([self.C0] + [c.weight.shape[0] for c in self.convs]), dtype=F32, device=dev)
        self.gp_partial = torch.zeros(1024, dtype=F32, device=dev)
        self.gp_out = torch.zeros(3, dtype=F32, device=dev)
        self.fused_img = FUSED_IMG and self.C0 == 64 and self.Cimg <= 3 and self.size >= 16
        self._img_in = {}            # tag -> (x, y, mode, eps_dev): the image operand of the pass, for the layer-0 wgrad
        self.sync = GradSync(module)
        # critic step: real / fake passes share their weight- and input-gradient launches (see backward_pair)
        self.bufs.pair("real.col", "fake.col")
        self.bufs.pair("real.da0", "fake.da0")
        for l in range(0, self.n + 1):
            self.bufs.pair(f"real.h{l}", f"fake.h{l}")
            self.bufs.pair(f"real.dh{l}", f"fake.dh{l}")
            if l >= 1:
                self.bufs.pair(f"real.da{l}", f"fake.da{l}")
        # gradient penalty: the two weight-gradient terms of layer l, (da_l (x) A_dh_{l-1}) + (T_l (x) h_{l-1}), are one
        # contraction over the concatenated pixel axis
        for l in range(1, self.n + 1):
            self.bufs.pair(f"gp.da{l}", f"gp.Aa{l}" if l == self.n else f"gp.T{l}")
            self.bufs.pair(f"gp.Adh{l - 1}", f"gp.h{l - 1}")
        self.pack()

    def pack(self, full=True):
        """See GeneratorEngine.pack: full=False right after the fused Adam step (w_down already re-emitted)."""
        if self.fused_img:                   # the other fused image-side kernels read the fp32 parameter itself
            self.w_up_img0 = ops.img_conv_up_pack(self.conv0.weight.detach(), getattr(self, "w_up_img0", None))
        else:
            ops.pack_edge(self.conv0.weight.detach(), self.w_col0)
            ops.pack_edge_t(self.conv0.weight.detach(), self.w_colT0)
        if full:
            for p in _to_native(self._native_params):
                self.sync.rebind(p)
            for c, wd in zip(self.convs, self.w_down):
                ops.cast_pad_bf16(ops.phys2d(c.weight.detach()), out=wd)
        for c, wd, wu, w9 in zip(self.convs, self.w_down, self.w_upk, self.w_up9):
            if wu is not None:
                ops.pack_up_from_down(wd, wu, c.weight.shape[1])
            if w9 is not None:
                ops.pack_up9_from_down(wd, c.weight.shape[1], out=w9)
        ops.pack_head(self.head.weight.detach(), self.w_head)

    def _wup(self, l, npix=0):
        return _up_operand(self, l, npix)

    # ------------------------------------------------------------------ forward
    def forward(self, x, tag="d", training=True, mix=None):
        """x: fp32 NCHW image [B, Cimg, S, S] -> critic output fp32 [B].
        mix = (y, eps_dev): run on the gradient-penalty interpolate eps*x + (1-eps)*y (src/wgan_loss.py:376-380), which is
        formed while the first layer stages the image -- never materialised."""
        g = self.bufs.get
        S = self.size
        B = x.shape[0]
        H = S // 2
        npix = B * H * H
        y, eps_dev = mix if mix is not None else (None, None)
        mode = 1 if mix is not None else 0
        self._img_in[tag] = (x, y, mode, eps_dev)
        h = g(f"{tag}.h0", (B, H, H, self.C0))
        if self.fused_img:
            ops.img_conv_down(x, self.conv0.weight.detach(), h, y=y, mode=mode, eps_dev=eps_dev,
                              bias=self.conv0.bias.detach(), slope=SLOPE)
        else:
            col = g(f"{tag}.col", (npix, 64))
            ops.im2col_img(x, col, y=y, mode=mode, eps_dev=eps_dev)
            ops.gemm_nt(col, self.w_col0, out=h.view(npix, self.C0), col_shift=self.conv0.bias.detach(), slope=SLOPE)
        for l, (c, bn) in enumerate(zip(self.convs, self.bns), start=1):
            Cp = c.weight.shape[0]
            H //= 2
            a = g(f"{tag}.a{l}", (B, H, H, Cp))
            sws = bn.stats_ws(training)
            ops.conv_down(h, self.w_down[l - 1], out=a, stats=sws)
            h = g(f"{tag}.h{l}", (B, H, H, Cp))
            bn.forward(a, h, B * H * H, training, tag=tag, stats=sws)
        a6 = g(f"{tag}.a6", (B,), F32)
        out = g(f"{tag}.out", (B,), F32)
        ops.head_fwd(h, self.w_head, B, 16 * self.Cn, SLOPE, a6, out)
        return out

    # ------------------------------------------------------------------ first-order backward
    def backward(self, B, dout_const, tag="d", params=False, acc=0.0, want_dimg=False, keep_du=False, final=False,
                 dout=None):
        """Backward of sum_b dout_const * out[b] (or sum_b dout[b] * out[b] with a device vector `dout`) through the
        pass saved under `tag`.

        params: also produce parameter gradients (acc=1.0 accumulates onto existing .grad).
        want_dimg: return dL/d(image) as fp32 NCHW.   keep_du: keep per-layer du / da (gradient-penalty step 2).
        """
        g = self.bufs.get
        n = self.n
        H = 4
        hn = g(f"{tag}.h{n}", (B, H, H, self.Cn))
        a6 = g(f"{tag}.a6", (B,), F32)
        da6 = g(f"{tag}.da6", (B,), F32)
        dh = g(f"{tag}.dh{n}", (B, H, H, self.Cn))
        ops.head_bwd_data(a6, dout_const, self.w_head, B, 16 * self.Cn, SLOPE, da6, dh, dout=dout)
        if params:
            ops.head_wgrad(da6, hn, B, 16 * self.Cn, self.Cn, _grad_of(self.head.weight), acc)
            if final:
                self.sync.layer_done(self.head.weight)
        fuse = FUSED_BWD and not keep_du     # the gradient-penalty pass keeps dh AND du of every layer: unfused
        fuse0 = FUSED_LRELU and not keep_du
        fused = None
        for l in range(n, 0, -1):
            c, bn = self.convs[l - 1], self.bns[l - 1]
            Cp, Cs = c.weight.shape[0], c.weight.shape[1]
            a = g(f"{tag}.a{l}", (B, H, H, Cp))
            da = g(f"{tag}.da{l}", (B, H, H, Cp))
            du = g(f"{tag}.du{l}", (B, H, H, Cp)) if keep_du else None
            bn.backward(dh, a, da, B * H * H, param_grads=params, acc_gamma=acc, acc_beta=acc, du_out=du, tag=tag,
                        fused=fused)
            hprev = g(f"{tag}.h{l - 1}", (B, 2 * H, 2 * H, Cs))
            if params:
                ops.conv_wgrad(da, hprev, _grad_of(c.weight), beta=acc)
                if final:
                    self.sync.layer_done(c.weight, bn.mod.weight, bn.mod.bias)
            H *= 2
            fused, aux = None, None
            if l >= 2:
                dh = g(f"{tag}.dh{l - 1}", (B, H, H, Cs))
                if fuse:
                    fused = self.bns[l - 2].bwd_ws()
                    if fused is not None:
                        aux = self.bns[l - 2].aux(g(f"{tag}.a{l - 1}", (B, H, H, Cs)), tag)
            elif fuse0:      # layer 0 has no BatchNorm: the LeakyReLU mask goes into the epilogue, output is da0
                dh = g(f"{tag}.da0", (B, H, H, self.C0))
                aux = ("lrelu", g(f"{tag}.h0", (B, H, H, self.C0)), SLOPE)
            else:
                dh = g(f"{tag}.dh0", (B, H, H, Cs))
            ops.conv_up(da, self._wup(l, da.shape[0] * da.shape[1] * da.shape[2]), Cs, out=dh, stats=fused, aux=aux)
        h0 = g(f"{tag}.h0", (B, H, H, self.C0))
        da0 = g(f"{tag}.da0", (B, H, H, self.C0))
        npix = B * H * H
        if not fuse0:
            ops.lrelu_bwd(dh, h0, SLOPE, da0, npix, self.C0)
        if params:
            self._wgrad0(da0, tag, acc, acc)
            if final:
                self.sync.layer_done(self.conv0.weight, self.conv0.bias)
        if want_dimg:
            dimg = g(f"{tag}.dimg", (B, self.Cimg, 2 * H, 2 * H), F32)
            if self.fused_img:
                ops.img_conv_up(da0, self.w_up_img0, dimg, Cimg=self.Cimg)
            else:
                colimg = g("bwd.colimg", (npix, 16 * self.Cimg), F32)
                ops.conv_up_img_col(da0, self.w_colT0, self.Cimg, colimg, dimg)
            return dimg
        return None

    def _wgrad0(self, da0, tag, acc_w, acc_b, img=None):
        """Layer-0 weight (+ bias, when acc_b is not None) gradient of the pass saved under `tag`:
        conv0.weight.grad = acc_w * grad + da0 (x) image operand of that pass (or the explicit `img` tuple)."""
        x, y, mode, eps_dev, mul_dev = (img if img is not None else self._img_in[tag] + (None,))
        B, H = da0.shape[0], da0.shape[1]
        npix = B * H * H
        gb = _grad_of(self.conv0.bias) if acc_b is not None else None
        if self.fused_img:
            ops.img_conv_wgrad(da0, x, _grad_of(self.conv0.weight), y=y, mode=mode, eps_dev=eps_dev, mul_dev=mul_dev,
                               acc=acc_w, dbias=gb, acc_bias=acc_b or 0.0)
            return
        if gb is not None:
            ops.col_sum(da0, npix, self.C0, self.tmpC, gb, acc_b)
        col = self.bufs.get("bwd.col0", (npix, 64))
        ops.im2col_img(x, col, y=y, mode=mode, eps_dev=eps_dev, mul_dev=mul_dev)
        dcol = self.bufs.get("bwd.dcol", (self.C0, 64), F32)
        ops.gemm_tn(da0.view(npix, self.C0), col, out=dcol)
        ops.unpack_edge_grad(dcol, _grad_of(self.conv0.weight), acc=acc_w)

    def backward_pair(self, B, passes=(("real", -1.0), ("fake", 1.0))):
        """Critic-step backward of sum_b c_real*out_real[b] + c_fake*out_fake[b] (passes = ((tag, c*B), ...)) with
        parameter gradients (overwriting .grad): BatchNorm backward runs per pass (its statistics are per pass), the
        weight gradient and the input gradient of every conv layer run ONCE over the concatenated 2B batch."""
        g = self.bufs.get
        n = self.n
        (ta, _), (tb, _) = passes
        fused = [None, None]
        H = 4
        for i, (tag, c) in enumerate(passes):
            hn = g(f"{tag}.h{n}", (B, H, H, self.Cn))
            a6 = g(f"{tag}.a6", (B,), F32)
            da6 = g(f"{tag}.da6", (B,), F32)
            dh = g(f"{tag}.dh{n}", (B, H, H, self.Cn))
            ops.head_bwd_data(a6, c / B, self.w_head, B, 16 * self.Cn, SLOPE, da6, dh)
            ops.head_wgrad(da6, hn, B, 16 * self.Cn, self.Cn, _grad_of(self.head.weight), float(i > 0))
        self.sync.layer_done(self.head.weight)
        for l in range(n, 0, -1):
            c_, bn = self.convs[l - 1], self.bns[l - 1]
            Cp, Cs = c_.weight.shape[0], c_.weight.shape[1]
            for i, (tag, _) in enumerate(passes):
                a = g(f"{tag}.a{l}", (B, H, H, Cp))
                da = g(f"{tag}.da{l}", (B, H, H, Cp))
                dh = g(f"{tag}.dh{l}", (B, H, H, Cp))
                bn.backward(dh, a, da, B * H * H, param_grads=True, acc_gamma=float(i > 0), acc_beta=float(i > 0), tag=tag,
                            fused=fused[i])
            da2 = self.bufs.joint(f"{ta}.da{l}", (B, H, H, Cp))
            hprev2 = self.bufs.joint(f"{ta}.h{l - 1}", (B, 2 * H, 2 * H, Cs))
            ops.conv_wgrad(da2, hprev2, _grad_of(c_.weight))
            self.sync.layer_done(c_.weight, bn.mod.weight, bn.mod.bias)
            H *= 2
            fused = [None, None]
            bnp = self.bns[l - 2] if l >= 2 else None
            if bnp is not None and bnp.bwd_ws() is not None:
                # BatchNorm statistics are per pass: one input-gradient launch per pass, each storing du of its pass
                # and summing it (the joint launch could not: its two halves normalise with different vectors)
                for i, (tag, _) in enumerate(passes):
                    fused[i] = bnp.bwd_ws(slot=1 + i)
                    da_t = g(f"{tag}.da{l}", (B, H // 2, H // 2, Cp))
                    ops.conv_up(da_t, self._wup(l, B * (H // 2) * (H // 2)), Cs, out=g(f"{tag}.dh{l - 1}", (B, H, H, Cs)),
                                stats=fused[i], aux=bnp.aux(g(f"{tag}.a{l - 1}", (B, H, H, Cs)), tag))
            elif l == 1 and FUSED_LRELU:
                # layer 0 has no BatchNorm: joint launch with the LeakyReLU mask in the epilogue, output is da0
                ops.conv_up(da2, self._wup(l, da2.shape[0] * da2.shape[1] * da2.shape[2]), Cs,
                            out=self.bufs.joint(f"{ta}.da0", (B, H, H, self.C0)),
                            aux=("lrelu", self.bufs.joint(f"{ta}.h0", (B, H, H, self.C0)), SLOPE))
            else:
                dh2 = self.bufs.joint(f"{ta}.dh{l - 1}", (B, H, H, Cs))
                ops.conv_up(da2, self._wup(l, da2.shape[0] * da2.shape[1] * da2.shape[2]), Cs, out=dh2)
        npix = B * H * H
        h0 = self.bufs.joint(f"{ta}.h0", (B, H, H, self.C0))
        dh0 = self.bufs.joint(f"{ta}.dh0", (B, H, H, self.C0))
        da0 = self.bufs.joint(f"{ta}.da0", (B, H, H, self.C0))
        if not FUSED_LRELU:
            ops.lrelu_bwd(dh0, h0, SLOPE, da0, 2 * npix, self.C0)
        # layer-0 weight / bias gradient: one launch per pass (their image operands are different tensors)
        for i, (tag, _) in enumerate(passes):
            self._wgrad0(g(f"{tag}.da0", (B, H, H, self.C0)), tag, float(i > 0), float(i > 0))
        self.sync.layer_done(self.conv0.weight, self.conv0.bias)

    # ------------------------------------------------------------------ gradient penalty
    def gradient_penalty(self, real, fake, eps_dev, lambd=10.0, tag="gp"):
        """WassersteinGradientPenaltyVAE core (src/wgan_loss.py:32-44, 376-388): writes d(lambda*P)/d(theta_D)
        into .grad (overwriting) and returns the device tensor [P, seed, ||g||]."""
        g = self.bufs.get
        n = self.n
        S = self.size
        B = real.shape[0]
        H0 = S // 2
        npix0 = B * H0 * H0
        # step 1: forward on x_hat = eps*real + (1-eps)*fake (train-mode BN, running stats updated)
        self.forward(real, tag=tag, mix=(fake, eps_dev))
        # step 2: g = d(sum out)/d(x_hat), keeping du_l, da_l and the BN backward sums
        grad_x = self.backward(B, 1.0, tag=tag, params=False, want_dimg=True, keep_du=True)
        ops.gp_norm(grad_x, lambd, self.gp_partial, self.gp_out)
        seed = self.gp_out[1:2]
        # step 3: adjoint sweep bottom -> top, seeded with A_g = seed * g
        da0 = g(f"{tag}.da0", (B, H0, H0, self.C0))
        h0 = g(f"{tag}.h0", (B, H0, H0, self.C0))
        A_dh = g(f"{tag}.Adh0", (B, H0, H0, self.C0))
        self._wgrad0(da0, tag, 0.0, None, img=(grad_x, None, 0, None, seed))      # da0 (x) A_g
        if self.fused_img:                   # A_dh0 = conv(A_g; W0) * lrelu'(h0), mask applied in the store loop
            ops.img_conv_down(grad_x, self.conv0.weight.detach(), A_dh, mul_dev=seed, mask_src=h0, mask_slope=SLOPE)
        else:
            col_g = g("bwd.col0", (npix0, 64))                                     # still holds im2col(seed * grad_x)
            A_da0 = g(f"{tag}.Ada0", (B, H0, H0, self.C0))
            ops.gemm_nt(col_g, self.w_col0, out=A_da0.view(npix0, self.C0))
            ops.lrelu_bwd(A_da0, h0, SLOPE, A_dh, npix0, self.C0)
        H = H0
        for l in range(1, n + 1):
            c, bn = self.convs[l - 1], self.bns[l - 1]
            Cp = c.weight.shape[0]
            H //= 2
            M = B * H * H
            da = g(f"{tag}.da{l}", (B, H, H, Cp))
            du = g(f"{tag}.du{l}", (B, H, H, Cp))
            a = g(f"{tag}.a{l}", (B, H, H, Cp))
            # (the da_l (x) A_dh_{l-1} weight-gradient term is contracted in step 4, jointly with T_l (x) h_{l-1})
            ggI = g(f"{tag}.ggI{l}", (B, H, H, Cp))
            ops.conv_down(A_dh, self.w_down[l - 1], out=ggI)
            # bn.bsums still holds S(du), S(du*xhat) of THIS layer from step 2
            bs = bn.st(tag)
            ops.bn_gp_reduce(ggI, a, du, bs.mean, bs.rstd, M, Cp, bs.q)
            A_dh = g(f"{tag}.Adh{l}", (B, H, H, Cp))
            A_a = g(f"{tag}.Aa{l}", (B, H, H, Cp))
            ops.bn_gp_apply(ggI, a, du, bs.mean, bs.rstd, bn.mod.weight, bs.scale, bs.shift, SLOPE, bs.bsums, bs.q, M,
                            Cp, A_dh, A_a, _grad_of(bn.mod.weight), 0.0)
        da6 = g(f"{tag}.da6", (B,), F32)
        ops.head_wgrad(da6, A_dh, B, 16 * self.Cn, self.Cn, _grad_of(self.head.weight), 0.0)
        self.sync.layer_done(self.head.weight)
        # step 4: ordinary backward over the forward graph, seeded by the A_a terms
        T = g(f"{tag}.Aa{n}", (B, 4, 4, self.Cn))
        _grad_of(self.bns[n - 1].mod.bias).zero_()
        H = 4
        for l in range(n, 0, -1):
            c = self.convs[l - 1]
            Cp, Cs = c.weight.shape[0], c.weight.shape[1]
            lo2 = self.bufs.joint(f"{tag}.da{l}", (B, H, H, Cp))                  # [da_l ; T_l]
            hi2 = self.bufs.joint(f"{tag}.Adh{l - 1}", (B, 2 * H, 2 * H, Cs))     # [A_dh_{l-1} ; h_{l-1}]
            ops.conv_wgrad(lo2, hi2, _grad_of(c.weight))
            # conv_l's weight is final now; BN_l's gamma/beta were finalised by step 3 (l = n) or the previous turn
            self.sync.layer_done(c.weight, self.bns[l - 1].mod.weight, self.bns[l - 1].mod.bias)
            H *= 2
            wup = self._wup(l, T.shape[0] * T.shape[1] * T.shape[2])
            if l - 1 >= 1:
                bnp = self.bns[l - 2]
                a = g(f"{tag}.a{l - 1}", (B, H, H, Cs))
                A_h = g(f"{tag}.Ah{l - 1}", (B, H, H, Cs))
                fused = bnp.bwd_ws()
                ops.conv_up(T, wup, Cs, out=A_h, stats=fused, aux=bnp.aux(a, tag) if fused is not None else None)
                A_a = g(f"{tag}.Aa{l - 1}", (B, H, H, Cs))
                Tn = g(f"{tag}.T{l - 1}", (B, H, H, Cs))
                # note: overwrites bnp.bsums (step-2 sums of layer l-1 were consumed in step 3 already)
                bnp.backward(A_h, a, Tn, B * H * H, param_grads=True, acc_gamma=1.0, acc_beta=0.0, add=A_a, tag=tag,
                             fused=fused)
                T = Tn
            else:
                A_a0 = g(f"{tag}.Aa0", (B, H, H, self.C0))
                if FUSED_LRELU:
                    ops.conv_up(T, wup, Cs, out=A_a0, aux=("lrelu", h0, SLOPE))
                else:
                    A_h = g(f"{tag}.Ah{l - 1}", (B, H, H, Cs))
                    ops.conv_up(T, wup, Cs, out=A_h)
                    ops.lrelu_bwd(A_h, h0, SLOPE, A_a0, npix0, self.C0)
                self._wgrad0(A_a0, tag, 1.0, 0.0)                   # + A_a0 (x) x_hat, and the bias gradient
                self.sync.layer_done(self.conv0.weight, self.conv0.bias)
        return self.gp_out


# ====================================================================================================== encoder
class EncoderEngine:
    """Frozen, eval-mode betaVAE.encode -> z_mean (src/betaVAE.py:102-107 as used at src/wgan_loss.py:97):
    3 x [Linear + eval BatchNorm1d + LeakyReLU(0.01)] folded into GEMM epilogues, then the z_mu Linear."""

    def __init__(self, vae):
        enc = vae.encoder.encoder
        dev = next(vae.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("EncoderEngine needs the betaVAE on a CUDA device (there is no CPU path)")
        self.device = dev
        self.vae = vae
        self.layers = []
        self.bufs = _Bufs(dev)
        self.in_features = enc[1][0].weight.shape[1]
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        """Re-derive packed weights / folded BN vectors from the module (call after loading a checkpoint)."""
        enc = self.vae.encoder.encoder
        self.layers = []
        for blk in list(enc)[1:]:
            lin, bn = blk[0], blk[1]
            K = lin.weight.shape[1]
            Kp = (K + 63) // 64 * 64
            w = ops.cast_pad_bf16(lin.weight.detach().contiguous(), Kp)
            scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).float().contiguous()
            shift = ((lin.bias - bn.running_mean) * scale + bn.bias).float().contiguous()
            self.layers.append((w, scale, shift, blk[2].negative_slope, lin.weight.shape[0], Kp))
        mu, lv = self.vae.z_mu, self.vae.z_logvar
        self.w_mu = ops.cast_pad_bf16(mu.weight.detach().contiguous(), mu.weight.shape[1])
        self.b_mu = mu.bias.detach().float().contiguous()
        self.w_lv = ops.cast_pad_bf16(lv.weight.detach().contiguous(), lv.weight.shape[1])
        self.b_lv = lv.bias.detach().float().contiguous()

    def encode(self, x, want_all=False):
        """x: fp32 [B, F] on the device -> z_mean fp32 [B, Z] (and z_logvar, last hidden when want_all)."""
        B = x.shape[0]
        g = self.bufs.get
        Kp0 = self.layers[0][5]
        h = g("x", (B, Kp0))
        ops.cast_pad_bf16(x.contiguous(), Kp0, out=h)
        for i, (w, scale, shift, slope, N, Kp) in enumerate(self.layers):
            Np = (N + 63) // 64 * 64        # next layer's K: keep the row stride padded, pad columns stay zero
            o = g(f"h{i}", (B, Np), zero=True)
            ops.gemm_nt(h, w, out=o, col_scale=scale, col_shift=shift, slope=slope, N=N)
            h = o
        z = g("z", (B, self.w_mu.shape[0]), F32)
        ops.gemm_nt(h, self.w_mu, out=z, col_shift=self.b_mu)
        if not want_all:
            return z
        lv = g("zlv", (B, self.w_lv.shape[0]), F32)
        ops.gemm_nt(h, self.w_lv, out=lv, col_shift=self.b_lv)
        return z, lv, h[:, :self.layers[-1][4]].float()


# ====================================================================================================== decoder
class DecoderEngine:
    """Eval-mode betaVAE decoder (src/betaVAE.py:80-92, used by decode / forward / sample, :109-143):
    [Linear + eval BatchNorm1d + LeakyReLU(0.01)] per hidden layer folded into GEMM epilogues, then Linear + Tanh
    (tanh in the fp32 epilogue)."""

    def __init__(self, vae):
        dev = next(vae.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("DecoderEngine needs the betaVAE on a CUDA device (there is no CPU path)")
        self.device, self.vae = dev, vae
        self.bufs = _Bufs(dev)
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        blocks = list(self.vae.decoder)
        self.layers = []
        for blk in blocks[:-1]:
            lin, bn = blk[0], blk[1]
            Kp = (lin.weight.shape[1] + 63) // 64 * 64
            w = ops.cast_pad_bf16(lin.weight.detach().contiguous(), Kp)
            scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).float().contiguous()
            shift = ((lin.bias - bn.running_mean) * scale + bn.bias).float().contiguous()
            self.layers.append((w, scale, shift, blk[2].negative_slope, lin.weight.shape[0], Kp))
        last = blocks[-1][0]
        self.F = last.weight.shape[0]
        self.Fp = (self.F + 3) // 4 * 4
        self.w_last = ops.cast_pad_bf16(last.weight.detach().contiguous(), (last.weight.shape[1] + 63) // 64 * 64)
        self.b_last = last.bias.detach().float().contiguous()

    def decode(self, z):
        """z: fp32 [B, z_dim] on the device -> tanh(decoder(z)) fp32 [B, in_channels]."""
        B = z.shape[0]
        g = self.bufs.get
        K0 = self.layers[0][5] if self.layers else self.w_last.shape[1]
        h = g("z", (B, K0), zero=True)
        ops.cast_pad_bf16(z.contiguous(), K0, out=h)
        for i, (w, scale, shift, slope, N, Kp) in enumerate(self.layers):
            Np = (N + 63) // 64 * 64
            o = g(f"h{i}", (B, Np), zero=True)
            ops.gemm_nt(h, w, out=o, col_scale=scale, col_shift=shift, slope=slope, N=N)
            h = o
        out = g("out", (B, self.Fp), F32)
        ops.gemm_nt(h, self.w_last, out=out, col_shift=self.b_last, N=self.F, tanh=True)
        return out[:, :self.F].clone()


# ====================================================================================================== VAE training
class VAETrainEngine:
    """betaVAE training step (BASELINE config 5; inner step of train_betaVAE, src/betaVAE.py:216-236):
    Dropout -> 3x[Linear, BatchNorm1d(train), LeakyReLU(0.01)] -> (z_mu | z_logvar) -> reparametrise ->
    2x[Linear, BN1d, LeakyReLU] -> Linear + Tanh; loss = MSE + beta * KLD (betaVAEloss, :145-162); backward; gradients
    land in the flat buffer for the fused Adam.  Every Linear is a tcgen05 GEMM (forward NT, input gradient NN read
    MN-major from the same bf16 weight copy, weight gradient TN with split-K); BN1d reuses the BN2d kernels with M = B."""

    def __init__(self, vae):
        dev = next(vae.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("VAETrainEngine needs the betaVAE on a CUDA device (there is no CPU path)")
        self.vae, self.device = vae, dev
        self.bufs = _Bufs(dev)
        enc = list(vae.encoder.encoder)[1:]
        dec = list(vae.decoder)
        self.p_drop = vae.encoder.encoder[0][0].p
        self.blocks = []                      # (linear, _BN) for the 3 encoder + 2 decoder hidden blocks
        for blk in enc + dec[:-1]:
            self.blocks.append((blk[0], _BN(blk[1], dev, slope=blk[2].negative_slope)))
        self.n_enc = len(enc)
        self.last = dec[-1][0]
        self.F = self.last.weight.shape[0]
        self.Z = vae.z_mu.weight.shape[0]
        self.Fp = (self.F + 7) // 8 * 8
        self.w = [torch.empty(l.weight.shape[0], (l.weight.shape[1] + 7) // 8 * 8, dtype=BF16, device=dev)
                  for l, _ in self.blocks]
        self.w_last = torch.empty(self.F, self.last.weight.shape[1], dtype=BF16, device=dev)
        self.w_cat = torch.empty(2 * self.Z, self.Z, dtype=BF16, device=dev)
        self.b_cat = torch.empty(2 * self.Z, dtype=F32, device=dev)
        self.partial = torch.zeros(2, 256, dtype=F32, device=dev)
        self.out3 = torch.zeros(3, dtype=F32, device=dev)
        self.sync = GradSync(vae)
        # nn.Linear weights whose bf16 operand copy has the SAME element order (no column padding): the fused Adam
        # step re-emits the copy while it updates the fp32 master, so no cast pass follows the optimiser step
        self._shadowed = set()
        pairs = [(l.weight, w) for (l, _), w in zip(self.blocks, self.w)]
        pairs += [(self.last.weight, self.w_last), (vae.z_mu.weight, self.w_cat[:self.Z]),
                  (vae.z_logvar.weight, self.w_cat[self.Z:])]
        for p, w in pairs:
            # same element order, or the same rows with a zero-padded K (Adam then writes the copy row by row:
            # rg_adam_build_table_pitched) -- e.g. encoder layer 0, [6000, 19198] against a [6000, 19200] operand
            if w.is_contiguous() and w.shape[0] == p.shape[0] and w.shape[1] >= p.shape[1]:
                p._rg_shadow = w
                self._shadowed.add(id(p))
        self.pack()

    def pack(self, full=True):
        """full=False (right after the fused Adam step): only the copies Adam does not re-emit (padded layouts, the
        concatenated head bias)."""
        def cast(p, cols, out):
            if full or id(p) not in self._shadowed:
                ops.cast_pad_bf16(p.detach(), cols, out=out)
        for (l, _), w in zip(self.blocks, self.w):
            cast(l.weight, w.shape[1], w)
        cast(self.last.weight, self.w_last.shape[1], self.w_last)
        cast(self.vae.z_mu.weight, self.Z, self.w_cat[:self.Z])
        cast(self.vae.z_logvar.weight, self.Z, self.w_cat[self.Z:])
        with torch.no_grad():
            self.b_cat[:self.Z].copy_(self.vae.z_mu.bias)
            self.b_cat[self.Z:].copy_(self.vae.z_logvar.bias)

    def step(self, x, beta, keep_mask=None, eps=None):
        """Forward + backward of one batch x fp32 [B, F] (device).  keep_mask: fp32 0/1 [B, F] dropout keep mask,
        eps: fp32 [B, Z] reparametrisation noise (both drawn with torch's device RNG when omitted).
        Returns the device tensor [total, reconstruction, kl]; gradients are in p.grad (flat buffer)."""
        g = self.bufs.get
        B, F, Z = x.shape[0], self.F, self.Z
        if keep_mask is None:
            keep_mask = (torch.rand(B, F, device=self.device) >= self.p_drop).float()
        if eps is None:
            eps = torch.randn(B, Z, device=self.device)
        # ---------------- forward
        xd = g("xd", (B, self.Fp))
        ops.mul_cast_pad_bf16(x, xd, mul=keep_mask, scale=1.0 / (1.0 - self.p_drop))
        h, hs, As = xd, [xd], []
        for i, ((lin, bn), w) in enumerate(zip(self.blocks, self.w)):
            if i == self.n_enc:                 # latent sits between encoder and decoder
                mulv = g("mulv", (B, 2 * Z), F32)
                ops.gemm_nt(h, self.w_cat, out=mulv, col_shift=self.b_cat)
                z = g("z", (B, Z))
                n_kld = ops.vae_reparam(mulv, eps, z, self.partial[1])
                h_lat, h = h, z
                hs.append(z)
            N, K = lin.weight.shape
            a = g(f"a{i}", (B, N))
            ops.gemm_nt(h, w, out=a, col_shift=lin.bias.detach(), K=K)
            hn = g(f"h{i}", (B, N))
            bn.forward(a, hn, B, training=True, tag="vae")
            As.append(a)
            h = hn
            hs.append(hn)
        pre = g("pre", (B, self.Fp), F32)
        ops.gemm_nt(h, self.w_last, out=pre, col_shift=self.last.bias.detach(), N=F)
        dpre = g("dpre", (B, self.Fp))
        n_sse = ops.vae_recon(pre, x, 2.0 / (B * F), dpre, self.partial[0])
        ops.vae_loss_finalize(self.partial[0], n_sse, self.partial[1], n_kld, B, F, beta, self.out3)
        # ---------------- backward
        tmp = g("tmpcols", (max(self.Fp, 2 * Z, 8192),), F32)
        ops.gemm_tn(dpre, h, out=_grad_of(self.last.weight), M=F)
        ops.col_sum(dpre, B, self.Fp, tmp, tmp, 0.0)
        with torch.no_grad():
            _grad_of(self.last.bias).copy_(tmp[:F])
        dh = g("dh_last", (B, self.last.weight.shape[1]))
        ops.gemm_nn(dpre, self.w_last, out=dh, K=F)
        self.sync.layer_done(self.last.weight, self.last.bias)
        hidx = len(hs) - 2                       # hs[hidx] is the input of the block being differentiated
        for i in range(len(self.blocks) - 1, -1, -1):
            (lin, bn), w = self.blocks[i], self.w[i]
            N, K = lin.weight.shape
            da = g(f"da{i}", (B, N))
            bn.backward(dh, As[i], da, B, param_grads=True, tag="vae")
            hin = hs[hidx]
            ops.gemm_tn(da, hin, out=_grad_of(lin.weight), N=K)
            ops.col_sum(da, B, N, tmp, _grad_of(lin.bias), 0.0)
            self.sync.layer_done(lin.weight, lin.bias, bn.mod.weight, bn.mod.bias)
            if i > 0:
                dh = g(f"dh{i}", (B, K))
                ops.gemm_nn(da, w, out=dh, N=K)
            hidx -= 1
            if i == self.n_enc:                 # crossing the latent: dz -> (d_mu | d_logvar) -> encoder output
                dcat = g("dcat", (B, 2 * Z))
                ops.vae_latent_grad(dh, mulv, eps, beta / B, dcat)
                ops.gemm_tn(dcat[:, :Z], h_lat, out=_grad_of(self.vae.z_mu.weight))
                ops.gemm_tn(dcat[:, Z:], h_lat, out=_grad_of(self.vae.z_logvar.weight))
                ops.col_sum(dcat, B, 2 * Z, tmp, tmp, 0.0)
                with torch.no_grad():
                    _grad_of(self.vae.z_mu.bias).copy_(tmp[:Z])
                    _grad_of(self.vae.z_logvar.bias).copy_(tmp[Z:2 * Z])
                self.sync.layer_done(self.vae.z_mu.weight, self.vae.z_mu.bias, self.vae.z_logvar.weight,
                                     self.vae.z_logvar.bias)
                dh = g("dh_lat", (B, Z))
                ops.gemm_nn(dcat, self.w_cat, out=dh)
                hidx -= 1
        return self.out3
