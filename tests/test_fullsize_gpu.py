"""GPU parity at the REAL BASELINE.json config-2 shapes (gan_run_lung.json: 256x256 tiles, 2048-dim latent, 19198
protein-coding genes, 64..2048 channels, per-GPU batch 64), complementing the 32/64-pixel minis of test_train_gpu.py:

  * every contraction of every link at B = 64 against fp32 torch convolutions on the GPU (TF32 off) evaluated on
    the SAME bf16-rounded operands: F.conv2d / F.conv_transpose2d / torch.nn.grad.conv2d_weight -- for every operand
    layout the engines use; the generator's layer 0 (rg_gemm_nn at N = 32768) and its weight gradient; both
    image-side layers with their weight gradient; the critic head; BatchNorm backward at M = 64*128^2 and the
    BatchNorm double backward of the gradient penalty at M = 64*64^2 against float64 formulas (SURVEY Appendix C,
    the form tests/test_gp_math_cpu.py checks against autograd);
  * one whole iteration (G step, critic step, GP step) at B = 64 through the reference-facing train_ops against
    the oracle's train_ops executed in fp32 by torch on the same GPU (the oracle is device-agnostic plain torch;
    oracle.draw_noise is redirected so the CPU RNG stream is consumed exactly as on the CPU), with torch's own
    bf16 autocast deviation as the yardstick -- the B = 8 variant of this against the CPU oracle AND the
    reference's own golden values is test_train_gpu.py::test_train_steps_match_oracle[full256];
  * DCGANUpGenerator forward at 256x256; betaVAE.encode and one betaVAE train step at 19198 genes / batch 128.

Tolerances are those of the small-shape tests (stated there and in DESIGN.md section 4).
"""
import contextlib
import copy
import os
import tempfile

import pytest
import torch
import torch.nn.functional as F
from torch.optim import Adam

from oracle import ref_oracle as O

pytestmark = pytest.mark.gpu

CH = [64, 128, 256, 512, 1024, 2048]
B64 = 64


@pytest.fixture(autouse=True)
def _fp32_reference_math():
    """The torch reference must be true fp32 (no TF32 tensor-core shortcuts)."""
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    torch.cuda.empty_cache()


def _relmax(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten().to(a.device)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten().to(a.device)
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()


def _bf(t):
    return t.to(torch.bfloat16).float()


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _nchw(x):
    return x.float().permute(0, 3, 1, 2)


def _randn(shape, seed, dev, scale=1.0):
    g = torch.Generator(device=dev).manual_seed(seed)
    return _bf(torch.randn(shape, generator=g, device=dev) * scale)


# ------------------------------------------------------------------------------------------------ contractions
@pytest.mark.parametrize("layer", [1, 2, 3, 4, 5])
def test_link_contractions_vs_torch_fp32(cuda_dev, layer):
    """Link `layer` of the 256x256 networks at B = 64: fprop / dgrad / wgrad in every operand layout vs torch fp32."""
    from rnagan_b200 import ops
    Cs, Cp = CH[layer - 1], CH[layer]
    h = 128 >> layer
    x = _randn((B64, Cs, 2 * h, 2 * h), 100 + layer, cuda_dev)                 # hi side (NCHW for torch)
    y = _randn((B64, Cp, h, h), 200 + layer, cuda_dev)                         # lo side
    Wt = _randn((Cp, Cs, 4, 4), 300 + layer, cuda_dev, (16 * Cs) ** -0.5)
    xn, yn = _nhwc(x), _nhwc(y)
    w_down, w_up = ops.pack_link(Wt, want_up=Cs <= 128)
    # critic fprop / generator dgrad
    ref = F.conv2d(x, Wt, stride=2, padding=1)
    out = ops.conv_down(xn, w_down)
    assert out.shape == (B64, h, h, Cp)
    assert _relmax(_nchw(out), ref) < 8e-3 and _rel(_nchw(out), ref) < 4e-3
    # generator fprop / critic dgrad, every B-operand form
    ref = F.conv_transpose2d(y, Wt, stride=2, padding=1)
    operands = [w_down] + ([w_up] if Cs <= 128 else []) + ([ops.pack_up9_from_down(w_down, Cs)] if Cs == 64 else [])
    for w in operands:
        out = ops.conv_up(yn, w, Cs)
        assert out.shape == (B64, 2 * h, 2 * h, Cs)
        assert _relmax(_nchw(out), ref) < 8e-3 and _rel(_nchw(out), ref) < 4e-3, w.dim()
    del out
    # weight gradient, torch layout and the engines' native (channels_last) layout
    ref = torch.nn.grad.conv2d_weight(x, (Cp, Cs, 4, 4), y, stride=2, padding=1)
    dW = torch.empty(Cp, Cs, 4, 4, device=cuda_dev)
    dWn = torch.empty(Cp, Cs, 4, 4, device=cuda_dev).contiguous(memory_format=torch.channels_last)
    for d in (dW, dWn):
        ops.conv_wgrad(yn, xn, d)
        assert _relmax(d, ref) < 2e-3, d.stride()
    assert torch.equal(dW, dWn)
    # the joint [2B] launch of the critic / gradient-penalty steps (K = 2*B*h*h pixels)
    x2, y2 = torch.cat([xn, xn.flip(0)]), torch.cat([yn, yn.flip(0)])
    ops.conv_wgrad(y2, x2, dWn)
    assert _relmax(dWn, 2 * ref) < 2e-3


def test_generator_layer0_full_size(cuda_dev):
    """G.0 = ConvTranspose2d(2048, 2048, 4, 1, 0) on the 1x1 latent: rg_gemm_nn with the MN-major bf16 image of the
    native weight (N = 16*2048 = 32768), and rg_proj_wgrad into the native gradient layout."""
    from rnagan_b200 import ops
    E, C0 = 2048, 2048
    z = _randn((B64, E), 1, cuda_dev)
    Wt = _randn((E, C0, 4, 4), 2, cuda_dev, E ** -0.5).contiguous(memory_format=torch.channels_last)
    ref = F.conv_transpose2d(z.view(B64, E, 1, 1), Wt)                          # [B, C0, 4, 4]
    w_projkn = ops.cast_pad_bf16(ops.phys2d(Wt))
    assert w_projkn.shape == (E, 16 * C0)
    out = ops.gemm_nn(z.to(torch.bfloat16), w_projkn).view(B64, 4, 4, C0)
    assert _relmax(_nchw(out), ref) < 8e-3 and _rel(_nchw(out), ref) < 4e-3
    da0 = _randn((B64, C0, 4, 4), 3, cuda_dev)
    refw = torch.einsum("be,bchw->echw", z, da0)
    for dW in (torch.empty(E, C0, 4, 4, device=cuda_dev),
               torch.empty(E, C0, 4, 4, device=cuda_dev).contiguous(memory_format=torch.channels_last)):
        ops.proj_wgrad(z.to(torch.bfloat16), _nhwc(da0), dW)
        assert _relmax(dW, refw) < 2e-3


def test_image_side_layers_full_size(cuda_dev):
    """D.0 = Conv2d(3, 64, 4, 2, 1) + LeakyReLU on [64, 3, 256, 256] (im2col + GEMM with bias/activation epilogue),
    its weight / bias / input gradients, and G.6 = ConvTranspose2d(64, 3, 4, 2, 1) + Tanh (GEMM + col2im) with its
    tanh-backward im2col -- 1 M output rows each."""
    from rnagan_b200 import ops
    S, H, C0 = 256, 128, 64
    npix = B64 * H * H
    g = torch.Generator(device=cuda_dev).manual_seed(5)
    x = torch.rand(B64, 3, S, S, generator=g, device=cuda_dev) * 2 - 1
    W0 = _randn((C0, 3, 4, 4), 6, cuda_dev, 48 ** -0.5)
    b0 = torch.randn(C0, generator=g, device=cuda_dev) * 0.1
    col = torch.empty(npix, 64, dtype=torch.bfloat16, device=cuda_dev)
    ops.im2col_img(x, col)
    h0 = ops.gemm_nt(col, ops.pack_edge(W0), col_shift=b0, slope=0.2).view(B64, H, H, C0)
    ref = F.leaky_relu(F.conv2d(_bf(x), W0, b0, stride=2, padding=1), 0.2)
    assert _relmax(_nchw(h0), ref) < 8e-3 and _rel(_nchw(h0), ref) < 4e-3
    # weight / bias gradient of D.0 from da0
    da0 = _randn((B64, C0, H, H), 7, cuda_dev)
    dcol = ops.gemm_tn(_nhwc(da0).view(npix, C0), col)
    dW0 = torch.empty(C0, 3, 4, 4, device=cuda_dev)
    ops.unpack_edge_grad(dcol, dW0, acc=0.0)
    refw = torch.nn.grad.conv2d_weight(_bf(x), (C0, 3, 4, 4), da0, stride=2, padding=1)
    assert _relmax(dW0, refw) < 2e-3
    db0, tmp = torch.zeros(C0, device=cuda_dev), torch.zeros(C0, device=cuda_dev)
    ops.col_sum(_nhwc(da0), npix, C0, tmp, db0, 0.0)
    assert _relmax(db0, da0.sum(dim=(0, 2, 3))) < 1e-3
    # input gradient of D.0 (the gradient penalty's g = d out / d x_hat) and G.6 forward share rg_col2im_img
    wT = ops.pack_edge_t(W0, torch.zeros(48, C0, dtype=torch.bfloat16, device=cuda_dev))
    colf = torch.empty(npix, 48, device=cuda_dev)
    dimg = torch.empty(B64, 3, S, S, device=cuda_dev)
    ops.conv_up_img_col(_nhwc(da0), wT, 3, colf, dimg)
    assert _relmax(dimg, F.conv_transpose2d(da0, W0, stride=2, padding=1)) < 2e-3
    bias = torch.randn(3, generator=g, device=cuda_dev) * 0.1
    Wl = _randn((C0, 3, 4, 4), 8, cuda_dev, 0.05)
    wTl = ops.pack_edge_t(Wl, torch.zeros(48, C0, dtype=torch.bfloat16, device=cuda_dev))
    hn = _randn((B64, C0, H, H), 9, cuda_dev)
    img = torch.empty(B64, 3, S, S, device=cuda_dev)
    ops.conv_up_img_col(_nhwc(hn), wTl, 3, colf, img, bias=bias, act_tanh=True)
    ref_img = torch.tanh(F.conv_transpose2d(hn, Wl, bias=bias, stride=2, padding=1))
    assert (img - ref_img).abs().max().item() < 2e-3
    # G.6 backward: d(pre-tanh) in im2col form (mode 2), bias gradient, weight gradient, input gradient
    d_img = torch.randn(B64, 3, S, S, generator=g, device=cuda_dev)
    dpre = d_img * (1 - img * img)
    ops.im2col_img(d_img, col, y=img, mode=2)
    dbias = torch.zeros(3, device=cuda_dev)
    ops.img_channel_sum(d_img, dbias, y=img, mode=2, acc=0.0)
    assert _relmax(dbias, dpre.sum(dim=(0, 2, 3))) < 1e-3
    dcol = ops.gemm_tn(_nhwc(hn).view(npix, C0), col)
    dWl = torch.empty(C0, 3, 4, 4, device=cuda_dev)
    ops.unpack_edge_grad(dcol, dWl, acc=0.0)
    # ConvTranspose2d(64->3) weight [64, 3, 4, 4]: its gradient is conv2d_weight with the roles of input/output swapped
    refwl = torch.nn.grad.conv2d_weight(_bf(dpre), (C0, 3, 4, 4), hn, stride=2, padding=1)
    assert _relmax(dWl, refwl) < 3e-3
    dh = ops.gemm_nt(col, ops.pack_edge(Wl)).view(B64, H, H, C0)
    assert _relmax(_nchw(dh), F.conv2d(_bf(dpre), Wl, stride=2, padding=1)) < 1e-2


def test_critic_head_full_size(cuda_dev):
    """disc = Conv2d(2048, 1, 4, 1, 0) + LeakyReLU on [64, 2048, 4, 4]: forward, input gradient, weight gradient."""
    from rnagan_b200 import ops
    Cn = 2048
    h5 = _randn((B64, Cn, 4, 4), 11, cuda_dev)
    Wh = _randn((1, Cn, 4, 4), 12, cuda_dev, (16 * Cn) ** -0.5).float()
    w_head = torch.empty(16 * Cn, device=cuda_dev)
    ops.pack_head(Wh, w_head)
    a6, out = torch.empty(B64, device=cuda_dev), torch.empty(B64, device=cuda_dev)
    ops.head_fwd(_nhwc(h5), w_head, B64, 16 * Cn, 0.2, a6, out)
    ref_a = F.conv2d(h5, Wh).view(B64)
    assert (a6 - ref_a).abs().max().item() < 2e-3 * ref_a.abs().max().item() + 1e-4
    assert (out - F.leaky_relu(ref_a, 0.2)).abs().max().item() < 2e-3 * ref_a.abs().max().item() + 1e-4
    da6 = torch.empty(B64, device=cuda_dev)
    dh5 = torch.empty(B64, 4, 4, Cn, dtype=torch.bfloat16, device=cuda_dev)
    ops.head_bwd_data(a6, 1.0 / B64, w_head, B64, 16 * Cn, 0.2, da6, dh5)
    ref_da6 = torch.where(ref_a > 0, torch.ones_like(ref_a), torch.full_like(ref_a, 0.2)) / B64
    assert (da6 - ref_da6).abs().max().item() < 1e-6
    ref_dh5 = ref_da6.view(B64, 1, 1, 1) * Wh
    assert _relmax(_nchw(dh5), ref_dh5) < 8e-3
    dWh = torch.zeros(1, Cn, 4, 4, device=cuda_dev)
    ops.head_wgrad(da6, _nhwc(h5), B64, 16 * Cn, Cn, dWh, 0.0)
    assert _relmax(dWh, (ref_da6.view(B64, 1, 1, 1) * h5).sum(0, keepdim=True)) < 2e-3


# ------------------------------------------------------------------------------------------------ BatchNorm passes
def _bn_inputs(M, C, seed, dev):
    g = torch.Generator(device=dev).manual_seed(seed)
    a = (torch.randn(M, C, generator=g, device=dev) * 1.3 + 0.2).to(torch.bfloat16)
    af = a.double()
    mean = af.mean(0)
    var = af.var(0, unbiased=False)
    rstd = (var + 1e-5).rsqrt()
    gamma = 1 + 0.2 * torch.randn(C, generator=g, device=dev, dtype=torch.float64)
    beta = 0.1 * torch.randn(C, generator=g, device=dev, dtype=torch.float64)
    scale = gamma * rstd
    shift = beta - mean * scale
    f = lambda t: t.float().contiguous()
    return a, af, g, (mean, rstd, gamma, beta, scale, shift), tuple(map(f, (mean, rstd, gamma, beta, scale, shift)))


@pytest.mark.parametrize("M,C", [(B64 * 128 * 128, 64), (B64 * 64 * 64, 128), (B64 * 16, 2048)])
def test_bn_backward_full_size(cuda_dev, M, C):
    """rg_bn_act / rg_bn_bwd_reduce / rg_bn_bwd_apply / rg_bn_param_grads at the largest generator BatchNorm
    (M = 64*128^2 rows, C = 64), the largest critic one and the 2048-channel 4x4 one vs float64 formulas."""
    from rnagan_b200 import ops
    slope = 0.2
    a, af, g, (mean, rstd, gamma, beta, scale, shift), (meanf, rstdf, gammaf, betaf, scalef, shiftf) = \
        _bn_inputs(M, C, 21 + C, cuda_dev)
    h = torch.empty(M, C, dtype=torch.bfloat16, device=cuda_dev)
    ops.bn_act(a, scalef, shiftf, slope, h, M, C)
    u = af * scale + shift
    assert _relmax(h, F.leaky_relu(u, slope)) < 8e-3
    dh = (torch.randn(M, C, generator=g, device=cuda_dev)).to(torch.bfloat16)
    mask = torch.where(u > 0, 1.0, slope)
    du = dh.double() * mask
    xhat = (af - mean) * rstd
    s1, s2 = du.sum(0), (du * xhat).sum(0)
    sums = torch.zeros(2, C, device=cuda_dev)
    ops.bn_bwd_reduce(dh, a, meanf, rstdf, scalef, shiftf, slope, M, C, sums)
    tol = 2e-3 * (du.abs().sum(0).max().item())
    assert (sums[0].double() - s1).abs().max().item() <= tol and (sums[1].double() - s2).abs().max().item() <= tol
    da = torch.empty(M, C, dtype=torch.bfloat16, device=cuda_dev)
    ops.bn_bwd_apply(dh, a, None, meanf, rstdf, scalef, shiftf, slope, sums, M, C, da, None)
    ref_da = scale * (du - s1 / M - xhat * s2 / M)
    assert _relmax(da, ref_da) < 8e-3 and _rel(da, ref_da) < 4e-3
    dgam, dbet = torch.zeros(C, device=cuda_dev), torch.zeros(C, device=cuda_dev)
    ops.bn_param_grads(sums, dgam, dbet, C, 0.0, 0.0)
    assert (dgam.double() - s2).abs().max().item() <= tol and (dbet.double() - s1).abs().max().item() <= tol


@pytest.mark.parametrize("M,C", [(B64 * 64 * 64, 128), (B64 * 16, 2048)])
def test_bn_gradient_penalty_passes_full_size(cuda_dev, M, C):
    """rg_bn_gp_reduce / rg_bn_gp_apply (BatchNorm double backward inside the gradient penalty) at the critic's
    largest and widest BatchNorm layers vs the float64 formulas of tests/test_gp_math_cpu.py (which are checked
    against autograd's double backward there)."""
    from rnagan_b200 import ops
    slope = 0.2
    a, af, g, (mean, rstd, gamma, beta, scale, shift), (meanf, rstdf, gammaf, betaf, scalef, shiftf) = \
        _bn_inputs(M, C, 31 + C, cuda_dev)
    u = af * scale + shift
    mask = torch.where(u > 0, 1.0, slope)
    xhat = (af - mean) * rstd
    gO16 = (torch.randn(M, C, generator=g, device=cuda_dev) * mask.float()).to(torch.bfloat16)     # du of step 2
    ggI16 = torch.randn(M, C, generator=g, device=cuda_dev).to(torch.bfloat16)
    gO, ggI = gO16.double(), ggI16.double()
    s1, s2 = gO.sum(0), (gO * xhat).sum(0)
    q1, q2, q3 = ggI.sum(0), (ggI * xhat).sum(0), (ggI * gO).sum(0)
    s = torch.stack([s1, s2]).float().contiguous()
    q = torch.zeros(3, C, device=cuda_dev)
    ops.bn_gp_reduce(ggI16, a, gO16, meanf, rstdf, M, C, q)
    tol = 2e-3 * ggI.abs().sum(0).max().item()
    for got, ref in zip(q, (q1, q2, q3)):
        assert (got.double() - ref).abs().max().item() <= tol
    A_dh = torch.empty(M, C, dtype=torch.bfloat16, device=cuda_dev)
    A_a = torch.empty(M, C, dtype=torch.bfloat16, device=cuda_dev)
    dgam = torch.zeros(C, device=cuda_dev)
    qf = torch.stack([q1, q2, q3]).float().contiguous()
    ops.bn_gp_apply(ggI16, a, gO16, meanf, rstdf, gammaf, scalef, shiftf, slope, s, qf, M, C, A_dh, A_a, dgam, 0.0)
    r = rstd
    ref_Adh = gamma * r * (ggI - q1 / M - xhat * q2 / M) * mask
    ref_Aa = gamma * r * r / M * (xhat * (q1 * s1 / M - q3 + 3 * s2 * q2 / M) + q2 * (s1 / M - gO) + s2 * (q1 / M - ggI))
    ref_dg = r * (q3 - (q1 * s1 + q2 * s2) / M)
    assert _relmax(A_dh, ref_Adh) < 8e-3 and _rel(A_dh, ref_Adh) < 4e-3
    assert _relmax(A_a, ref_Aa) < 1e-2 and _rel(A_a, ref_Aa) < 5e-3
    assert (dgam.double() - ref_dg).abs().max().item() <= 2e-3 * ref_dg.abs().max().item() + 1e-6


# ------------------------------------------------------------------------------------------------ whole iteration
@contextlib.contextmanager
def _oracle_on(dev):
    """Run the oracle's step functions with its networks on `dev`: the CPU uniform(-0.3, 0.3) draw keeps consuming the
    global CPU generator exactly as in the reference and is then moved to the device."""
    orig = O.draw_noise
    O.draw_noise = lambda b, d: orig(b, d).to(dev)
    try:
        yield
    finally:
        O.draw_noise = orig


def _build_job(size, feats, dev):
    from rnagan_b200 import dcgan, wgan_loss
    from rnagan_b200.trainer import Trainer
    lrelu, tanh = torch.nn.LeakyReLU(0.2), torch.nn.Tanh()
    oG = O.OracleGenerator(2048, size, 3, 64, nonlinearity=lrelu, last_nonlinearity=tanh).train()
    oD = O.OracleCritic(size, 3, 64, nonlinearity=lrelu, last_nonlinearity=lrelu).train()
    oV = O.OracleVAE(feats, beta=0.005).eval()
    O.reinit_(oV, 13); O.reinit_(oG, 11); O.reinit_(oD, 12)
    ckpt = os.path.join(tempfile.mkdtemp(), "vae.pt")
    torch.save(oV.state_dict(), ckpt)
    net = {
        "generator": {"name": dcgan.DCGANGenerator,
                      "args": {"encoding_dims": 2048, "out_channels": 3, "step_channels": 64, "out_size": size,
                               "nonlinearity": torch.nn.LeakyReLU(0.2), "last_nonlinearity": torch.nn.Tanh()},
                      "optimizer": {"name": Adam, "args": {"lr": 0.0001, "betas": (0.5, 0.999)}}},
        "discriminator": {"name": dcgan.DCGANDiscriminator,
                          "args": {"in_size": size, "in_channels": 3, "step_channels": 64,
                                   "nonlinearity": torch.nn.LeakyReLU(0.2),
                                   "last_nonlinearity": torch.nn.LeakyReLU(0.2)},
                          "optimizer": {"name": Adam, "args": {"lr": 0.0004, "betas": (0.5, 0.999)}}},
    }
    losses = [wgan_loss.WassersteinGeneratorLossVAE(ckpt, feats), wgan_loss.WassersteinDiscriminatorLossVAE(ckpt, feats),
              wgan_loss.WassersteinGradientPenaltyVAE(ckpt, feats)]
    os.unlink(ckpt)
    tr = Trainer(net, losses, device=dev, sample_size=64, epochs=1, devices=[0])
    tr.generator.train(); tr.discriminator.train()
    return oG.to(dev), oD.to(dev), oV.to(dev), tr


def test_config2_iteration_b64_matches_fp32_oracle(cuda_dev):
    """BASELINE config 2 exactly (lung shapes, per-GPU batch 64): each of the three train_ops started from the
    oracle's state; losses, every parameter gradient (cosine, judged against torch's own bf16-autocast deviation on
    the same step), post-step weights, BatchNorm buffers and the CPU RNG stream."""
    size, feats, B = 256, 19198, B64
    dev = cuda_dev
    oG, oD, oV, tr = _build_job(size, feats, dev)
    data_cpu = O.make_batch(B, feats, size, 14)
    data_dev = {k: v.to(dev) for k, v in data_cpu.items()}
    tr.real_inputs, tr.batch_size = data_cpu, B
    og = Adam(oG.parameters(), lr=1e-4, betas=(0.5, 0.999))
    od = Adam(oD.parameters(), lr=4e-4, betas=(0.5, 0.999))
    names = list(tr.losses.keys())
    steps = [(0, O.g_step, og, oG, "generator"), (1, O.critic_step, od, oD, "discriminator"),
             (2, O.gp_step, od, oD, "discriminator")]
    report = []
    torch.manual_seed(99)
    with _oracle_on(dev):
        for which, fn, opt, onet, mname in steps:
            # torch's own bf16 deviation on this step (yardstick), on copies, RNG restored
            st = torch.get_rng_state()
            cG, cD = copy.deepcopy(oG), copy.deepcopy(oD)
            copt = Adam((cG if which == 0 else cD).parameters(), lr=1e-4, betas=(0.5, 0.999))
            with torch.autocast("cuda", dtype=torch.bfloat16):
                fn(cG, cD, copt, oV, data_dev)
            anet = cG if which == 0 else cD
            torch.set_rng_state(st)
            # same starting state on both sides
            tr.generator.load_state_dict(oG.state_dict())
            tr.discriminator.load_state_dict(oD.state_dict())
            if len(og.state_dict()["state"]):
                tr.optimizer_generator.load_state_dict(og.state_dict())
            if len(od.state_dict()["state"]):
                tr.optimizer_discriminator.load_state_dict(od.state_dict())
            v_ref = fn(oG, oD, opt, oV, data_dev)
            st_after = torch.get_rng_state()
            torch.set_rng_state(st)
            v = tr._call(names[which])
            v = float(v.item()) if torch.is_tensor(v) else float(v)
            assert torch.equal(torch.get_rng_state(), st_after)
            assert abs(v - v_ref) <= 0.02 + 0.02 * abs(v_ref), f"step{which}: {v} vs oracle {v_ref}"
            mnet = getattr(tr, mname)
            c_bf16, pairs = 1.0, []
            for (n, po), (_, pm), (_, pa) in zip(onet.named_parameters(), mnet.named_parameters(),
                                                 anet.named_parameters()):
                if po.grad is None or po.grad.norm() == 0:
                    assert pm.grad is None or pm.grad.abs().max().item() == 0.0, f"step{which} {n}: expected zero grad"
                    continue
                c_bf16 = min(c_bf16, _cos(pa.grad, po.grad))
                pairs.append((n, _cos(pm.grad, po.grad)))
            bound = 1.0 - 2.0 * (1.0 - c_bf16) - 0.01
            worst = min(pairs, key=lambda t: t[1])
            report.append((which, v, v_ref, worst, c_bf16))
            for n, c in pairs:
                assert c >= bound, f"step{which} {n}: cosine {c:.4f} < {bound:.4f} (torch-bf16 worst {c_bf16:.4f})"
            for (n, po), (_, pm) in zip(onet.named_parameters(), mnet.named_parameters()):
                assert _rel(pm.data, po.data) <= 2e-2, f"step{which} weight {n}"
            for (n, bo), (_, bm) in zip(onet.named_buffers(), mnet.named_buffers()):
                if n.endswith("num_batches_tracked"):
                    assert int(bo) == int(bm), f"step{which} {n}"
                else:
                    assert _rel(bm.float(), bo.float()) <= 5e-2, f"step{which} buffer {n}"
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):          # measured cosines, quoted in DESIGN.md section 4
        with open(os.path.join(out_dir, "fullsize_parity_b64.txt"), "w") as f:
            for which, v, v_ref, worst, c_bf16 in report:
                f.write(f"step{which}: loss {v:.6f} oracle {v_ref:.6f}; worst gradient cosine {worst[1]:.5f} "
                        f"({worst[0]}); torch-bf16 worst cosine {c_bf16:.5f}\n")


def test_synthesis_full_size_matches_oracle(cuda_dev):
    """generate_images semantics (src/gan_utils.py:197-244) at 256x256 / 19198 genes: one profile, chunks of 10,
    train-mode BN; and one profile per row at chunk 32 through generate_tiles."""
    from rnagan_b200 import gan_utils
    size, feats = 256, 19198
    oG, oD, oV, tr = _build_job(size, feats, cuda_dev)
    tr.generator.load_state_dict(oG.state_dict())
    vae = tr.losses["WassersteinGeneratorLossVAE"]._encoder(cuda_dev)
    prof = torch.randn(1, feats, generator=torch.Generator().manual_seed(15))
    n = 24
    with _oracle_on(cuda_dev):
        torch.manual_seed(5)
        ref = O.synth_tiles(oG, oV, prof.to(cuda_dev), n)
    torch.manual_seed(5)
    got = gan_utils.generate_images(tr, gene_exp=prof, sample_size=n, betavae=vae)
    assert got.shape == (n, size, size, 3)
    d = got.astype("float64") - ref.astype("float64")
    assert (d ** 2).sum() ** 0.5 / (ref.astype("float64") ** 2).sum() ** 0.5 <= 3e-2 and abs(d).max() <= 0.12


def test_up_generator_forward_full_size(cuda_dev):
    """DCGANUpGenerator (src/dcgan.py:8-99) at 256x256, train-mode BN, vs the oracle in fp32 on the GPU."""
    from rnagan_b200 import dcgan
    oU = O.OracleUpGenerator(2048, 256, 3, 64, nonlinearity=torch.nn.LeakyReLU(0.2), last_nonlinearity=torch.nn.Tanh())
    O.reinit_(oU, 21)
    oU = oU.to(cuda_dev).train()
    G = dcgan.DCGANUpGenerator(2048, 256, 3, 64, nonlinearity=torch.nn.LeakyReLU(0.2),
                               last_nonlinearity=torch.nn.Tanh()).to(cuda_dev)
    G.load_state_dict(oU.state_dict())
    G.train()
    z = torch.randn(8, 2048, generator=torch.Generator().manual_seed(22)).to(cuda_dev)
    with torch.no_grad():
        ref = oU(z)
    out = G(z)
    assert out.shape == ref.shape == (8, 3, 256, 256)
    assert _rel(out, ref) <= 3e-2
    for (n, bo), (_, bm) in zip(oU.named_buffers(), G.named_buffers()):
        if n.endswith("num_batches_tracked"):
            assert int(bo) == int(bm) == 1
        else:
            assert _rel(bm.float(), bo.float()) <= 2e-2, n


# ------------------------------------------------------------------------------------------------ betaVAE, 19198 genes
def test_vae_encode_and_train_step_full_size(cuda_dev):
    """betaVAE over all 19198 protein-coding genes (betavae_tissues.json shapes, batch 128): eval-mode encode (the
    GAN's conditioning path, ragged K = 19198 inside `encode`) and one training step (config 5) vs the oracle."""
    from rnagan_b200 import betaVAE as bv
    feats, B, beta = 19198, 128, 0.0005
    oV = O.OracleVAE(feats, beta=beta)
    O.reinit_(oV, 23)
    vae = bv.betaVAE(feats, 2048, [6000, 4000, 2048], [4000, 6000], beta=beta)
    vae.load_state_dict(oV.state_dict())
    oV = oV.to(cuda_dev)
    vae = vae.to(cuda_dev)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, feats, generator=g).to(cuda_dev)
    keep = (torch.rand(B, feats, generator=g) >= 0.5).float().to(cuda_dev)
    eps = torch.randn(B, 2048, generator=g).to(cuda_dev)
    oV.eval(); vae.eval()
    with torch.no_grad():
        zm, zl, h = oV.encode(x)
    gm, gl, gh = vae.encode(x)
    assert _rel(gm, zm) <= 2e-2 and _rel(gl, zl) <= 2e-2 and _rel(gh, h) <= 2e-2
    oV.train(); vae.train()
    oo = torch.optim.Adam(oV.parameters(), lr=5e-5)
    om = torch.optim.Adam(vae.parameters(), lr=5e-5)
    ref = O.vae_train_step_explicit(oV, oo, x, beta, keep, eps)
    out3 = bv.train_step(vae, om, x, beta, keep_mask=keep, eps=eps).cpu()
    for got, key in zip(out3.tolist(), ("total_loss", "reconstruction_loss", "kl_loss")):
        assert abs(got - ref[key]) <= 2e-2 * abs(ref[key]) + 1e-4, key
    for (n, po), (_, pm) in zip(oV.named_parameters(), vae.named_parameters()):
        a, b = pm.grad.double().flatten(), po.grad.double().flatten()
        if n.endswith(".0.bias") and not n.startswith("decoder.2"):
            assert b.norm().item() <= 1e-3 and a.norm().item() <= 1e-2, n      # exact gradient is 0 (BN follows)
            continue
        if b.norm() < 1e-12:
            continue
        assert _cos(a, b) >= 0.98, f"{n}: cosine {_cos(a, b):.4f}"
        assert _rel(pm.data, po.data) <= 1e-2, n
    for (n, bo), (_, bm) in zip(oV.named_buffers(), vae.named_buffers()):
        if n.endswith("num_batches_tracked"):
            assert int(bo) == int(bm)
        else:
            assert _rel(bm.float(), bo.float()) <= 2e-2, n
