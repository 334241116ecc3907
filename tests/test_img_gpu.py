"""GPU parity of the fused image-side convolutions (csrc/rg_img.cu: rg_img_conv_up / rg_img_conv_down /
rg_img_conv_wgrad) against fp32 torch ops on the same bf16-rounded operands: tiles smaller than the CTA tile (8x8,
16x16), several tiles per image, 1 / 3 / 4 image channels, every prologue mode (plain * scalar, gradient-penalty
interpolate, tanh backward), every output format (fp32 NCHW, fp32 NHWC unit range, uint8 NHWC, BGR), the fused
LeakyReLU-backward mask and the fused bias gradient; and the BASELINE config-2 size (B = 64, 256x256)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_reference_math():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _bf(t):
    return t.to(torch.bfloat16).float()


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _nchw(x):
    return x.float().permute(0, 3, 1, 2)


def _relmax(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


CASES = [(3, 16, 3), (2, 32, 3), (5, 64, 3), (2, 128, 1), (2, 64, 4), (64, 256, 3)]


@pytest.mark.parametrize("B,S,Cimg", CASES)
def test_img_conv_up(cuda_dev, B, S, Cimg):
    from rnagan_b200 import ops
    H = S // 2
    g = torch.Generator(device=cuda_dev).manual_seed(B + S + Cimg)
    lo = _bf(torch.randn(B, 64, H, H, generator=g, device=cuda_dev))
    Wt = _bf(torch.randn(64, Cimg, 4, 4, generator=g, device=cuda_dev) * 0.05)
    bias = torch.randn(Cimg, generator=g, device=cuda_dev) * 0.1
    lon = _nhwc(lo)
    ref = F.conv_transpose2d(lo, Wt, stride=2, padding=1)
    out = torch.full((B, Cimg, S, S), float("nan"), device=cuda_dev)
    ops.img_conv_up(lon, Wt, out)
    assert _relmax(out, ref) < 2e-3
    ref_t = torch.tanh(ref + bias.view(1, -1, 1, 1))
    ops.img_conv_up(lon, Wt, out, bias=bias, act_tanh=True)
    assert (out - ref_t).abs().max().item() < 2e-3
    unit = torch.full((B, S, S, Cimg), float("nan"), device=cuda_dev)
    ops.img_conv_up(lon, Wt, unit, bias=bias, act_tanh=True, unit_nhwc=True)
    assert torch.equal(unit, ((out + 1.0) * 0.5).permute(0, 2, 3, 1).contiguous())
    for bgr in (False, True):
        u8 = torch.zeros(B, S, S, Cimg, dtype=torch.uint8, device=cuda_dev)
        ops.img_conv_up(lon, Wt, u8, bias=bias, act_tanh=True, u8=True, bgr=bgr)
        want = (unit.cpu().numpy() * np.float32(255)).astype(np.uint8)
        assert np.array_equal(u8.cpu().numpy(), want[..., ::-1] if bgr else want)


@pytest.mark.parametrize("B,S,Cimg", CASES)
def test_img_conv_down(cuda_dev, B, S, Cimg):
    from rnagan_b200 import ops
    H = S // 2
    g = torch.Generator(device=cuda_dev).manual_seed(B + S + Cimg + 1)
    x = torch.rand(B, Cimg, S, S, generator=g, device=cuda_dev) * 2 - 1
    y = torch.tanh(torch.randn(B, Cimg, S, S, generator=g, device=cuda_dev))
    Wt = _bf(torch.randn(64, Cimg, 4, 4, generator=g, device=cuda_dev) * 0.2)
    bias = torch.randn(64, generator=g, device=cuda_dev) * 0.1
    out = torch.full((B, H, H, 64), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    # critic layer 0: conv + bias + LeakyReLU on the bf16-rounded image
    ops.img_conv_down(x, Wt, out, bias=bias, slope=0.2)
    ref = F.leaky_relu(F.conv2d(_bf(x), Wt, bias, stride=2, padding=1), 0.2)
    assert _relmax(_nchw(out), ref) < 8e-3
    # gradient-penalty interpolate (src/wgan_loss.py:376-380)
    eps = torch.tensor([0.3], device=cuda_dev)
    ops.img_conv_down(x, Wt, out, y=y, mode=1, eps_dev=eps, bias=bias, slope=0.2)
    ref = F.leaky_relu(F.conv2d(_bf(0.3 * x + 0.7 * y), Wt, bias, stride=2, padding=1), 0.2)
    assert _relmax(_nchw(out), ref) < 8e-3
    # tanh backward (generator output layer's input gradient), no bias / activation
    ops.img_conv_down(x, Wt, out, y=y, mode=2)
    ref = F.conv2d(_bf(x * (1 - y * y)), Wt, stride=2, padding=1)
    assert _relmax(_nchw(out), ref) < 8e-3
    # scalar multiplier + LeakyReLU-backward mask (the gradient penalty's adjoint sweep through layer 0)
    mul = torch.tensor([1.7], device=cuda_dev)
    h0 = torch.randn(B, H, H, 64, generator=g, device=cuda_dev).to(torch.bfloat16)
    ops.img_conv_down(x, Wt, out, mul_dev=mul, mask_src=h0, mask_slope=0.2)
    pre = _bf(F.conv2d(_bf(x * 1.7), Wt, stride=2, padding=1))
    ref = pre * torch.where(_nchw(h0) > 0, 1.0, 0.2)
    assert _relmax(_nchw(out), ref) < 1e-2


@pytest.mark.parametrize("B,S,Cimg", CASES)
def test_img_conv_wgrad(cuda_dev, B, S, Cimg):
    from rnagan_b200 import ops
    H = S // 2
    g = torch.Generator(device=cuda_dev).manual_seed(B + S + Cimg + 2)
    x = torch.rand(B, Cimg, S, S, generator=g, device=cuda_dev) * 2 - 1
    y = torch.tanh(torch.randn(B, Cimg, S, S, generator=g, device=cuda_dev))
    act = _bf(torch.randn(B, 64, H, H, generator=g, device=cuda_dev))
    actn = _nhwc(act)
    dW = torch.full((64, Cimg, 4, 4), float("nan"), device=cuda_dev)
    db = torch.full((64,), float("nan"), device=cuda_dev) if Cimg <= 3 else None
    ops.img_conv_wgrad(actn, x, dW, dbias=db)
    ref = torch.nn.grad.conv2d_weight(_bf(x), (64, Cimg, 4, 4), act, stride=2, padding=1)
    assert _relmax(dW, ref) < 2e-3
    if db is not None:
        assert _relmax(db, act.sum(dim=(0, 2, 3))) < 2e-3
    # accumulate (acc = 1) with the interpolated image, bias gradient accumulated too
    eps = torch.tensor([0.25], device=cuda_dev)
    ops.img_conv_wgrad(actn, x, dW, y=y, mode=1, eps_dev=eps, acc=1.0, dbias=db, acc_bias=1.0)
    ref2 = torch.nn.grad.conv2d_weight(_bf(0.25 * x + 0.75 * y), (64, Cimg, 4, 4), act, stride=2, padding=1)
    assert _relmax(dW, ref + ref2) < 2e-3
    if db is not None:
        assert _relmax(db, 2 * act.sum(dim=(0, 2, 3))) < 2e-3
    # tanh backward x scalar: the ConvTranspose2d weight gradient of the generator's output layer
    mul = torch.tensor([0.5], device=cuda_dev)
    ops.img_conv_wgrad(actn, x, dW, y=y, mode=2, mul_dev=mul)
    ref3 = torch.nn.grad.conv2d_weight(_bf(0.5 * x * (1 - y * y)), (64, Cimg, 4, 4), act, stride=2, padding=1)
    assert _relmax(dW, ref3) < 2e-3
    # fixed summation order: bit-reproducible
    dW2 = torch.empty_like(dW)
    ops.img_conv_wgrad(actn, x, dW2, y=y, mode=2, mul_dev=mul)
    assert torch.equal(dW, dW2)


@pytest.mark.parametrize("B,S", [(3, 32), (2, 128)])
def test_img_conv_up_fused_batchnorm_is_bit_identical(cuda_dev, B, S):
    """bn=(scale, shift, slope): the output kernel applies the last hidden block's BatchNorm + LeakyReLU to the staged tile
    with rg_bn_act's arithmetic and rounding, so it equals rg_bn_act followed by the plain kernel BIT for bit (including
    the zero padding outside the image), and a generator forward with keep=False equals the keep=True one."""
    from rnagan_b200 import dcgan, ops
    H = S // 2
    g = torch.Generator(device=cuda_dev).manual_seed(B + S)
    a = torch.randn(B, H, H, 64, generator=g, device=cuda_dev).to(torch.bfloat16)
    scale = (torch.rand(64, generator=g, device=cuda_dev) + 0.5).contiguous()
    shift = (torch.randn(64, generator=g, device=cuda_dev) * 0.3).contiguous()
    Wt = _bf(torch.randn(64, 3, 4, 4, generator=g, device=cuda_dev) * 0.05)
    bias = torch.randn(3, generator=g, device=cuda_dev) * 0.1
    h = torch.empty_like(a)
    ops.bn_act(a, scale, shift, 0.2, h, B * H * H, 64)
    ref = torch.empty(B, 3, S, S, device=cuda_dev)
    out = torch.empty(B, 3, S, S, device=cuda_dev)
    ops.img_conv_up(h, Wt, ref, bias=bias, act_tanh=True)
    ops.img_conv_up(a, Wt, out, bias=bias, act_tanh=True, bn=(scale, shift, 0.2))
    assert torch.equal(out, ref)
    G = dcgan.DCGANGenerator(2048, S, 3, 64, nonlinearity=torch.nn.LeakyReLU(0.2),
                             last_nonlinearity=torch.nn.Tanh()).to(cuda_dev).train()
    eng = G._engine()
    lat = torch.randn(4, 2048, generator=g, device=cuda_dev).to(torch.bfloat16)
    img_keep = eng.forward(lat, tag="t1", training=True, keep=True).clone()
    img_fused = eng.forward(lat, tag="t2", training=True, keep=False).clone()
    assert torch.equal(img_keep, img_fused)
