"""GPU parity of the tcgen05 tile engine (C ABI: rg_conv_down / rg_conv_up / rg_conv_up_img / rg_conv_wgrad /
rg_proj_wgrad / rg_gemm_nt / rg_gemm_tn) against fp32 torch ops evaluated on the SAME bf16-rounded operands, so the
only difference is accumulation order: tolerance 2e-3 of the output scale for fp32 outputs, bf16 rounding
(2^-8 relative) on top of that for bf16 outputs.
"""
import os
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = a.float().cpu()
    b = b.float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _bf(t):
    return t.to(torch.bfloat16).float()


def _nhwc(x):   # NCHW fp32 -> NHWC bf16
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _nchw(x):   # NHWC -> NCHW fp32
    return x.float().permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (256, 128, 64), (64, 64, 64), (16, 2048, 2048),
                                   (1000, 200, 192), (384, 6000, 640)])
def test_gemm_nt(cuda_dev, M, N, K):
    from rnagan_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = _bf(torch.randn(M, K, generator=g)).to(cuda_dev)
    Bw = _bf(torch.randn(N, K, generator=g)).to(cuda_dev)
    ref = A @ Bw.t()
    out = ops.gemm_nt(A.to(torch.bfloat16), Bw.to(torch.bfloat16), out_f32=True)
    assert _rel(out, ref) < 2e-3
    out16 = ops.gemm_nt(A.to(torch.bfloat16), Bw.to(torch.bfloat16))
    assert _rel(out16, ref) < 8e-3


def test_gemm_nt_epilogue(cuda_dev):
    from rnagan_b200 import ops
    M, N, K = 64, 6000, 19200
    g = torch.Generator(device="cpu").manual_seed(5)
    A = _bf(torch.randn(M, K, generator=g)).to(cuda_dev)
    Bw = _bf(torch.randn(N, K, generator=g) * 0.01).to(cuda_dev)
    sc = (torch.rand(N, generator=g) + 0.5).to(cuda_dev)
    sh = torch.randn(N, generator=g).to(cuda_dev)
    ref = F.leaky_relu((A @ Bw.t()) * sc + sh, 0.01)
    out = ops.gemm_nt(A.to(torch.bfloat16), Bw.to(torch.bfloat16), col_scale=sc, col_shift=sh, slope=0.01,
                      out_f32=True)
    assert _rel(out, ref) < 2e-3


@pytest.mark.parametrize("B,H,W,Cs,Cp", [(16, 16, 16, 256, 512), (8, 4, 4, 1024, 2048), (4, 64, 64, 64, 128),
                                         (2, 32, 32, 128, 256), (16, 8, 8, 128, 64)])
def test_conv_down(cuda_dev, B, H, W, Cs, Cp):
    """lo = conv2d(hi, W[Cp,Cs,4,4], stride 2, pad 1) -- critic forward / generator dgrad."""
    from rnagan_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(B + H + Cs + Cp)
    x = _bf(torch.randn(B, Cs, 2 * H, 2 * W, generator=g)).to(cuda_dev)
    Wt = _bf(torch.randn(Cp, Cs, 4, 4, generator=g) * 0.05).to(cuda_dev)
    ref = F.conv2d(x, Wt, stride=2, padding=1)
    w_down, _ = ops.pack_link(Wt, want_up=False)
    out = ops.conv_down(_nhwc(x), w_down)
    assert out.shape == (B, H, W, Cp)
    assert _rel(_nchw(out), ref) < 8e-3


@pytest.mark.parametrize("B,H,W,Cp,Cs", [(16, 16, 16, 512, 256), (8, 4, 4, 2048, 1024), (4, 64, 64, 128, 64),
                                         (2, 32, 32, 256, 128), (16, 8, 8, 64, 128)])
def test_conv_up(cuda_dev, B, H, W, Cp, Cs):
    """hi = conv_transpose2d(lo, W[Cp,Cs,4,4], stride 2, pad 1) -- generator forward / critic dgrad."""
    from rnagan_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(B + H + Cs + Cp + 1)
    x = _bf(torch.randn(B, Cp, H, W, generator=g)).to(cuda_dev)
    Wt = _bf(torch.randn(Cp, Cs, 4, 4, generator=g) * 0.05).to(cuda_dev)
    ref = F.conv_transpose2d(x, Wt, stride=2, padding=1)
    w_down, w_up = ops.pack_link(Wt)
    for w in (w_down, w_up):      # MN-major read of w_down, and the K-major w_up copy
        out = ops.conv_up(_nhwc(x), w, Cs)
        assert out.shape == (B, 2 * H, 2 * W, Cs)
        assert _rel(_nchw(out), ref) < 8e-3


@pytest.mark.parametrize("B,H,W,Cp", [(4, 32, 32, 64), (2, 128, 128, 64)])
def test_conv_up_img(cuda_dev, B, H, W, Cp):
    from rnagan_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(B + H)
    x = _bf(torch.randn(B, Cp, H, W, generator=g)).to(cuda_dev)
    Wt = _bf(torch.randn(Cp, 3, 4, 4, generator=g) * 0.05).to(cuda_dev)
    bias = torch.randn(3, generator=g).to(cuda_dev)
    ref = torch.tanh(F.conv_transpose2d(x, Wt, bias=bias, stride=2, padding=1))
    _, w_up = ops.pack_link(Wt, want_down=False)
    out = ops.conv_up_img(_nhwc(x), w_up, 3, bias=bias, act_tanh=True)
    assert (out - ref).abs().max().item() < 2e-3
    ref2 = F.conv_transpose2d(x, Wt, stride=2, padding=1)
    out2 = ops.conv_up_img(_nhwc(x), w_up, 3)
    assert _rel(out2, ref2) < 2e-3


@pytest.mark.parametrize("B,H,W,Cs,Cp", [(16, 16, 16, 256, 512), (8, 4, 4, 1024, 2048), (4, 64, 64, 64, 128),
                                         (2, 32, 32, 128, 256), (16, 8, 8, 128, 64)])
def test_conv_wgrad(cuda_dev, B, H, W, Cs, Cp):
    """dW[p,s,kh,kw] = sum lo[b,i,j,p] hi[b,2i-1+kh,2j-1+kw,s] == conv2d weight gradient."""
    from rnagan_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(B + H + Cs + Cp + 2)
    hi = _bf(torch.randn(B, Cs, 2 * H, 2 * W, generator=g)).to(cuda_dev)
    lo = _bf(torch.randn(B, Cp, H, W, generator=g)).to(cuda_dev)
    ref = torch.nn.grad.conv2d_weight(hi, (Cp, Cs, 4, 4), lo, stride=2, padding=1)
    dW = torch.empty(Cp, Cs, 4, 4, device=cuda_dev)
    ops.conv_wgrad(_nhwc(lo), _nhwc(hi), dW)
    assert _rel(dW, ref) < 2e-3
    # accumulate + scale
    scale = torch.tensor([0.5], device=cuda_dev)
    ops.conv_wgrad(_nhwc(lo), _nhwc(hi), dW, alpha=2.0, alpha_dev=scale, beta=1.0)
    assert _rel(dW, 2 * ref) < 2e-3
    # native (channels_last) gradient layout: what the engines use -- direct epilogue or native split-K reduce
    dWn = torch.empty(Cp, Cs, 4, 4, device=cuda_dev).contiguous(memory_format=torch.channels_last)
    assert ops.is_native4(dWn)
    ops.conv_wgrad(_nhwc(lo), _nhwc(hi), dWn)
    assert _rel(dWn, ref) < 2e-3
    ops.conv_wgrad(_nhwc(lo), _nhwc(hi), dWn, alpha=2.0, alpha_dev=scale, beta=1.0)
    assert _rel(dWn, 2 * ref) < 2e-3


def test_proj_wgrad_and_fwd(cuda_dev):
    """generator layer 0: ConvTranspose2d(E, C0, 4, 1, 0) on a 1x1 input as a plain GEMM, and its weight gradient."""
    from rnagan_b200 import ops
    B, E, C0 = 16, 256, 128
    g = torch.Generator(device="cpu").manual_seed(11)
    z = _bf(torch.randn(B, E, generator=g)).to(cuda_dev)
    Wt = _bf(torch.randn(E, C0, 4, 4, generator=g) * 0.05).to(cuda_dev)
    ref = F.conv_transpose2d(z.view(B, E, 1, 1), Wt)           # [B, C0, 4, 4]
    wp = ops.pack_proj(Wt)
    out = ops.gemm_nt(z.to(torch.bfloat16), wp, out_f32=True).view(B, 4, 4, C0)
    assert _rel(_nchw(out), ref) < 2e-3
    da0 = _bf(torch.randn(B, C0, 4, 4, generator=g)).to(cuda_dev)
    refw = torch.einsum("be,bchw->echw", z, da0)
    dW = torch.empty(E, C0, 4, 4, device=cuda_dev)
    ops.proj_wgrad(z.to(torch.bfloat16), _nhwc(da0), dW)
    assert _rel(dW, refw) < 2e-3


@pytest.mark.parametrize("R,M,N", [(4096, 64, 64), (100000, 64, 64), (64, 128, 256), (777, 256, 128),
                                   (128, 1024, 1998), (128, 1024, 199), (128, 512, 2048)])
def test_gemm_tn(cuda_dev, R, M, N):
    """C = A^T B (nn.Linear weight gradients, image-side wgrad): split-K + reduce, the direct epilogue (one unit covers
    all rows; N = 1998: ragged 8-byte-aligned rows, float2 stores; N = 199: odd, falls back to the reduce path), and
    alpha / beta accumulation in place."""
    from rnagan_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(R + M + N)
    A = _bf(torch.randn(R, M, generator=g)).to(cuda_dev)
    Bm = _bf(torch.randn(R, N, generator=g)).to(cuda_dev)
    ref = A.double().t() @ Bm.double()
    Ab, Bb = A.to(torch.bfloat16), Bm.to(torch.bfloat16)
    if N % 8:                      # rows of the bf16 operands need a pitch that is a multiple of 8 elements
        Np = (N + 7) // 8 * 8
        Bb = torch.zeros(R, Np, dtype=torch.bfloat16, device=cuda_dev)
        Bb[:, :N] = Bm.to(torch.bfloat16)
        Bb = Bb[:, :N]
    out = ops.gemm_tn(Ab, Bb, N=N)
    assert out.shape == (M, N) and _rel(out.double(), ref) < 2e-3
    ops.gemm_tn(Ab, Bb, out=out, alpha=0.5, beta=1.0, N=N)
    assert _rel(out.double(), 1.5 * ref) < 2e-3


def test_pack_edge(cuda_dev):
    from rnagan_b200 import ops
    Wt = torch.randn(64, 3, 4, 4, device=cuda_dev)
    wc = ops.pack_edge(Wt).float().view(64, 16, 4)
    ref = _bf(Wt).permute(0, 2, 3, 1).reshape(64, 16, 3)
    assert torch.equal(wc[:, :, :3], ref)
    assert wc[:, :, 3].abs().max().item() == 0


@pytest.mark.parametrize("B,H,W,Cp", [(4, 32, 32, 64), (2, 128, 128, 64)])
def test_conv_up_img_col(cuda_dev, B, H, W, Cp):
    """image-side transposed conv as dgrad-form GEMM + col2im (rg_gemm_nt + rg_col2im_img)."""
    from rnagan_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(B + H + 7)
    x = _bf(torch.randn(B, Cp, H, W, generator=g)).to(cuda_dev)
    Wt = _bf(torch.randn(Cp, 3, 4, 4, generator=g) * 0.05).to(cuda_dev)
    bias = torch.randn(3, generator=g).to(cuda_dev)
    ref = torch.tanh(F.conv_transpose2d(x, Wt, bias=bias, stride=2, padding=1))
    wT = ops.pack_edge_t(Wt, torch.zeros(48, Cp, dtype=torch.bfloat16, device=cuda_dev))
    col = torch.empty(B * H * W, 48, device=cuda_dev)
    out = torch.empty(B, 3, 2 * H, 2 * W, device=cuda_dev)
    ops.conv_up_img_col(_nhwc(x), wT, 3, col, out, bias=bias, act_tanh=True)
    assert (out - ref).abs().max().item() < 2e-3
    ops.conv_up_img_col(_nhwc(x), wT, 3, col, out)
    assert _rel(out, F.conv_transpose2d(x, Wt, stride=2, padding=1)) < 2e-3


def test_img_channel_sum_and_im2col(cuda_dev):
    from rnagan_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn(5, 3, 64, 64, generator=g).to(cuda_dev)
    y = torch.tanh(torch.randn(5, 3, 64, 64, generator=g)).to(cuda_dev)
    out = torch.zeros(3, device=cuda_dev)
    ops.img_channel_sum(x, out, y=y, mode=2)
    ref = (x * (1 - y * y)).sum(dim=(0, 2, 3))
    assert _rel(out, ref) < 1e-4
    # im2col: col[pix][tap*4+c] equals unfold of the (eps-mixed) image
    eps = torch.tensor([0.3], device=cuda_dev)
    col = torch.empty(5 * 32 * 32, 64, dtype=torch.bfloat16, device=cuda_dev)
    ops.im2col_img(x, col, y=y, mode=1, eps_dev=eps)
    mixed = 0.3 * x + 0.7 * y
    unf = F.unfold(mixed, kernel_size=4, stride=2, padding=1)            # [B, 3*16, L] index c*16+tap
    unf = unf.view(5, 3, 16, 32 * 32).permute(0, 3, 2, 1)                   # [B, L, tap, c]
    got = col.float().view(5, 32 * 32, 16, 4)
    assert (got[..., :3] - _bf(unf)).abs().max().item() < 1e-6
    assert got[..., 3].abs().max().item() == 0


@pytest.mark.parametrize("M,N,K", [(128, 6000, 19198), (128, 2048, 4000), (32, 300, 6000), (128, 4000, 2048)])
def test_gemm_nn_and_ragged_k(cuda_dev, M, N, K):
    """C = A[M,K] @ W[K,N] with W row-major (nn.Linear input gradient), K and N not multiples of 64."""
    from rnagan_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    Kp = (K + 7) // 8 * 8
    A = torch.zeros(M, Kp)
    A[:, :K] = _bf(torch.randn(M, K, generator=g))
    Np = (N + 7) // 8 * 8
    W = torch.zeros(K, Np)
    W[:, :N] = _bf(torch.randn(K, N, generator=g) * 0.05)
    ref = A[:, :K] @ W[:, :N]
    out = ops.gemm_nn(A.to(cuda_dev).to(torch.bfloat16), W.to(cuda_dev).to(torch.bfloat16), out_f32=True, N=N, K=K)
    assert _rel(out, ref) < 2e-3
    # NT with ragged K (true K in the tensor map, zero-filled tail)
    Wt = torch.zeros(N, Kp)
    Wt[:, :K] = _bf(torch.randn(N, K, generator=g) * 0.05)
    out2 = ops.gemm_nt(A.to(cuda_dev).to(torch.bfloat16), Wt.to(cuda_dev).to(torch.bfloat16), out_f32=True, K=K)
    assert _rel(out2, A[:, :K] @ Wt[:, :K].t()) < 2e-3
    # TN with a ragged N (weight gradient [M_out, N_in] with N_in = K here)
    dY = _bf(torch.randn(M, 200, generator=g)).to(cuda_dev).to(torch.bfloat16)
    dW = ops.gemm_tn(dY, A.to(cuda_dev).to(torch.bfloat16), N=K)
    assert dW.shape == (200, K)
    assert _rel(dW, dY.float().t() @ A[:, :K].to(cuda_dev)) < 2e-3


@pytest.mark.parametrize("B,H,W,Cs,Cp", [(16, 16, 16, 256, 512), (4, 64, 64, 64, 128), (3, 8, 8, 64, 128),
                                         (8, 4, 4, 128, 64), (5, 16, 16, 128, 256)])
def test_conv_fused_stats(cuda_dev, B, H, W, Cs, Cp):
    """The per-CTA channel sums / sums of squares a convolution epilogue leaves in stats_ws (BatchNorm batch
    statistics without a second pass) equal the sums over the stored bf16 output; odd tile counts and batches that
    do not fill a tile included.  Also checks the output itself on these shapes (CTA-pair path with an odd M tail)."""
    from rnagan_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(B + H + Cs + Cp + 5)
    Wt = _bf(torch.randn(Cp, Cs, 4, 4, generator=g) * 0.05).to(cuda_dev)
    w_down, w_up = ops.pack_link(Wt)
    # down
    x = _bf(torch.randn(B, Cs, 2 * H, 2 * W, generator=g)).to(cuda_dev)
    ws = ops.stats_ws(Cp, cuda_dev, slot=7)
    ws.fill_(float("nan"))          # every row must be (re)written
    out = ops.conv_down(_nhwc(x), w_down, stats=ws)
    assert _rel(_nchw(out), F.conv2d(x, Wt, stride=2, padding=1)) < 8e-3
    o = out.float().view(-1, Cp)
    got = ws.sum(0)
    assert _rel(got[0], o.sum(0)) < 1e-4 and _rel(got[1], (o * o).sum(0)) < 1e-4
    # up (both weight layouts)
    y = _bf(torch.randn(B, Cp, H, W, generator=g)).to(cuda_dev)
    ws2 = ops.stats_ws(Cs, cuda_dev, slot=7)
    for w in (w_down, w_up):
        ws2.fill_(float("nan"))
        out = ops.conv_up(_nhwc(y), w, Cs, stats=ws2)
        assert _rel(_nchw(out), F.conv_transpose2d(y, Wt, stride=2, padding=1)) < 8e-3
        o = out.float().view(-1, Cs)
        got = ws2.sum(0)
        assert _rel(got[0], o.sum(0)) < 1e-4 and _rel(got[1], (o * o).sum(0)) < 1e-4


def test_bn_finalize_partials(cuda_dev):
    from rnagan_b200 import ops
    C, M = 128, 4096
    g = torch.Generator(device="cpu").manual_seed(3)
    a = (torch.randn(M, C, generator=g) * 2 + 0.5).to(torch.bfloat16).to(cuda_dev)
    parts = ops.stats_ws(C, cuda_dev, slot=8)
    P = parts.shape[0]
    af = a.float()
    with torch.no_grad():
        rows = torch.arange(M, device=cuda_dev) % P
        parts.zero_()
        parts[:, 0].index_add_(0, rows, af)
        parts[:, 1].index_add_(0, rows, af * af)
    gamma = (torch.rand(C, generator=g) + 0.5).to(cuda_dev)
    beta = torch.randn(C, generator=g).to(cuda_dev)
    rm, rv = torch.zeros(C, device=cuda_dev), torch.ones(C, device=cuda_dev)
    nbt = torch.zeros((), dtype=torch.int64, device=cuda_dev)
    f = lambda *s: torch.zeros(*s, device=cuda_dev)
    sums, mean, rstd, scale, shift = f(2, C), f(C), f(C), f(C), f(C)
    ops.bn_finalize_partials(parts, gamma, beta, M, C, 1e-5, 0.1, rm, rv, nbt, sums, mean, rstd, scale, shift)
    m_ref, v_ref = af.mean(0), af.var(0, unbiased=False)
    assert _rel(mean, m_ref) < 1e-4 and _rel(rstd, (v_ref + 1e-5).rsqrt()) < 1e-3
    assert _rel(scale, gamma * (v_ref + 1e-5).rsqrt()) < 1e-3
    assert _rel(rm, 0.1 * m_ref) < 1e-4 and _rel(rv, 0.9 + 0.1 * af.var(0, unbiased=True)) < 1e-3
    assert int(nbt.item()) == 1 and _rel(sums[0], af.sum(0)) < 1e-4


def test_pack_up_from_down_and_adam_shadow(cuda_dev):
    """The K-major w_up copy derived from the bf16 w_down equals the one packed from the fp32 weight, and the fused
    Adam step re-emits the bf16 image of a channels_last weight (= w_down) while updating the fp32 master."""
    from rnagan_b200 import ops
    Cp, Cs = 128, 64
    g = torch.Generator(device="cpu").manual_seed(21)
    Wt = _bf(torch.randn(Cp, Cs, 4, 4, generator=g) * 0.05).to(cuda_dev)
    w_down, w_up = ops.pack_link(Wt)
    w_up2 = torch.zeros_like(w_up)
    ops.pack_up_from_down(w_down, w_up2, Cs)
    assert torch.equal(w_up, w_up2)
    Wn = Wt.contiguous(memory_format=torch.channels_last)
    assert torch.equal(ops.cast_pad_bf16(ops.phys2d(Wn)), w_down)
    # Adam with a shadow
    p = (torch.randn(Cp, Cs, 4, 4, generator=g) * 0.05).to(cuda_dev).contiguous(memory_format=torch.channels_last)
    gr = torch.randn(Cp, Cs, 4, 4, generator=g).to(cuda_dev).contiguous(memory_format=torch.channels_last)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    shadow = torch.zeros(Cp, 16 * Cs, dtype=torch.bfloat16, device=cuda_dev)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3, betas=(0.5, 0.999))
    pr.grad = gr.clone()
    opt.step()
    tab = ops.AdamTable([p], [gr], [m], [v], [shadow])
    tab.step(1e-3, 0.5, 0.999, 1e-8, 1)
    assert _rel(p, pr.detach()) < 1e-5
    assert torch.equal(shadow, ops.cast_pad_bf16(ops.phys2d(p)))


@pytest.mark.parametrize("B,H,W,Cp", [(4, 64, 64, 128), (2, 16, 16, 64), (3, 8, 8, 192), (16, 32, 32, 128)])
def test_conv_up_merged_phases(cuda_dev, B, H, W, Cp):
    """rg_conv_up with the merged-phase operand (Cs == 64: all four output phases in one tile) equals
    conv_transpose2d, and its fused statistics equal the sums over the stored output."""
    from rnagan_b200 import ops
    Cs = 64
    g = torch.Generator(device="cpu").manual_seed(B + H + Cp + 9)
    x = _bf(torch.randn(B, Cp, H, W, generator=g)).to(cuda_dev)
    Wt = _bf(torch.randn(Cp, Cs, 4, 4, generator=g) * 0.05).to(cuda_dev)
    ref = F.conv_transpose2d(x, Wt, stride=2, padding=1)
    w_down, _ = ops.pack_link(Wt, want_up=False)
    w9 = ops.pack_up9_from_down(w_down, Cs)
    ws = ops.stats_ws(Cs, cuda_dev, slot=9)
    ws.fill_(float("nan"))
    out = ops.conv_up(_nhwc(x), w9, Cs, stats=ws)
    assert out.shape == (B, 2 * H, 2 * W, Cs)
    assert _rel(_nchw(out), ref) < 8e-3
    o = out.float().view(-1, Cs)
    got = ws.sum(0)
    assert _rel(got[0], o.sum(0)) < 1e-4 and _rel(got[1], (o * o).sum(0)) < 1e-4


@pytest.mark.parametrize("B,H,W,Cs,Cp", [(8, 16, 16, 128, 256), (4, 32, 32, 64, 128), (3, 8, 8, 64, 128)])
def test_fused_bn_backward_epilogue(cuda_dev, B, H, W, Cs, Cp):
    """rg_conv_down / rg_conv_up with rg_epilogue_aux mode 2: the stored output is du = dh * lrelu'(scale*a + shift)
    and the partial sums are S(du), S(du * xhat) -- compared with the same quantities from torch ops."""
    from rnagan_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(B + H + Cs + Cp + 13)
    Wt = _bf(torch.randn(Cp, Cs, 4, 4, generator=g) * 0.05).to(cuda_dev)
    w_down, w_up = ops.pack_link(Wt)
    slope = 0.2

    def check(out, dh_ref_nchw, a_nhwc, C, ws, mean, rstd, scale, shift):
        dh = _nhwc(dh_ref_nchw).float().view(-1, C)
        a = a_nhwc.float().view(-1, C)
        u = a * scale + shift
        du = torch.where(u > 0, dh, dh * slope)
        o = out.float().view(-1, C)
        assert _rel(o, du) < 1e-2
        xhat = (a - mean) * rstd
        got = ws.sum(0)
        assert _rel(got[0], o.sum(0)) < 1e-3
        assert _rel(got[1], (o * xhat).sum(0)) < 1e-3

    def params(C):
        mean = torch.randn(C, generator=g).to(cuda_dev) * 0.1
        rstd = (torch.rand(C, generator=g) + 0.5).to(cuda_dev)
        gamma = (torch.rand(C, generator=g) + 0.5).to(cuda_dev)
        beta = torch.randn(C, generator=g).to(cuda_dev) * 0.1
        scale = gamma * rstd
        shift = beta - mean * scale
        return mean, rstd, scale.contiguous(), shift.contiguous()

    # down: out [B,H,W,Cp]
    x = _bf(torch.randn(B, Cs, 2 * H, 2 * W, generator=g)).to(cuda_dev)
    a = torch.randn(B, H, W, Cp, generator=g).to(cuda_dev).to(torch.bfloat16)
    mean, rstd, scale, shift = params(Cp)
    ws = ops.stats_ws(Cp, cuda_dev, slot=11)
    ws.fill_(float("nan"))
    out = ops.conv_down(_nhwc(x), w_down, stats=ws, aux=("bn", a, mean, rstd, scale, shift, slope))
    check(out, F.conv2d(x, Wt, stride=2, padding=1), a, Cp, ws, mean, rstd, scale, shift)
    # up: out [B,2H,2W,Cs] (K-major operand; merged-phase operand when Cs == 64 and the grid is large enough)
    y = _bf(torch.randn(B, Cp, H, W, generator=g)).to(cuda_dev)
    a2 = torch.randn(B, 2 * H, 2 * W, Cs, generator=g).to(cuda_dev).to(torch.bfloat16)
    mean, rstd, scale, shift = params(Cs)
    ws2 = ops.stats_ws(Cs, cuda_dev, slot=11)
    ref_up = F.conv_transpose2d(y, Wt, stride=2, padding=1)
    operands = [w_down] + ([w_up] if Cs <= 128 else [])
    if Cs == 64 and B * H * W >= 256:
        operands.append(ops.pack_up9_from_down(w_down, Cs))
    for w in operands:
        ws2.fill_(float("nan"))
        out = ops.conv_up(_nhwc(y), w, Cs, stats=ws2, aux=("bn", a2, mean, rstd, scale, shift, slope))
        check(out, ref_up, a2, Cs, ws2, mean, rstd, scale, shift)
        # mode 1: plain LeakyReLU backward from the stored activation
        out1 = ops.conv_up(_nhwc(y), w, Cs, aux=("lrelu", a2, slope))
        dh = _nhwc(ref_up).float()
        ref1 = torch.where(a2.float() > 0, dh, dh * slope)
        assert _rel(out1.float(), ref1) < 1e-2


@pytest.mark.parametrize("layer", [1, 2, 3, 4, 5])
def test_full_size_adjoint_identities(cuda_dev, layer):
    """BASELINE config 2 layer shapes (B = 64, 256x256 images), where a CPU reference would take minutes: the three
    contractions of a link must be mutually adjoint,
        <conv_down(x; W), y> = <x, conv_up(y; W)>          (dgrad is the transpose of fprop)
        <wgrad(y, x), V>     = <conv_down(x; V), y>        (wgrad is the derivative w.r.t. the weight)
    for every weight layout the engines use (w_down K-major / MN-major, w_up, merged-phase w_up9, native gradient).
    bf16 outputs round each element to 2^-9 relative; the inner products over >= 2M elements average that out."""
    from rnagan_b200 import ops
    B = 64
    chans = [64, 128, 256, 512, 1024, 2048]
    Cs, Cp = chans[layer - 1], chans[layer]
    h = 128 >> layer
    g = torch.Generator(device="cpu").manual_seed(40 + layer)
    x = torch.randn(B, 2 * h, 2 * h, Cs, generator=g).to(cuda_dev).to(torch.bfloat16)      # hi side
    y = torch.randn(B, h, h, Cp, generator=g).to(cuda_dev).to(torch.bfloat16)              # lo side
    scale = (16 * Cs) ** -0.5
    Wt = _bf(torch.randn(Cp, Cs, 4, 4, generator=g) * scale).to(cuda_dev)
    V = _bf(torch.randn(Cp, Cs, 4, 4, generator=g) * scale).to(cuda_dev)
    w_down, w_up = ops.pack_link(Wt, want_up=Cs <= 128)
    v_down, _ = ops.pack_link(V, want_up=False)

    def dot(a, b):
        return (a.double() * b.double()).sum().item()

    lhs = dot(ops.conv_down(x, w_down), y)
    operands = [w_down] + ([w_up] if Cs <= 128 else []) + ([ops.pack_up9_from_down(w_down, Cs)] if Cs == 64 else [])
    for w in operands:
        rhs = dot(x, ops.conv_up(y, w, Cs))
        assert abs(lhs - rhs) <= 2e-3 * (abs(lhs) + abs(rhs)) + 1e-2 * (x.numel() ** 0.5), (layer, w.dim(), lhs, rhs)
    ref_w = dot(ops.conv_down(x, v_down), y)
    dW = torch.empty(Cp, Cs, 4, 4, device=cuda_dev)
    dWn = torch.empty(Cp, Cs, 4, 4, device=cuda_dev).contiguous(memory_format=torch.channels_last)
    for d in (dW, dWn):
        ops.conv_wgrad(y, x, d)
        got = dot(d, V)
        assert abs(got - ref_w) <= 2e-3 * (abs(got) + abs(ref_w)) + 1e-2 * (x.numel() ** 0.5), (layer, got, ref_w)
    assert torch.equal(dW, dWn)          # both layouts come from the same fixed-order reduction
