"""Micro-benchmark of the HBM-bound kernels (BatchNorm passes, LeakyReLU backward, image-side im2col / col2im, Adam) at
the B=64, 256x256 layer shapes: CUDA events per launch, L2 flushed between launches, algorithmic bytes (DESIGN.md
section 3.2) / time against the measured HBM peak of MEASURED_PEAKS.json.
Usage: python tools/hbm_bench.py [filter] [--once]     (--once: one launch per case, for `ncu -k ...` captures)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rnagan_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
BF, F32 = torch.bfloat16, torch.float32
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ONCE = "--once" in sys.argv
BIG_ONLY = "--big" in sys.argv        # only the largest BatchNorm shape (ncu captures)
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def timeit(fn, reps=9):
    if ONCE:
        flush.zero_()
        fn()
        torch.cuda.synchronize()
        return float("nan")
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda._sleep(400000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print(f"{name:52s} {ms * 1e3:8.1f} us  {nbytes / 1e6:8.1f} MB  {gbs:8.1f} GB/s  {gbs / PEAK:5.2f} of HBM peak", flush=True)


def main(filt=""):
    B = 64
    shapes = [(B * 128 * 128, 64), (B * 64 * 64, 128), (B * 32 * 32, 256), (B * 16 * 16, 512), (B * 16, 2048)]
    for M, C in (shapes[:1] if BIG_ONLY else shapes):
        tag = f"[{M}x{C}]"
        g = torch.Generator(device=dev).manual_seed(M + C)
        a = torch.randn(M, C, generator=g, device=dev).to(BF)
        dh = torch.randn(M, C, generator=g, device=dev).to(BF)
        gO = torch.randn(M, C, generator=g, device=dev).to(BF)
        out1, out2 = torch.empty_like(a), torch.empty_like(a)
        mean = a.float().mean(0).contiguous()
        rstd = (a.float().var(0, unbiased=False) + 1e-5).rsqrt().contiguous()
        gamma = torch.ones(C, device=dev)
        scale, shift = (gamma * rstd).contiguous(), (-mean * rstd).contiguous()
        sums, q = torch.zeros(2, C, device=dev), torch.zeros(3, C, device=dev)
        dgam = torch.zeros(C, device=dev)
        el = M * C
        cases = [
            ("bn_act (2r+2w)", lambda: ops.bn_act(a, scale, shift, 0.2, out1, M, C), 4 * el),
            ("bn_bwd_reduce (4r)", lambda: ops.bn_bwd_reduce(dh, a, mean, rstd, scale, shift, 0.2, M, C, sums), 4 * el),
            ("bn_bwd_apply (4r+2w)", lambda: ops.bn_bwd_apply(dh, a, None, mean, rstd, scale, shift, 0.2, sums, M, C, out1,
                                                              None), 6 * el),
            ("bn_bwd_apply +du (4r+4w)", lambda: ops.bn_bwd_apply(dh, a, None, mean, rstd, scale, shift, 0.2, sums, M, C,
                                                                  out1, out2), 8 * el),
            ("lrelu_bwd (4r+2w)", lambda: ops.lrelu_bwd(dh, a, 0.2, out1, M, C), 6 * el),
            ("bn_gp_reduce (6r)", lambda: ops.bn_gp_reduce(dh, a, gO, mean, rstd, M, C, q), 6 * el),
            ("bn_gp_apply (6r+4w)", lambda: ops.bn_gp_apply(dh, a, gO, mean, rstd, gamma, scale, shift, 0.2, sums, q, M, C,
                                                            out1, out2, dgam, 0.0), 10 * el),
        ]
        for name, fn, nb in cases:
            if filt in name:
                report(f"{name} {tag}", timeit(fn), nb)
        del a, dh, gO, out1, out2
    # image side (B = 64, 3 x 256 x 256 <-> 64 x 128 x 128 x 64)
    S, H, C0 = 256, 128, 64
    npix = B * H * H
    x = torch.rand(B, 3, S, S, device=dev) * 2 - 1
    y = torch.tanh(torch.randn(B, 3, S, S, device=dev))
    col = torch.empty(npix, 64, dtype=BF, device=dev)
    colf = torch.randn(npix, 48, device=dev)
    img = torch.empty(B, 3, S, S, device=dev)
    eps = torch.tensor([0.3], device=dev)
    bias = torch.zeros(3, device=dev)
    u8 = torch.empty(B, S, S, 3, dtype=torch.uint8, device=dev)
    L = ops._lib.lib()
    img_cases = [
        ("im2col_img plain (12r+128w B/px)", lambda: ops.im2col_img(x, col), B * 3 * S * S * 4 + npix * 128),
        ("im2col_img eps-mix (24r+128w)", lambda: ops.im2col_img(x, col, y=y, mode=1, eps_dev=eps),
         2 * B * 3 * S * S * 4 + npix * 128),
        ("col2im_img tanh (192r+48w B/px)", lambda: ops._lib.check(L.rg_col2im_img(
            colf.data_ptr(), 48, bias.data_ptr(), 1, B, 3, H, H, img.data_ptr(), ops._st()), "col2im"), npix * (192 + 48)),
        ("col2im_img tanh u8 (192r+12w)", lambda: ops._lib.check(L.rg_col2im_img(
            colf.data_ptr(), 48, bias.data_ptr(), 1 | 4, B, 3, H, H, u8.data_ptr(), ops._st()), "col2im"), npix * (192 + 12)),
    ]
    lo = torch.randn(B, H, H, C0, device=dev).to(BF)
    W0 = torch.randn(C0, 3, 4, 4, device=dev) * 0.1
    W0p = ops.img_conv_up_pack(W0)
    b64 = torch.zeros(C0, device=dev)
    out_lo = torch.empty(B, H, H, C0, dtype=BF, device=dev)
    unit = torch.empty(B, S, S, 3, device=dev)
    dW = torch.empty(C0, 3, 4, 4, device=dev)
    db = torch.empty(C0, device=dev)
    img_cases += [
        ("img_conv_up fused tanh NCHW (128r+48w B/px)", lambda: ops.img_conv_up(lo, W0p, img, bias=bias, act_tanh=True, Cimg=3),
         npix * (128 + 48)),
        ("img_conv_up fused tanh unit NHWC (128r+48w)", lambda: ops.img_conv_up(lo, W0p, unit, bias=bias, act_tanh=True,
                                                                              unit_nhwc=True, Cimg=3), npix * (128 + 48)),
        ("img_conv_up fused tanh u8 (128r+12w)", lambda: ops.img_conv_up(lo, W0p, u8, bias=bias, act_tanh=True, u8=True, Cimg=3),
         npix * (128 + 12)),
        ("img_conv_down fused plain (48r+128w B/px)", lambda: ops.img_conv_down(x, W0, out_lo, bias=b64, slope=0.2),
         npix * (48 + 128)),
        ("img_conv_down fused eps-mix (96r+128w)", lambda: ops.img_conv_down(x, W0, out_lo, y=y, mode=1, eps_dev=eps,
                                                                         bias=b64, slope=0.2), npix * (96 + 128)),
        ("img_conv_wgrad fused +dbias (128r+48r B/px)", lambda: ops.img_conv_wgrad(lo, x, dW, dbias=db), npix * (128 + 48)),
    ]
    for name, fn, nb in img_cases:
        if filt in name:
            report(name, timeit(fn), nb)
    # Adam over the generator's 112 M parameters (30 B / parameter with the bf16 shadow)
    if filt in "adam":
        n = 112 * 1000 * 1000
        p = torch.randn(n, device=dev)
        gr = torch.randn(n, device=dev)
        m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        sh = torch.empty(n, dtype=BF, device=dev)
        tab = ops.AdamTable([p], [gr], [m], [v], [sh])
        report("adam_step 112M params (16r+12w+2w B/param)", timeit(lambda: tab.step(1e-4, 0.5, 0.999, 1e-8, 3)), 30 * n)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    main(args[0] if args else "")
