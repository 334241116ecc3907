"""Micro-benchmark of the tcgen05 tile engine on the layer shapes of the B=64, 256x256 workload (CUDA events, L2
flushed between launches).  Usage: python tools/gemm_bench.py [filter]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rnagan_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=8):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda._sleep(400000)      # ~0.2 ms of GPU idle-spin: the host enqueues fn() behind it
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, flops, bytes_):
    print(f"{name:44s} {ms * 1e3:8.1f} us  {flops / ms / 1e9:8.1f} TF/s  {bytes_ / ms / 1e6:8.1f} GB/s", flush=True)


def main(filt=""):
    B = 64
    chans = [64, 128, 256, 512, 1024, 2048]
    H = 128
    cases = []
    for i in range(5):
        Cs, Cp = chans[i], chans[i + 1]
        h = H >> (i + 1)
        cases.append((f"L{i + 1}", B, h, h, Cs, Cp))
    for name, B_, h, w, Cs, Cp in cases:
        hi = torch.randn(B_, 2 * h, 2 * w, Cs, device=dev).to(BF)
        lo = torch.randn(B_, h, w, Cp, device=dev).to(BF)
        W = torch.randn(Cp, Cs, 4, 4, device=dev) * 0.05
        wd, wu = ops.pack_link(W)
        fl = 2.0 * B_ * h * w * Cp * 16 * Cs
        by = (hi.numel() + lo.numel()) * 2 + W.numel() * 2
        dW = torch.empty_like(W)
        out_lo, out_hi = torch.empty_like(lo), torch.empty_like(hi)
        if filt in f"down {name}":
            report(f"conv_down {name} {Cs}->{Cp} @{h}", timeit(lambda: ops.conv_down(hi, wd, out=out_lo)), fl, by)
        if filt in f"up {name}":
            report(f"conv_up   {name} {Cp}->{Cs} @{h} (w_down, MN-major B)", timeit(lambda: ops.conv_up(lo, wd, Cs, out=out_hi)), fl, by)
            if Cs <= 128:
                report(f"conv_up   {name} {Cp}->{Cs} @{h} (w_up, K-major B)", timeit(lambda: ops.conv_up(lo, wu, Cs, out=out_hi)), fl, by)
            if Cs == 64:
                w9 = ops.pack_up9_from_down(wd, Cs)
                report(f"conv_up   {name} {Cp}->{Cs} @{h} (w_up9, merged phases)", timeit(lambda: ops.conv_up(lo, w9, Cs, out=out_hi)), fl, by)
        if filt in f"wgrad {name}":
            report(f"conv_wgrad {name} (torch layout)", timeit(lambda: ops.conv_wgrad(lo, hi, dW)), fl, by + W.numel() * 2)
            dWn = torch.empty_like(W).contiguous(memory_format=torch.channels_last)
            report(f"conv_wgrad {name} (native layout)", timeit(lambda: ops.conv_wgrad(lo, hi, dWn)), fl, by + W.numel() * 2)
            report(f"conv_wgrad {name} (native, beta=1)", timeit(lambda: ops.conv_wgrad(lo, hi, dWn, beta=1.0)), fl, by + W.numel() * 6)
    # image-side / projection GEMMs
    M = B * 128 * 128
    A = torch.randn(M, 64, device=dev).to(BF)
    Wc = (torch.randn(64, 64, device=dev) * 0.1).to(BF)
    bias = torch.randn(64, device=dev)
    o16 = torch.empty(M, 64, dtype=BF, device=dev)
    o32 = torch.empty(M, 48, device=dev)
    if filt in "edge nt":
        report("gemm_nt [1M,64]x[64,64] bf16 plain", timeit(lambda: ops.gemm_nt(A, Wc, out=o16)), 2.0 * M * 64 * 64, M * 256)
        report("gemm_nt [1M,64]x[64,64] bf16 +bias", timeit(lambda: ops.gemm_nt(A, Wc, out=o16, col_shift=bias)), 2.0 * M * 64 * 64, M * 256)
        report("gemm_nt [1M,64]x[64,64] bf16 +bias+lrelu", timeit(lambda: ops.gemm_nt(A, Wc, out=o16, col_shift=bias, slope=0.2)), 2.0 * M * 64 * 64, M * 256)
        report("gemm_nt [1M,64]x[64,64] bf16 lrelu only", timeit(lambda: ops.gemm_nt(A, Wc, out=o16, slope=0.2)), 2.0 * M * 64 * 64, M * 256)
        report("gemm_nt [1M,64]x[48,64] fp32", timeit(lambda: ops.gemm_nt(A, Wc[:48].contiguous(), out=o32, N=48)), 2.0 * M * 48 * 64, M * (128 + 192))
    if filt in "edge tn":
        C = torch.empty(64, 64, device=dev)
        report("gemm_tn [1M,64]^T[1M,64]", timeit(lambda: ops.gemm_tn(A, o16, out=C)), 2.0 * M * 64 * 64, M * 256)
    if filt in "proj":
        z = torch.randn(B, 2048, device=dev).to(BF)
        Wp = (torch.randn(16 * 2048, 2048, device=dev) * 0.02).to(BF)
        a0 = torch.empty(B, 16 * 2048, dtype=BF, device=dev)
        report("gemm_nt proj [64,2048]x[32768,2048]", timeit(lambda: ops.gemm_nt(z, Wp, out=a0)), 2.0 * B * 2048 * 32768, Wp.numel() * 2)
        da0 = torch.randn(B, 4, 4, 2048, device=dev).to(BF)
        dW = torch.empty(2048, 2048, 4, 4, device=dev)
        report("proj_wgrad", timeit(lambda: ops.proj_wgrad(z, da0, dW)), 2.0 * B * 2048 * 32768, dW.numel() * 4)
    if filt in "ew im2col":
        img = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
        img2 = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
        col = torch.empty(M, 64, dtype=BF, device=dev)
        eps = torch.tensor([0.3], device=dev)
        report("im2col mode0", timeit(lambda: ops.im2col_img(img, col)), 0, img.numel() * 4 + M * 128)
        report("im2col mode1 (interp)", timeit(lambda: ops.im2col_img(img, col, y=img2, mode=1, eps_dev=eps)), 0, img.numel() * 8 + M * 128)
        colf = torch.randn(M, 48, device=dev)
        out = torch.empty(B, 3, 256, 256, device=dev)
        bias3 = torch.zeros(3, device=dev)
        from rnagan_b200 import _lib
        report("col2im (+bias+tanh)", timeit(lambda: _lib.check(_lib.lib().rg_col2im_img(colf.data_ptr(), 48, bias3.data_ptr(), 1, B, 3, 128, 128, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "x")), 0, M * 192 + out.numel() * 4)
        a = torch.randn(M, 64, device=dev).to(BF)
        h = torch.empty_like(a)
        sums = torch.zeros(2, 64, device=dev)
        sc, sh = torch.ones(64, device=dev), torch.zeros(64, device=dev)
        report("bn_stats [1M,64]", timeit(lambda: ops.bn_stats(a, M, 64, sums)), 0, M * 128)
        report("bn_act [1M,64]", timeit(lambda: ops.bn_act(a, sc, sh, 0.2, h, M, 64)), 0, M * 256)
        report("bn_bwd_reduce [1M,64]", timeit(lambda: ops.bn_bwd_reduce(h, a, sh, sc, sc, sh, 0.2, M, 64, sums)), 0, M * 256)
        report("bn_bwd_apply [1M,64]", timeit(lambda: ops.bn_bwd_apply(h, a, None, sh, sc, sc, sh, 0.2, sums, M, 64, h)), 0, M * 384)
    if filt in "encoder":
        x = torch.randn(B, 19200, device=dev).to(BF)
        W1 = (torch.randn(6000, 19200, device=dev) * 0.01).to(BF)
        sc, sh = torch.ones(6000, device=dev), torch.zeros(6000, device=dev)
        o = torch.zeros(B, 6016, dtype=BF, device=dev)
        report("encoder L1 [64,19200]x[6000,19200]", timeit(lambda: ops.gemm_nt(x, W1, out=o, col_scale=sc, col_shift=sh, slope=0.01, N=6000)), 2.0 * B * 19200 * 6000, W1.numel() * 2)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "")
