"""In-kernel role profile of the forward tile engine: where the producer / MMA issuer / epilogue of each CTA spend
their cycles (clock64 totals written by the kernel when rg_debug_set_prof is armed).  Usage: python tools/gemm_prof.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rnagan_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16


def run(name, fn):
    prof = torch.zeros(148, 12, dtype=torch.int64, device=dev)
    fn(); fn()
    torch.cuda.synchronize()
    _lib.lib().rg_debug_set_prof(prof.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    _lib.lib().rg_debug_set_prof(None)
    p = prof.cpu().double()
    act = p[p[:, 6] > 0]
    lead = p[p[:, 3] > 0]
    m = lambda t, i: t[:, i].mean().item() / 1e3
    print(f"{name:40s} {e0.elapsed_time(e1) * 1e3:7.1f} us | kclk: prod tot {m(act, 0):7.1f} wait_empty {m(act, 1):7.1f} "
          f"issue {m(act, 2):7.1f} | mma tot {m(lead, 3):7.1f} wait_full {m(lead, 4):7.1f} wait_tempty {m(lead, 5):7.1f} | "
          f"epi tot {m(act, 6):7.1f} wait_tfull {m(act, 7):7.1f} wait_store {m(act, 8):7.1f} tiles {m(act, 9) * 1e3:5.1f}",
          flush=True)


def main():
    B = 64
    chans = [64, 128, 256, 512, 1024, 2048]
    for i in range(5):
        Cs, Cp = chans[i], chans[i + 1]
        h = 128 >> (i + 1)
        hi = torch.randn(B, 2 * h, 2 * h, Cs, device=dev).to(BF)
        lo = torch.randn(B, h, h, Cp, device=dev).to(BF)
        W = torch.randn(Cp, Cs, 4, 4, device=dev) * 0.05
        wd, wu = ops.pack_link(W)
        out_lo, out_hi = torch.empty_like(lo), torch.empty_like(hi)
        run(f"conv_down L{i + 1}", lambda: ops.conv_down(hi, wd, out=out_lo))
        run(f"conv_up L{i + 1} w_down", lambda: ops.conv_up(lo, wd, Cs, out=out_hi))
        if Cs <= 128:
            run(f"conv_up L{i + 1} w_up", lambda: ops.conv_up(lo, wu, Cs, out=out_hi))
        if Cs == 64:
            w9 = ops.pack_up9_from_down(wd, Cs)
            run(f"conv_up L{i + 1} w_up9 (merged)", lambda: ops.conv_up(lo, w9, Cs, out=out_hi))
            st = ops.stats_ws(Cs, dev, slot=3)
            run(f"conv_up L{i + 1} w_up9 (merged) + stats", lambda: ops.conv_up(lo, w9, Cs, out=out_hi, stats=st))


if __name__ == "__main__":
    main()
