"""SM clock / power under a sustained run of one conv kernel (nvidia-smi sampled every 100 ms)."""
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rnagan_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
B, h, Cs, Cp = 64, 16, 256, 512
hi = torch.randn(B, 2 * h, 2 * h, Cs, device=dev).to(BF)
W = torch.randn(Cp, Cs, 4, 4, device=dev) * 0.05
wd, _ = ops.pack_link(W, want_up=False)
out = torch.empty(B, h, h, Cp, dtype=BF, device=dev)
for _ in range(10):
    ops.conv_down(hi, wd, out=out)
torch.cuda.synchronize()
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader",
                        "-lms", "100"], stdout=subprocess.PIPE, text=True)
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
t0 = time.time()
n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < secs:
    for _ in range(200):
        ops.conv_down(hi, wd, out=out)
    n += 200
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
smi.terminate()
lines = smi.stdout.read().strip().splitlines()
ms = e0.elapsed_time(e1)
print(f"{n} launches in {ms:.0f} ms: {ms / n * 1e3:.1f} us/launch = {2.0 * B * h * h * Cp * 16 * Cs / (ms / n) / 1e9:.0f} TF/s sustained")
print("clock samples (MHz, W, power-cap):", lines[::4])
