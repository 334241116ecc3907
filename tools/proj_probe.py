import sys, torch
sys.path.insert(0, '.')
from rnagan_b200 import ops
dev = torch.device('cuda:0')
B = 64
z = torch.randn(B, 2048, device=dev).to(torch.bfloat16)
da0 = torch.randn(B, 4, 4, 2048, device=dev).to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, dW in (("torch layout", torch.empty(2048, 2048, 4, 4, device=dev)),
                 ("native (channels_last)", torch.empty(2048, 2048, 4, 4, device=dev).contiguous(memory_format=torch.channels_last))):
    ops.proj_wgrad(z, da0, dW); ops.proj_wgrad(z, da0, dW)
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.proj_wgrad(z, da0, dW); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"proj_wgrad {name}: {ts[3]*1e3:.1f} us  {dW.numel()*4/ts[3]/1e6:.0f} GB/s")
