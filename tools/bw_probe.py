import torch
dev = torch.device("cuda:0")
def timeit(fn, reps=10):
    fn(); fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)
x = torch.randn(1 << 29, device=dev, dtype=torch.bfloat16)  # 1 GiB
y = torch.empty_like(x)
t = timeit(lambda: y.copy_(x)); print(f"copy 1GiB: {2*x.numel()*2/t/1e6:.0f} GB/s")
t = timeit(lambda: x.sum()); print(f"sum  1GiB (read only): {x.numel()*2/t/1e6:.0f} GB/s")
t = timeit(lambda: y.zero_()); print(f"zero 1GiB (write only): {x.numel()*2/t/1e6:.0f} GB/s")
a = torch.randn(1 << 20, 64, device=dev, dtype=torch.bfloat16)
t = timeit(lambda: a.float().sum(0)); print(f"[1M,64] float().sum(0): {t*1e3:.1f} us")
t = timeit(lambda: a.sum(0, dtype=torch.float32)); print(f"[1M,64] sum(0,f32): {t*1e3:.1f} us = {a.numel()*2/t/1e6:.0f} GB/s")
z = torch.randn(1<<26, device=dev, dtype=torch.bfloat16)  # 128 MiB
t = timeit(lambda: z.sum()); print(f"sum 128MiB: {z.numel()*2/t/1e6:.0f} GB/s")
