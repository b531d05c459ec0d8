import torch, time
n = 4096*900
d = torch.randn(n, device="cuda"); h = torch.empty(n).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=50):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def one():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
def two():
    m = n // 2
    with torch.cuda.stream(s1): h[:m].copy_(d[:m], non_blocking=True)
    with torch.cuda.stream(s2): h[m:].copy_(d[m:], non_blocking=True)
def four():
    m = n // 4
    for i, s in enumerate((s1, s2, s1, s2)):
        with torch.cuda.stream(s): h[i*m:(i+1)*m].copy_(d[i*m:(i+1)*m], non_blocking=True)
for name, fn in (("1 stream", one), ("2 streams", two), ("4 chunks / 2 streams", four)):
    dt = t(fn); print(f"{name}: {dt*1e6:.1f} us  {n*4/dt/1e9:.1f} GB/s")
