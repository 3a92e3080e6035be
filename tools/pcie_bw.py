import torch, time
x = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(x, non_blocking=True)), ("D2H", lambda: x.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize()
    print(name, "%.1f GB/s" % (10 * (256 << 20) / (time.perf_counter() - t) / 1e9))
