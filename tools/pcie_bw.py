import torch, time
n = 2_000_000
h = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device='cuda')
for name, fn in (('H2D', lambda: d.copy_(h, non_blocking=True)), ('D2H', lambda: h.copy_(d, non_blocking=True))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(name, '%.3f ms per 16 MB -> %.1f GB/s' % (ms, 16e6 / ms / 1e6))
