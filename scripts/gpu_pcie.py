import torch, time
dev = torch.device('cuda')
n = 4194304
for mb in (8, 32, 134, 189):
    h = torch.empty(mb * 1000 * 1000 // 4, dtype=torch.float32).pin_memory(); d = torch.empty_like(h, device=dev)
    for name, fn in (('H2D', lambda: d.copy_(h, non_blocking=True)), ('D2H', lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): fn()
        torch.cuda.synchronize(); el = (time.perf_counter() - t0) / 10
        print('%s %4d MB: %.1f GB/s' % (name, mb, mb / 1e3 / el))
hi = torch.empty(134 * 250000, dtype=torch.float32).pin_memory(); di = torch.empty_like(hi, device=dev)
ho = torch.empty(189 * 250000, dtype=torch.float32).pin_memory(); do = torch.empty_like(ho, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): di.copy_(hi, non_blocking=True)
    with torch.cuda.stream(s2): ho.copy_(do, non_blocking=True)
for _ in range(3): both()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): both()
torch.cuda.synchronize(); el = (time.perf_counter() - t0) / 10
print('duplex 134 MB H2D + 189 MB D2H: %.2f ms -> max %.3e steps/s' % (el * 1e3, n / el))
