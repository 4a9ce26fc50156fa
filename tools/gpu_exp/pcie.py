import time, torch
n = 4096 * 1088
d = torch.zeros(n, dtype=torch.float32, device='cuda')
h = torch.zeros(n, dtype=torch.float32).pin_memory()
for name, f in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
    for _ in range(3): f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50): f()
    torch.cuda.synchronize()
    el = (time.perf_counter() - t0) / 50
    print(name, "%.1f MB in %.3f ms = %.1f GB/s" % (n * 4 / 1e6, el * 1e3, n * 4 / el / 1e9))
# 4 chunks on 4 streams
ss = [torch.cuda.Stream() for _ in range(4)]
c = n // 4
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    for k, s in enumerate(ss):
        with torch.cuda.stream(s):
            h[k * c:(k + 1) * c].copy_(d[k * c:(k + 1) * c], non_blocking=True)
    torch.cuda.synchronize()
print("D2H 4 chunks/4 streams %.3f ms" % ((time.perf_counter() - t0) / 50 * 1e3))
