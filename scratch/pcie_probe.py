"""Pinned host <-> device copy rates of this box (what bounds the e2e leg)."""
import torch, time
n = 402653184 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
for name, src, dst in (("h2d", h, d), ("d2h", d, h)):
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dst.copy_(src, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(name, "%.2f ms for 403 MB = %.1f GB/s" % (ms, 0.402653184 / ms * 1e3))
# both directions at once
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.float32).pin_memory(); d2 = torch.empty(n, dtype=torch.float32, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
print("duplex %.2f ms for 403 MB each way" % ms)
