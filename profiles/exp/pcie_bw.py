"""Experiment: H2D / D2H bandwidth of this box from pinned memory (what bounds the host-buffer entry points)."""
import time, torch
dev = torch.device("cuda")
for mb in (32, 256):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
        print("%s %d MB: %.1f GB/s" % (name, mb, n / dt / 1e9))
# both directions at once
h2 = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); d2 = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("both directions, 256 MB each: %.1f GB/s per direction" % ((256 << 20) / dt / 1e9))
