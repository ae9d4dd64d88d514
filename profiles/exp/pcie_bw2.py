"""Experiment: why 32 MB H2D copies run at less than half the link rate on this box."""
import time, torch
dev = torch.device("cuda")
n = 32 << 20
hs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(16)]
for h in hs: h.fill_(1)
ds = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(4)]
def run(label, fn, total):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    print("%s: %.1f GB/s" % (label, total / (time.perf_counter() - t0) / 1e9))
run("16 distinct 32 MB sources, one stream", lambda: [ds[i % 4].copy_(hs[i], non_blocking=True) for i in range(16)], 16 * n)
run("same 32 MB source 16 times", lambda: [ds[i % 4].copy_(hs[0], non_blocking=True) for i in range(16)], 16 * n)
s = [torch.cuda.Stream() for _ in range(2)]
def two():
    for i in range(16):
        with torch.cuda.stream(s[i % 2]): ds[i % 4].copy_(hs[i], non_blocking=True)
run("16 distinct sources over two streams", two, 16 * n)
big = torch.empty(16 * n, dtype=torch.uint8).pin_memory(); big.fill_(1); dbig = torch.empty(16 * n, dtype=torch.uint8, device=dev)
run("one 512 MB copy", lambda: dbig.copy_(big, non_blocking=True), 16 * n)
run("512 MB source in 16 slices of 32 MB", lambda: [dbig[i * n:(i + 1) * n].copy_(big[i * n:(i + 1) * n], non_blocking=True) for i in range(16)], 16 * n)
run("512 MB source in 128 slices of 4 MB", lambda: [dbig[i * (n // 8):(i + 1) * (n // 8)].copy_(big[i * (n // 8):(i + 1) * (n // 8)], non_blocking=True) for i in range(128)], 16 * n)
