"""Pinned host -> device copy bandwidth in the shapes the batch entry point uses (33.5 MB volumes): one stream, two
streams, and with a concurrent device -> host stream."""
import torch, time
dev = torch.device("cuda")
n, sz = 32, 128 * 512 * 512
h = [torch.empty(sz, dtype=torch.uint8).pin_memory() for _ in range(n)]
d = [torch.empty(sz, dtype=torch.uint8, device=dev) for _ in range(n)]
hd = [torch.empty(sz // 8, dtype=torch.uint8).pin_memory() for _ in range(n)]
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def run(streams, down=False):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        with torch.cuda.stream(streams[i % len(streams)]):
            d[i].copy_(h[i], non_blocking=True)
        if down:
            with torch.cuda.stream(s3):
                hd[i].copy_(d[i][:sz // 8], non_blocking=True)
    torch.cuda.synchronize(); return n * sz / (time.perf_counter() - t0) / 1e9
for _ in range(2): run([s1])
print("H2D 1 stream : %.1f GB/s" % run([s1]))
print("H2D 2 streams: %.1f GB/s" % run([s1, s2]))
print("H2D 1 stream + D2H of 1/8: %.1f GB/s (H2D bytes only)" % run([s1], True))
