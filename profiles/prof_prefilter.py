"""Driver for ncu captures of the prefilter kernels: one warm call + one profiled call each on a 128x512x512 uint16 volume."""
import sys
import torch
sys.path.insert(0, ".")
from b200seg import prefilter
v = torch.randint(0, 4096, (128, 512, 512), device="cuda", dtype=torch.int32).to(torch.uint16)
for _ in range(3):
    g = prefilter.gaussian_filter(v, 1)
    m = prefilter.median_filter(g, 3)
torch.cuda.synchronize()
print("ok", int(m.sum()))
