import sys, numpy as np, torch
sys.path.insert(0, '.')
from b200seg import prefilter
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for shape in ((128, 512, 512), (59, 350, 640)):
    for dt in (torch.uint16, torch.uint8):
        v = torch.randint(0, 255, shape, device='cuda', dtype=torch.int32).to(dt)
        V = v.numel()
        tg = t(lambda: prefilter.gaussian_filter(v, 1)); tm = t(lambda: prefilter.median_filter(v, 3))
        print(shape, dt, "gauss %.3f ms (%.0f GB/s alg)  median %.3f ms (%.0f GB/s alg)" % (tg, 2*v.element_size()*V/tg/1e6, tm, 2*v.element_size()*V/tm/1e6))
v = torch.randint(0, 4000, (128,512,512), device='cuda', dtype=torch.int32).to(torch.uint16)
print("zscore %.3f ms" % t(lambda: prefilter.zscore_norm(v)))
p = torch.rand((14, 32, 128, 128), device='cuda')
print("prm_u8 %.3f ms" % t(lambda: prefilter.prm_to_uint8(p)))
