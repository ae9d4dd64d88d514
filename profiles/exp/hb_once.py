import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch, bench, b200seg
from b200seg.binarization import set_host_batch_out
cases = [bench.make_case(2000 + i) for i in range(2)]
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
keep, hc = [], []
for i in range(4):
    c = cases[i % 2]; d = {}
    for k in ("volume", "dets", "boxes", "prm", "crop_off"):
        t = pin(c[k]); keep.append(t); d[k] = t.numpy()
    hc.append(d)
segs = [torch.zeros(bench.SHAPE, dtype=torch.uint16).pin_memory().numpy() for _ in range(4)]
set_host_batch_out(2)
for _ in range(2):
    b200seg.postproc_soma_host_batch(hc, bench.NMS_THRESH, seg_out=segs)
