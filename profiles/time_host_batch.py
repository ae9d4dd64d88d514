"""Times b200seg_postproc_soma_host_batch (the e2e leg of bench.py) on 64 volumes of 128x512x512 built from a few
distinct synthetic cases, for each `host_batch_out` state, with the library's own wait breakdown (B200SEG_HB_TRACE).
    python profiles/time_host_batch.py [n_distinct_cases]"""
import os, sys, time
os.environ["B200SEG_HB_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, bench, b200seg
from b200seg.binarization import set_host_batch_out, set_host_batch_mode, host_batch_traffic
nd = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cases = [bench.make_case(2000 + i) for i in range(nd)]
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
keep, hc = [], []
for i in range(64):
    c = cases[i % nd]
    d = {}
    for k in ("volume", "dets", "boxes", "prm", "crop_off"):
        t = pin(c[k]); keep.append(t); d[k] = t.numpy()
    hc.append(d)
segs_t = [torch.empty(bench.SHAPE, dtype=torch.uint16).pin_memory() for _ in range(64)]
segs = [t.numpy() for t in segs_t]
V = int(np.prod(bench.SHAPE))
for mode, state in ((3, 2), (7, 0), (7, 2), (7, 1)):
    set_host_batch_mode(mode)
    set_host_batch_out(state)
    for _ in range(2):
        b200seg.postproc_soma_host_batch(hc, bench.NMS_THRESH, seg_out=segs)
    if state == 1:
        for s in segs: s[...] = 0
    t0 = time.perf_counter()
    n = 3 if state != 1 else 1
    for _ in range(n):
        b200seg.postproc_soma_host_batch(hc, bench.NMS_THRESH, seg_out=segs)
    dt = (time.perf_counter() - t0) / n
    h2d, d2h = host_batch_traffic()
    print("host_batch_mode=%d host_batch_out=%d: %.1f ms per 64 volumes = %.1f Gvox/s (up %.0f MB, down %.0f MB)" % (mode, state, dt * 1e3, 64 * V / dt / 1e9, h2d / 1e6, d2h / 1e6), flush=True)
set_host_batch_out(0)
set_host_batch_mode(7)
