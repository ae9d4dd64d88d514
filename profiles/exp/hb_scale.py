"""Host batch entry point at a given volume count and host thread count (B200SEG_HOST_THREADS in the environment), with the
library's wait breakdown (B200SEG_HB_TRACE=1 in the environment): python profiles/exp/hb_scale.py n_volumes [mode[,mode...]] [seconds per mode]
Several copies can run side by side (CUDA_VISIBLE_DEVICES=i LOCAL_WORLD_SIZE=8) to load the host like an 8-rank job."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, bench, b200seg
from b200seg.binarization import set_host_batch_out, set_host_batch_mode, host_batch_traffic
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 64
modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [7]
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
nd = min(8, nv)
cases = [bench.make_case(2000 + i) for i in range(nd)]
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
keep, hc = [], []
for i in range(nv):
    c = cases[i % nd]
    d = {}
    for k in ("volume", "dets", "boxes", "prm", "crop_off"):
        t = pin(c[k]); keep.append(t); d[k] = t.numpy()
    hc.append(d)
segs_t = [torch.empty(bench.SHAPE, dtype=torch.uint16).pin_memory() for _ in range(nv)]
segs = [t.numpy() for t in segs_t]
V = int(np.prod(bench.SHAPE))
set_host_batch_out(2)
for mode in modes:
    set_host_batch_mode(mode)
    for _ in range(3):
        b200seg.postproc_soma_host_batch(hc, bench.NMS_THRESH, seg_out=segs)
    t0 = time.perf_counter()
    n = 0
    while n < 5 or time.perf_counter() - t0 < secs:
        b200seg.postproc_soma_host_batch(hc, bench.NMS_THRESH, seg_out=segs)
        n += 1
    dt = (time.perf_counter() - t0) / n
    print("%.1f mode=%d threads=%s volumes=%d: %.2f ms per call = %.1f Gvox/s, %.3f ms per volume (%d calls)" %
          (time.time() % 1000, mode, os.environ.get("B200SEG_HOST_THREADS", "auto"), nv, dt * 1e3, nv * V / dt / 1e9, dt * 1e3 / nv, n), flush=True)
set_host_batch_mode(7)
