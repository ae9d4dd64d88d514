"""Experiment: transfer schemes of b200seg_postproc_soma_host_batch (b200seg_set_option("host_batch_mode")) on pinned buffers."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np, torch, b200seg
from b200seg import synth
from b200seg.binarization import set_host_batch_mode, host_batch_traffic
mode = int(sys.argv[1]); shape = tuple(int(x) for x in sys.argv[2].split("x")); nv = int(sys.argv[3])
nb = int(sys.argv[4]) if len(sys.argv) > 4 else 200
set_host_batch_mode(mode)
t0 = time.perf_counter()
cases = [synth.postproc_case(3000 + i, shape=shape, n_blobs=nb, n_dup=int(2.5 * nb), n_false=nb // 2) for i in range(2)]
print("synth %.1f s" % (time.perf_counter() - t0), flush=True)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
hc = [{k: pin(c[k]) for k in ("volume", "dets", "boxes", "prm", "crop_off")} for c in cases]
e2e = [{k: h[k].numpy() for k in h} for h in hc] * (nv // 2)
segs = [torch.empty(shape, dtype=torch.uint16).pin_memory().numpy() for _ in range(nv)]
print("pinned", flush=True)
for _ in range(2):
    b200seg.postproc_soma_host_batch(e2e, 0.23, seg_out=segs)
    print("warm", flush=True)
t0 = time.perf_counter()
for _ in range(3):
    out = b200seg.postproc_soma_host_batch(e2e, 0.23, seg_out=segs)
dt = (time.perf_counter() - t0) / 3
V = shape[0] * shape[1] * shape[2]
print("mode", mode, "ms/step %.2f" % (dt * 1e3), "Gvox/s %.2f" % (nv * V / dt / 1e9), "traffic", host_batch_traffic(), "dets", cases[0]["dets"].shape,
      "keep", out[0]["n_keep"], flush=True)
