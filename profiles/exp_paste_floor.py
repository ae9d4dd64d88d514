"""Experiment: how much of paste_labels_kernel's time is the pure zero-fill walk (no instances) vs torch's fill of the same bytes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from b200seg import _lib

L = _lib.lib()
dev = torch.device("cuda")
nv, S, H, W, n_max = 8, 128, 512, 512, 800
seg = torch.empty((nv, S, H, W), dtype=torch.int16, device=dev)
det_off = torch.zeros(nv + 1, dtype=torch.int32, device=dev)            # no instance in any volume
boxes = torch.zeros((1, 6), dtype=torch.int32, device=dev)
ids = torch.ones(n_max, dtype=torch.int16, device=dev)
masks = torch.zeros(64, dtype=torch.uint8, device=dev)
moff = torch.zeros(2, dtype=torch.int64, device=dev)
surv = torch.zeros(n_max, dtype=torch.uint8, device=dev)
wsb = L.b200seg_paste_labels_workspace_bytes(nv, S, H, W, n_max)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(it):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / it


def paste():
    _lib.check(L.b200seg_paste_labels_dev(_lib.ptr(seg), nv, S, H, W, _lib.ptr(det_off), n_max, _lib.ptr(boxes), _lib.ptr(ids), _lib.ptr(masks),
                                          _lib.ptr(moff), None, None, _lib.ptr(surv), _lib.ptr(ws), wsb, _lib.current_stream()), "paste")


print("paste, no instances (bin + labels kernels + bitmap memset): %.4f ms" % timed(paste))
print("torch zero_ of the same 537 MB: %.4f ms" % timed(lambda: seg.zero_()))
