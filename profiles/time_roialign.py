"""Times RoIAlign3D forward / backward (fp32, bf16) on BASELINE config 4 with the bench's method (CUDA events, L2 flushed
before every call).  `python profiles/time_roialign.py [P]` prints one JSON line; used while tuning roialign3d.cu."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import b200seg
from b200seg import synth
from b200seg.roi_align_3d import roialign3d_forward, roialign3d_backward

P = int(sys.argv[1]) if len(sys.argv) > 1 else 7
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
feat, rois = synth.roialign_case(1004)
f, r = torch.from_numpy(feat).to(dev), torch.from_numpy(rois).to(dev)
R, C = rois.shape[0], feat.shape[1]
g = torch.randn((R, C, P, P, P), device=dev)
fb, gb = f.bfloat16(), g.bfloat16()
peak = bench.hbm_peak()[0]
res = {"P": P}
for name, fn, e in (("fwd_f32", lambda: roialign3d_forward(f, r, P, P, P, 0.125, 2), 4),
                    ("bwd_f32", lambda: roialign3d_backward(g, r, feat.shape, 0.125, 2), 4),
                    ("fwd_bf16", lambda: roialign3d_forward(fb, r, P, P, P, 0.125, 2), 2),
                    ("bwd_bf16", lambda: roialign3d_backward(gb, r, feat.shape, 0.125, 2), 2)):
    ms = bench.time_op(torch, fn, 20, flush)
    nbytes = (R * C * P ** 3 + feat.size) * e + 28 * R
    res[name] = {"ms": round(ms, 4), "gbs": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / peak, 3),
                 "Mrois_s": round(R / ms / 1e3, 2)}
print(json.dumps(res))
