"""Times the Mask R-CNN tile flow (bench ops entry) and the share of its two RoIAlign3D forward calls."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, bench, b200seg, numpy as np
from b200seg.maskrcnn_flow import TileFlow
from b200seg.roi_align_3d import roialign3d_forward
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flow = TileFlow(tile=(64, 200, 200), C=256, dets_per_im=300, seed=1)
g = torch.Generator(device="cpu").manual_seed(0)
feat = torch.randn((1, 256, 8, 25, 25), generator=g).to(dev)
A = flow.anchors.shape[0]
cls = torch.rand((1, A, 8, 25, 25), generator=g).to(dev)
box = (torch.randn((1, 6 * A, 8, 25, 25), generator=g) * 0.3).to(dev)
out = flow.run(feat, cls, box)
ms = bench.time_op(torch, lambda: flow.run(feat, cls, box), 10, flush)
rois, probs, keep_idx, counts, cap = flow.gp.forward_device(cls, box, flow.im_info)
ms_ra = bench.time_op(torch, lambda: roialign3d_forward(feat, rois, 7, 7, 7, 0.125, 2), 20, flush)
w = (rois[:, 4:7] - rois[:, 1:4]).cpu().numpy() / 8
print("flow eager %.3f ms; RoIAlign 7^3 on %d proposal rows %.3f ms; roi size (feature voxels) median %s max %s" % (ms, rois.shape[0], ms_ra, np.median(w, 0), w.max(0)))
