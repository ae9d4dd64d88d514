"""Device-resident 3D Mask R-CNN inference flow for ONE tile (BASELINE configs[1], lib/core/test.py:54-173 with
configs/cell_tracking_baseline: 64x200x200 tile, stride 8, 256 channels, 35 anchors, RPN top-1000 / NMS 0.15, RoIAlign3D 7^3
sr 2 for both heads, TEST.NMS 0.15, SCORE_THRESH 0.05, DETECTIONS_PER_IM 300, 14^3 class-agnostic masks).

The backbone, the RPN convolutions and the two heads are the reference's PyTorch model and are NOT the product: here they are
random-init stand-ins (`StandInHeads`).  What this module chains -- without leaving the GPU, without a host synchronisation
and with fixed shapes, so that the whole tile is one CUDA graph -- are the operator-layer steps between them:

    GenerateProposalsOp_3d.forward_device   (generate_proposals_3d.py:20-192)      b200seg_generate_proposals_dev
    RoIAlignFunction_3d 7^3                 (model_builder.py:280-281)             b200seg_roialign3d_fwd_dev
    bbox_transform_3d + clip                (boxes_3d.py:167-225)                  a few elementwise torch ops
    box_results_with_nms_and_limit          (core/test.py:806-878)                 b200seg_nms3d_dev + top-k threshold
    RoIAlignFunction_3d 7^3 on the detections (im_detect_mask, test.py:396-426)    b200seg_roialign3d_fwd_dev
    segm_results                            (core/test.py:886-945)                 b200seg_segm_paste_dev

Data-dependent counts stay on the device: every stage works on its capacity (1000 proposals, 300 detections); rows beyond a
count carry a score of -1 and an empty box, never pass a threshold, and -- because greedy NMS only lets a box suppress LOWER
scored ones -- cannot change which valid rows are kept.  `TileFlow.run` is what `box_results.box_results_with_nms_and_limit`
and `segm.segm_results` compute when fed the same scores / boxes / masks (tests/test_gpu_parity.py checks exactly that).
"""
import numpy as np

from . import _lib
from .generate_proposals_3d import GenerateProposalsOp_3d
from .roi_align_3d import roialign3d_forward
from .segm import gauss_table


def cell_tracking_anchors():
    """35 anchors like configs/cell_tracking_baseline (7 sizes x 5 (xy, z) aspect pairs), [x1,y1,z1,x2,y2,z2] about the origin."""
    out = []
    for s in (10, 27, 33, 38, 42, 46, 50):
        for rxy, rz in ((1.0, 0.5), (0.5, 0.5), (2.0, 0.5), (0.2, 0.5), (3.0, 2.0)):
            w, h, d = s * np.sqrt(rxy), s / np.sqrt(rxy), s * rz
            out.append([-(w - 1) / 2, -(h - 1) / 2, -(d - 1) / 2, (w - 1) / 2, (h - 1) / 2, (d - 1) / 2])
    return np.asarray(out, np.float32)


class StandInHeads(object):
    """Random-init stand-ins for the reference's box head / mask head (plain PyTorch, not the product): average pool + two
    linear maps for (class scores, box deltas); a 1x1x1 convolution + x2 nearest upsampling + sigmoid for the 14^3 masks."""

    def __init__(self, C, ncls, device, seed=0):
        import torch
        g = torch.Generator().manual_seed(seed)
        r = lambda *s: (torch.randn(*s, generator=g) / np.sqrt(s[-1])).to(device)
        self.w_cls, self.w_box, self.w_mask = r(ncls, C), r(6 * ncls, C) * 0.5, r(1, C)
        self.b_cls = torch.tensor([0.0] + [0.3] * (ncls - 1), device=device)

    def box(self, pooled):                                  # [R,C,7,7,7] -> softmax scores [R,ncls], deltas [R,6*ncls]
        import torch
        f = pooled.float().mean(dim=(2, 3, 4))
        return torch.softmax(f @ self.w_cls.t() * 4.0 + self.b_cls, dim=1), f @ self.w_box.t()

    def mask(self, pooled):                                 # [n,C,7,7,7] -> [n,1,14,14,14] probabilities
        import torch
        m = torch.einsum("ncdhw,kc->nkdhw", pooled.float(), self.w_mask)
        return torch.sigmoid(torch.nn.functional.interpolate(m, scale_factor=2, mode="nearest") * 2.0)


class TileFlow(object):
    def __init__(self, tile=(64, 200, 200), C=256, ncls=2, stride=8, pre_nms_topN=1000, post_nms_topN=1000, rpn_nms=0.15,
                 score_thresh=0.05, nms=0.15, dets_per_im=300, bbox_reg_weights=(10., 10., 10., 5., 5., 5.), M=14,
                 thresh_binarize=0.5, device="cuda", seed=0):
        import torch
        self.torch = torch
        self.tile, self.C, self.ncls, self.stride, self.M = tuple(tile), C, ncls, stride, M
        self.score_thresh, self.nms, self.dpi, self.thresh_binarize = float(score_thresh), float(nms), int(dets_per_im), float(thresh_binarize)
        self.dev = torch.device(device)
        self.anchors = cell_tracking_anchors()
        self.gp = GenerateProposalsOp_3d(self.anchors, 1.0 / stride, pre_nms_topN=pre_nms_topN, post_nms_topN=post_nms_topN,
                                         nms_thresh=rpn_nms)
        self.heads = StandInHeads(C, ncls, self.dev, seed)
        self.w = torch.tensor(bbox_reg_weights, device=self.dev)
        self.clip = float(np.log(1000. / 16.))
        S, H, W = self.tile
        self.im_info = np.array([[S, H, W, 1.0]], np.float32)
        self.bound = torch.tensor([W - 1, H - 1, S - 1, W - 1, H - 1, S - 1], dtype=torch.float32, device=self.dev)
        self.tab = torch.from_numpy(gauss_table(M)).to(self.dev)
        self.whs = torch.tensor([W, H, S], dtype=torch.int64, device=self.dev)
        self._nms_ws = None
        self.crops = None                                   # worst case: every detection covers the whole tile

    # boxes_3d.py:167-225 in torch (float32 like the numpy original; exp differs from numpy's by <= 1 ulp)
    def _decode(self, rois6, deltas_j):
        torch = self.torch
        wdt = rois6[:, 3:6] - rois6[:, 0:3] + 1.0
        ctr = rois6[:, 0:3] + 0.5 * wdt
        d = deltas_j / self.w
        pc = d[:, 0:3] * wdt + ctr
        ps = torch.exp(torch.clamp(d[:, 3:6], max=self.clip)) * wdt
        b = torch.cat([pc - 0.5 * ps, pc + 0.5 * ps - 1.0], dim=1)
        return torch.minimum(torch.clamp(b, min=0.0), self.bound)          # clip_tiled_boxes_3d

    def run(self, features, rpn_cls_prob, rpn_bbox_pred):
        """features [1,C,S/8,H/8,W/8], rpn_cls_prob [1,A,...], rpn_bbox_pred [1,6A,...] (cuda).  Returns a dict of DEVICE tensors:
        dets [dpi,7] (x1..z2,score; rows >= n_dets are zero), n_dets int32[1], keep_rows int64 [dpi] (proposal index of each
        detection), masks [dpi,1,M,M,M], crops uint8 (packed mask crops), crop_off int64 [dpi+1], boxes_i32 [dpi,6] expanded."""
        torch = self.torch
        L = _lib.lib()
        S, H, W = self.tile
        rois, probs, keep_idx, counts, cap = self.gp.forward_device(rpn_cls_prob, rpn_bbox_pred, self.im_info)
        row = torch.arange(cap, device=self.dev)
        valid = row < counts[0]
        rois = torch.where(valid[:, None], rois, torch.zeros_like(rois))
        pooled = roialign3d_forward(features, rois, 7, 7, 7, 1.0 / self.stride, 2)
        scores, deltas = self.heads.box(pooled)
        j = 1                                               # NUM_CLASSES = 2: one foreground class
        boxes = self._decode(rois[:, 1:7], deltas[:, 6 * j:6 * j + 6])
        sc = torch.where(valid & (scores[:, j] > self.score_thresh), scores[:, j], torch.full_like(scores[:, j], -1.0))
        dets = torch.cat([boxes, sc[:, None]], dim=1).contiguous()
        # ---- box_results_with_nms_and_limit (core/test.py:828-878) on the full capacity ----------------------
        if self._nms_ws is None:
            self._off = torch.tensor([0, cap], dtype=torch.int32, device=self.dev)
            self._nms_ws = torch.empty(L.b200seg_nms3d_workspace_bytes(1, cap), dtype=torch.uint8, device=self.dev)
            self._keep = torch.empty(cap, dtype=torch.int64, device=self.dev)
            self._cnt = torch.empty(1, dtype=torch.int32, device=self.dev)
        _lib.check(L.b200seg_nms3d_dev(_lib.ptr(dets), _lib.ptr(self._off), 1, cap, float(np.float32(self.nms)), 0, _lib.ptr(self._keep),
                                       _lib.ptr(self._cnt), None, _lib.ptr(self._nms_ws), self._nms_ws.numel(), _lib.current_stream()), "nms3d_dev")
        kept = torch.zeros(cap + 1, dtype=torch.bool, device=self.dev)
        kept.scatter_(0, torch.where(row < self._cnt[0], self._keep, torch.full_like(self._keep, cap)), True)     # entries beyond the count -> slot `cap`
        kept = kept[:cap] & (sc > self.score_thresh)
        s_kept = torch.where(kept, sc, torch.full_like(sc, -1.0))
        # limit to DETECTIONS_PER_IM: image_thresh = the dpi-th largest kept score (np.sort(...)[-dpi]); ties at the threshold stay
        top = torch.topk(s_kept, min(self.dpi, cap)).values
        n_kept = kept.sum()
        thr = torch.where(n_kept > self.dpi, top[-1], torch.full_like(top[-1], -1.0))
        final = kept & (s_kept >= thr)
        # compact in proposal order (cls_boxes rows keep the NMS's ascending index order): stable sort of the flags
        order = torch.sort((~final).to(torch.int8), stable=True).indices[:self.dpi]
        n_dets = torch.clamp(final.sum(), max=self.dpi).to(torch.int32).reshape(1)
        ok = torch.arange(self.dpi, device=self.dev) < n_dets[0]
        out_dets = torch.where(ok[:, None], dets[order], torch.zeros_like(dets[order]))
        # ---- mask branch (im_detect_mask: rois = detections, RoIAlign 7^3, head -> M^3) -------------------------
        mrois = torch.cat([torch.zeros_like(out_dets[:, :1]), out_dets[:, :6]], dim=1).contiguous()
        masks = self.heads.mask(roialign3d_forward(features, mrois, 7, 7, 7, 1.0 / self.stride, 2)).contiguous()
        # ---- segm_results (core/test.py:886-945): expand by (M+2)/M, int32 truncation, resize + threshold + clip ------
        b = out_dets[:, :6]                                  # expand_boxes (boxes_3d.py:271-292): float32 arithmetic, then int32 truncation
        half = (b[:, 3:6] - b[:, 0:3]) * 0.5
        ctr = (b[:, 3:6] + b[:, 0:3]) * 0.5
        half = half * ((self.M + 2.0) / self.M)
        bi = torch.cat([ctr - half, ctr + half], dim=1).to(torch.int32)
        bi = torch.where(ok[:, None], bi, torch.full_like(bi, -4))          # unused rows: empty after clipping
        lo = torch.clamp(bi[:, 0:3].long(), min=0)
        hi = torch.minimum(bi[:, 3:6].long() + 1, self.whs)
        ext = torch.clamp(hi - lo, min=0)
        vol = torch.where((ext > 0).all(dim=1) & ok, ext.prod(dim=1), torch.zeros_like(ext[:, 0]))
        crop_off = torch.zeros(self.dpi + 1, dtype=torch.int64, device=self.dev)
        crop_off[1:] = torch.cumsum(vol, 0)
        if self.crops is None:
            self.crops = torch.empty(self.dpi * S * H * W + 16, dtype=torch.uint8, device=self.dev)
            self._midx = torch.arange(self.dpi, dtype=torch.int32, device=self.dev)
        bi = bi.contiguous()
        _lib.check(L.b200seg_segm_paste_dev(_lib.ptr(masks), _lib.ptr(self._midx), _lib.ptr(bi), self.dpi, self.M, _lib.ptr(self.tab),
                                            float(np.float32(self.thresh_binarize)), S, H, W, _lib.ptr(self.crops), _lib.ptr(crop_off),
                                            _lib.current_stream()), "segm_paste_dev")
        return dict(dets=out_dets, n_dets=n_dets, keep_rows=order, masks=masks, crops=self.crops, crop_off=crop_off, boxes_i32=bi,
                    rois=rois, n_rois=counts, scores=scores, boxes=boxes, valid=valid)
