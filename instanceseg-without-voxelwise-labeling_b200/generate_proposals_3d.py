"""RPN proposal generation behind the reference's name (lib/modeling/generate_proposals_3d.py:12-177).

GenerateProposalsOp_3d(anchors, spatial_scale)(rpn_cls_prob, rpn_bbox_pred, im_info) -> (rois, roi_probs, scores_keep_idx)
as numpy arrays, like the reference; `.forward_device(...)` keeps everything on the GPU.  The settings the reference
reads from the global cfg (cfg.TRAIN / cfg.TEST .RPN_PRE_NMS_TOP_N, .RPN_POST_NMS_TOP_N, .RPN_NMS_THRESH, .RPN_MIN_SIZE)
are read from `core.config.cfg` when that module is importable (i.e. inside the reference tree) and otherwise from
the constructor arguments (defaults = lib/core/config.py:138-155, 211-224)."""
import ctypes as C

import numpy as np

from . import _lib


class GenerateProposalsOp_3d(object):
    def __init__(self, anchors, spatial_scale, pre_nms_topN=12000, post_nms_topN=2000, nms_thresh=0.7, min_size=0, training=False):
        self._anchors = np.ascontiguousarray(anchors, dtype=np.float32).reshape(-1, 6)
        self._num_anchors = self._anchors.shape[0]
        self._feat_stride = 1. / spatial_scale
        self.training = training
        self._settings = dict(pre_nms_topN=int(pre_nms_topN), post_nms_topN=int(post_nms_topN), nms_thresh=float(nms_thresh),
                              min_size=float(min_size))

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def settings(self):
        try:
            from core.config import cfg                        # inside the reference tree: same source as :105-110
            k = cfg["TRAIN" if self.training else "TEST"]
            return dict(pre_nms_topN=int(k.RPN_PRE_NMS_TOP_N), post_nms_topN=int(k.RPN_POST_NMS_TOP_N),
                        nms_thresh=float(k.RPN_NMS_THRESH), min_size=float(k.RPN_MIN_SIZE))
        except ImportError:
            return dict(self._settings)

    def forward_device(self, rpn_cls_prob, rpn_bbox_pred, im_info):
        """cuda tensors in; returns (rois [n_images*cap, 7], probs [n_images*cap], keep_idx [n_images*cap] int64, counts [n_images] int32,
        cap): image i owns rows [i*cap, i*cap + counts[i])."""
        import torch
        L = _lib.lib()
        s = self.settings()
        sc = rpn_cls_prob.detach().to(device="cuda", dtype=torch.float32).contiguous()
        dl = rpn_bbox_pred.detach().to(device="cuda", dtype=torch.float32).contiguous()
        N, A, S, H, W = sc.shape
        if A != self._num_anchors or tuple(dl.shape) != (N, 6 * A, S, H, W):
            raise ValueError("generate_proposals: scores %s / deltas %s do not match %d anchors" % (tuple(sc.shape), tuple(dl.shape), self._num_anchors))
        info = np.ascontiguousarray(im_info.detach().cpu().numpy() if hasattr(im_info, "detach") else im_info, dtype=np.float32).reshape(N, 4)
        args = (A, S, H, W, s["pre_nms_topN"], s["post_nms_topN"], s["nms_thresh"])
        cap = L.b200seg_generate_proposals_capacity(*args)
        if cap < 0:
            raise _lib.B200SegError("generate_proposals: unsupported geometry A=%d S=%d H=%d W=%d" % (A, S, H, W))
        ws_bytes = L.b200seg_generate_proposals_workspace_bytes(*args)
        d = sc.device
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=d)
        rois = torch.empty((N * cap, 7), dtype=torch.float32, device=d)
        probs = torch.empty(N * cap, dtype=torch.float32, device=d)
        keep_idx = torch.empty(N * cap, dtype=torch.int64, device=d)
        counts = torch.empty(N, dtype=torch.int32, device=d)
        _lib.check(L.b200seg_generate_proposals_dev(_lib.ptr(sc), _lib.ptr(dl), info.ctypes.data_as(C.c_void_p), N, A, S, H, W,
                                                    self._anchors.ctypes.data_as(C.c_void_p), self._feat_stride,
                                                    s["pre_nms_topN"], s["post_nms_topN"], s["nms_thresh"], s["min_size"],
                                                    _lib.ptr(rois), _lib.ptr(probs), _lib.ptr(keep_idx), _lib.ptr(counts),
                                                    _lib.ptr(ws), ws_bytes, _lib.current_stream()), "generate_proposals")
        return rois, probs, keep_idx, counts, cap

    def forward(self, rpn_cls_prob, rpn_bbox_pred, im_info):
        rois, probs, keep_idx, counts, cap = self.forward_device(rpn_cls_prob, rpn_bbox_pred, im_info)
        n = counts.cpu().numpy()
        r, p, k = rois.cpu().numpy(), probs.cpu().numpy(), keep_idx.cpu().numpy()
        out_r = np.concatenate([r[i * cap:i * cap + n[i]] for i in range(len(n))], axis=0) if len(n) else np.empty((0, 7), np.float32)
        out_p = np.concatenate([p[i * cap:i * cap + n[i]] for i in range(len(n))], axis=0)[:, None] if len(n) else np.empty((0, 1), np.float32)
        last = len(n) - 1                                        # the reference returns the LAST image's indices (:99-102)
        return out_r, out_p, k[last * cap:last * cap + n[last]].copy()

    __call__ = forward
